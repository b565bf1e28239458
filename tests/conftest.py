import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib_path():
    from kzg_grandsums_study_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def curve(lib_path):
    """the device-backed curve object; fails loudly when the extension or the GPU is missing"""
    from kzg_grandsums_study_b200.curve import getCurveFromName
    c = getCurveFromName("bn128")
    yield c
    c.terminate()


TAU_SEED = 1001


@pytest.fixture(scope="session")
def tau():
    from oracle.py import inputs
    return inputs.tau_from_seed(TAU_SEED)


@pytest.fixture(scope="session")
def ptau_factory(curve, tau, tmp_path_factory):
    """synthetic .ptau files of a given power, made by the device SRS generator (G2 points from the oracle)"""
    import ctypes as C
    from kzg_grandsums_study_b200._lib import as_ptr
    from oracle.py import bn254 as bn
    cache = {}
    d = tmp_path_factory.mktemp("ptau")

    def make(power):
        if power in cache:
            return cache[power]
        path = str(d / ("synthetic_%02d.ptau" % power))
        n_pts = 1 << (power + 1)
        srs = C.c_void_p()
        curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n_pts, C.byref(srs)))
        g2_one = bn.g2_to_bytes(bn.G2_GEN)
        g2_tau = bn.g2_to_bytes(bn.g2_mul(bn.G2_GEN, tau))
        curve.check(curve.lib.kzg_srs_write_ptau(curve.ctx, srs, power, as_ptr(g2_one), as_ptr(g2_tau), path.encode()))
        curve.lib.kzg_srs_free(curve.ctx, srs)
        cache[power] = path
        return path

    return make
