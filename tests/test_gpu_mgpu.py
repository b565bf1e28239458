"""GPU: the multi-GPU MSM inside one process (kzg_mgpu_*, csrc/mgpu.cu) -- what a single-process host such as the N-API
addon calls for G1.multiExpAffine over a sharded SRS (polynomial.js:1106-1115).  On a one-GPU box the same code runs with
several contexts on device 0 (devices = [0, 0, 0]): every shard, the threaded enqueue, the peer copies of the partials and
the final sum are exercised; with two or more GPUs visible the real split runs as well."""
import ctypes as C

import numpy as np
import pytest

from oracle.py import bn254 as bn, inputs

pytestmark = pytest.mark.gpu
R = bn.R


def _closed_form(tau, scalars):
    return bn.g1_to_bytes(bn.g1_mul_gen(sum(s * pow(tau, i, R) for i, s in enumerate(scalars)) % R))


def _device_sets():
    import torch
    sets = [[0], [0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append([0, 1])
    if torch.cuda.device_count() >= 4:
        sets.append([0, 1, 2, 3])
    return sets


class Mgpu:
    def __init__(self, devices):
        from kzg_grandsums_study_b200 import _lib
        self.lib = _lib.load()
        self.h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        rc = self.lib.kzg_mgpu_create(arr, len(devices), C.byref(self.h))
        assert rc == 0, rc

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.kzg_mgpu_last_error(self.h).decode())

    def close(self):
        self.lib.kzg_mgpu_destroy(self.h)


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_mgpu_msm_vs_closed_form(devices, tau):
    from kzg_grandsums_study_b200._lib import as_ptr
    m = Mgpu(devices)
    try:
        n = 5003                                   # not divisible by the device count: uneven shards
        assert m.lib.kzg_mgpu_device_count(m.h) == len(devices)
        m.check(m.lib.kzg_mgpu_srs_generate(m.h, as_ptr(tau.to_bytes(32, "little")), n))
        assert m.lib.kzg_mgpu_srs_len(m.h) == n
        pos = 0
        for g in range(len(devices)):
            first, count = C.c_uint64(), C.c_uint64()
            m.check(m.lib.kzg_mgpu_shard(m.h, g, C.byref(first), C.byref(count)))
            assert first.value == pos
            pos += count.value
        assert pos == n
        scalars = inputs.random_column(77, n)
        host = bn.fr_vec_to_std_bytes(scalars)
        out = bytearray(64)
        m.check(m.lib.kzg_mgpu_srs_msm_host(m.h, as_ptr(host), n, as_ptr(out)))
        assert bytes(out) == _closed_form(tau, scalars)
        # fewer scalars than SRS points: the trailing devices get empty shards
        for k in (0, 1, n // 3 + 1):
            m.check(m.lib.kzg_mgpu_srs_msm_host(m.h, as_ptr(host), k, as_ptr(out)))
            assert bytes(out) == _closed_form(tau, scalars[:k]), k
        # resident scalars
        m.check(m.lib.kzg_mgpu_scalars_upload(m.h, as_ptr(host), n))
        out2 = bytearray(64)
        for _ in range(2):
            m.check(m.lib.kzg_mgpu_srs_msm(m.h, as_ptr(out2)))
            assert bytes(out2) == _closed_form(tau, scalars)
        with pytest.raises(RuntimeError, match="more scalars than SRS points"):
            m.check(m.lib.kzg_mgpu_srs_msm_host(m.h, as_ptr(host + host), 2 * n, as_ptr(out)))
    finally:
        m.close()


def test_mgpu_ptau_shards_and_large_host_msm(tau, ptau_factory):
    """SRS shards read from a .ptau (prover.js:15-16,83-85) == generated shards; a 2^22-point host-scalar MSM over three
    contexts (each shard large enough for the piecewise upload and the affine rounds) == the single-context result"""
    import torch
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    from kzg_grandsums_study_b200.curve import getCurveFromName
    curve = getCurveFromName("bn128")
    m = Mgpu([0, 0, 0] if torch.cuda.device_count() < 3 else [0, 1, 2])
    try:
        path = ptau_factory(12)
        n = 1 << 13
        m.check(m.lib.kzg_mgpu_srs_load_ptau(m.h, path.encode(), n))
        scal = synthetic.random_fr_std(31, n)
        out = bytearray(64)
        m.check(m.lib.kzg_mgpu_srs_msm_host(m.h, as_ptr(scal), n, as_ptr(out)))
        srs, _ = curve.load_srs(path, n)
        want = bytearray(64)
        curve.check(curve.lib.kzg_srs_msm_host(curve.ctx, srs, 0, as_ptr(scal), n, as_ptr(want)))
        assert bytes(out) == bytes(want)
        # large: pinned through kzg_host_register
        n = 1 << 22
        m.check(m.lib.kzg_mgpu_srs_generate(m.h, as_ptr(tau.to_bytes(32, "little")), n))
        scal = np.ascontiguousarray(synthetic.random_fr_std(32, n))
        assert m.lib.kzg_host_register(as_ptr(scal), scal.nbytes) == 0
        try:
            m.check(m.lib.kzg_mgpu_srs_msm_host(m.h, as_ptr(scal), n, as_ptr(out)))
        finally:
            assert m.lib.kzg_host_unregister(as_ptr(scal)) == 0
        s1 = C.c_void_p()
        curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(s1)))
        curve.check(curve.lib.kzg_srs_precompute(curve.ctx, s1, 0))
        curve.check(curve.lib.kzg_srs_msm_host(curve.ctx, s1, 0, as_ptr(scal), n, as_ptr(want)))
        curve.lib.kzg_srs_free(curve.ctx, s1)
        assert bytes(out) == bytes(want)
    finally:
        m.close()
