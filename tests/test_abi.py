"""CPU: the C-ABI library builds, loads and exports every symbol include/kzgb200.h declares; the host-side
helpers (Keccak, byte conversions) agree with the oracle.  No device compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.py import bn254 as bn, inputs, keccak

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kzgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kzg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_path):
    lib = C.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header(lib_path):
    from kzg_grandsums_study_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_no_cpu_fallback(lib_path):
    """without a CUDA device the context cannot be created (and nothing else can run)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kzg_grandsums_study_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.kzg_ctx_create(0, None, C.byref(h)) != 0
    from kzg_grandsums_study_b200.curve import Curve
    with pytest.raises(Exception, match="no CPU fallback"):
        Curve(0)


def test_host_keccak_matches_oracle(lib_path):
    from kzg_grandsums_study_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for ln in [0, 1, 31, 32, 64, 135, 136, 137, 271, 272, 273, 1000]:
        data = rng.integers(0, 256, ln, dtype=np.uint8).tobytes()
        out = bytearray(32)
        lib.kzg_keccak256(_lib.as_ptr(data), ln, _lib.as_ptr(out))
        assert bytes(out) == keccak.keccak256(data), ln
    out = bytearray(32)
    lib.kzg_keccak256(None, 0, _lib.as_ptr(out))
    assert out.hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"


def test_host_rpr_conversions(lib_path):
    from kzg_grandsums_study_b200 import _lib
    lib = _lib.load()
    g = inputs.SplitMix64(3)
    for _ in range(20):
        x = g.fr()
        out = bytearray(32)
        lib.kzg_fr_to_rpr_be(_lib.as_ptr(bn.fr_to_mont_bytes(x)), _lib.as_ptr(out))
        assert bytes(out) == x.to_bytes(32, "big")
        P = bn.g1_mul_gen(x)
        out64 = bytearray(64)
        lib.kzg_g1_to_rpr_uncompressed(_lib.as_ptr(bn.g1_to_bytes(P)), _lib.as_ptr(out64))
        assert bytes(out64) == bn.g1_to_rpr_uncompressed(bn.g1_to_bytes(P))
    out64 = bytearray(64)
    lib.kzg_g1_to_rpr_uncompressed(_lib.as_ptr(bytes(64)), _lib.as_ptr(out64))
    assert bytes(out64) == bn.g1_to_rpr_uncompressed(bytes(64))
    # hash -> Fr: values above r (up to 2^256 - 1) must be reduced
    for h in [bytes(32), b"\xff" * 32, (bn.R).to_bytes(32, "big"), (bn.R - 1).to_bytes(32, "big"),
              (5 * bn.R + 7).to_bytes(32, "big"), keccak.keccak256(b"x")]:
        out = bytearray(32)
        lib.kzg_fr_from_hash_be(_lib.as_ptr(h), _lib.as_ptr(out))
        assert bytes(out) == bn.fr_to_mont_bytes(int.from_bytes(h, "big") % bn.R)


def test_synthetic_generator_matches_oracle():
    from kzg_grandsums_study_b200 import synthetic
    for seed, n in [(1, 5), (6, 300), (12345, 1500)]:
        a = synthetic.random_fr_std(seed, n)
        assert a.tobytes() == inputs.to_std_bytes(inputs.random_column(seed, n))
    assert synthetic.tau_from_seed(1001) == inputs.tau_from_seed(1001)
    assert list(synthetic.permutation(4, 50)) == inputs.permutation(4, 50)
