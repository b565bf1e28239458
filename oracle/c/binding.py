"""ctypes binding of oracle/c/libkzgoracle.so -- ORACLE, test infrastructure only (tests/, smoke(), bench.py's CPU legs)."""
import ctypes as C
import os
import time

from . import build as _build

_lib = None
vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        l = C.CDLL(path)
        l.ko_max_threads.restype = i32
        l.ko_g1_msm.argtypes = [vp, vp, sz, vp, i32]
        l.ko_srs_generate.argtypes = [vp, C.c_uint64, sz, vp, i32]
        l.ko_fr_to_mont.argtypes = [vp, vp, sz, i32]
        l.ko_fr_from_mont.argtypes = [vp, vp, sz, i32]
        l.ko_fr_ntt.argtypes = [vp, i32, i32, i32]
        l.ko_fr_batch_inverse.argtypes = [vp, sz, i32]
        l.ko_keccak256.argtypes = [vp, sz, vp]
        l.ko_prove.argtypes = [i32, vp, sz, C.POINTER(vp), C.POINTER(vp), vp, vp, i32, i32, vp, vp, i32]
        l.ko_prove.restype = i32
        _lib = l
    return _lib


def _buf(b):
    return (C.c_char * len(b)).from_buffer_copy(b) if isinstance(b, (bytes, bytearray)) else b


def max_threads():
    return int(lib().ko_max_threads())


def msm(bases, scalars_std, threads=0):
    n = len(scalars_std) // 32
    out = C.create_string_buffer(64)
    lib().ko_g1_msm(_buf(bases), _buf(scalars_std), n, out, threads)
    return out.raw


def srs_generate(tau, n, first=0, threads=0):
    out = C.create_string_buffer(64 * n)
    lib().ko_srs_generate(_buf(int(tau).to_bytes(32, "little")), first, n, out, threads)
    return out.raw


def to_mont(b, threads=0):
    out = C.create_string_buffer(len(b))
    lib().ko_fr_to_mont(_buf(b), out, len(b) // 32, threads)
    return out.raw


def from_mont(b, threads=0):
    out = C.create_string_buffer(len(b))
    lib().ko_fr_from_mont(_buf(b), out, len(b) // 32, threads)
    return out.raw


def ntt(b, inverse=False, threads=0):
    n = len(b) // 32
    buf = C.create_string_buffer(bytes(b), len(b))
    lib().ko_fr_ntt(buf, n.bit_length() - 1, 1 if inverse else 0, threads)
    return buf.raw


def batch_inverse(b, threads=0):
    buf = C.create_string_buffer(bytes(b), len(b))
    lib().ko_fr_batch_inverse(buf, len(b) // 32, threads)
    return buf.raw


def keccak256(data):
    out = C.create_string_buffer(32)
    lib().ko_keccak256(_buf(bytes(data)), len(data), out)
    return out.raw


ERRORS = {-1: "Polynomial does not divide", -2: "polynomial is not well calculated", -3: "Polynomial is not divisible",
          -4: "bad arguments"}


def prove(kind, srs_bytes, cols_f, cols_t, sel_f=None, sel_t=None, threads=0):
    """kind: 'gs' | 'gp'; columns: lists of n*32 B standard-form buffers; selectors Montgomery or None.
    -> (proof_bytes: commitments then evaluations in key order, challenges: dict of 32 B Montgomery)"""
    k = len(cols_f)
    n = len(cols_f[0]) // 32
    nbits = n.bit_length() - 1
    gs = kind == "gs"
    selected = sel_f is not None
    ncm = 2 * k + (2 if selected else 0) + 4
    nev = (2 * k if gs else k) + (2 if selected else 0) + 1
    keep = [_buf(c) for c in cols_f] + [_buf(c) for c in cols_t]
    pf = (vp * k)(*[C.cast(b, vp) for b in keep[:k]])
    pt = (vp * k)(*[C.cast(b, vp) for b in keep[k:]])
    sf = _buf(sel_f) if selected else None
    st = _buf(sel_t) if selected else None
    proof = C.create_string_buffer(64 * ncm + 32 * nev)
    ch = C.create_string_buffer(160)
    srs = _buf(srs_bytes)
    rc = lib().ko_prove(0 if gs else 1, srs, len(srs_bytes) // 64, pf, pt, sf, st, nbits, k, proof, ch, threads)
    if rc != 0:
        raise ValueError(ERRORS.get(rc, "error %d" % rc))
    names = ["beta", "gamma", "alpha", "xi", "v"]
    return proof.raw, {nm: ch.raw[32 * i:32 * i + 32] for i, nm in enumerate(names)}


def bench_msm(log_n, tau_seed, scalar_seed, threads=None):
    """time one 2^log_n-point MSM (SRS + scalars generated outside the timed region) -> (Mpts/s, seconds, threads)"""
    from ..py import inputs
    import numpy as np
    n = 1 << log_n
    threads = threads or max_threads()
    tau = inputs.tau_from_seed(tau_seed)
    bases = C.create_string_buffer(64 * n)
    lib().ko_srs_generate(_buf(int(tau).to_bytes(32, "little")), 0, n, bases, threads)
    try:
        from kzg_grandsums_study_b200 import synthetic      # same generator, vectorised (agreement is a CPU test)
        scal = synthetic.random_fr_std(scalar_seed, n).tobytes()
    except Exception:
        scal = inputs.to_std_bytes(inputs.random_column(scalar_seed, n))
    sb = _buf(scal)
    out = C.create_string_buffer(64)
    t0 = time.perf_counter()
    lib().ko_g1_msm(bases, sb, n, out, threads)
    dt = time.perf_counter() - t0
    return n / dt / 1e6, dt, threads


def bench_prove(log_n, tau_seed, seed, threads=None, kind="gs"):
    """time one plain grand-sum / grand-product proof at n = 2^log_n on the host cores (SRS generated outside the
    timed region) -> (seconds, threads, sha256 of the proof bytes)"""
    import hashlib
    from ..py import inputs
    n = 1 << log_n
    threads = threads or max_threads()
    tau = inputs.tau_from_seed(tau_seed)
    srs = C.create_string_buffer(64 * 2 * n)
    lib().ko_srs_generate(_buf(int(tau).to_bytes(32, "little")), 0, 2 * n, srs, threads)
    from kzg_grandsums_study_b200 import synthetic
    f = synthetic.random_fr_std(seed, n)
    t = f[synthetic.permutation(seed, n)]
    fb, tb = _buf(f.tobytes()), _buf(t.tobytes())
    pf = (vp * 1)(C.cast(fb, vp))
    pt = (vp * 1)(C.cast(tb, vp))
    proof = C.create_string_buffer(64 * 6 + 32 * 3)
    ch = C.create_string_buffer(160)
    t0 = time.perf_counter()
    rc = lib().ko_prove(0 if kind == "gs" else 1, srs, 2 * n, pf, pt, None, None, log_n, 1, proof, ch, threads)
    dt = time.perf_counter() - t0
    if rc != 0:
        raise ValueError(ERRORS.get(rc, "error %d" % rc))
    return dt, threads, hashlib.sha256(proof.raw).hexdigest()
