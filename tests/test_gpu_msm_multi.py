"""GPU: several MSMs over one SRS window table as ONE pipeline (kzg_commit_many, SURVEY.md 8f-4: the reference commits
[F],[T] and [Wxi],[Wxiw] one after the other, prover.js:161-162,409-410), the chunked batched-affine rounds with their
inversions on a side stream, and the host-scalar partial of the multi-GPU MSM.  Every result against the closed form
p(tau) G1 from the oracle and against the one-at-a-time path."""
import ctypes as C

import numpy as np
import pytest

from oracle.py import bn254 as bn

pytestmark = pytest.mark.gpu
R = bn.R


def _srs(curve, tau, n, table_c=None):
    from kzg_grandsums_study_b200._lib import as_ptr
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    if table_c is not None:
        curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, table_c))
    return srs


def _commit_many(curve, srs, bufs):
    from kzg_grandsums_study_b200._lib import as_ptr
    k = len(bufs)
    handles = (C.c_void_p * k)(*[b.handle for b in bufs])
    out = bytearray(64 * k)
    curve.check(curve.lib.kzg_commit_many(curve.ctx, srs, handles, k, as_ptr(out)))
    return [bytes(out[64 * i:64 * i + 64]) for i in range(k)]


def _commit(curve, srs, buf):
    from kzg_grandsums_study_b200._lib import as_ptr
    out = bytearray(64)
    curve.check(curve.lib.kzg_commit(curve.ctx, srs, buf.handle, as_ptr(out)))
    return bytes(out)


def _closed_form(tau, coeffs):
    return bn.g1_to_bytes(bn.g1_mul_gen(sum(c * pow(tau, i, R) for i, c in enumerate(coeffs)) % R))


@pytest.mark.parametrize("table_c", [5, 11])
@pytest.mark.parametrize("rounds", [0, 2, 5])
def test_commit_many_vs_oracle(curve, tau, table_c, rounds):
    """k polynomials of DIFFERENT lengths (one empty, one all-zero, one with every coefficient equal, one longer than
    the others) through one merged pipeline == closed form == k separate commits; with and without affine rounds"""
    from oracle.py import inputs
    n = 3000
    srs = _srs(curve, tau, n, table_c)
    try:
        polys = [inputs.random_column(31, 1500), inputs.random_column(32, 3000), [0] * 700, [7] * 2048, [],
                 [R - 1] * 33, inputs.random_column(33, 1)]
        bufs = [curve.to_device(bn.fr_vec_to_mont_bytes(p)) for p in polys]
        want = [_closed_form(tau, p) for p in polys]
        curve.set_option("aff_rounds", rounds)
        for merge in (1, 0):
            curve.set_option("msm_merge", merge)
            assert _commit_many(curve, srs, bufs) == want, (table_c, rounds, merge)
        assert [_commit(curve, srs, b) for b in bufs] == want
        # more jobs than one pipeline takes (8): two merged groups on the two lanes
        curve.set_option("msm_merge", 1)
        many = bufs + bufs[:5]
        assert _commit_many(curve, srs, many) == want + want[:5]
    finally:
        curve.set_option("aff_rounds", -1)
        curve.set_option("msm_merge", -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)


def test_commit_many_without_table_and_degree_check(curve, tau):
    """no window table: the jobs cannot be merged and run on the two lanes, same results; a polynomial whose degree
    exceeds the SRS is refused with the reference's message (prover.js:79-81), trailing zeros beyond the SRS are fine"""
    from kzg_grandsums_study_b200 import KzgError
    from oracle.py import inputs
    n = 500
    srs = _srs(curve, tau, n)
    try:
        polys = [inputs.random_column(41, 500), inputs.random_column(42, 100), inputs.random_column(43, 400) + [0] * 300]
        bufs = [curve.to_device(bn.fr_vec_to_mont_bytes(p)) for p in polys]
        assert _commit_many(curve, srs, bufs) == [_closed_form(tau, p) for p in polys]
        too_long = curve.to_device(bn.fr_vec_to_mont_bytes([1] * 501))
        with pytest.raises(KzgError, match="not sufficiently large"):
            _commit_many(curve, srs, bufs + [too_long])
    finally:
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("log_n,k", [(18, 2), (20, 2), (19, 5)])
def test_commit_many_large_matches_single_and_closed_form(curve, tau, log_n, k):
    """sizes where the merged list crosses the affine-round threshold although the single lists do not"""
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200.polynomial import Polynomial
    n = 1 << log_n
    srs = _srs(curve, tau, n, 0)
    try:
        bufs = [curve.to_device(synthetic.random_fr_std(700 + i, n if i != 1 else n // 2 + 17).tobytes()) for i in range(k)]
        got = _commit_many(curve, srs, bufs)
        for i, b in enumerate(bufs):
            p_tau = bn.fr_from_mont_bytes(Polynomial(b, curve).evaluate(bn.fr_to_mont_bytes(tau)))
            assert got[i] == bn.g1_to_bytes(bn.g1_mul_gen(p_tau)), i
        curve.set_option("msm_merge", 0)
        assert _commit_many(curve, srs, bufs) == got
        assert [_commit(curve, srs, b) for b in bufs] == got
    finally:
        curve.set_option("msm_merge", -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("chunks", [1, 2, 3, 7, 16])
def test_affine_round_chunking(curve, tau, chunks):
    """the rounds launched in 1 .. 16 chunks (inversions on the side stream) give the same point; 2^18 points so that
    every chunk count really splits the thread range"""
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    from kzg_grandsums_study_b200.polynomial import Polynomial
    n = 1 << 18
    srs = _srs(curve, tau, n, 12)
    try:
        scal = synthetic.random_fr_std(808, n)
        buf = curve.to_device(scal.tobytes())
        p_tau = bn.fr_from_mont_bytes(Polynomial(buf, curve).evaluate(bn.fr_to_mont_bytes(tau)))
        want = bn.g1_to_bytes(bn.g1_mul_gen(p_tau))
        curve.set_option("aff_rounds", 3)
        curve.set_option("aff_chunks", chunks)
        assert _commit(curve, srs, buf) == want
    finally:
        curve.set_option("aff_rounds", -1)
        curve.set_option("aff_chunks", -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("log_n", [12, 22])
def test_host_partial_equals_resident_partial(curve, tau, log_n):
    """kzg_srs_msm_host_partial (scalars in host memory, piecewise upload from 2^21 points on) leaves the same group
    element as kzg_srs_msm_partial; combined with a second shard it gives the closed form of the whole MSM"""
    import torch
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    from kzg_grandsums_study_b200.polynomial import Polynomial
    lib, ctx = curve.lib, curve.ctx
    n = 1 << log_n
    half = n // 2 + 5
    scal = synthetic.random_fr_std(909, n)
    srs_a = srs_b = None
    try:
        srs_a = C.c_void_p()
        srs_b = C.c_void_p()
        curve.check(lib.kzg_srs_generate_range(ctx, as_ptr(tau.to_bytes(32, "little")), 0, half, C.byref(srs_a)))
        curve.check(lib.kzg_srs_generate_range(ctx, as_ptr(tau.to_bytes(32, "little")), half, n - half, C.byref(srs_b)))
        curve.check(lib.kzg_srs_precompute(ctx, srs_a, 0))
        curve.check(lib.kzg_srs_precompute(ctx, srs_b, 0))
        parts = torch.zeros(2 * 16, dtype=torch.int64, device="cuda")
        host_a = torch.from_numpy(scal[:half].copy().view(np.uint8).reshape(-1)).pin_memory()
        curve.check(lib.kzg_srs_msm_host_partial(ctx, srs_a, 0, as_ptr(host_a), half, C.c_void_p(parts.data_ptr())))
        dev_b = curve.to_device(scal[half:].tobytes())
        curve.check(lib.kzg_srs_msm_partial(ctx, srs_b, 0, dev_b.handle, n - half, C.c_void_p(parts.data_ptr() + 128)))
        out = bytearray(64)
        curve.check(lib.kzg_g1_partials_combine(ctx, C.c_void_p(parts.data_ptr()), 2, as_ptr(out)))
        # closed form through the Montgomery-coefficient route: commit-independent Horner evaluation on the device
        mont = curve.Fr.batchToMontgomery(curve.to_device(scal.tobytes()))
        p_tau = bn.fr_from_mont_bytes(Polynomial(mont, curve).evaluate(bn.fr_to_mont_bytes(tau)))
        assert bytes(out) == bn.g1_to_bytes(bn.g1_mul_gen(p_tau))
        # and the host partial alone == the resident partial of the same shard
        one = bytearray(64)
        two = bytearray(64)
        curve.check(lib.kzg_g1_partials_combine(ctx, C.c_void_p(parts.data_ptr()), 1, as_ptr(one)))
        dev_a = curve.to_device(scal[:half].tobytes())
        curve.check(lib.kzg_srs_msm(ctx, srs_a, 0, dev_a.handle, half, as_ptr(two)))
        assert bytes(one) == bytes(two)
    finally:
        for h in (srs_a, srs_b):
            if h:
                lib.kzg_srs_free(ctx, h)


def test_stream_ordering_entry_points(curve):
    """kzg_stream_wait_ctx / kzg_ctx_wait_stream accept any stream of the device (a side stream, the legacy default
    stream) and leave both streams usable; the multi-rank MSM relies on them around its all-gather"""
    import torch
    from kzg_grandsums_study_b200._lib import as_ptr
    lib, ctx = curve.lib, curve.ctx
    side = torch.cuda.Stream()
    buf = curve.alloc(1 << 20)
    one = bn.fr_to_mont_bytes(1)
    curve.check(lib.kzg_buf_fill(ctx, buf.handle, 0, 1 << 20, as_ptr(one)))
    for stream in (C.c_void_p(side.cuda_stream), None):
        curve.check(lib.kzg_stream_wait_ctx(ctx, stream))
        with torch.cuda.stream(side):
            torch.zeros(1 << 20, device="cuda").sum()
        curve.check(lib.kzg_ctx_wait_stream(ctx, stream))
    curve.sync()
    side.synchronize()
    raw = buf.tobytes()
    assert raw[:32] == one and raw[-32:] == one


def test_two_contexts_on_two_devices(tau):
    """every entry point switches to its context's device.  With two GPUs visible, two contexts in one
    process interleave calls and both produce the oracle's commitment."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    from kzg_grandsums_study_b200.curve import Curve
    from oracle.py import inputs
    a, b = Curve(0), Curve(1)
    try:
        n = 1 << 12
        coeffs = inputs.random_column(5150, n)
        want = _closed_form(tau, coeffs)
        sa, sb = _srs(a, tau, n, 0), _srs(b, tau, n, 0)
        ba = a.to_device(bn.fr_vec_to_mont_bytes(coeffs))
        bb = b.to_device(bn.fr_vec_to_mont_bytes(coeffs))
        for _ in range(3):
            assert _commit(a, sa, ba) == want
            assert _commit(b, sb, bb) == want
        assert torch.cuda.current_device() == 0
        a.lib.kzg_srs_free(a.ctx, sa)
        b.lib.kzg_srs_free(b.ctx, sb)
        del ba, bb
    finally:
        a.terminate()
        b.terminate()


@pytest.mark.parametrize("log_n,table_c,cuts", [(21, 0, None), (22, 0, None), (21, 14, None), (21, 0, (4, 20)), (21, 13, (9, 10))])
def test_host_scalar_pieces_share_one_reduction(curve, tau, log_n, table_c, cuts):
    """kzg_srs_msm_host cuts the scalars into two pieces (three from 2^24 points on; forced here with the cut knobs: at
    a / 64 and b / 64 of the points) whose uploads hide behind compute; with a window table every piece but the last hands
    its folded bucket sums to the last piece's reduction (host_link = 1, default) instead of reducing on its own (0).
    Same point either way, equal to the resident-scalar MSM -- also with skewed scalars whose buckets are empty in one
    piece and full in the other."""
    import torch
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    lib, ctx = curve.lib, curve.ctx
    n = 1 << log_n
    srs = _srs(curve, tau, n, table_c)
    try:
        scal = synthetic.random_fr_std(1234 + log_n, n)
        skew = scal.copy()
        skew[: n // 4] = 0                      # the whole first piece contributes nothing
        skew[n // 2:, 1:] = 0                   # 64-bit scalars: the upper windows are empty in the second piece
        for data in (scal, skew):
            host = torch.from_numpy(data.view(np.uint8).reshape(-1).copy()).pin_memory()
            dev = curve.to_device(data.tobytes())
            want = bytearray(64)
            curve.check(lib.kzg_srs_msm(ctx, srs, 0, dev.handle, n, as_ptr(want)))
            for link in (1, 0):
                curve.set_option("host_link", link)
                if cuts:
                    curve.set_option("host_cut_a", cuts[0])
                    curve.set_option("host_cut_b", cuts[1])
                got = bytearray(64)
                curve.check(lib.kzg_srs_msm_host(ctx, srs, 0, as_ptr(host), n, as_ptr(got)))
                assert bytes(got) == bytes(want), (log_n, table_c, link)
    finally:
        curve.set_option("host_link", -1)
        curve.set_option("host_cut_a", -1)
        curve.set_option("host_cut_b", -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("rounds", [1, 3])
def test_linked_pieces_with_affine_rounds(curve, tau, rounds):
    """three linked pieces whose walks run over the dense lists the batched-affine rounds leave (msm_accumulate_kernel
    <DIRECT, CARRY>: at the library's own sizes only pieces of 2^24-point MSMs get there): forced rounds, cuts at 8 / 64 and
    24 / 64, uniform and skewed scalars -- the same point as the resident-scalar MSM without rounds"""
    import torch
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    lib, ctx = curve.lib, curve.ctx
    n = 1 << 21
    srs = _srs(curve, tau, n, 12)
    try:
        scal = synthetic.random_fr_std(77 + rounds, n)
        skew = scal.copy()
        skew[n // 8: n // 4] = 0                # the middle piece starts with nothing
        skew[n // 2:, 1:] = 0                   # 64-bit scalars: most buckets empty in the last piece
        for data in (scal, skew):
            host = torch.from_numpy(data.view(np.uint8).reshape(-1).copy()).pin_memory()
            dev = curve.to_device(data.tobytes())
            want = bytearray(64)
            curve.set_option("aff_rounds", 0)
            curve.check(lib.kzg_srs_msm(ctx, srs, 0, dev.handle, n, as_ptr(want)))
            curve.set_option("aff_rounds", rounds)
            curve.set_option("host_cut_a", 8)
            curve.set_option("host_cut_b", 24)
            got = bytearray(64)
            curve.check(lib.kzg_srs_msm_host(ctx, srs, 0, as_ptr(host), n, as_ptr(got)))
            assert bytes(got) == bytes(want), rounds
    finally:
        for k in ("aff_rounds", "host_cut_a", "host_cut_b"):
            curve.set_option(k, -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)
