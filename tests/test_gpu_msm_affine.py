"""Batched-affine bucket rounds of the MSM (csrc/msm.cu: msm_aff_forward / fq_batch_inverse / msm_aff_backward).

The rounds switch themselves on from 2^22 points (48 M bucket entries); here they are forced on (KZGB200_AFF_ROUNDS)
at sizes the oracle can check, through the C ABI, for both MSM flavours.  The result replaces G1.multiExpAffine + G1.toAffine
(reference src/polynomial/polynomial.js:1106-1115) and must be the same canonical affine bytes whatever the number
of rounds.  Exceptional pairs (operand at infinity, P + P, P + (-P)) get their own SRS.
"""
import ctypes as C

import pytest

from oracle.py import bn254 as bn, inputs

pytestmark = pytest.mark.gpu
R = bn.R


def _srs_msm(curve, srs, scalars, n):
    from kzg_grandsums_study_b200._lib import as_ptr
    buf = curve.to_device(bn.fr_vec_to_std_bytes(scalars))
    out = bytearray(64)
    curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out)))
    return bytes(out)


@pytest.mark.parametrize("flavour,c", [("table", 4), ("table", 9), ("table", 13), ("raw", 5), ("raw", 11)])
@pytest.mark.parametrize("rounds", ["1", "2", "3", "6"])
def test_affine_rounds_vs_closed_form(curve, tau, flavour, c, rounds, monkeypatch):
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    n = 2600
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        if flavour == "table":
            curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, c))
        else:
            curve.check(curve.lib.kzg_msm_set_window(curve.ctx, c))
        scal = synthetic.random_fr_std(6100 + c, n)
        cases = [[sum(int(scal[i, j]) << (64 * j) for j in range(4)) for i in range(n)],
                 [5] * n,                                   # one bucket holds everything: a deep pairwise tree
                 [0] * n,                                   # no entries at all
                 [R - 1 - (i % 2) for i in range(n)],
                 [(i % 7) + 1 for i in range(n)]]           # seven buckets, odd and even sizes
        curve.set_option("aff_rounds", int(rounds))
        for scalars in cases:
            expect = sum(s * pow(tau, i, R) for i, s in enumerate(scalars)) % R
            want = bn.g1_to_bytes(bn.g1_mul_gen(expect))
            assert _srs_msm(curve, srs, scalars, n) == want, (flavour, c, rounds, scalars[0])
            mont = curve.to_device(bn.fr_vec_to_mont_bytes(scalars))
            out = bytearray(64)
            curve.check(curve.lib.kzg_commit(curve.ctx, srs, mont.handle, as_ptr(out)))
            assert bytes(out) == want, (flavour, c, rounds, "commit")
    finally:
        curve.set_option("aff_rounds", -1)
        curve.check(curve.lib.kzg_msm_set_window(curve.ctx, 0))
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("rounds", ["1", "2", "4"])
def test_affine_rounds_exceptional_pairs(curve, rounds, monkeypatch):
    """an SRS of G, infinity, G, -G repeated: every pair of a bucket is a doubling, a cancellation or has an operand at
    infinity, and the sums that come out of one round meet again in the next"""
    from kzg_grandsums_study_b200._lib import as_ptr
    G = bn.g1_to_bytes((1, 2))
    negG = bn.g1_to_bytes((1, bn.Q - 2))
    inf = bytes(64)
    n = 1200
    layouts = [((G + inf + G + negG) * 300, [1, 0, 1, -1] * 300),
               (G * n, [1] * n),                             # nothing but doublings, level after level
               ((G + negG) * 600, [1, -1] * 600),            # nothing but cancellations
               ((inf + inf + G) * 400, [0, 0, 1] * 400)]
    for pts, coeffs in layouts:
        srs = C.c_void_p()
        curve.check(curve.lib.kzg_srs_from_host(curve.ctx, as_ptr(pts), n, C.byref(srs)))
        try:
            for table_c in (0, 5):
                curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, table_c))
                curve.set_option("aff_rounds", int(rounds))
                for scalars in ([7] * n, inputs.random_column(3, n), [R - 1] * n, [(i % 5) + 1 for i in range(n)]):
                    expect = sum(s * c for s, c in zip(scalars, coeffs)) % R
                    assert _srs_msm(curve, srs, scalars, n) == bn.g1_to_bytes(bn.g1_mul_gen(expect)), (table_c, scalars[0])
                curve.set_option("aff_rounds", -1)
        finally:
            curve.set_option("aff_rounds", -1)
            curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("log_n", [18, 20])
def test_affine_rounds_match_the_xyzz_walk(curve, tau, log_n, monkeypatch):
    """forced rounds over millions of entries (multi-level batch inversion, the partition sort in front): same bytes as
    the plain XYZZ walk (KZGB200_AFF_ROUNDS=0), as the library's own choice, and as the closed form p(tau) G1"""
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    from kzg_grandsums_study_b200.polynomial import Polynomial
    n = 1 << log_n
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        p = Polynomial(synthetic.random_fr_std(91 + log_n, n).tobytes(), curve)
        p_tau = bn.fr_from_mont_bytes(p.evaluate(bn.fr_to_mont_bytes(tau)))
        want = bn.g1_to_bytes(bn.g1_mul_gen(p_tau))
        got = {}
        for rounds in ("0", "2", "4", None):
            if rounds is None:
                curve.set_option("aff_rounds", -1)
            else:
                curve.set_option("aff_rounds", int(rounds))
            got[rounds] = p.multiExponentiation(srs)
        assert all(v == want for v in got.values()), {k: v == want for k, v in got.items()}
    finally:
        curve.set_option("aff_rounds", -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)


def test_msm_plan_reports_the_rounds(curve, monkeypatch):
    """kzg_msm_plan: what an n-point MSM will do -- no rounds for small inputs, some for 2^24 points, and the override"""
    c, w, r = C.c_uint32(), C.c_uint32(), C.c_uint32()
    curve.set_option("aff_rounds", -1)
    for n, want in ((1, 0), (1 << 16, 0), (1 << 20, 0)):
        curve.check(curve.lib.kzg_msm_plan(curve.ctx, None, n, 0, C.byref(c), C.byref(w), C.byref(r)))
        assert r.value == want and w.value == -(-257 // c.value), (n, c.value, w.value, r.value)
    curve.check(curve.lib.kzg_msm_plan(curve.ctx, None, 1 << 24, 0, C.byref(c), C.byref(w), C.byref(r)))
    assert 1 <= r.value <= 6
    g_c, g_w = C.c_uint32(), C.c_uint32()
    curve.check(curve.lib.kzg_msm_geometry(curve.ctx, None, 1 << 24, 0, C.byref(g_c), C.byref(g_w)))
    assert (g_c.value, g_w.value) == (c.value, w.value)
    curve.set_option("aff_rounds", 2)
    curve.check(curve.lib.kzg_msm_plan(curve.ctx, None, 1000, 1, C.byref(c), C.byref(w), C.byref(r)))
    assert r.value == 2 and w.value == -(-255 // c.value)
    curve.set_option("aff_rounds", -1)


@pytest.mark.parametrize("kind", ["gs", "gp"])
@pytest.mark.parametrize("rounds", ["1", "3"])
def test_whole_proofs_with_forced_rounds(kind, rounds, curve, tau, ptau_factory, monkeypatch):
    """every commitment of a proof through the affine rounds (both lanes of the prover, Montgomery-source scalars):
    the proof stays byte-identical to the oracle's (reference src/grandsum|grandproduct/mset_eq_kzg_prover.js)"""
    from kzg_grandsums_study_b200.grandproduct import mset_eq_kzg_grandproduct_prover
    from kzg_grandsums_study_b200.grandsum import mset_eq_kzg_grandsum_prover
    from kzg_grandsums_study_b200.polynomial import Evaluations
    from oracle.py import protocol as pr
    nbits = 9
    n = 1 << nbits
    path = ptau_factory(nbits)
    f = inputs.random_column(40 + int(rounds), n)
    t = inputs.rotate_right(f)
    fb, tb = bn.fr_vec_to_std_bytes(f), bn.fr_vec_to_std_bytes(t)
    gpu = mset_eq_kzg_grandsum_prover if kind == "gs" else mset_eq_kzg_grandproduct_prover
    cpu = pr.grandsum_prover if kind == "gs" else pr.grandproduct_prover
    want = cpu(pr.TrapdoorSrs(tau, nbits), [fb], [tb])
    curve.set_option("aff_rounds", int(rounds))
    try:
        got = gpu(path, Evaluations(fb, curve), Evaluations(tb, curve))
    finally:
        curve.set_option("aff_rounds", -1)
    assert pr.proof_bytes(got) == pr.proof_bytes(want)


@pytest.mark.parametrize("pattern", ["uniform", "skewed", "two_values"])
def test_default_rounds_at_2_22_points_known_answer(curve, tau, pattern):
    """2^22 points over the window table: the rounds are on by the library's own choice.  Uniform scalars, three
    quarters of the scalars equal (one bucket with millions of entries: a pairwise tree over one run, then the
    block-tier collapse of its partial sums) and scalars taking two values only.  Known answer (sum s_i tau^i) G1 with
    the sum evaluated by the device Horner kernel (an independent path; bench.py's check)."""
    import numpy as np
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    n = 1 << 22
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, 0))
        c, w, r = C.c_uint32(), C.c_uint32(), C.c_uint32()
        curve.check(curve.lib.kzg_msm_plan(curve.ctx, srs, n, 0, C.byref(c), C.byref(w), C.byref(r)))
        assert r.value >= 1, "the batched-affine rounds should be on at 2^22 points"
        scal = synthetic.random_fr_std(77, n).copy()
        if pattern == "skewed":
            scal[: 3 * n // 4] = np.array([5, 0, 0, 0], dtype=np.uint64)
        elif pattern == "two_values":
            scal[0::2] = np.array([(1 << 20) + 3, 0, 0, 0], dtype=np.uint64)
            scal[1::2] = scal[1]
        buf = curve.to_device(scal.tobytes())
        out = bytearray(64)
        curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out)))
        # sum s_i tau^i: the standard-form scalars re-read as Montgomery residues have value s_i / 2^256, and the result
        # comes back as Montgomery bytes, so the raw little-endian integer is the sum itself
        ev = bytearray(32)
        tau_m = (tau << 256) % R
        curve.check(curve.lib.kzg_poly_evaluate(curve.ctx, buf.handle, as_ptr(tau_m.to_bytes(32, "little")), as_ptr(ev)))
        k = int.from_bytes(ev, "little") % R
        assert bytes(out) == bn.g1_to_bytes(bn.g1_mul_gen(k)), pattern
    finally:
        curve.lib.kzg_srs_free(curve.ctx, srs)
