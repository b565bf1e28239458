"""GPU: the Lagrange-basis SRS (kzg_srs_lagrange).  [L_i(tau)]_1 against the oracle's closed form
L_i(tau) G1 (the synthetic tau is known), and the property it exists for: committing a polynomial's EVALUATIONS over the
Lagrange SRS gives the same point as committing its coefficients (iNTT of the evaluations) over the monomial SRS --
the reference commits F, T, S through the iNTT (prover.js:151-162, grandsum.js:61)."""
import ctypes as C

import pytest

from oracle.py import bn254 as bn, inputs

pytestmark = pytest.mark.gpu
R = bn.R


def _lagrange_at(tau, n_bits):
    """L_i(tau) = (tau^n - 1) w^i / (n (tau - w^i))"""
    n = 1 << n_bits
    w = bn.FR_W[n_bits]
    zh = (pow(tau, n, R) - 1) % R
    n_inv = pow(n, -1, R)
    out = []
    wi = 1
    for _ in range(n):
        out.append(zh * wi % R * n_inv % R * pow((tau - wi) % R, -1, R) % R)
        wi = wi * w % R
    return out


@pytest.mark.parametrize("n_bits", [0, 1, 2, 5, 9])
def test_lagrange_srs_vs_closed_form(curve, tau, n_bits):
    from kzg_grandsums_study_b200._lib import as_ptr
    lib, ctx = curve.lib, curve.ctx
    n = 1 << n_bits
    srs, lag = C.c_void_p(), C.c_void_p()
    curve.check(lib.kzg_srs_generate(ctx, as_ptr(tau.to_bytes(32, "little")), n + 3, C.byref(srs)))
    try:
        curve.check(lib.kzg_srs_lagrange(ctx, srs, n_bits, C.byref(lag)))
        assert lib.kzg_srs_len(lag) == n
        raw = bytearray(64 * n)
        curve.check(lib.kzg_srs_download(ctx, lag, 0, n, as_ptr(raw)))
        want = b"".join(bn.g1_to_bytes(bn.g1_mul_gen(v)) for v in _lagrange_at(tau, n_bits))
        assert bytes(raw) == want
    finally:
        if lag:
            lib.kzg_srs_free(ctx, lag)
        lib.kzg_srs_free(ctx, srs)


@pytest.mark.parametrize("n_bits", [4, 12, 16])
def test_commit_of_evaluations_equals_commit_of_coefficients(curve, tau, n_bits):
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    lib, ctx = curve.lib, curve.ctx
    n = 1 << n_bits
    srs, lag = C.c_void_p(), C.c_void_p()
    curve.check(lib.kzg_srs_generate(ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        curve.check(lib.kzg_srs_lagrange(ctx, srs, n_bits, C.byref(lag)))
        curve.check(lib.kzg_srs_precompute(ctx, srs, 0))
        curve.check(lib.kzg_srs_precompute(ctx, lag, 0))
        evals = curve.to_device(synthetic.random_fr_std(1200 + n_bits, n).tobytes())   # any residues < r are Montgomery values
        coefs = curve.Fr.ifft(evals)
        a, b = bytearray(64), bytearray(64)
        curve.check(lib.kzg_commit(ctx, lag, evals.handle, as_ptr(a)))
        curve.check(lib.kzg_commit(ctx, srs, coefs.handle, as_ptr(b)))
        assert bytes(a) == bytes(b) and bytes(a) != bytes(64)
        # the all-ones evaluation vector is the constant polynomial 1: sum_i [L_i(tau)] = G1
        ones = curve.to_device(bn.fr_to_mont_bytes(1) * n)
        curve.check(lib.kzg_commit(ctx, lag, ones.handle, as_ptr(a)))
        assert bytes(a) == bn.g1_to_bytes((1, 2))
    finally:
        if lag:
            lib.kzg_srs_free(ctx, lag)
        lib.kzg_srs_free(ctx, srs)


def test_lagrange_needs_enough_points(curve, tau):
    from kzg_grandsums_study_b200 import KzgError
    from kzg_grandsums_study_b200._lib import as_ptr
    srs, lag = C.c_void_p(), C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), 7, C.byref(srs)))
    try:
        with pytest.raises(KzgError, match="not enough monomial points"):
            curve.check(curve.lib.kzg_srs_lagrange(curve.ctx, srs, 3, C.byref(lag)))
    finally:
        curve.lib.kzg_srs_free(curve.ctx, srs)


def test_host_layer_commit_from_evaluations(curve, ptau_factory):
    """Evaluations.commit(curve.load_lagrange_srs(ptau, nBits)) == Polynomial.fromEvaluations(...).multiExponentiation(srs):
    the [F], [T] of a proof from their evaluations, no iNTT"""
    from kzg_grandsums_study_b200.polynomial import Evaluations, Polynomial
    nbits = 10
    path = ptau_factory(nbits)
    f = inputs.random_column(321, 1 << nbits)
    ev = Evaluations(bn.fr_vec_to_mont_bytes(f), curve)
    lag = curve.load_lagrange_srs(path, nbits)
    srs = curve.load_srs(path, 2 << nbits)
    assert ev.commit(lag) == Polynomial.fromEvaluations(ev.eval, curve).multiExponentiation(srs)
    assert curve.load_lagrange_srs(path, nbits) is lag or curve.load_lagrange_srs(path, nbits).value == lag.value
