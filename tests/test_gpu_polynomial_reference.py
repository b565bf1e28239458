"""GPU: the known-answer tests of the reference's own test/polynomial.test.js, restated against the device-backed
Polynomial class -- the subset whose methods the provers use: degree (:31-70), isEqual
(:72-93), getCoef (:95-102), length (:104-114), evaluate (:116-124), add / sub with a blinding value (:126-166),
mulScalar (:168-179), addScalar / subScalar (:181-205), multiply (:207-220)."""
import random

import pytest

pytestmark = pytest.mark.gpu


def random_polynomial(min_degree, max_degree, curve, rng):
    from kzg_grandsums_study_b200.polynomial import Polynomial
    degree = rng.randint(min_degree, max_degree)
    return Polynomial(b"".join(curve.Fr.e(rng.randrange(curve.r)) for _ in range(degree + 1)), curve)


def test_should_return_the_correct_degree(curve):
    from kzg_grandsums_study_b200.polynomial import Polynomial
    Fr = curve.Fr
    rnd = lambda: Fr.e(random.Random(1).randrange(1, curve.r))
    assert Polynomial(b"", curve).degree() == 0                       # no coefficients => degree 0
    assert Polynomial(rnd(), curve).degree() == 0                     # one coefficient => degree 0
    assert Polynomial(rnd() + rnd(), curve).degree() == 1             # two coefficients => degree 1
    assert Polynomial(rnd() + Fr.zero, curve).degree() == 0           # the greatest is zero => degree 0
    assert Polynomial(rnd() + Fr.zero + Fr.zero, curve).degree() == 0
    assert Polynomial(rnd() + Fr.zero + Fr.one, curve).degree() == 2


def test_should_check_if_two_polynomials_are_equal(curve):
    from kzg_grandsums_study_b200.polynomial import Polynomial
    Fr = curve.Fr
    pol1 = random_polynomial(10, 30, curve, random.Random(2))
    assert pol1.isEqual(pol1)
    pol2 = Polynomial(pol1.tobytes(), curve)
    assert pol1.isEqual(pol2)
    pol3 = Polynomial(pol1.tobytes() + Fr.zero, curve)                # one more (zero) coefficient
    assert pol1.isEqual(pol3)
    pol4 = Polynomial(Fr.one + pol1.tobytes()[32:] + Fr.zero, curve)
    assert not pol1.isEqual(pol4)


def test_should_get_the_correct_coefficient_and_length(curve):
    pol = random_polynomial(10, 30, curve, random.Random(3))
    raw = pol.tobytes()
    assert pol.length() == len(raw) // 32
    for i in range(pol.length()):
        assert pol.getCoef(i) == raw[32 * i:32 * i + 32]
    assert pol.getCoef(pol.length()) == curve.Fr.zero                 # polynomial.js:178-186: out of range reads as zero


def test_should_evaluate_a_polynomial(curve):
    from kzg_grandsums_study_b200.polynomial import Polynomial
    Fr = curve.Fr
    pol = Polynomial(b"".join(Fr.e(i) for i in range(4)), curve)      # x + 2x^2 + 3x^3
    assert pol.evaluate(Fr.e(2)) == Fr.e(34)


@pytest.mark.parametrize("op", ["add", "sub"])
def test_should_add_and_sub_a_polynomial_with_blinding(curve, op):
    Fr = curve.Fr
    rng = random.Random(5 if op == "add" else 6)
    p1 = random_polynomial(10, 30, curve, rng)
    p2 = random_polynomial(10, 30, curve, rng)
    l1, l2 = p1.length(), p2.length()
    c1, c2 = p1.tobytes(), p2.tobytes()
    blinding = Fr.e(rng.randrange(curve.r))
    getattr(p1, op)(p2, blinding)
    assert p1.length() == max(l1, l2)
    f = Fr.add if op == "add" else Fr.sub
    for i in range(p1.length()):
        v1 = c1[32 * i:32 * i + 32] if i < l1 else Fr.zero
        v2 = c2[32 * i:32 * i + 32] if i < l2 else Fr.zero
        assert p1.getCoef(i) == f(v1, Fr.mul(v2, blinding))


def test_should_mul_add_sub_a_scalar(curve):
    Fr = curve.Fr
    rng = random.Random(7)
    for op, f in (("mulScalar", Fr.mul), ("addScalar", Fr.add), ("subScalar", Fr.sub)):
        p = random_polynomial(10, 30, curve, rng)
        clone = p.tobytes()
        scalar = Fr.e(rng.randrange(curve.r))
        getattr(p, op)(scalar)
        for i in range(p.length()):
            c = clone[32 * i:32 * i + 32]
            want = f(c, scalar) if (op == "mulScalar" or i == 0) else c
            assert p.getCoef(i) == want


def test_should_multiply_by_a_polynomial(curve):
    from kzg_grandsums_study_b200.polynomial import Polynomial
    Fr = curve.Fr
    # (2x^3 - 3x^2 + 2)(x^2 + 3x) = 2x^5 + 3x^4 - 9x^3 + 2x^2 + 6x
    pol1 = Polynomial.fromCoefficientsArray([Fr.e(2), Fr.e(0), Fr.e(-3), Fr.e(2)], curve)
    pol2 = Polynomial.fromCoefficientsArray([Fr.e(0), Fr.e(3), Fr.one], curve)
    want = Polynomial.fromCoefficientsArray([Fr.zero, Fr.e(6), Fr.e(2), Fr.e(-9), Fr.e(3), Fr.e(2)], curve)
    assert want.isEqual(pol1.multiply(pol2))
