"""Host scalar helpers, mirror of reference src/polynomial/polynomial_utils.js:1-19 (O(log n) field operations)."""


def computeZHEvaluation(curve, x, nBits):
    """Z_H(x) = x^(2^nBits) - 1   (polynomial_utils.js:1-10)"""
    Fr = curve.Fr
    xn = x
    for _ in range(nBits):
        xn = Fr.square(xn)
    return Fr.sub(xn, Fr.one)


def computeL1Evaluation(curve, x, ZHx, nBits):
    """L_1(x) = Z_H(x) / (n (x - 1))   (polynomial_utils.js:12-19)"""
    Fr = curve.Fr
    n = Fr.e(2 ** nBits)
    w = Fr.one
    return Fr.div(Fr.mul(w, ZHx), Fr.mul(n, Fr.sub(x, w)))
