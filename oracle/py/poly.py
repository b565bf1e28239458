"""Polynomial / Evaluations algebra over Fr -- ORACLE, test infrastructure only.

Restates the subset of reference src/polynomial/polynomial.js and src/polynomial/evaluations.js that
the provers use.  Coefficients / evaluations are held as lists of standard-form Python ints; the
32-byte Montgomery-LE buffers of the reference are produced at the boundary (bn254.fr_vec_to_mont_bytes).
Every value is a canonical field element, so the int representation cannot change a proof byte.

Deviation (documented): `multiply` computes the mathematically correct product for any
operands; the reference is only correct for full-degree operands, which is the parity domain.
"""
from . import bn254 as bn

R = bn.R


def ntt(a, inverse=False):
    """Fr.fft / Fr.ifft [dep]: out[k] = sum_j a[j] w^(jk), w = FR_W[log2 n], natural order in and out;
    ifft is the exact inverse (scaled by n^-1).  Call sites: polynomial.js:34, evaluations.js:18."""
    n = len(a)
    if n == 1:
        return list(a)
    lg = n.bit_length() - 1
    if 1 << lg != n:
        raise ValueError("fft must be multiple of 2")
    a = list(a)
    # bit reversal
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j |= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    w_n = bn.FR_W[lg]
    if inverse:
        w_n = pow(w_n, -1, R)
    length = 2
    while length <= n:
        wl = pow(w_n, n // length, R)
        half = length >> 1
        tw = [1] * half
        for k in range(1, half):
            tw[k] = tw[k - 1] * wl % R
        for s in range(0, n, length):
            for k in range(half):
                u = a[s + k]
                v = a[s + k + half] * tw[k] % R
                a[s + k] = (u + v) % R
                a[s + k + half] = (u - v) % R
        length <<= 1
    if inverse:
        ninv = pow(n, -1, R)
        a = [x * ninv % R for x in a]
    return a


def batch_inverse(v):
    """Fr.batchInverse [dep]: Montgomery trick; zero elements map to zero (grandsum.js:41)."""
    out = [0] * len(v)
    acc = 1
    pref = []
    for x in v:
        pref.append(acc)
        if x:
            acc = acc * x % R
    inv = pow(acc, -1, R)
    for i in range(len(v) - 1, -1, -1):
        if v[i]:
            out[i] = inv * pref[i] % R
            inv = inv * v[i] % R
    return out


class Polynomial:
    """polynomial.js:25 -- coefficient form; `coef` is a list of ints, len = buffer length."""

    def __init__(self, coef):
        self.coef = list(coef)

    @staticmethod
    def from_evaluations(evals):                      # polynomial.js:33-37
        return Polynomial(ntt(evals, inverse=True))

    @staticmethod
    def zero(length):                                 # polynomial.js:63-66
        return Polynomial([0] * length)

    @staticmethod
    def lagrange1(power):                             # polynomial.js:68-78
        e = [0] * (1 << power)
        e[0] = 1
        return Polynomial.from_evaluations(e)

    def clone(self):                                  # polynomial.js:80-82
        return Polynomial(self.coef)

    def length(self):                                 # polynomial.js:198-206
        return len(self.coef)

    def degree(self):                                 # polynomial.js:212-226
        for i in range(len(self.coef) - 1, 0, -1):
            if self.coef[i]:
                return i
        return 0

    def evaluate(self, x):                            # polynomial.js:228-238 (Horner)
        res = 0
        if not self.coef:                             # empty buffer (divZh of the zero polynomial): the zero polynomial
            return 0
        for i in range(self.degree(), -1, -1):
            res = (self.coef[i] + res * x) % R
        return res

    def _binop(self, other, sign):                    # polynomial.js:276-350 (result takes the longer length)
        n = max(len(self.coef), len(other.coef))
        a = self.coef + [0] * (n - len(self.coef))
        b = other.coef + [0] * (n - len(other.coef))
        self.coef = [(x + sign * y) % R for x, y in zip(a, b)]
        return self

    def add(self, other):
        return self._binop(other, 1)

    def sub(self, other):
        return self._binop(other, -1)

    def multiply(self, other):                        # polynomial.js:352-376
        da, db = self.degree(), other.degree()
        new_len = 1
        while new_len < da + db + 1:
            new_len <<= 1
        fa = ntt(self.coef[:da + 1] + [0] * (new_len - da - 1))
        fb = ntt(other.coef[:db + 1] + [0] * (new_len - db - 1))
        self.coef = ntt([x * y % R for x, y in zip(fa, fb)], inverse=True)
        return self

    def shift_omega(self):                            # polynomial.js:378-393: p(X) -> p(wX)
        ev = Evaluations.from_polynomial(self, 1).eval
        self.coef = ntt(ev[1:] + ev[:1], inverse=True)
        return self

    def mul_scalar(self, v):                          # polynomial.js:395-406
        self.coef = [c * v % R for c in self.coef]
        return self

    def add_scalar(self, v):                          # polynomial.js:408-414
        self.coef[0] = (self.coef[0] + v) % R
        return self

    def sub_scalar(self, v):                          # polynomial.js:416-422
        self.coef[0] = (self.coef[0] - v) % R
        return self

    def div_by_x_sub_value(self, v):                  # polynomial.js:814-851
        n = len(self.coef)
        q = [0] * n
        q[n - 2] = self.coef[n - 1]
        for i in range(n - 3, -1, -1):
            q[i] = (self.coef[i + 1] + v * q[i + 1]) % R
        if self.coef[0] != (-v) * q[0] % R:
            raise ValueError("Polynomial does not divide")
        self.coef = q
        return self

    def div_zh(self, domain_size):                    # polynomial.js:853-888
        n = domain_size
        deg = self.degree()
        m = deg + 1 - n
        length = 0 if deg < n else (1 if m == 1 else 1 << (m - 1).bit_length())   # 2**ceil(log2(m))
        a = list(self.coef)
        ext = len(a) // n
        for i in range(n):
            a[i] = (-a[i]) % R
        for i in range(n, n * ext):
            a[i] = (a[i - n] - a[i]) % R
            if i > n * (ext - 1) - ext and a[i]:
                raise ValueError("Polynomial is not divisible")
        p = Polynomial(a)
        d = p.degree()
        out = [0] * length
        out[:d + 1] = a[:d + 1]
        self.coef = out[:length] if length else []
        return self

    def to_mont_bytes(self):
        return bn.fr_vec_to_mont_bytes(self.coef)


class Evaluations:
    """evaluations.js:5 -- evaluation form; `eval` is a list of ints."""

    def __init__(self, ev):
        self.eval = list(ev)

    @staticmethod
    def from_polynomial(p, extension):                # evaluations.js:12-21
        n = len(p.coef)
        size = 1
        while size < n:
            size <<= 1
        size *= extension
        return Evaluations(ntt(p.coef + [0] * (size - n)))

    def length(self):
        return len(self.eval)
