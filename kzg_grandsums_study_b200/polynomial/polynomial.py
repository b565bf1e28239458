"""Polynomial: coefficient-form polynomial over a device buffer.

Mirror of the part of reference src/polynomial/polynomial.js:25-1116 that the provers use (SURVEY.md
section 2, row 5): fromEvaluations, zero, Lagrange1, clone, length, degree, evaluate, add, sub, multiply,
shiftOmega, mulScalar, addScalar, subScalar, divByXSubValue, divZh, multiExponentiation, plus the small
accessors.  `coef` is a public field like in the reference; here it is a DeviceBuffer (`.tobytes()` /
`.slice()` give the reference's 32-byte Montgomery-LE coefficients).  Methods that mutate `this` in the
reference mutate `self` here and return it.
"""
import ctypes as C

from .._lib import as_ptr


class Polynomial:
    def __init__(self, coefficients, curve):                    # polynomial.js:26-31
        self.curve = curve
        self.coef = curve.to_device(coefficients)
        self.Fr = curve.Fr
        self.G1 = curve.G1

    # ---- constructors -------------------------------------------------------------------------------
    @staticmethod
    def fromEvaluations(buffer, curve):                         # polynomial.js:33-37
        return Polynomial(curve.Fr.ifft(buffer), curve)

    @staticmethod
    def fromCoefficientsArray(array, curve):                    # polynomial.js:39-49
        return Polynomial(b"".join(bytes(a) for a in array), curve)

    @staticmethod
    def fromPolynomial(polynomial, curve):                      # polynomial.js:51-61
        return Polynomial(polynomial.coef.clone(), curve)

    @staticmethod
    def zero(length, curve):                                    # polynomial.js:63-66
        return Polynomial(curve.alloc(length), curve)

    @staticmethod
    def Lagrange1(power, curve):                                # polynomial.js:68-78
        out = C.c_void_p()
        curve.check(curve.lib.kzg_poly_lagrange1(curve.ctx, power, C.byref(out)))
        return Polynomial(curve.wrap(out), curve)

    def clone(self):                                            # polynomial.js:80-82
        return Polynomial.fromPolynomial(self, self.curve)

    # ---- accessors ----------------------------------------------------------------------------------
    def tobytes(self):
        return self.coef.tobytes()

    def length(self):                                           # polynomial.js:198-206
        return self.coef.length()

    def degree(self):                                           # polynomial.js:212-226
        d = C.c_uint64()
        self.curve.check(self.curve.lib.kzg_poly_degree(self.curve.ctx, self.coef.handle, C.byref(d)))
        return int(d.value)

    def getCoef(self, index):                                   # polynomial.js:178-186
        if index > self.length() - 1:
            return self.Fr.zero
        return self.coef.slice(index * 32, (index + 1) * 32)

    def setCoef(self, index, value):                            # polynomial.js:188-196
        if index > self.length() - 1:
            raise IndexError("Coef index is not available")
        self.curve.check(self.curve.lib.kzg_buf_upload(self.curve.ctx, self.coef.handle, index, as_ptr(bytes(value)), 1))

    def isEqual(self, polynomial):                              # polynomial.js:84-95
        degree = self.degree()
        if degree != polynomial.degree():
            return False
        return self.coef.slice(0, (degree + 1) * 32) == polynomial.coef.slice(0, (degree + 1) * 32)

    def evaluate(self, point):                                  # polynomial.js:228-238
        out = bytearray(32)
        self.curve.check(self.curve.lib.kzg_poly_evaluate(self.curve.ctx, self.coef.handle, as_ptr(bytes(point)), as_ptr(out)))
        return bytes(out)

    # ---- arithmetic ---------------------------------------------------------------------------------
    def _replace(self, handle):
        self.coef = self.curve.wrap(handle)
        return self

    def add(self, polynomial, blindingValue=None):              # polynomial.js:276-312
        other = polynomial
        if blindingValue is not None:
            other = polynomial.clone().mulScalar(blindingValue)
        out = C.c_void_p()
        self.curve.check(self.curve.lib.kzg_poly_add(self.curve.ctx, self.coef.handle, other.coef.handle, C.byref(out)))
        return self._replace(out)

    def sub(self, polynomial, blindingValue=None):              # polynomial.js:314-350
        other = polynomial
        if blindingValue is not None:
            other = polynomial.clone().mulScalar(blindingValue)
        out = C.c_void_p()
        self.curve.check(self.curve.lib.kzg_poly_sub(self.curve.ctx, self.coef.handle, other.coef.handle, C.byref(out)))
        return self._replace(out)

    def multiply(self, polynomial):                             # polynomial.js:352-376
        out = C.c_void_p()
        self.curve.check(self.curve.lib.kzg_poly_multiply(self.curve.ctx, self.coef.handle, polynomial.coef.handle, C.byref(out)))
        return self._replace(out)

    def shiftOmega(self):                                       # polynomial.js:378-393
        out = C.c_void_p()
        self.curve.check(self.curve.lib.kzg_poly_shift_omega(self.curve.ctx, self.coef.handle, C.byref(out)))
        return self._replace(out)

    def mulScalar(self, value):                                 # polynomial.js:395-406
        self.curve.check(self.curve.lib.kzg_poly_mul_scalar(self.curve.ctx, self.coef.handle, as_ptr(bytes(value))))
        return self

    def addScalar(self, value):                                 # polynomial.js:408-414
        self.curve.check(self.curve.lib.kzg_poly_add_scalar(self.curve.ctx, self.coef.handle, as_ptr(bytes(value))))
        return self

    def subScalar(self, value):                                 # polynomial.js:416-422
        self.curve.check(self.curve.lib.kzg_poly_sub_scalar(self.curve.ctx, self.coef.handle, as_ptr(bytes(value))))
        return self

    def divByXSubValue(self, value):                            # polynomial.js:814-851
        out = C.c_void_p()
        self.curve.check(self.curve.lib.kzg_poly_div_x_sub_value(self.curve.ctx, self.coef.handle, as_ptr(bytes(value)), C.byref(out)))
        return self._replace(out)

    def divZh(self, domainSize):                                # polynomial.js:853-888
        out = C.c_void_p()
        self.curve.check(self.curve.lib.kzg_poly_div_zh(self.curve.ctx, self.coef.handle, domainSize, C.byref(out)))
        return self._replace(out)

    def multiExponentiation(self, PTau, name=None):             # polynomial.js:1106-1115
        """PTau: an SRS handle from curve.load_srs (device-resident [tau^i]_1); returns the 64 B affine commitment"""
        out = bytearray(64)
        srs = PTau[0] if isinstance(PTau, tuple) else PTau
        self.curve.check(self.curve.lib.kzg_commit(self.curve.ctx, srs, self.coef.handle, as_ptr(out)))
        return bytes(out)

    def print(self):                                            # polynomial.js:1089-1104
        terms = []
        for i in range(self.degree(), -1, -1):
            c = self.Fr.toString(self.getCoef(i))
            if c != "0":
                terms.append(c + ("" if i == 0 else " x" if i == 1 else " x^%d" % i))
        print(" + ".join(terms))
