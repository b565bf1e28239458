"""Evaluations: evaluation-form vector, mirror of reference src/polynomial/evaluations.js:5-138.

`eval` is either host bytes (what callers construct, e.g. getRandomEvals) or a device buffer (what the
library returns); bulk work runs on the device through the C ABI.
"""
import ctypes as C
import random as _random

from ..curve import DeviceBuffer
from .._lib import as_ptr


class Evaluations:
    def __init__(self, evaluations, curve):                     # evaluations.js:6-10
        self.eval = evaluations
        self.curve = curve
        self.Fr = curve.Fr

    # ---- constructors -------------------------------------------------------------------------------
    @staticmethod
    def fromPolynomial(polynomial, extension, curve):           # evaluations.js:12-21
        out = C.c_void_p()
        coef = curve.to_device(polynomial.coef)
        curve.check(curve.lib.kzg_fr_extend_ntt(curve.ctx, coef.handle, extension, C.byref(out)))
        return Evaluations(curve.wrap(out), curve)

    def commit(self, lagrange_srs):
        """[p(tau)]_1 of the polynomial these are the evaluations of (Montgomery form, on H), committed DIRECTLY over the
        Lagrange-basis SRS of curve.load_lagrange_srs(...) -- the same 64 bytes as
        Polynomial.fromEvaluations(...).multiExponentiation(monomial SRS), without the iNTT"""
        d = self.curve.to_device(self.eval)
        out = bytearray(64)
        self.curve.check(self.curve.lib.kzg_commit(self.curve.ctx, lagrange_srs, d.handle, as_ptr(out)))
        return bytes(out)

    @staticmethod
    def fromArray(array, curve):                                # evaluations.js:23-29
        return Evaluations(b"".join(bytes(a) for a in array), curve)

    @staticmethod
    def fromEvals(evals):                                       # evaluations.js:31-33
        return Evaluations(evals.tobytes(), evals.curve)

    @staticmethod
    def getOneEvals(length, curve):                             # evaluations.js:35-41
        return Evaluations(curve.Fr.one * length, curve)

    @staticmethod
    def getZeroEvals(length, curve):                            # evaluations.js:43-49
        return Evaluations(bytes(32 * length), curve)

    @staticmethod
    def getRandomEvals(length, curve):                          # evaluations.js:51-57
        return Evaluations(b"".join(curve.Fr.random() for _ in range(length)), curve)

    @staticmethod
    def getRandomBinEvals(length, curve):                       # evaluations.js:59-66
        return Evaluations(b"".join(curve.Fr.one if _random.getrandbits(1) else curve.Fr.zero for _ in range(length)), curve)

    # ---- accessors ----------------------------------------------------------------------------------
    def tobytes(self):
        if isinstance(self.eval, DeviceBuffer):
            return self.eval.tobytes()
        if isinstance(self.eval, (bytes, bytearray)):
            return bytes(self.eval)
        if hasattr(self.eval, "data_ptr"):          # torch tensor (e.g. pinned host memory)
            return self.eval.cpu().numpy().tobytes()
        return memoryview(self.eval).tobytes()       # numpy array or any buffer

    def host_buffer(self):
        """the evaluations as something as_ptr() can hand to the C ABI WITHOUT copying when they already sit in
        host memory (bytes, bytearray, numpy array, pinned torch tensor); device buffers are downloaded"""
        if isinstance(self.eval, DeviceBuffer):
            return self.eval.tobytes()
        if isinstance(self.eval, memoryview):
            return self.eval.tobytes()
        return self.eval

    def _nbytes(self):
        e = self.eval
        if isinstance(e, DeviceBuffer):
            return e.byteLength
        if isinstance(e, (bytes, bytearray)):
            return len(e)
        if hasattr(e, "data_ptr"):
            return e.numel() * e.element_size()
        return memoryview(e).nbytes

    def getEvaluation(self, index):                             # evaluations.js:68-74
        if (index + 1) * 32 > self.length() * 32:
            raise IndexError("Evaluations.getEvaluation() out of bounds")
        if isinstance(self.eval, DeviceBuffer):
            return self.eval.slice(index * 32, (index + 1) * 32)
        if isinstance(self.eval, (bytes, bytearray)):
            return bytes(self.eval[index * 32:(index + 1) * 32])
        return self.tobytes()[index * 32:(index + 1) * 32]

    def getEvaluationSequence(self, start, end):                # evaluations.js:76-88
        if start > end:
            raise IndexError("Evaluations.getEvaluationSequence() start index is greater than end index")
        if start == end:
            raise IndexError("Use Evaluations.getEvaluation() instead")
        if end > self.length() - 1:
            raise IndexError("Evaluations.getEvaluationSequence() end index is out of bounds")
        if isinstance(self.eval, DeviceBuffer):
            return self.eval.slice(start * 32, end * 32)
        if isinstance(self.eval, (bytes, bytearray)):
            return bytes(self.eval[start * 32:end * 32])
        return self.tobytes()[start * 32:end * 32]

    def setEvaluation(self, index, value):                      # evaluations.js:90-96
        if index > self.length() - 1:
            raise IndexError("Evaluation index is out of bounds")
        if isinstance(self.eval, DeviceBuffer):
            self.curve.check(self.curve.lib.kzg_buf_upload(self.curve.ctx, self.eval.handle, index, as_ptr(bytes(value)), 1))
        else:
            b = bytearray(self.tobytes())
            b[index * 32:(index + 1) * 32] = bytes(value)
            self.eval = b

    def length(self):                                           # evaluations.js:99-108
        nbytes = self._nbytes()
        if nbytes % 32:
            raise ValueError("Polynomial evaluations buffer has incorrect size")
        return nbytes // 32

    def isEqual(self, other):                                   # evaluations.js:110-116
        if self.length() != other.length():
            return False
        return self.tobytes() == other.tobytes()

    def _all_equal(self, value):
        if isinstance(self.eval, DeviceBuffer):
            out = C.c_int()
            self.curve.check(self.curve.lib.kzg_buf_all_equal(self.curve.ctx, self.eval.handle, as_ptr(value), C.byref(out)))
            return bool(out.value)
        return self.tobytes() == value * self.length()

    def isAllZeros(self):                                       # evaluations.js:118-121
        return self._all_equal(self.Fr.zero)

    def isAllOnes(self):                                        # evaluations.js:123-129
        return self._all_equal(self.Fr.one)

    def print(self, name="f"):                                  # evaluations.js:131-135
        for i in range(self.length()):
            print("%s(w^%d) = %s" % (name, i, self.Fr.toString(self.getEvaluation(i))))
