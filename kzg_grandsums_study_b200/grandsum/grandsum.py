"""ComputeSGrandSumPolynomial -- drop-in for reference src/grandsum/grandsum.js:6-62."""
import ctypes as C

from ..polynomial.polynomial import Polynomial
from .._lib import as_ptr


def ComputeSGrandSumPolynomial(evalsF, evalsT, evalsSelF, evalsSelT, challenge, curve):
    """evaluations (Montgomery) in, coefficients of S out; raises the reference's
    "The grand-sum polynomial S is not well calculated" when the multisets differ."""
    f = curve.to_device(evalsF.eval)
    t = curve.to_device(evalsT.eval)
    sf = curve.to_device(evalsSelF.eval) if evalsSelF is not None else None
    st = curve.to_device(evalsSelT.eval) if evalsSelT is not None else None
    out = C.c_void_p()
    curve.check(curve.lib.kzg_grandsum_build(curve.ctx, f.handle, t.handle, sf.handle if sf else None,
                                             st.handle if st else None, as_ptr(bytes(challenge)), C.byref(out)))
    return Polynomial(curve.wrap(out), curve)
