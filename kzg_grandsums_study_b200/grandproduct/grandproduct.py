"""ComputeZGrandProductPolynomial -- drop-in for reference src/grandproduct/grandproduct.js:6-57."""
import ctypes as C

from ..polynomial.polynomial import Polynomial
from .._lib import as_ptr


def ComputeZGrandProductPolynomial(evalsF, evalsT, evalsSelF, evalsSelT, isSelected, challenge, curve):
    """evaluations (Montgomery) in, coefficients of Z out; raises the reference's
    "The grand-product polynomial Z is not well calculated" when the multisets differ."""
    f = curve.to_device(evalsF.eval)
    t = curve.to_device(evalsT.eval)
    sf = curve.to_device(evalsSelF.eval) if (isSelected and evalsSelF is not None) else None
    st = curve.to_device(evalsSelT.eval) if (isSelected and evalsSelT is not None) else None
    out = C.c_void_p()
    curve.check(curve.lib.kzg_grandproduct_build(curve.ctx, f.handle, t.handle, sf.handle if sf else None,
                                                 st.handle if st else None, as_ptr(bytes(challenge)), C.byref(out)))
    return Polynomial(curve.wrap(out), curve)
