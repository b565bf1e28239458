// Drop-in for reference src/grandproduct/grandproduct.js (grandproduct.js:6-57), same argument list: terms -> batch inverse -> running accumulation ->
// wrap check -> iNTT as one device pipeline.  Evaluations in (Montgomery), coefficients of the accumulator polynomial out;
// throws the reference's "... is not well calculated" error when the multisets differ.  Selectors that are null (or all
// ones) mean "plain argument".
"use strict";
const { Polynomial } = require("../polynomial/polynomial.js");

module.exports.ComputeZGrandProductPolynomial = async function ComputeZGrandProductPolynomial(evalsF, evalsT, evalsSelF, evalsSelT, isSelected, challenge, curve) {
    const a = curve.addon;
    const selected = Boolean(isSelected && evalsSelF && evalsSelT) && !(evalsSelF.isAllOnes() && evalsSelT.isAllOnes());
    const hs = [curve.upload(evalsF.eval), curve.upload(evalsT.eval),
        selected ? curve.upload(evalsSelF.eval) : null, selected ? curve.upload(evalsSelT.eval) : null];
    let out = null;
    try {
        out = a.kzg_grandproduct_build(curve.ctx, hs[0], hs[1], hs[2], hs[3], challenge);
        return new Polynomial(curve.download(out), curve);
    } finally {
        curve.free(...hs.filter(Boolean), out);
    }
};
