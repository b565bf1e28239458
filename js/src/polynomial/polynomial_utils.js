// Drop-in for reference src/polynomial/polynomial_utils.js:1-19 (O(1) host field arithmetic).
"use strict";
module.exports.computeZHEvaluation = function computeZHEvaluation(curve, x, nBits) {
    const Fr = curve.Fr;
    let xn = x;
    for (let i = 0; i < nBits; i++) xn = Fr.square(xn);
    return Fr.sub(xn, Fr.one);
};
module.exports.computeL1Evaluation = function computeL1Evaluation(curve, x, ZHx, nBits) {
    const Fr = curve.Fr;
    const n = Fr.e(2 ** nBits);
    return Fr.div(ZHx, Fr.mul(n, Fr.sub(x, Fr.one)));
};
