// Drop-in surface of the reference (SURVEY.md 8b) on top of the B200 backend: same module layout, names, argument
// meaning, proof key order and error strings as /src of xavi-pinsach/kzg-grandsums-study; the bulk work runs in
// libkzgb200.so through addon/kzgb200_napi.cc.  There is no CPU fallback: without the addon `require` throws.
"use strict";
const { getCurveFromName, getCurveFromQ } = require("./src/curve.js");
module.exports = {
    getCurveFromName,
    getCurveFromQ,
    Polynomial: require("./src/polynomial/polynomial.js").Polynomial,
    Evaluations: require("./src/polynomial/evaluations.js").Evaluations,
    Keccak256Transcript: require("./src/Keccak256Transcript.js").Keccak256Transcript,
    readPTauHeader: require("./src/ptau_utils.js").readPTauHeader,
    mset_eq_kzg_grandsum_prover: require("./src/grandsum/mset_eq_kzg_prover.js"),
    mset_eq_kzg_grandsum_verifier: require("./src/grandsum/mset_eq_kzg_verifier.js"),
    mset_eq_kzg_grandproduct_prover: require("./src/grandproduct/mset_eq_kzg_prover.js"),
    mset_eq_kzg_grandproduct_verifier: require("./src/grandproduct/mset_eq_kzg_verifier.js"),
    ComputeSGrandSumPolynomial: require("./src/grandsum/grandsum.js").ComputeSGrandSumPolynomial,
    ComputeZGrandProductPolynomial: require("./src/grandproduct/grandproduct.js").ComputeZGrandProductPolynomial,
};
