// Drop-in for reference src/grandproduct/mset_eq_kzg_verifier.js:9 -- (pTauFilename, proof, nBits) -> Promise<bool>.
"use strict";
const { verify } = require("../verifier_common.js");
module.exports = async function mset_eq_kzg_grandproduct_verifier(pTauFilename, proof, nBits, logger) {
    return verify("gp", pTauFilename, proof, nBits, logger);
};
