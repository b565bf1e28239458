// Drop-in for reference src/grandsum/mset_eq_kzg_verifier.js:9 -- (pTauFilename, proof, nBits) -> Promise<bool>.
"use strict";
const { verify } = require("../verifier_common.js");
module.exports = async function mset_eq_kzg_grandsum_verifier(pTauFilename, proof, nBits, logger) {
    return verify("gs", pTauFilename, proof, nBits, logger);
};
