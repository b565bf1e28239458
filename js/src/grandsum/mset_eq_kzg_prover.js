// Drop-in for reference src/grandsum/mset_eq_kzg_prover.js:12 -- same call, same proof object, byte-identical values.
"use strict";
const prove = require("../prover_common.js");
module.exports = async function mset_eq_kzg_grandsum_prover(pTauFilename, evalsFs, evalsTs, evalsSelF = null, evalsSelT = null, options) {
    return prove("gs", pTauFilename, evalsFs, evalsTs, evalsSelF, evalsSelT, options);
};
