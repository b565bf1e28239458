// Drop-in for reference src/grandproduct/mset_eq_kzg_prover.js:12 -- same call, same proof object, byte-identical values.
"use strict";
const prove = require("../prover_common.js");
module.exports = async function mset_eq_kzg_grandproduct_prover(pTauFilename, evalsFs, evalsTs, evalsSelF = null, evalsSelT = null, options) {
    return prove("gp", pTauFilename, evalsFs, evalsTs, evalsSelF, evalsSelT, options);
};
