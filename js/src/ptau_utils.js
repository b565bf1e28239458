// Drop-in for reference src/ptau_utils.js:3-24.  The reference parses the header through a binfileutils file
// descriptor; here the library validates the container (magic "ptau", version <= 1, exactly one header section, n8 == 32,
// q == BN254 q, header size -- the reference's own messages) and the caller gets the same {curve, power, ceremonyPower}.
// `fd` may be the reference's descriptor (its .fileName is used) or a path.
"use strict";
const { getCurveFromName } = require("./curve.js");

module.exports.readPTauHeader = async function readPTauHeader(fd, sections, device = 0) {
    const fileName = typeof fd === "string" ? fd : fd.fileName;
    if (sections !== undefined && sections !== null) {
        if (!sections[1]) throw new Error(fileName + ": File has no  header");
        if (sections[1].length > 1) throw new Error(fileName + ": File has more than one header");
    }
    const curve = await getCurveFromName("bn128", device);
    const r = curve.addon.kzg_ptau_read_header(curve.ctx, fileName);
    return { curve, power: r.power, ceremonyPower: r.ceremony_power };
};

// [tau]_2: the second G2 point of section 3 (verifier.js:18-19), 128 bytes Montgomery-LE
module.exports.readTauG2 = async function readTauG2(fileName, device = 0) {
    const curve = await getCurveFromName("bn128", device);
    return Uint8Array.from(curve.addon.kzg_ptau_read_tau_g2(curve.ctx, fileName));
};
