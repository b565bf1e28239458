// Shared host driver of the two provers: the argument normalisation and checks of the reference
// (src/grandsum/mset_eq_kzg_prover.js:12-81 == src/grandproduct/mset_eq_kzg_prover.js:12-81, same error strings), the
// Keccak transcript schedule (SURVEY.md A.1) and the five fused device rounds behind the addon (kzg_prover_round1..5).
// The proof object has the reference's shape and KEY INSERTION ORDER (prover.js:95,161-162,173-174,229,284,301-316,409-410).
"use strict";
const { Keccak256Transcript } = require("./Keccak256Transcript.js");
const { Evaluations } = require("./polynomial/evaluations.js");
const { readPTauHeader } = require("./ptau_utils.js");
const { asU8 } = require("./curve.js");

module.exports = async function prove(kind, pTauFilename, evalsFs, evalsTs, evalsSelF = null, evalsSelT = null, options = {}) {
    const gs = kind === "gs";
    const { curve, power: nBitsPTau } = await readPTauHeader(pTauFilename, null, options.device || 0);   // :15-16
    const a = curve.addon;

    if (!Array.isArray(evalsFs)) evalsFs = [evalsFs];                                                   // :22-27
    if (!Array.isArray(evalsTs)) evalsTs = [evalsTs];
    if (evalsFs.length !== evalsTs.length) throw new Error("The lengths of the two vector multisets must be the same.");
    const nPols = evalsFs.length;
    if (nPols === 0) throw new Error("The number of multisets must be greater than 0.");
    for (let i = 0; i < nPols; i++) {                                                                   // :39-45
        if (evalsFs[i].length() !== evalsTs[i].length()) throw new Error(`The ${i}-th multiset buffers must have the same length.`);
        if (evalsFs[i].length() !== evalsFs[0].length()) throw new Error("The multiset buffers must all have the same length.");
    }
    // :48-68 -- selectors that are not given mean "all ones"; the reference materialises them and then finds out
    let isSelected = false;
    if (evalsSelF !== null || evalsSelT !== null) {
        if (evalsSelF === null) evalsSelF = Evaluations.getOneEvals(evalsFs[0].length(), curve);
        if (evalsSelT === null) evalsSelT = Evaluations.getOneEvals(evalsTs[0].length(), curve);
        if (evalsSelF.length() !== evalsSelT.length()) throw new Error("The selection buffers must have the same length.");
        if (evalsSelF.length() !== evalsFs[0].length()) throw new Error("The selection buffers must have the same length as the multiset buffers.");
        isSelected = !(evalsSelF.isAllOnes() && evalsSelT.isAllOnes());
        if (evalsSelF.isAllZeros() && evalsSelT.isAllZeros() && options.logger)
            options.logger.warn("The selection buffers are all zeros. The argument is trivially satisfied.");
    }
    const length = evalsFs[0].length();
    const nBits = length > 0 ? Math.ceil(Math.log2(length)) : 0;                                        // :70-71
    const domainSize = 2 ** nBits;
    if (length !== domainSize) throw new Error("Polynomial length must be a power of two.");            // :74-76
    if (nBitsPTau < nBits) throw new Error("The Powers of Tau file is not sufficiently large to commit the polynomials.");

    const { srs } = curve.loadSrs(pTauFilename, domainSize * 2);                                        // :83-85, device-resident
    const isVector = nPols > 1;
    const acc = gs ? "S" : "Z";
    const prover = a.kzg_prover_create(curve.ctx, srs, gs ? a.KZG_GRANDSUM : a.KZG_GRANDPRODUCT, nBits, nPols, isSelected ? 1 : 0);
    try {
        const proof = { evaluations: {}, commitments: {} };
        const Cm = proof.commitments, Ev = proof.evaluations;
        const transcript = new Keccak256Transcript(curve);
        const challenges = {};
        const slice = (buf, off, len) => Uint8Array.from(buf.subarray(off, off + len));
        const fName = (i) => (isVector ? `F${i}` : "F"), tName = (i) => (isVector ? `T${i}` : "T");

        // ---- round 1: witness polynomials and their commitments (:144-179)
        const out1 = a.kzg_prover_round1(prover, evalsFs.map((e) => asU8(e.eval)), evalsTs.map((e) => asU8(e.eval)),
            isSelected ? asU8(evalsSelF.eval) : null, isSelected ? asU8(evalsSelT.eval) : null);
        for (let i = 0; i < nPols; i++) {
            Cm[fName(i)] = slice(out1, 128 * i, 64);
            Cm[tName(i)] = slice(out1, 128 * i + 64, 64);
        }
        if (isSelected) {
            Cm.selF = slice(out1, 128 * nPols, 64);
            Cm.selT = slice(out1, 128 * nPols + 64, 64);
        }
        // ---- round 2: the grand-sum / grand-product polynomial (:181-231)
        for (let i = 0; i < nPols; i++) {
            transcript.addPolCommitment(Cm[fName(i)]);
            transcript.addPolCommitment(Cm[tName(i)]);
        }
        if (isSelected) {
            transcript.addPolCommitment(Cm.selF);
            transcript.addPolCommitment(Cm.selT);
        }
        let beta = null;
        if (isVector) {
            beta = challenges.beta = transcript.getChallenge();
            transcript.addFieldElement(beta);
        }
        const gamma = (challenges.gamma = transcript.getChallenge());
        Cm[acc] = Uint8Array.from(a.kzg_prover_round2(prover, beta, gamma));
        // ---- round 3: the quotient polynomial (:233-286)
        transcript.addFieldElement(gamma);
        transcript.addPolCommitment(Cm[acc]);
        const alpha = (challenges.alpha = transcript.getChallenge());
        Cm.Q = Uint8Array.from(a.kzg_prover_round3(prover, alpha));
        // ---- round 4: evaluations (:288-318)
        transcript.addFieldElement(alpha);
        transcript.addPolCommitment(Cm.Q);
        const xi = (challenges.xi = transcript.getChallenge());
        const out4 = a.kzg_prover_round4(prover, xi);
        let pos = 0;
        const next = () => slice(out4, 32 * pos++, 32);
        for (let i = 0; i < nPols; i++) {
            Ev[isVector ? `f${i}xi` : "fxi"] = next();
            if (gs) Ev[isVector ? `t${i}xi` : "txi"] = next();
        }
        if (isSelected) {
            Ev.selFxi = next();
            Ev.selTxi = next();
        }
        Ev[gs ? "sxiw" : "zxiw"] = next();
        // ---- round 5: opening proofs (:320-413)
        transcript.addFieldElement(xi);
        for (const key of Object.keys(Ev)) transcript.addFieldElement(Ev[key]);
        const v = (challenges.v = transcript.getChallenge());
        const out5 = a.kzg_prover_round5(prover, v);
        Cm.Wxi = slice(out5, 0, 64);
        Cm.Wxiw = slice(out5, 64, 64);
        // side effect of the reference (:147-148): the callers' F / T evaluations are now in Montgomery form
        for (let i = 0; i < nPols; i++) {
            for (const [which, ev] of [[0, evalsFs[i]], [1, evalsTs[i]]]) {
                const h = a.kzg_prover_take_evals(prover, i, which);
                try {
                    ev.eval = curve.download(h);
                } finally {
                    curve.free(h);
                }
            }
        }
        if (options.trace) options.trace.challenges = challenges;
        return proof;
    } finally {
        a.kzg_prover_destroy(prover);
    }
};
