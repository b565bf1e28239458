// Shared host driver of the two verifiers -- drop-in for reference src/grandsum/mset_eq_kzg_verifier.js:9-313 and
// src/grandproduct/mset_eq_kzg_verifier.js:9-299: same validity checks, same transcript, same verdict (a bad proof
// returns false, it never throws).  Field arithmetic with BigInt on the host; the group side of the pairing check,
//     A = [Wxi] + u [Wxiw]        B = xi [Wxi] + u xi w [Wxiw] + [F] - [E]
// is expanded into ONE linear combination of the proof's commitments per side (the same expansion as the Python layer's
// verify_batch, kzg_grandsums_study_b200/_verifier_common.py::linear_terms) and evaluated by the device MSM; the final
// 2-pairing product is ffjavascript's curve.pairingEq, as in the reference -- the pairing stays a host-side library call
// (optionalDependency: it is present wherever the reference runs).
"use strict";
const { Keccak256Transcript } = require("./Keccak256Transcript.js");
const { getCurveFromName, R, leToBig } = require("./curve.js");
const { readTauG2 } = require("./ptau_utils.js");

const mod = (x) => ((x % R) + R) % R;
function modinv(a) {
    let [r0, r1, s0, s1] = [R, mod(a), 0n, 1n];
    while (r1 !== 0n) {
        const q = r0 / r1;
        [r0, r1, s0, s1] = [r1, r0 - q * r1, s1, s0 - q * s1];
    }
    return mod(s0);
}

let ffCurvePromise = null;
function pairingCurve() {
    if (!ffCurvePromise) {
        let ff;
        try {
            ff = require("ffjavascript");
        } catch (e) {
            throw new Error("the verifier's pairing check uses ffjavascript's curve.pairingEq (as the reference does): npm install ffjavascript");
        }
        ffCurvePromise = ff.getCurveFromName("bn128");
    }
    return ffCurvePromise;
}

function validG1(curve, bytes) {
    // canonical coordinates (raw integer < q) and on the curve y^2 = x^3 + 3 (cofactor 1), or infinity
    const Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583n;
    if (!(bytes instanceof Uint8Array) || bytes.length !== 64) return false;
    if (bytes.every((b) => b === 0)) return true;
    const xm = leToBig(bytes.subarray(0, 32)), ym = leToBig(bytes.subarray(32, 64));
    if (xm >= Q || ym >= Q) return false;
    const rinv = (() => {                     // 2^-256 mod q
        let [r0, r1, s0, s1] = [Q, (1n << 256n) % Q, 0n, 1n];
        while (r1 !== 0n) {
            const q = r0 / r1;
            [r0, r1, s0, s1] = [r1, r0 - q * r1, s1, s0 - q * s1];
        }
        return ((s0 % Q) + Q) % Q;
    })();
    const x = (xm * rinv) % Q, y = (ym * rinv) % Q;
    return (y * y - x * x * x - 3n) % Q === 0n;
}

// steps 1-5 (validity, challenges, Z_H(xi), L_1(xi), r_0) and the expansion of steps 6-9 into linear forms
async function analyse(kind, pTauFilename, proof, nBits, logger) {
    const gs = kind === "gs";
    const curve = await getCurveFromName("bn128");
    const Fr = curve.Fr;
    const Cm = proof.commitments, Ev = proof.evaluations;
    const acc = gs ? "S" : "Z", accEval = gs ? "sxiw" : "zxiw";
    const nFi = Object.keys(Cm).filter((k) => k.match(/^F\d/)).length;                 // verifier.js:23-28
    const nPols = nFi > 0 ? nFi : 1;
    const isVector = nPols > 1;
    const isSelected = Object.keys(Cm).filter((k) => k.match(/^selF/)).length === 1;
    const fName = (i) => (isVector ? `F${i}` : "F"), tName = (i) => (isVector ? `T${i}` : "T");
    const fEv = (i) => (isVector ? `f${i}xi` : "fxi"), tEv = (i) => (isVector ? `t${i}xi` : "txi");
    const fail = (msg) => {
        if (logger) logger.error(msg);
        return null;
    };
    // STEP 1: commitments are valid G1 elements (:50)
    const names = [];
    for (let i = 0; i < nPols; i++) names.push(fName(i), tName(i));
    if (isSelected) names.push("selF", "selT");
    names.push(acc, "Q", "Wxi", "Wxiw");
    for (const nm of names) if (!Cm[nm] || !validG1(curve, Cm[nm])) return fail(`${nm} is not a valid G1 element`);
    // STEP 2: evaluations are valid field elements (:61)
    const evNames = [];
    for (let i = 0; i < nPols; i++) {
        evNames.push(fEv(i));
        if (gs) evNames.push(tEv(i));
    }
    evNames.push(accEval);
    for (const nm of evNames) if (!Ev[nm] || Ev[nm].length !== 32 || leToBig(Ev[nm]) >= R) return fail(`${nm} is not a valid field element`);
    if (isSelected) for (const nm of ["selFxi", "selTxi"]) if (!Ev[nm] || Ev[nm].length !== 32) return fail(`missing evaluation ${nm}`);
    const val = {};
    for (const k of Object.keys(Ev)) val[k] = Fr.toObject(Ev[k]);
    // STEP 3: challenges (:246-312)
    const tr = new Keccak256Transcript(curve);
    for (let i = 0; i < nPols; i++) {
        tr.addPolCommitment(Cm[fName(i)]);
        tr.addPolCommitment(Cm[tName(i)]);
    }
    if (isSelected) {
        tr.addPolCommitment(Cm.selF);
        tr.addPolCommitment(Cm.selT);
    }
    let betaB = null;
    if (isVector) {
        betaB = tr.getChallenge();
        tr.addFieldElement(betaB);
    }
    const gammaB = tr.getChallenge();
    tr.addFieldElement(gammaB);
    tr.addPolCommitment(Cm[acc]);
    const alphaB = tr.getChallenge();
    tr.addFieldElement(alphaB);
    tr.addPolCommitment(Cm.Q);
    const xiB = tr.getChallenge();
    tr.addFieldElement(xiB);
    for (let i = 0; i < nPols; i++) {
        tr.addFieldElement(Ev[fEv(i)]);
        if (gs) tr.addFieldElement(Ev[tEv(i)]);
    }
    if (isSelected) {
        tr.addFieldElement(Ev.selFxi);
        tr.addFieldElement(Ev.selTxi);
    }
    tr.addFieldElement(Ev[accEval]);
    const vB = tr.getChallenge();
    tr.addFieldElement(vB);
    tr.addPolCommitment(Cm.Wxi);
    tr.addPolCommitment(Cm.Wxiw);
    const uB = tr.getChallenge();
    const beta = isVector ? Fr.toObject(betaB) : 0n;
    const [gamma, alpha, xi, v, u] = [gammaB, alphaB, xiB, vB, uB].map((b) => Fr.toObject(b));
    // STEP 4: Z_H(xi), L_1(xi)
    let xn = xi;
    for (let i = 0; i < nBits; i++) xn = mod(xn * xn);
    const ZHxi = mod(xn - 1n);
    const L1xi = mod(ZHxi * modinv(mod((1n << BigInt(nBits)) * mod(xi - 1n))));
    // STEP 5: r0 (:78-111 / grand-product :78-97)
    let r0 = 0n;
    if (isSelected) {
        r0 = mod((r0 + val.selTxi - val.selTxi * val.selTxi) * alpha);
        r0 = mod((r0 + val.selFxi - val.selFxi * val.selFxi) * alpha);
    }
    let fxi = 0n, txi = 0n;
    for (let i = nPols - 1; i >= 0; i--) {
        fxi = mod(fxi * beta + val[fEv(i)]);
        if (gs) txi = mod(txi * beta + val[tEv(i)]);
    }
    let fxig = mod(fxi + gamma);
    const txig = mod(txi + gamma);
    if (gs) {
        let r01 = mod(val.sxiw * fxig % R * txig);
        r01 = isSelected ? mod(r01 + val.selTxi * fxig - val.selFxi * txig) : mod(r01 + fxi - txi);
        r0 = mod((r0 + r01) * alpha);
    } else {
        let r01 = val.zxiw;
        r01 = isSelected ? mod(r01 * mod((gamma - 1n) * val.selTxi + 1n)) : mod(r01 * gamma);
        r0 = mod((r0 + r01) * alpha - L1xi);
        if (isSelected) fxig = mod((fxig - 1n) * val.selFxi + 1n);
    }
    // steps 6-9 as linear forms over the commitments
    const B = new Map();
    const add = (name, c) => B.set(name, mod((B.get(name) || 0n) + c));
    if (gs) {
        add(acc, L1xi - alpha * fxig % R * txig + u);
    } else {
        add(acc, L1xi - alpha * fxig + u);
        let c = mod(alpha * val.zxiw);
        if (isSelected) c = mod(c * val.selTxi);
        let bp = 1n;
        for (let i = 0; i < nPols; i++) {
            add(tName(i), c * bp);
            bp = mod(bp * beta);
        }
    }
    add("Q", -ZHxi);
    let vp = v;
    for (let i = 0; i < nPols; i++) {
        add(fName(i), vp);
        vp = mod(vp * v);
    }
    if (gs) for (let i = 0; i < nPols; i++) {
        add(tName(i), vp);
        vp = mod(vp * v);
    }
    if (isSelected) {
        add("selF", vp);
        vp = mod(vp * v);
        add("selT", vp);
    }
    let E = 0n;
    if (isSelected) {
        E = mod(E + val.selTxi);
        E = mod(E * v + val.selFxi);
    }
    if (gs) for (let i = nPols - 1; i >= 0; i--) E = mod(E * v + val[tEv(i)]);
    for (let i = nPols - 1; i >= 0; i--) E = mod(E * v + val[fEv(i)]);
    E = mod(E * v + u * val[accEval]);
    E = mod(E - r0);
    const w = Fr.toObject(Fr.w[nBits]);
    add("Wxi", xi);
    add("Wxiw", u * xi % R * w);
    const A = new Map([["Wxi", 1n], ["Wxiw", u]]);
    return { curve, Cm, A, B, gen: mod(-E), challenges: { beta, gamma, alpha, xi, v, u } };
}

async function verify(kind, pTauFilename, proof, nBits, logger) {
    const st = await analyse(kind, pTauFilename, proof, nBits, logger);
    if (!st) return false;
    const { curve, Cm, A, B, gen } = st;
    const Fr = curve.Fr, G1 = curve.G1;
    const side = async (terms, extra) => {
        const pts = [], ks = [];
        for (const [name, c] of terms) {
            pts.push(Cm[name]);
            ks.push(Fr.e(c));
        }
        if (extra !== undefined) {
            pts.push(G1.one);
            ks.push(Fr.e(extra));
        }
        return G1.linearCombination(pts, ks);
    };
    const Apt = await side(A), Bpt = await side(B, gen);
    const ff = await pairingCurve();
    const X2 = await readTauG2(pTauFilename);
    const isValid = await ff.pairingEq(ff.G1.neg(Apt), X2, Bpt, ff.G2.one);
    if (logger) (isValid ? logger.info : logger.error).call(logger, isValid ? "> VERIFICATION OK" : "> VERIFICATION FAILED");
    return isValid;
}

module.exports = { verify, analyse };
