// The reference's own protocol tests (test/mset_eq_kzg_grandsum.test.js:24-104, test/mset_eq_kzg_grandproduct.test.js) against
// the GPU-backed drop-in modules: prove, verify, assert.ok(isValid) -- for a machine with Node, the addon and a B200.
//   PTAU=tmp/synthetic_11.ptau npx mocha js/test/mset_eq.test.js     (python bench/ref_node/make_ptau.py writes the file)
"use strict";
const assert = require("assert");
const path = require("path");
const kzg = require("../index.js");

const pTauFilename = process.env.PTAU || path.join("tmp", "synthetic_11.ptau");
const rnd = (lo, hi) => lo + Math.floor(Math.random() * (hi - lo + 1));

function rotated(evalsF) {                   // T = F rotated right by one (test/...test.js:28-30)
    const evalsT = kzg.Evaluations.fromEvals(evalsF);
    evalsT.setEvaluation(1, evalsF.getEvaluationSequence(0, evalsF.length() - 1));
    evalsT.setEvaluation(0, evalsF.getEvaluation(evalsF.length() - 1));
    return evalsT;
}

for (const [name, prover, verifier] of [
    ["grand-sum", kzg.mset_eq_kzg_grandsum_prover, kzg.mset_eq_kzg_grandsum_verifier],
    ["grand-product", kzg.mset_eq_kzg_grandproduct_prover, kzg.mset_eq_kzg_grandproduct_verifier]]) {
    describe(`Protocols based on ${name}s and KZG (B200 backend)`, function () {
        let curve;
        before(async () => { curve = await kzg.getCurveFromName("bn128"); });
        after(async () => { await curve.terminate(); });

        it("Should proof and verify a standard multiset equality", async () => {
            const nBits = rnd(1, 10);
            const evalsF = kzg.Evaluations.getRandomEvals(2 ** nBits, curve);
            const evalsT = rotated(evalsF);
            const proof = await prover(pTauFilename, evalsF, evalsT);
            assert.deepStrictEqual(Object.keys(proof.commitments), ["F", "T", name === "grand-sum" ? "S" : "Z", "Q", "Wxi", "Wxiw"]);
            assert.ok(await verifier(pTauFilename, proof, nBits));
        });

        it("Should proof and verify a vector multiset equality", async () => {
            const nBits = rnd(1, 10), nPols = rnd(2, 10);
            const evalsF = [], evalsT = [];
            for (let i = 0; i < nPols; i++) {
                evalsF.push(kzg.Evaluations.getRandomEvals(2 ** nBits, curve));
                evalsT.push(rotated(evalsF[i]));
            }
            const proof = await prover(pTauFilename, evalsF, evalsT);
            assert.ok(await verifier(pTauFilename, proof, nBits));
        });

        it("Should proof and verify a selected multiset equality", async () => {
            const nBits = rnd(1, 10), n = 2 ** nBits;
            const evalsF = kzg.Evaluations.getRandomEvals(n, curve);
            const evalsT = rotated(evalsF);
            const selF = kzg.Evaluations.getOneEvals(n, curve), selT = kzg.Evaluations.getOneEvals(n, curve);
            selF.setEvaluation(n - 1, curve.Fr.zero);
            selT.setEvaluation(0, curve.Fr.zero);
            const proof = await prover(pTauFilename, evalsF, evalsT, selF, selT);
            assert.ok(await verifier(pTauFilename, proof, nBits));
        });

        it("Should reject a tampered proof and refuse different multisets", async () => {
            const nBits = 4;
            const evalsF = kzg.Evaluations.getRandomEvals(2 ** nBits, curve);
            const proof = await prover(pTauFilename, evalsF, rotated(evalsF));
            const key = Object.keys(proof.evaluations)[0];
            proof.evaluations[key] = curve.Fr.add(proof.evaluations[key], curve.Fr.one);
            assert.strictEqual(await verifier(pTauFilename, proof, nBits), false);
            const other = kzg.Evaluations.getRandomEvals(2 ** nBits, curve);
            await assert.rejects(prover(pTauFilename, kzg.Evaluations.getRandomEvals(2 ** nBits, curve), other), /is not well calculated/);
        });
    });
}
