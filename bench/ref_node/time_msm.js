#!/usr/bin/env node
// The reference's own CPU path for BASELINE config 5: time curve.G1.multiExpAffine (what Polynomial.multiExponentiation
// calls, src/polynomial/polynomial.js:1106-1115) and, optionally, a whole grand-sum proof, on THIS machine's cores with
// the ffjavascript WASM worker pool -- the baseline north_star names and the image of this repo cannot run (no Node).
//
//   node bench/ref_node/time_msm.js --ref /path/to/kzg-grandsums-study --ptau tmp/synthetic_23.ptau [--logs 16,18,20,22,24] [--prove 20]
//
// Prints one JSON line per size: {"what":"msm","log_n":..,"ms":..,"mpts_s":..,"threads":..,"result":"<64-byte affine hex>"}.
// `result` is the canonical affine point: bench.py's known answer for seed 6 must equal it.
"use strict";
const fs = require("fs");
const os = require("os");
const path = require("path");
const C = require("./common.js");

async function main() {
    const argv = process.argv.slice(2);
    const ref = C.referenceRoot(argv);
    const ptau = C.arg(argv, "ptau", "tmp/synthetic_23.ptau");
    const logs = C.arg(argv, "logs", "16,18,20").split(",").map(Number);
    const proveLog = C.arg(argv, "prove", null);
    const ff = require(require.resolve("ffjavascript", { paths: [ref] }));
    const { readBinFile } = require(require.resolve("@iden3/binfileutils", { paths: [ref] }));
    const curve = await ff.getCurveFromName("bn128");
    const threads = Math.min(os.cpus().length, 64);
    const { fd, sections } = await readBinFile(ptau, "ptau", 1, 1 << 22, 1 << 24);
    for (const logN of logs) {
        const n = 2 ** logN;
        const bases = new ff.BigBuffer(n * 64);
        await fd.readToBuffer(bases, 0, n * 64, sections[2][0].p);
        const scal = C.randomColumn(6, n);                   // SURVEY.md 8d: C5 uses seed 6, standard-form scalars
        const scalars = new ff.BigBuffer(n * 32);
        scalars.set(scal, 0);
        const t0 = process.hrtime.bigint();
        const res = await curve.G1.multiExpAffine(bases, scalars, undefined, "msm");
        const ms = Number(process.hrtime.bigint() - t0) / 1e6;
        console.log(JSON.stringify({ what: "msm", log_n: logN, ms, mpts_s: n / ms / 1e3, threads,
                                     result: C.hex(curve.G1.toAffine(res)) }));
    }
    await fd.close();
    if (proveLog !== null) {
        const nbits = Number(proveLog);
        const { Evaluations } = require(path.join(ref, "src/polynomial/evaluations.js"));
        const prover = require(path.join(ref, "src/grandsum/mset_eq_kzg_prover.js"));
        const n = 2 ** nbits;
        const seed = nbits === 22 ? 5 : 4;                   // C4: seeds 4 (2^20) and 5 (2^22), T = permutation of F
        const f = C.randomColumn(seed, n);
        const t = C.permute(f, C.permutation(seed, n));
        const file = C.arg(argv, "prove-ptau", `tmp/synthetic_${String(nbits).padStart(2, "0")}.ptau`);
        const t0 = process.hrtime.bigint();
        const proof = await prover(file, new Evaluations(f, curve), new Evaluations(t, curve));
        const ms = Number(process.hrtime.bigint() - t0) / 1e6;
        const bytes = [];
        for (const key of Object.keys(proof.commitments)) bytes.push(Buffer.from(proof.commitments[key]));
        for (const key of Object.keys(proof.evaluations)) bytes.push(Buffer.from(proof.evaluations[key]));
        const sha = require("crypto").createHash("sha256").update(Buffer.concat(bytes)).digest("hex");
        console.log(JSON.stringify({ what: "grandsum_prove", nbits, ms, threads, proof_sha256: sha }));
    }
    await curve.terminate();
}

main().catch((e) => { console.error(e); process.exit(2); });
