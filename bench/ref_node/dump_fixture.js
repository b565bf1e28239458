#!/usr/bin/env node
// Run the UNMODIFIED reference provers (src/grandsum/mset_eq_kzg_prover.js:12, src/grandproduct/mset_eq_kzg_prover.js:12)
// on the seeded synthetic cases of tests/golden/make_golden.py and write what they produce as tests/golden/ref_<name>.json
// -- the same schema as the oracle-made fixtures next to them, plus "source": "reference-node".  With such a file in the
// tree, tests/test_reference_fixtures.py byte-compares the oracle (CPU run) and the CUDA path (GPU run) against GENUINE
// ffjavascript output: that is what lifts "parity unpinned".
//
//   node bench/ref_node/dump_fixture.js --ref /path/to/kzg-grandsums-study --ptau-dir tmp [--cases all|name,name] [--out tests/golden]
//
// The .ptau files come from  python bench/ref_node/make_ptau.py --out-dir tmp  (synthetic tau of seed 1001, same bytes the
// GPU tests use).  The Fiat-Shamir challenges are captured by wrapping Keccak256Transcript.prototype.getChallenge at run
// time; no reference file is edited.
"use strict";
const fs = require("fs");
const path = require("path");
const C = require("./common.js");

const TAU_SEED = 1001;
const CASES = [   // name, kind, seed, nbits, k, selected, rotate, ptau_power  -- tests/golden/make_golden.py::CASES + C3
    ["gs_c1_n8", "gs", 1, 8, 1, false, true, 11],
    ["gp_c1_n8", "gp", 1, 8, 1, false, true, 11],
    ["gs_c2_n11", "gs", 2, 11, 1, false, true, 11],
    ["gp_c2_n11", "gp", 2, 11, 1, false, true, 11],
    ["gs_vec_sel_n6_k3", "gs", 21, 6, 3, true, false, 6],
    ["gp_vec_sel_n6_k3", "gp", 21, 6, 3, true, false, 6],
    ["gs_perm_n9", "gs", 31, 9, 1, false, false, 9],
    ["gp_perm_n9", "gp", 31, 9, 1, false, false, 9],
    ["gs_c3_n16_k4", "gs", 3, 16, 4, true, false, 16],
    ["gp_c3_n16_k4", "gp", 3, 16, 4, true, false, 16],
];

async function main() {
    const argv = process.argv.slice(2);
    const ref = C.referenceRoot(argv);
    const ptauDir = C.arg(argv, "ptau-dir", "tmp");
    const outDir = C.arg(argv, "out", path.resolve(__dirname, "..", "..", "tests", "golden"));
    const want = C.arg(argv, "cases", "all");
    const refRequire = (p) => require(path.join(ref, p));
    const ff = require(require.resolve("ffjavascript", { paths: [ref] }));
    const ffVersion = require(require.resolve("ffjavascript/package.json", { paths: [ref] })).version;
    const { Evaluations } = refRequire("src/polynomial/evaluations.js");
    const provers = { gs: refRequire("src/grandsum/mset_eq_kzg_prover.js"), gp: refRequire("src/grandproduct/mset_eq_kzg_prover.js") };
    const verifiers = { gs: refRequire("src/grandsum/mset_eq_kzg_verifier.js"), gp: refRequire("src/grandproduct/mset_eq_kzg_verifier.js") };
    const { Keccak256Transcript } = refRequire("src/Keccak256Transcript.js");

    const curve = await ff.getCurveFromName("bn128");
    let captured = [];
    const original = Keccak256Transcript.prototype.getChallenge;
    Keccak256Transcript.prototype.getChallenge = function () {
        const ch = original.call(this);
        captured.push(ch);
        return ch;
    };

    let failures = 0;
    for (const [name, kind, seed, nbits, k, selected, rotate, power] of CASES) {
        if (want !== "all" && !want.split(",").includes(name)) continue;
        const ptau = path.join(ptauDir, "synthetic_" + String(power).padStart(2, "0") + ".ptau");
        if (!fs.existsSync(ptau)) { console.error(name + ": missing " + ptau + " (run make_ptau.py)"); failures++; continue; }
        const c = C.buildCase(curve, { seed, nbits, k, selected, rotate });
        const ev = (b) => new Evaluations(b.slice(), curve);
        const fs_ = k === 1 ? ev(c.colsF[0]) : c.colsF.map(ev);
        const ts_ = k === 1 ? ev(c.colsT[0]) : c.colsT.map(ev);
        captured = [];
        const t0 = process.hrtime.bigint();
        const proof = selected ? await provers[kind](ptau, fs_, ts_, ev(c.selF), ev(c.selT)) : await provers[kind](ptau, fs_, ts_);
        const ms = Number(process.hrtime.bigint() - t0) / 1e6;
        // the prover derives beta (vector arguments only), gamma, alpha, xi, v in this order (prover.js:181-413)
        const names = (k > 1 ? ["beta"] : []).concat(["gamma", "alpha", "xi", "v"]);
        const challenges = {};
        names.forEach((nm, i) => { challenges[nm] = "0x" + curve.Fr.toObject(captured[i]).toString(16); });
        const ok = await verifiers[kind](ptau, proof, nbits);
        const bytes = [];
        for (const key of Object.keys(proof.commitments)) bytes.push(Buffer.from(proof.commitments[key]));
        for (const key of Object.keys(proof.evaluations)) bytes.push(Buffer.from(proof.evaluations[key]));
        const out = {
            source: "reference-node", ffjavascript: ffVersion, node: process.version, reference_verifier_accepts: ok,
            prove_ms: ms, kind, seed, nbits, k, selected, rotate, ptau_power: power, tau_seed: TAU_SEED,
            ptau_sha256: C.sha256File(ptau),
            commitment_keys: Object.keys(proof.commitments), evaluation_keys: Object.keys(proof.evaluations),
            challenges, proof_bytes: Buffer.concat(bytes).toString("hex"),
        };
        fs.writeFileSync(path.join(outDir, "ref_" + name + ".json"), JSON.stringify(out, null, 1));
        // immediate verdict against the oracle-made fixture of the same case, when there is one
        const mine = path.join(outDir, name + ".json");
        let verdict = "no oracle fixture to compare";
        if (fs.existsSync(mine)) {
            const o = JSON.parse(fs.readFileSync(mine));
            const same = o.proof_bytes === out.proof_bytes && JSON.stringify(o.commitment_keys) === JSON.stringify(out.commitment_keys) &&
                JSON.stringify(o.evaluation_keys) === JSON.stringify(out.evaluation_keys) &&
                Object.keys(o.challenges).every((nm) => BigInt(o.challenges[nm]) === BigInt(out.challenges[nm]));
            verdict = same ? "IDENTICAL to the oracle fixture" : "DIFFERS from the oracle fixture";
            if (!same) failures++;
        }
        console.log(`${name}: reference proved in ${ms.toFixed(1)} ms, verifier ${ok ? "accepts" : "REJECTS"}; ${verdict}`);
    }
    Keccak256Transcript.prototype.getChallenge = original;
    await curve.terminate();
    console.log(failures === 0 ? "parity: green" : `parity: RED (${failures} case(s))`);
    process.exit(failures === 0 ? 0 : 1);
}

main().catch((e) => { console.error(e); process.exit(2); });
