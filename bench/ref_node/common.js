// Shared helpers of the Node-side landing gear (bench/ref_node): the synthetic input generator of SURVEY.md 8(d),
// restated in JavaScript so that a machine with Node builds the SAME columns as oracle/py/inputs.py and
// kzg_grandsums_study_b200/synthetic.py from a seed, and the fixture schema of tests/golden/*.json.
//
// Nothing here touches the reference: dump_fixture.js / time_msm.js `require` the unmodified reference tree given by
// --ref (default: ../../../reference or $KZG_REFERENCE) and the ffjavascript installed next to it.
"use strict";
const crypto = require("crypto");
const fs = require("fs");
const path = require("path");

const R = 21888242871839275222246405745257275088548364400416034343698204186575808495617n;
const M64 = (1n << 64n) - 1n;

class SplitMix64 {
    constructor(seed) { this.s = BigInt(seed) & M64; }
    next() {
        this.s = (this.s + 0x9E3779B97F4A7C15n) & M64;
        let z = this.s;
        z = ((z ^ (z >> 30n)) * 0xBF58476D1CE4E5B9n) & M64;
        z = ((z ^ (z >> 27n)) * 0x94D049BB133111EBn) & M64;
        return z ^ (z >> 31n);
    }
    // one Fr element: 4 consecutive outputs as LE limbs, top two bits of limb 3 cleared, rejected when >= r
    fr() {
        for (;;) {
            const l0 = this.next(), l1 = this.next(), l2 = this.next(), l3 = this.next();
            const x = l0 | (l1 << 64n) | (l2 << 128n) | ((l3 & 0x3FFFFFFFFFFFFFFFn) << 192n);
            if (x < R) return x;
        }
    }
}

function toLE32(x) {
    const out = new Uint8Array(32);
    for (let i = 0; i < 32; i++) { out[i] = Number(x & 0xFFn); x >>= 8n; }
    return out;
}

// n standard-form little-endian scalars (what the provers take for F / T, like Evaluations.getRandomEvals)
function randomColumn(seed, n) {
    const g = new SplitMix64(seed);
    const buf = new Uint8Array(32 * n);
    for (let i = 0; i < n; i++) buf.set(toLE32(g.fr()), 32 * i);
    return buf;
}

function tauFromSeed(seed) { return new SplitMix64(seed).fr(); }

// Fisher-Yates driven by SplitMix64(seed): for i = n-1 .. 1: j = next() % (i+1); swap(i, j)
function permutation(seed, n) {
    const g = new SplitMix64(seed);
    const p = new Array(n);
    for (let i = 0; i < n; i++) p[i] = i;
    for (let i = n - 1; i > 0; i--) {
        const j = Number(g.next() % BigInt(i + 1));
        const t = p[i]; p[i] = p[j]; p[j] = t;
    }
    return p;
}

function rotateRight(col) {          // T = F rotated right by one (test/mset_eq_kzg_grandsum.test.js:28-30)
    const n = col.length / 32;
    const out = new Uint8Array(col.length);
    out.set(col.subarray(32 * (n - 1)), 0);
    out.set(col.subarray(0, 32 * (n - 1)), 32);
    return out;
}

function permute(col, perm) {
    const out = new Uint8Array(col.length);
    for (let i = 0; i < perm.length; i++) out.set(col.subarray(32 * perm[i], 32 * perm[i] + 32), 32 * i);
    return out;
}

// the case construction of tests/golden/make_golden.py::columns (seeded, no randomness of its own)
function buildCase(curve, c) {
    const n = 1 << c.nbits;
    const colsF = [], colsT = [];
    for (let i = 0; i < c.k; i++) colsF.push(randomColumn(c.seed * 100 + i, n));
    if (c.selected || c.rotate) {
        for (let i = 0; i < c.k; i++) colsT.push(rotateRight(colsF[i]));
    } else {
        const perm = permutation(c.seed, n);
        for (let i = 0; i < c.k; i++) colsT.push(permute(colsF[i], perm));
    }
    let selF = null, selT = null;
    if (c.selected) {                // selF[n-1] = 0, selT[0] = 0 (test/...test.js:68-71); selectors are MONTGOMERY values
        selF = new Uint8Array(32 * n);
        selT = new Uint8Array(32 * n);
        for (let i = 0; i < n; i++) {
            if (i !== n - 1) selF.set(curve.Fr.one, 32 * i);
            if (i !== 0) selT.set(curve.Fr.one, 32 * i);
        }
    }
    return { colsF, colsT, selF, selT };
}

const hex = (u8) => Buffer.from(u8).toString("hex");
const sha256File = (p) => crypto.createHash("sha256").update(fs.readFileSync(p)).digest("hex");

function referenceRoot(argv) {
    const i = argv.indexOf("--ref");
    const root = i >= 0 ? argv[i + 1] : (process.env.KZG_REFERENCE || path.resolve(__dirname, "..", "..", "..", "reference"));
    if (!fs.existsSync(path.join(root, "src", "grandsum", "mset_eq_kzg_prover.js")))
        throw new Error("reference tree not found at " + root + " (pass --ref /path/to/kzg-grandsums-study)");
    return root;
}

function arg(argv, name, dflt) {
    const i = argv.indexOf("--" + name);
    return i >= 0 ? argv[i + 1] : dflt;
}

module.exports = { R, SplitMix64, toLE32, randomColumn, tauFromSeed, permutation, rotateRight, permute, buildCase, hex,
                   sha256File, referenceRoot, arg };
