// The `curve` object the reference gets from ffjavascript's getCurveFromName("bn128") (test/*.test.js:14-21,
// src/ptau_utils.js:13), rebuilt on the addon: scalar Fr helpers on 32-byte Montgomery little-endian buffers (BigInt on the
// host -- O(1) work per proof), bulk Fr / G1 calls on the device.  One context per process and device; curve.terminate()
// releases it (ffjavascript's curve.terminate()).
"use strict";
const fs = require("fs");
const path = require("path");

function loadAddon() {
    const candidates = [
        process.env.KZGB200_ADDON,
        path.join(__dirname, "..", "..", "addon", "build", "Release", "kzgb200.node"),
        path.join(__dirname, "..", "build", "Release", "kzgb200.node"),
    ].filter(Boolean);
    for (const c of candidates) if (fs.existsSync(c)) return require(c);
    throw new Error("kzgb200 addon not built (cd addon && node-gyp rebuild); the prover hot path has no CPU fallback. Looked in: " +
        candidates.join(", "));
}

const Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583n;
const R = 21888242871839275222246405745257275088548364400416034343698204186575808495617n;
const MONT = (1n << 256n) % R;

function modpow(b, e, m) {
    let r = 1n;
    b %= m;
    while (e > 0n) {
        if (e & 1n) r = (r * b) % m;
        b = (b * b) % m;
        e >>= 1n;
    }
    return r;
}
const MONT_INV = modpow(MONT, R - 2n, R);

function leToBig(buf) {
    let x = 0n;
    for (let i = buf.length - 1; i >= 0; i--) x = (x << 8n) | BigInt(buf[i]);
    return x;
}
function bigToLE(x, n) {
    const out = new Uint8Array(n);
    for (let i = 0; i < n; i++) {
        out[i] = Number(x & 0xFFn);
        x >>= 8n;
    }
    return out;
}

class Fr {
    constructor(curve) {
        this._curve = curve;
        this.n8 = 32;
        this.n64 = 4;
        this.p = R;
        this.zero = new Uint8Array(32);
        this.one = this.e(1n);
        this.negone = this.e(R - 1n);
        this.s = 28;
        // Fr.w[k] = 5^((r-1)/2^k): primitive 2^k-th roots of unity (ffjavascript convention, SURVEY.md B.1)
        this.w = new Array(29);
        let w = modpow(5n, (R - 1n) >> 28n, R);
        for (let k = 28; k >= 0; k--) {
            this.w[k] = this.e(w);
            w = (w * w) % R;
        }
    }
    // host scalar helpers ------------------------------------------------------------------------------------------
    e(x) {
        if (x instanceof Uint8Array) return x;
        let v = BigInt(x) % R;
        if (v < 0n) v += R;
        return bigToLE((v * MONT) % R, 32);
    }
    toObject(a) { return (leToBig(a) * MONT_INV) % R; }
    toString(a, radix = 10) { return this.toObject(a).toString(radix); }
    fromRprLE(buf, off = 0) { return this.e(leToBig(buf.subarray(off, off + 32))); }
    _bin(a, b, f) { return this.e(f(this.toObject(a), this.toObject(b))); }
    add(a, b) { return this._bin(a, b, (x, y) => x + y); }
    sub(a, b) { return this._bin(a, b, (x, y) => x - y + R); }
    mul(a, b) { return this._bin(a, b, (x, y) => x * y); }
    neg(a) { return this.e(R - this.toObject(a)); }
    square(a) { return this.mul(a, a); }
    inv(a) {
        const x = this.toObject(a);
        if (x === 0n) throw new Error("Division by zero");
        return this.e(modpow(x, R - 2n, R));
    }
    div(a, b) { return this.mul(a, this.inv(b)); }
    exp(a, k) { return this.e(modpow(this.toObject(a), BigInt(k), R)); }
    eq(a, b) { return Buffer.compare(Buffer.from(a), Buffer.from(b)) === 0; }
    isZero(a) { return this.eq(a, this.zero); }
    random() {
        for (;;) {
            const b = require("crypto").randomBytes(32);
            b[31] &= 0x3f;
            const x = leToBig(b);
            if (x < R) return this.e(x);
        }
    }
    toRprBE(buff, offset, a) {              // Fr.toRprBE(buffer, offset, element)  (Keccak256Transcript.js:45)
        const out = this._curve.addon.kzg_fr_to_rpr_be(a);
        buff.set(out, offset);
    }
    // bulk calls: device (each returns a fresh host buffer, like ffjavascript) ---------------------------------------
    async batchToMontgomery(buf) { return this._curve._unary(buf, "kzg_fr_to_mont"); }
    async batchFromMontgomery(buf) { return this._curve._unary(buf, "kzg_fr_from_mont"); }
    async batchInverse(buf) { return this._curve._unary(buf, "kzg_fr_batch_inverse"); }
    async fft(buf) { return this._curve._ntt(buf, 0); }
    async ifft(buf) { return this._curve._ntt(buf, 1); }
}

class G1 {
    constructor(curve) {
        this._curve = curve;
        this.F = { n8: 32, n64: 4 };
        this.zeroAffine = new Uint8Array(64);
        this.oneAffine = Uint8Array.from(Buffer.concat([Buffer.from(bigToLE((1n << 256n) % Q, 32)), Buffer.from(bigToLE((2n << 256n) % Q, 32))]));
        this.one = this.oneAffine;
        this.zero = this.zeroAffine;
    }
    // G1.toRprUncompressed(buffer, offset, point)  (Keccak256Transcript.js:42)
    toRprUncompressed(buff, offset, p) { buff.set(this._curve.addon.kzg_g1_to_rpr_uncompressed(p.subarray(0, 64)), offset); }
    // G1.multiExpAffine(bases, scalars) -> Jacobian triple (x, y, 1) | zeros (polynomial.js:1112)
    async multiExpAffine(bases, scalars) {
        const a = this._curve.addon;
        const n = Math.floor(scalars.byteLength / 32);
        const r = a.kzg_g1_msm_affine(this._curve.ctx, asU8(bases), asU8(scalars), n, 0);
        return Uint8Array.from(r.out_jacobian);
    }
    toAffine(jac) { return Uint8Array.from(jac.subarray(0, 64)); }
    isZero(p) { return p.subarray(0, 64).every((b) => b === 0); }
    eq(a, b) { return Buffer.compare(Buffer.from(this.toAffine(a)), Buffer.from(this.toAffine(b))) === 0; }
    // sum_i k_i P_i for a handful of points (verifier): one small device MSM; scalars are Montgomery Fr elements
    async linearCombination(points, scalarsMont) {
        const Fr = this._curve.Fr;
        const bases = Buffer.concat(points.map((p) => Buffer.from(p.subarray(0, 64))));
        const scal = Buffer.concat(scalarsMont.map((s) => Buffer.from(bigToLE(Fr.toObject(s), 32))));
        return this.toAffine(await this.multiExpAffine(bases, scal));
    }
    async timesFr(p, k) { return this.linearCombination([p], [k]); }
}

function asU8(b) {
    if (b instanceof Uint8Array) return b;
    if (b && typeof b.slice === "function" && b.byteLength !== undefined && b.buffers) return b.slice(0, b.byteLength);   // ffjavascript BigBuffer
    if (ArrayBuffer.isView(b)) return new Uint8Array(b.buffer, b.byteOffset, b.byteLength);
    return new Uint8Array(b);
}

class Curve {
    constructor(device = 0) {
        this.addon = loadAddon();
        this.name = "bn128";
        this.q = Q;
        this.r = R;
        this.device = device;
        this.ctx = this.addon.kzg_ctx_create(device, null);       // throws without a CUDA device: no CPU fallback
        this.Fr = new Fr(this);
        this.G1 = new G1(this);
        this.F1 = { n8: 32, n64: 4 };
        this._srs = new Map();
    }
    // host bytes -> device vector; caller frees
    upload(buf) {
        const u8 = asU8(buf);
        if (u8.byteLength % 32) throw new Error("buffer length is not a multiple of 32");
        const n = u8.byteLength / 32;
        const h = this.addon.kzg_buf_alloc(this.ctx, n);
        if (n) this.addon.kzg_buf_upload(this.ctx, h, 0, u8, n);
        return h;
    }
    download(h) {
        const n = Number(this.addon.kzg_buf_len(h));
        return Uint8Array.from(this.addon.kzg_buf_download(this.ctx, h, 0, n));
    }
    free(...hs) { for (const h of hs) if (h) this.addon.kzg_buf_free(this.ctx, h); }
    _unary(buf, fn) {
        const a = this.upload(buf);
        const out = this.addon.kzg_buf_alloc(this.ctx, this.addon.kzg_buf_len(a));
        try {
            this.addon[fn](this.ctx, a, out);
            return this.download(out);
        } finally {
            this.free(a, out);
        }
    }
    _ntt(buf, inverse) {
        const a = this.upload(buf);
        const out = this.addon.kzg_buf_alloc(this.ctx, this.addon.kzg_buf_len(a));
        try {
            this.addon.kzg_fr_ntt(this.ctx, a, out, inverse);     // throws "fft must be multiple of 2" like ffjavascript
            return this.download(out);
        } finally {
            this.free(a, out);
        }
    }
    // device-resident [tau^i]_1 with its window table, cached per (file identity, size)  (prover.js:15-16,83-85)
    loadSrs(ptauPath, nPoints) {
        const st = fs.statSync(ptauPath);
        const key = path.resolve(ptauPath) + ":" + nPoints;
        const ident = st.mtimeMs + ":" + st.size;
        let hit = this._srs.get(key);
        if (hit && hit.ident !== ident) {
            this.addon.kzg_srs_free(this.ctx, hit.srs);
            hit = undefined;
        }
        if (!hit) {
            const r = this.addon.kzg_srs_load_ptau(this.ctx, ptauPath, nPoints);
            this.addon.kzg_srs_precompute(this.ctx, r.out, 0);
            hit = { srs: r.out, power: r.power_out, ident };
            this._srs.set(key, hit);
        }
        return hit;
    }
    async terminate() {
        for (const v of this._srs.values()) this.addon.kzg_srs_free(this.ctx, v.srs);
        this._srs.clear();
        if (this.ctx) this.addon.kzg_ctx_destroy(this.ctx);
        this.ctx = null;
        CURVES.delete(this.device);
    }
}

const CURVES = new Map();
async function getCurveFromName(name = "bn128", device = 0) {
    if (!["bn128", "bn254", "altbn128"].includes(String(name).toLowerCase())) throw new Error("Curve not supported: " + name);
    if (!CURVES.has(device)) CURVES.set(device, new Curve(device));
    return CURVES.get(device);
}
async function getCurveFromQ(q, device = 0) {
    if (BigInt(q) !== Q) throw new Error("Curve not supported: " + q);
    return getCurveFromName("bn128", device);
}

module.exports = { getCurveFromName, getCurveFromQ, Curve, asU8, leToBig, bigToLE, Q, R };
