// Drop-in for the part of reference src/polynomial/polynomial.js that the provers, the verifiers and the reference's
// tests of those use (SURVEY.md 8a): fromEvaluations, fromCoefficientsArray, fromPolynomial, zero, Lagrange1, clone,
// isEqual, getCoef / setCoef, length, degree, evaluate, add, sub, mulScalar, addScalar, subScalar, multiply, shiftOmega,
// divByXSubValue, divZh, multiExponentiation.  `coef` is a PUBLIC host buffer, as in the reference (polynomial.test.js
// reads and writes it); every bulk method stages it on the device, runs the library kernel and returns / stores fresh host
// buffers -- the same "every call returns fresh buffers" contract as ffjavascript.  (The fused provers do not go through
// this class: they keep everything device-resident across the five rounds.)
// Error strings are the reference's ("Polynomial does not divide", "Polynomial is not divisible", ...).
"use strict";

class Polynomial {
    constructor(coefficients, curve, logger) {
        this.coef = coefficients;
        this.curve = curve;
        this.Fr = curve.Fr;
        this.G1 = curve.G1;
        this.logger = logger;
    }
    // ---- plumbing ---------------------------------------------------------------------------------------------------
    _with(handles, fn) {             // run fn(addon, ctx, ...handles), free every handle (inputs and results) afterwards
        try {
            return fn(this.curve.addon, this.curve.ctx);
        } finally {
            this.curve.free(...handles.filter(Boolean));
        }
    }
    static async fromEvaluations(buffer, curve, logger) {         // polynomial.js:33-37: iNTT, natural order, scaled by 1/n
        return new Polynomial(await curve.Fr.ifft(buffer), curve, logger);
    }
    static fromCoefficientsArray(array, curve, logger) {          // :39-48
        const Fr = curve.Fr;
        const buff = new Uint8Array(array.length * Fr.n8);
        for (let i = 0; i < array.length; i++) buff.set(array[i], i * Fr.n8);
        return new Polynomial(buff, curve, logger);
    }
    static fromPolynomial(polynomial, curve, logger) {            // :50-61
        return new Polynomial(polynomial.coef.slice(), curve, logger);
    }
    static zero(length, curve, logger) { return new Polynomial(new Uint8Array(length * curve.Fr.n8), curve, logger); }
    static async Lagrange1(power, curve, logger) {                // :68-78 (iNTT of e_0: every coefficient is 1/n)
        const h = curve.addon.kzg_poly_lagrange1(curve.ctx, power);
        try {
            return new Polynomial(curve.download(h), curve, logger);
        } finally {
            curve.free(h);
        }
    }
    clone() { return Polynomial.fromPolynomial(this, this.curve, this.logger); }
    isEqual(polynomial) {                                         // :84-94
        const degree = this.degree();
        if (degree !== polynomial.degree()) return false;
        for (let i = 0; i < degree + 1; i++) if (!this.Fr.eq(this.getCoef(i), polynomial.getCoef(i))) return false;
        return true;
    }
    getCoef(index) {
        const i_n8 = index * this.Fr.n8;
        if (i_n8 + this.Fr.n8 > this.coef.byteLength) return this.Fr.zero;
        return this.coef.slice(i_n8, i_n8 + this.Fr.n8);
    }
    setCoef(index, value) {
        if (index > this.length() - 1) throw new Error("Coef index is not available");
        this.coef.set(value, index * this.Fr.n8);
    }
    length() {
        const length = this.coef.byteLength / this.Fr.n8;
        if (length !== Math.floor(this.coef.byteLength / this.Fr.n8)) throw new Error("Polynomial coefficients buffer has incorrect size");
        return length;
    }
    degree() {                                                    // :212-226 (top-down zero scan, on the device)
        if (this.length() === 0) return 0;
        const h = this.curve.upload(this.coef);
        return this._with([h], (a, ctx) => Number(a.kzg_poly_degree(ctx, h)));
    }
    evaluate(point) {                                             // :228-238 (Horner)
        const h = this.curve.upload(this.coef);
        return this._with([h], (a, ctx) => Uint8Array.from(a.kzg_poly_evaluate(ctx, h, point)));
    }
    fastEvaluate(point) { return this.evaluate(point); }
    // ---- element-wise (the receiver is modified and returned, as in the reference) ----------------------------------
    _addsub(polynomial, blindingValue, fn) {
        const x = this.curve.upload(this.coef), y = this.curve.upload(polynomial.coef);
        let out = null;
        try {
            out = this.curve.addon[fn](this.curve.ctx, x, y);     // the result takes the longer length (:276-312)
            this.coef = this.curve.download(out);
        } finally {
            this.curve.free(x, y, out);
        }
        if (blindingValue !== undefined) this.coef.set(this.Fr[fn === "kzg_poly_add" ? "add" : "sub"](this.coef.subarray(0, 32), blindingValue), 0);
        return this;
    }
    add(polynomial, blindingValue) { return this._addsub(polynomial, blindingValue, "kzg_poly_add"); }
    sub(polynomial, blindingValue) { return this._addsub(polynomial, blindingValue, "kzg_poly_sub"); }
    _scalar(value, fn) {
        const h = this.curve.upload(this.coef);
        this._with([h], (a, ctx) => {
            a[fn](ctx, h, value);
            this.coef = this.curve.download(h);
        });
        return this;
    }
    mulScalar(value) { return this._scalar(value, "kzg_poly_mul_scalar"); }       // :395-406
    addScalar(value) { return this._scalar(value, "kzg_poly_add_scalar"); }       // :408-414
    subScalar(value) { return this._scalar(value, "kzg_poly_sub_scalar"); }       // :416-422
    // ---- products and divisions -------------------------------------------------------------------------------------
    _binary(other, fn) {
        const x = this.curve.upload(this.coef), y = other ? this.curve.upload(other.coef) : null;
        let out = null;
        try {
            out = y ? this.curve.addon[fn](this.curve.ctx, x, y) : this.curve.addon[fn](this.curve.ctx, x);
            return this.curve.download(out);
        } finally {
            this.curve.free(x, y, out);
        }
    }
    async multiply(polynomial) {                                  // :352-376 (receiver becomes the product)
        this.coef = this._binary(polynomial, "kzg_poly_multiply");
        return this;
    }
    async shiftOmega() {                                          // :378-393: p(X) -> p(w X)
        this.coef = this._binary(null, "kzg_poly_shift_omega");
        return this;
    }
    divByXSubValue(value) {                                       // :814-851 "Polynomial does not divide"
        const x = this.curve.upload(this.coef);
        let out = null;
        try {
            out = this.curve.addon.kzg_poly_div_x_sub_value(this.curve.ctx, x, value);
            this.coef = this.curve.download(out);
        } finally {
            this.curve.free(x, out);
        }
        return this;
    }
    divZh(domainSize) {                                           // :853-888 "Polynomial is not divisible"
        const x = this.curve.upload(this.coef);
        let out = null;
        try {
            out = this.curve.addon.kzg_poly_div_zh(this.curve.ctx, x, domainSize);
            this.coef = this.curve.download(out);
        } finally {
            this.curve.free(x, out);
        }
        return this;
    }
    truncate() {                                                  // :1006-1017
        const deg = this.degree();
        if (deg + 1 < this.length()) this.coef = this.coef.slice(0, (deg + 1) * this.Fr.n8);
    }
    // commit: sum coef_i [tau^i]_1 -> Jacobian (x, y, 1)  (polynomial.js:1106-1115).  PTau: 64-byte affine points (host
    // buffer / BigBuffer) OR an SRS handle from curve.loadSrs(...) (device-resident, with its window table).
    async multiExponentiation(PTau, name) {
        const a = this.curve.addon, ctx = this.curve.ctx;
        const n = this.degree() + 1;
        if (PTau && PTau.srs) {
            const h = this.curve.upload(this.coef);
            return this._with([h], () => {
                const aff = a.kzg_commit(ctx, PTau.srs, h);
                const jac = new Uint8Array(96);
                jac.set(aff, 0);
                if (!aff.every((b) => b === 0)) jac.set(this.G1.one.subarray(0, 32), 64);   // z = 1 (Montgomery)
                return jac;
            });
        }
        const scalars = await this.Fr.batchFromMontgomery(this.coef.subarray(0, n * this.Fr.n8));
        const bases = require("../curve.js").asU8(PTau).subarray(0, n * 64);
        return this.G1.multiExpAffine(bases, scalars, this.logger, name);
    }
    print() {
        let res = "";
        for (let i = this.degree(); i >= 0; i--) {
            const c = this.getCoef(i);
            if (!this.Fr.eq(this.Fr.zero, c)) res += (res ? " + " : "") + this.Fr.toString(c) + (i > 1 ? `x^${i}` : i > 0 ? "x" : "");
        }
        console.log(res);
    }
}
module.exports = { Polynomial };
