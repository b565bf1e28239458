// Drop-in for reference src/polynomial/evaluations.js:5-137.  `eval` is a public host buffer (callers read and replace
// it: prover.js:147,170); the bulk work (fromPolynomial's forward NTT) runs on the device.
"use strict";

class Evaluations {
    constructor(evaluations, curve) {
        this.eval = evaluations;
        this.curve = curve;
        this.Fr = curve.Fr;
    }
    // zero-pad to nextpow2(len) * extension, forward NTT over the bigger subgroup (evaluations.js:12-21)
    static async fromPolynomial(polynomial, extension, curve) {
        const a = curve.addon;
        const coef = curve.upload(polynomial.coef);
        let out = null;
        try {
            out = a.kzg_fr_extend_ntt(curve.ctx, coef, extension);
            return new Evaluations(curve.download(out), curve);
        } finally {
            curve.free(coef, out);
        }
    }
    static fromArray(array, curve) {
        const buffer = new Uint8Array(array.length * curve.Fr.n8);
        for (let i = 0; i < array.length; i++) buffer.set(array[i], i * curve.Fr.n8);
        return new Evaluations(buffer, curve);
    }
    static fromEvals(evals) { return new Evaluations(evals.eval.slice(), evals.curve); }
    static _filled(length, curve, value) {
        const buffer = new Uint8Array(length * curve.Fr.n8);
        for (let i = 0; i < length; i++) buffer.set(value, i * curve.Fr.n8);
        return new Evaluations(buffer, curve);
    }
    static getOneEvals(length, curve) { return Evaluations._filled(length, curve, curve.Fr.one); }
    static getZeroEvals(length, curve) { return Evaluations._filled(length, curve, curve.Fr.zero); }
    static getRandomEvals(length, curve) {
        const buffer = new Uint8Array(length * curve.Fr.n8);
        for (let i = 0; i < length; i++) buffer.set(curve.Fr.random(), i * curve.Fr.n8);
        return new Evaluations(buffer, curve);
    }
    static getRandomBinEvals(length, curve) {
        const buffer = new Uint8Array(length * curve.Fr.n8);
        for (let i = 0; i < length; i++) buffer.set(Math.floor(Math.random() * 2) === 1 ? curve.Fr.one : curve.Fr.zero, i * curve.Fr.n8);
        return new Evaluations(buffer, curve);
    }
    getEvaluation(index) {
        if ((index + 1) * this.Fr.n8 > this.eval.byteLength) throw new Error("Evaluations.getEvaluation() out of bounds");
        return this.eval.slice(index * this.Fr.n8, (index + 1) * this.Fr.n8);
    }
    getEvaluationSequence(start, end) {
        if (start > end) throw new Error("Evaluations.getEvaluationSequence() start index is greater than end index");
        else if (start === end) throw new Error("Use Evaluations.getEvaluation() instead");
        if (end > this.length() - 1) throw new Error("Evaluations.getEvaluationSequence() end index is out of bounds");
        return this.eval.slice(start * this.Fr.n8, end * this.Fr.n8);
    }
    setEvaluation(index, value) {
        if (index > this.length() - 1) throw new Error("Evaluation index is out of bounds");
        this.eval.set(value, index * this.Fr.n8);
    }
    length() {
        const length = this.eval.byteLength / this.Fr.n8;
        if (length !== Math.floor(this.eval.byteLength / this.Fr.n8)) throw new Error("Polynomial evaluations buffer has incorrect size");
        return length;
    }
    isEqual(other) {
        if (this.length() !== other.length()) return false;
        return Buffer.compare(Buffer.from(this.eval.buffer, this.eval.byteOffset, this.eval.byteLength),
            Buffer.from(other.eval.buffer, other.eval.byteOffset, other.eval.byteLength)) === 0;
    }
    _allEqual(value) {
        // bytewise compare with a broadcast element on the device for large vectors (evaluations.js:118-129)
        const n = this.length();
        if (n < 4096) {
            for (let i = 0; i < n; i++) if (!this.Fr.eq(this.eval.subarray(32 * i, 32 * i + 32), value)) return false;
            return true;
        }
        const h = this.curve.upload(this.eval);
        try {
            return this.curve.addon.kzg_buf_all_equal(this.curve.ctx, h, value) !== 0;
        } finally {
            this.curve.free(h);
        }
    }
    isAllZeros() { return this._allEqual(this.Fr.zero); }
    isAllOnes() { return this._allEqual(this.Fr.one); }
    print(name = "f") {
        for (let i = 0; i < this.length(); i++) console.log(`${name}(w^${i}) =`, this.Fr.toString(this.getEvaluation(i)));
    }
}
module.exports = { Evaluations };
