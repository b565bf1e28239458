// Drop-in for reference src/Keccak256Transcript.js:7-52 (stays on the HOST by mandate).  Same data model: every
// getChallenge() hashes ALL the data appended so far -- commitments as 64-byte big-endian standard-form affine points,
// scalars as 32-byte big-endian standard form -- with Keccak-256 (original padding, js-sha3's keccak256) and reduces the
// digest mod r.  Hash and byte conversions are the host helpers of libkzgb200.so, so no js-sha3 / ffjavascript is needed.
"use strict";
const POLYNOMIAL = 0;
const SCALAR = 1;

class Keccak256Transcript {
    constructor(curve) {
        this.curve = curve;
        this.G1 = curve.G1;
        this.Fr = curve.Fr;
        this.reset();
    }
    reset() { this.data = []; }
    addPolCommitment(polynomialCommitment) { this.data.push({ type: POLYNOMIAL, data: polynomialCommitment }); }
    addFieldElement(scalar) { this.data.push({ type: SCALAR, data: scalar }); }
    getChallenge() {
        if (0 === this.data.length) throw new Error("Keccak256Transcript: No data to generate a transcript");
        let nPolynomials = 0;
        let nScalars = 0;
        this.data.forEach((el) => (POLYNOMIAL === el.type ? nPolynomials++ : nScalars++));
        const buffer = new Uint8Array(nScalars * this.Fr.n8 + nPolynomials * this.G1.F.n8 * 2);
        let offset = 0;
        for (const el of this.data) {
            if (POLYNOMIAL === el.type) {
                this.G1.toRprUncompressed(buffer, offset, el.data);
                offset += this.G1.F.n8 * 2;
            } else {
                this.Fr.toRprBE(buffer, offset, el.data);
                offset += this.Fr.n8;
            }
        }
        const addon = this.curve.addon;
        const digest = addon.kzg_keccak256(buffer, buffer.byteLength);
        return Uint8Array.from(addon.kzg_fr_from_hash_be(digest));   // Fr.e(Scalar.fromRprBE(hash))
    }
}
module.exports = { Keccak256Transcript };
