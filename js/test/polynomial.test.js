// Known answers for the Polynomial / Evaluations drop-ins, small integers as in the reference's test/polynomial.test.js
// (evaluate :117-124, multiply :207-220, byXSubValue / divByXSubValue :255-262, Lagrange1 :360-371).
"use strict";
const assert = require("assert");
const kzg = require("../index.js");

describe("Polynomial (B200 backend)", function () {
    let curve, Fr;
    const P = (ints) => kzg.Polynomial.fromCoefficientsArray(ints.map((x) => Fr.e(x)), curve);
    const ints = (p, n) => Array.from({ length: n }, (_, i) => Fr.toObject(p.getCoef(i)));
    before(async () => {
        curve = await kzg.getCurveFromName("bn128");
        Fr = curve.Fr;
    });
    after(async () => { await curve.terminate(); });

    it("evaluate, degree, add, sub, scalars", async () => {
        const p = P([1, 2, 3, 0]);                                   // 1 + 2x + 3x^2
        assert.strictEqual(p.degree(), 2);
        assert.strictEqual(Fr.toObject(p.evaluate(Fr.e(2))), 17n);
        p.add(P([1, 1]));
        assert.deepStrictEqual(ints(p, 3), [2n, 3n, 3n]);
        p.sub(P([2]));
        p.mulScalar(Fr.e(2));
        p.addScalar(Fr.e(5));
        assert.deepStrictEqual(ints(p, 3), [5n, 6n, 6n]);
    });
    it("multiply, divByXSubValue, divZh", async () => {
        const p = P([1, 1]);                                          // (1 + x)
        await p.multiply(P([1, 1]));
        assert.deepStrictEqual(ints(p, 3), [1n, 2n, 1n]);
        const q = P([6, 5, 1]);                                       // (x + 2)(x + 3)
        q.divByXSubValue(Fr.e(-2));
        assert.deepStrictEqual(ints(q, 2), [3n, 1n]);
        assert.throws(() => P([1, 5, 1]).divByXSubValue(Fr.e(-2)), /Polynomial does not divide/);
        const z = P([-1, 0, 0, 0, 1, 0, 0, 0]);                       // x^4 - 1
        await z.multiply(P([7, 1]));
        z.divZh(4);
        assert.deepStrictEqual(ints(z, 2), [7n, 1n]);
    });
    it("fromEvaluations / Evaluations.fromPolynomial round trip, Lagrange1, shiftOmega", async () => {
        const evals = kzg.Evaluations.getRandomEvals(64, curve);
        const p = await kzg.Polynomial.fromEvaluations(evals.eval, curve);
        const back = await kzg.Evaluations.fromPolynomial(p, 1, curve);
        assert.ok(back.isEqual(evals));
        const l1 = await kzg.Polynomial.Lagrange1(6, curve);
        assert.strictEqual(Fr.toObject(l1.evaluate(Fr.one)), 1n);
        assert.strictEqual(Fr.toObject(l1.evaluate(Fr.w[6])), 0n);
        const x = Fr.e(12345);
        const want = p.evaluate(Fr.mul(x, Fr.w[6]));
        await p.shiftOmega();
        assert.ok(Fr.eq(p.evaluate(x), want));
    });
    it("commit: host PTau buffer and resident SRS give the same point", async function () {
        if (!process.env.PTAU) this.skip();
        const p = P([3, 1, 4, 1, 5, 9, 2, 6]);
        const srs = curve.loadSrs(process.env.PTAU, 16);
        const host = Uint8Array.from(curve.addon.kzg_srs_download(curve.ctx, srs.srs, 0, 16));
        const a = await p.multiExponentiation(host), b = await p.multiExponentiation(srs);
        assert.ok(curve.G1.eq(a, b));
    });
});
