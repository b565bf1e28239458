// BN254 G1 (y^2 = x^3 + 3 over Fq) group law for the MSM: affine inputs, extended-Jacobian "XYZZ"
// accumulators (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2).  Mixed add = 8M + 2S, the cheapest accumulate
// step available without inversions -- this is the per-(point, window) unit of work the roofline in
// DESIGN.md counts (10 modmul = 1360 wide MACs).
//
// Encodings match the reference's ffjavascript buffers: affine = x || y, 32-byte
// Montgomery-LE each, the point at infinity is 64 zero bytes.
#pragma once
#include "field.cuh"

namespace kzg {

struct alignas(16) G1Affine {
    Fq x, y;
};
struct alignas(16) G1XYZZ {
    Fq x, y, zz, zzz;
};

KZG_HD bool g1_affine_is_inf(const G1Affine& p) {
    return fp_is_zero(p.x) && fp_is_zero(p.y);
}
KZG_HD G1XYZZ xyzz_inf() {
    G1XYZZ r;
    r.x = fp_zero<FqP>();
    r.y = fp_zero<FqP>();
    r.zz = fp_zero<FqP>();
    r.zzz = fp_zero<FqP>();
    return r;
}
KZG_HD bool xyzz_is_inf(const G1XYZZ& p) {
    return fp_is_zero(p.zz);
}
KZG_HD G1XYZZ xyzz_from_affine(const G1Affine& p) {
    G1XYZZ r;
    if (g1_affine_is_inf(p)) return xyzz_inf();
    r.x = p.x;
    r.y = p.y;
    r.zz = fp_one<FqP>();
    r.zzz = fp_one<FqP>();
    return r;
}
KZG_HD G1Affine g1_affine_neg(const G1Affine& p) {
    G1Affine r;
    r.x = p.x;
    r.y = fp_neg(p.y);  // -0 = 0, so infinity stays infinity
    return r;
}

// 2 * (affine p), p != infinity  (mdbl-2008-s-1, a = 0)
KZG_HD G1XYZZ xyzz_dbl_affine(const G1Affine& p) {
    G1XYZZ r;
    Fq u = fp_dbl(p.y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(p.x, v);
    Fq xx = fp_sqr(p.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * p  (dbl-2008-s-1, a = 0)
KZG_HD G1XYZZ xyzz_dbl(const G1XYZZ& p) {
    if (xyzz_is_inf(p)) return p;
    G1XYZZ r;
    Fq u = fp_dbl(p.y);
    Fq v = fp_sqr(u);
    Fq w = fp_mul(u, v);
    Fq s = fp_mul(p.x, v);
    Fq xx = fp_sqr(p.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    r.x = fp_sub(fp_sqr(m), fp_dbl(s));
    r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
    r.zz = fp_mul(v, p.zz);
    r.zzz = fp_mul(w, p.zzz);
    return r;
}

// acc += affine p   (madd-2008-s); handles acc = inf, p = inf, p = +-acc
KZG_HD void xyzz_madd(G1XYZZ& acc, const G1Affine& p) {
    if (g1_affine_is_inf(p)) return;
    if (xyzz_is_inf(acc)) {
        acc.x = p.x;
        acc.y = p.y;
        acc.zz = fp_one<FqP>();
        acc.zzz = fp_one<FqP>();
        return;
    }
    Fq u2 = fp_mul(p.x, acc.zz);
    Fq s2 = fp_mul(p.y, acc.zzz);
    Fq pp_ = fp_sub(u2, acc.x);
    Fq r = fp_sub(s2, acc.y);
    if (fp_is_zero(pp_)) {
        if (fp_is_zero(r)) {
            acc = xyzz_dbl_affine(p);
        } else {
            acc = xyzz_inf();
        }
        return;
    }
    Fq pp = fp_sqr(pp_);
    Fq ppp = fp_mul(pp_, pp);
    Fq q = fp_mul(acc.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(acc.zz, pp);
    acc.zzz = fp_mul(acc.zzz, ppp);
}

// acc += b   (add-2008-s); handles infinities and b = +-acc
KZG_HD void xyzz_add(G1XYZZ& acc, const G1XYZZ& b) {
    if (xyzz_is_inf(b)) return;
    if (xyzz_is_inf(acc)) {
        acc = b;
        return;
    }
    Fq u1 = fp_mul(acc.x, b.zz);
    Fq u2 = fp_mul(b.x, acc.zz);
    Fq s1 = fp_mul(acc.y, b.zzz);
    Fq s2 = fp_mul(b.y, acc.zzz);
    Fq pp_ = fp_sub(u2, u1);
    Fq r = fp_sub(s2, s1);
    if (fp_is_zero(pp_)) {
        if (fp_is_zero(r)) {
            acc = xyzz_dbl(acc);
        } else {
            acc = xyzz_inf();
        }
        return;
    }
    Fq pp = fp_sqr(pp_);
    Fq ppp = fp_mul(pp_, pp);
    Fq q = fp_mul(u1, pp);
    Fq x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(q));
    Fq y3 = fp_sub(fp_mul(r, fp_sub(q, x3)), fp_mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(fp_mul(acc.zz, b.zz), pp);
    acc.zzz = fp_mul(fp_mul(acc.zzz, b.zzz), ppp);
}

// k * p for a small non-negative integer k (double-and-add, MSB first)
KZG_HD G1XYZZ xyzz_mul_small(const G1XYZZ& p, uint32_t k) {
    G1XYZZ r = xyzz_inf();
    int top = 31;
    while (top >= 0 && !((k >> top) & 1)) top--;  // skip the leading zeros: no work on the point at infinity
    for (int bit = top; bit >= 0; bit--) {
        r = xyzz_dbl(r);
        if ((k >> bit) & 1) xyzz_add(r, p);
    }
    return r;
}

// canonical affine form (one inversion); infinity -> 64 zero bytes
KZG_HD G1Affine xyzz_to_affine(const G1XYZZ& p) {
    G1Affine r;
    if (xyzz_is_inf(p)) {
        r.x = fp_zero<FqP>();
        r.y = fp_zero<FqP>();
        return r;
    }
    Fq inv = fp_inv(fp_mul(p.zz, p.zzz));
    r.x = fp_mul(p.x, fp_mul(inv, p.zzz));  // X / ZZ
    r.y = fp_mul(p.y, fp_mul(inv, p.zz));   // Y / ZZZ
    return r;
}

KZG_HD bool g1_affine_on_curve(const G1Affine& p) {
    if (g1_affine_is_inf(p)) return true;
    Fq three = fp_one<FqP>();
    three = fp_add(fp_dbl(three), three);
    Fq lhs = fp_sqr(p.y);
    Fq rhs = fp_add(fp_mul(fp_sqr(p.x), p.x), three);
    return fp_eq(lhs, rhs);
}

}  // namespace kzg
