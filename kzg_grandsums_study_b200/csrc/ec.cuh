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
    Fq y3 = fp_mul2_sub(r, fp_sub(q, x3), acc.y, ppp);  // R (Q - X3) - Y1 PPP under one reduction
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
    Fq y3 = fp_mul2_sub(r, fp_sub(q, x3), s1, ppp);
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

#if defined(__CUDACC__)
// ---- quad-lane group law ---------------------------------------------------------------------------------
// Four adjacent lanes (a "quad") hold ONE XYZZ point, lane j of the quad its coordinate j (0: X, 1: Y, 2: ZZ,
// 3: ZZZ), and cooperate on one addition or doubling: the 14 (9) field products are laid out in 4 (3) stages
// of independent products, one per lane, with the operands exchanged by shuffles inside the quad.  For a lone
// warp a field product costs ~860 cycles whatever the instruction-level parallelism (latency.cu), so the
// dependent chain of an addition drops from 14 to 4 products: ~3 x shorter latency for the tree sums and the
// double-and-add steps of the bucket-reduction tail, where nearly all lanes would otherwise idle.
// All lanes of the quad must call together (mask = the quad's four lanes); different quads may diverge.
__device__ __forceinline__ uint32_t quad_mask() {
    uint32_t lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    return 0xfu << (lane & 28u);
}
__device__ __forceinline__ Fq quad_get(uint32_t mask, const Fq& v, int src) {  // coordinate held by lane `src` of the quad
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(mask, v.l[i], src, 4);
    return r;
}
__device__ __forceinline__ Fq quad_xor(uint32_t mask, const Fq& v, int x) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(mask, v.l[i], x, 4);
    return r;
}
__device__ __forceinline__ Fq fq_select(bool c, const Fq& a, const Fq& b) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
__device__ __forceinline__ G1XYZZ quad_gather(uint32_t mask, const Fq& a) {
    G1XYZZ p;
    p.x = quad_get(mask, a, 0);
    p.y = quad_get(mask, a, 1);
    p.zz = quad_get(mask, a, 2);
    p.zzz = quad_get(mask, a, 3);
    return p;
}
__device__ __forceinline__ Fq quad_coord(const G1XYZZ& p, uint32_t j) {
    return j == 0 ? p.x : j == 1 ? p.y : j == 2 ? p.zz : p.zzz;
}
// the point at infinity in quad form: every coordinate zero
__device__ __forceinline__ Fq quad_load(const G1XYZZ* p, uint32_t j) { return fp_load<FqP>(reinterpret_cast<const Fq*>(p) + j); }
__device__ __forceinline__ void quad_store(G1XYZZ* p, uint32_t j, const Fq& a) { fp_store(reinterpret_cast<Fq*>(p) + j, a); }

// a (coordinate j of A) += b (coordinate j of B)
static __device__ __noinline__ void quad_add(Fq& a, const Fq& b, uint32_t j, uint32_t mask) {
    const bool a_inf = __shfl_sync(mask, (int)fp_is_zero(a), 2, 4) != 0;  // ZZ lives on lane 2
    const bool b_inf = __shfl_sync(mask, (int)fp_is_zero(b), 2, 4) != 0;
    if (b_inf) return;
    if (a_inf) {
        a = b;
        return;
    }
    // stage 1   lane 0: U1 = X1 ZZ2   lane 1: S1 = Y1 ZZZ2   lane 2: U2 = ZZ1 X2   lane 3: S2 = ZZZ1 Y2
    const Fq m1 = fp_mul(a, quad_xor(mask, b, 2));
    const Fq o1 = quad_xor(mask, m1, 2);
    const Fq d = j < 2 ? fp_sub(o1, m1) : fp_sub(m1, o1);  // lanes 0, 2: P = U2 - U1;   lanes 1, 3: R = S2 - S1
    const bool p_zero = __shfl_sync(mask, (int)fp_is_zero(d), 0, 4) != 0;
    if (p_zero) {  // same x: doubling or cancellation -- the scalar formula decides (rare)
        G1XYZZ pa = quad_gather(mask, a);
        const G1XYZZ pb = quad_gather(mask, b);
        xyzz_add(pa, pb);
        a = quad_coord(pa, j);
        return;
    }
    // stage 2   lane 0: PP = P^2   lane 1: RR = R^2   lane 2: ZZ1 ZZ2   lane 3: ZZZ1 ZZZ2
    const Fq m2 = fp_mul(j < 2 ? d : a, j < 2 ? d : b);
    // stage 3   lane 0: Q = U1 PP   (lane 1 idle)   lane 2: ZZ3 = ZZ1 ZZ2 PP   lane 3: PPP = P PP
    const Fq pp = quad_get(mask, m2, 0);
    const Fq p0 = quad_get(mask, d, 0);
    const Fq m3 = fp_mul(j == 0 ? m1 : j == 3 ? p0 : m2, pp);
    // X3 = RR - PPP - 2 Q on lane 0
    const Fq rr = quad_get(mask, m2, 1);
    const Fq ppp = quad_get(mask, m3, 3);
    const Fq x3 = fp_sub(fp_sub(rr, ppp), fp_dbl(m3));
    const Fq dq = quad_get(mask, fp_sub(m3, x3), 0);  // Q - X3
    const Fq s1 = quad_get(mask, m1, 1);
    // stage 4   (lane 0 idle)   lane 1: R (Q - X3)   lane 2: S1 PPP   lane 3: ZZZ3 = ZZZ1 ZZZ2 PPP
    const Fq m4 = fp_mul(j == 1 ? d : j == 2 ? s1 : m2, j == 1 ? dq : ppp);
    const Fq s1ppp = quad_get(mask, m4, 2);
    const Fq y3 = fp_sub(m4, s1ppp);
    a = j == 0 ? x3 : j == 1 ? y3 : j == 2 ? m3 : m4;
}

// a = 2 a
static __device__ __noinline__ void quad_dbl(Fq& a, uint32_t j, uint32_t mask) {
    const bool inf = __shfl_sync(mask, (int)fp_is_zero(a), 2, 4) != 0;
    if (inf) return;
    // stage 1   lane 0: XX = X^2   lane 1: V = U^2, U = 2 Y
    const Fq u = j == 1 ? fp_dbl(a) : a;
    const Fq m1 = fp_mul(u, u);
    const Fq v = quad_get(mask, m1, 1);
    const Fq m = quad_get(mask, fp_add(fp_dbl(m1), m1), 0);  // M = 3 XX
    // stage 2   lane 0: S = X V   lane 1: W = U V   lane 2: ZZ3 = ZZ V   lane 3: M^2
    const Fq m2 = fp_mul(j == 3 ? m : u, j == 3 ? m : v);
    const Fq mm = quad_get(mask, m2, 3);
    const Fq x3 = fp_sub(mm, fp_dbl(m2));  // lane 0
    const Fq w = quad_get(mask, m2, 1);
    // stage 3   lane 0: M (S - X3)   lane 1: W Y   (lane 2 idle)   lane 3: ZZZ3 = W ZZZ
    const Fq m3 = fp_mul(j == 0 ? m : w, j == 0 ? fp_sub(m2, x3) : a);
    const Fq y3a = quad_get(mask, m3, 0);
    const Fq y3 = fp_sub(y3a, m3);  // lane 1
    a = j == 0 ? x3 : j == 1 ? y3 : j == 2 ? m2 : m3;
}

// a = k a for a small non-negative integer k (double-and-add, MSB first; k uniform over the quad)
__device__ __forceinline__ void quad_mul_small(Fq& a, uint32_t k, uint32_t j, uint32_t mask) {
    const Fq p = a;
    a = fp_zero<FqP>();
    int top = 31;
    while (top >= 0 && !((k >> top) & 1)) top--;
#pragma unroll 1
    for (int bit = top; bit >= 0; bit--) {
        quad_dbl(a, j, mask);
        if ((k >> bit) & 1) quad_add(a, p, j, mask);
    }
}
#endif  // __CUDACC__

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
