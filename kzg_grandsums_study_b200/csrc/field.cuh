// BN254 Fq / Fr arithmetic: 8 x 32-bit limbs, Montgomery form (R = 2^256), fully reduced outputs.
//
// In-memory form is exactly the reference's: 32-byte little-endian Montgomery residue
// (ffjavascript Fr/Fq elements), so device buffers are byte-compatible with the
// `.coef` / `.eval` / ptau section-2 bytes and no transposition is needed at the boundary.
//
// Two multiplication paths:
//   * fp_mul_portable : operand-scanning CIOS with 64-bit temporaries; host + device; the
//     specification, unit-tested on the CPU against Python big ints.
//   * fp_mul (device) : even/odd split accumulators with mad.lo.cc / madc.hi.cc carry chains
//     (each lo/hi pair is one IMAD.WIDE.U32 with carry in SASS on sm_100a) -- 128 wide MACs
//     for the product + reduction, 8 IMAD for the quotient digits.  Checked against the
//     portable path on the device by kzg_selftest().
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KZG_HD __host__ __device__ __forceinline__
#define KZG_D __device__ __forceinline__
#else
#define KZG_HD inline
#define KZG_D inline
#endif

#ifndef KZG_FAST_MUL
#define KZG_FAST_MUL 1
#endif

namespace kzg {

struct FqP {
    static constexpr uint32_t INV = 0xe4866389u;  // -q^-1 mod 2^32
    KZG_HD static constexpr uint32_t mod(int i) {
        return i == 0 ? 0xd87cfd47u : i == 1 ? 0x3c208c16u : i == 2 ? 0x6871ca8du : i == 3 ? 0x97816a91u
             : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    KZG_HD static constexpr uint32_t r1(int i) {  // R mod q  (Montgomery one)
        return i == 0 ? 0xc58f0d9du : i == 1 ? 0xd35d438du : i == 2 ? 0xf5c70b3du : i == 3 ? 0x0a78eb28u
             : i == 4 ? 0x7879462cu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    KZG_HD static constexpr uint32_t r2(int i) {  // R^2 mod q
        return i == 0 ? 0x538afa89u : i == 1 ? 0xf32cfc5bu : i == 2 ? 0xd44501fbu : i == 3 ? 0xb5e71911u
             : i == 4 ? 0x0a417ff6u : i == 5 ? 0x47ab1effu : i == 6 ? 0xcab8351fu : 0x06d89f71u;
    }
};

struct FrP {
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
    KZG_HD static constexpr uint32_t mod(int i) {
        return i == 0 ? 0xf0000001u : i == 1 ? 0x43e1f593u : i == 2 ? 0x79b97091u : i == 3 ? 0x2833e848u
             : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    KZG_HD static constexpr uint32_t r1(int i) {
        return i == 0 ? 0x4ffffffbu : i == 1 ? 0xac96341cu : i == 2 ? 0x9f60cd29u : i == 3 ? 0x36fc7695u
             : i == 4 ? 0x7879462eu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    KZG_HD static constexpr uint32_t r2(int i) {
        return i == 0 ? 0xae216da7u : i == 1 ? 0x1bb8e645u : i == 2 ? 0xe35c59e3u : i == 3 ? 0x53fe3ab1u
             : i == 4 ? 0x53bb8085u : i == 5 ? 0x8c49833du : i == 6 ? 0x7f4e44a5u : 0x0216d0b1u;
    }
};

template <class P>
struct alignas(16) Fp {
    uint32_t l[8];
};
using Fq = Fp<FqP>;
using Fr = Fp<FrP>;

// ------------------------------------------------------------------------------------------------
// constants, predicates, load/store
// ------------------------------------------------------------------------------------------------
template <class P> KZG_HD Fp<P> fp_zero() {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
}
template <class P> KZG_HD Fp<P> fp_one() {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = P::r1(i);
    return r;
}
template <class P> KZG_HD Fp<P> fp_r2() {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = P::r2(i);
    return r;
}
template <class P> KZG_HD bool fp_is_zero(const Fp<P>& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.l[i];
    return o == 0;
}
template <class P> KZG_HD bool fp_eq(const Fp<P>& a, const Fp<P>& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i];
    return o == 0;
}
// a >= p ?
template <class P> KZG_HD bool fp_geq_mod(const uint32_t* a) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (a[i] > P::mod(i)) return true;
        if (a[i] < P::mod(i)) return false;
    }
    return true;
}

template <class P> KZG_HD Fp<P> fp_load(const void* p) {
    Fp<P> r;
#if defined(__CUDA_ARCH__)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
#else
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    for (int i = 0; i < 8; i++) r.l[i] = q[i];
#endif
    return r;
}
template <class P> KZG_HD void fp_store(void* p, const Fp<P>& r) {
#if defined(__CUDA_ARCH__)
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.l[0], r.l[1], r.l[2], r.l[3]);
    q[1] = make_uint4(r.l[4], r.l[5], r.l[6], r.l[7]);
#else
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
    for (int i = 0; i < 8; i++) q[i] = r.l[i];
#endif
}

// ------------------------------------------------------------------------------------------------
// add / sub / neg / dbl  (inputs < p, outputs < p)
// ------------------------------------------------------------------------------------------------
template <class P> KZG_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> s, d;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;\n\t"
        : "=r"(s.l[0]), "=r"(s.l[1]), "=r"(s.l[2]), "=r"(s.l[3]), "=r"(s.l[4]), "=r"(s.l[5]), "=r"(s.l[6]), "=r"(s.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    // p < 2^254 so a + b < 2^255: no carry out of limb 7.  d = s - p, keep d if no borrow.
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]), "=r"(d.l[6]), "=r"(d.l[7]),
          "=r"(borrow)
        : "r"(s.l[0]), "r"(s.l[1]), "r"(s.l[2]), "r"(s.l[3]), "r"(s.l[4]), "r"(s.l[5]), "r"(s.l[6]), "r"(s.l[7]),
          "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)), "r"(P::mod(6)), "r"(P::mod(7)));
#pragma unroll
    for (int i = 0; i < 8; i++) d.l[i] = borrow ? s.l[i] : d.l[i];
    return d;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.l[i] + b.l[i];
        s.l[i] = (uint32_t)c;
        c >>= 32;
    }
    int64_t bw = 0;
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)s.l[i] - (int64_t)P::mod(i);
        d.l[i] = (uint32_t)bw;
        bw >>= 32;
    }
    return bw ? s : d;
#endif
}

template <class P> KZG_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> d;
#if defined(__CUDA_ARCH__)
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]), "=r"(d.l[6]), "=r"(d.l[7]),
          "=r"(borrow)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    // borrow is 0 or 0xffffffff: add (p & borrow)
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;\n\t"
        : "+r"(d.l[0]), "+r"(d.l[1]), "+r"(d.l[2]), "+r"(d.l[3]), "+r"(d.l[4]), "+r"(d.l[5]), "+r"(d.l[6]), "+r"(d.l[7])
        : "r"(P::mod(0) & borrow), "r"(P::mod(1) & borrow), "r"(P::mod(2) & borrow), "r"(P::mod(3) & borrow),
          "r"(P::mod(4) & borrow), "r"(P::mod(5) & borrow), "r"(P::mod(6) & borrow), "r"(P::mod(7) & borrow));
    return d;
#else
    int64_t bw = 0;
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)a.l[i] - (int64_t)b.l[i];
        d.l[i] = (uint32_t)bw;
        bw >>= 32;
    }
    if (bw) {
        uint64_t c = 0;
        for (int i = 0; i < 8; i++) {
            c += (uint64_t)d.l[i] + P::mod(i);
            d.l[i] = (uint32_t)c;
            c >>= 32;
        }
    }
    return d;
#endif
}

template <class P> KZG_HD Fp<P> fp_neg(const Fp<P>& a) {
    return fp_sub(fp_zero<P>(), a);
}
template <class P> KZG_HD Fp<P> fp_dbl(const Fp<P>& a) {
    return fp_add(a, a);
}

// ------------------------------------------------------------------------------------------------
// Montgomery multiplication
// ------------------------------------------------------------------------------------------------
// final conditional subtraction: t in [0, 2p) -> [0, p)
template <class P> KZG_HD Fp<P> fp_final_sub(const uint32_t* t) {
    Fp<P> s, d;
#pragma unroll
    for (int i = 0; i < 8; i++) s.l[i] = t[i];
#if defined(__CUDA_ARCH__)
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]), "=r"(d.l[6]), "=r"(d.l[7]),
          "=r"(borrow)
        : "r"(s.l[0]), "r"(s.l[1]), "r"(s.l[2]), "r"(s.l[3]), "r"(s.l[4]), "r"(s.l[5]), "r"(s.l[6]), "r"(s.l[7]),
          "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)), "r"(P::mod(6)), "r"(P::mod(7)));
#pragma unroll
    for (int i = 0; i < 8; i++) d.l[i] = borrow ? s.l[i] : d.l[i];
    return d;
#else
    int64_t bw = 0;
    for (int i = 0; i < 8; i++) {
        bw += (int64_t)s.l[i] - (int64_t)P::mod(i);
        d.l[i] = (uint32_t)bw;
        bw >>= 32;
    }
    return bw ? s : d;
#endif
}

template <class P> KZG_HD Fp<P> fp_mul_portable(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.l[j] * b.l[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * P::INV;
        c = (uint64_t)m * P::mod(0) + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * P::mod(j) + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    return fp_final_sub<P>(t);
}

#if defined(__CUDA_ARCH__)
// ---- even/odd carry-chain building blocks (device only) ----------------------------------------
// acc[0..7] += {x0, x2, x4, x6} * y over limb positions 0..7 (one carry chain); the carry out of the
// chain is added to `top` (the word above, which lives in the *other* accumulator).
KZG_D void cmad4_top(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t y, uint32_t& top) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
          "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
}
// same, carry out dropped (provably zero by the 2p bound)
KZG_D void cmad4(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;\n\t"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
}
// lo[0] += carry_word (a word sitting at the same limb position in the other accumulator), then
// hi[j] = {x1,x3,x5,x7} * y + hi[j+2] (j = 0,2,4), hi[6,7] = x7*y + carry : the one-limb right
// shift of the accumulator pair after a reduction step, fused with the next row of products.
KZG_D void shift_mad4(uint32_t& lo0, uint32_t carry_word, uint32_t* hi, uint32_t x1, uint32_t x3, uint32_t x5, uint32_t x7, uint32_t y) {
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "madc.lo.cc.u32 %1, %10, %14, %3;\n\t"
        "madc.hi.cc.u32 %2, %10, %14, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %14, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %14, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %14, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %14, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %14, 0;\n\t"
        "madc.hi.u32 %8, %13, %14, 0;\n\t"
        : "+r"(lo0), "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7])
        : "r"(carry_word), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
}
#endif

// Even/odd Montgomery product.  `lo` holds limb positions 0..7, `hi` holds positions 1..8 of the
// running total; after each reduction step the total is divisible by 2^32, the pair is shifted one
// limb (roles swap: hi becomes the new lo) and the old lo, shifted by two limbs, is re-used as the
// new hi while the next row of odd products is accumulated into it.
template <class P> KZG_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && KZG_FAST_MUL
    uint32_t x[8], y[8];  // x: positions 0..7 ("lo"), y: positions 1..8 ("hi")
    // row 0: plain products
    asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;\n\t"
        : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
        : "r"(a.l[0]), "r"(a.l[2]), "r"(a.l[4]), "r"(a.l[6]), "r"(b.l[0]));
    asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;\n\t"
        : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7])
        : "r"(a.l[1]), "r"(a.l[3]), "r"(a.l[5]), "r"(a.l[7]), "r"(b.l[0]));
    {
        uint32_t m = x[0] * P::INV;
        cmad4(y, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        cmad4_top(x, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, y[7]);
    }
#pragma unroll
    for (int i = 1; i < 8; i++) {
        // roles: on odd i the pair is (lo = y, hi = x); on even i it is (lo = x, hi = y)
        uint32_t* lo = (i & 1) ? y : x;
        uint32_t* hi = (i & 1) ? x : y;
        // old-lo word 1 sits at the new position 0; old-lo words 2..7 become the new hi words 0..5
        shift_mad4(lo[0], hi[1], hi, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
        cmad4_top(lo, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i], hi[7]);
        uint32_t m = lo[0] * P::INV;
        cmad4(hi, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        cmad4_top(lo, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, hi[7]);
    }
    // after row 7 (odd): lo = y, hi = x.  result word j = hi[j] + lo[j+1] (lo[0] == 0)
    uint32_t t[8];
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;\n\t"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
    return fp_final_sub<P>(t);
#else
    return fp_mul_portable(a, b);
#endif
}

// a b + c d with ONE interleaved reduction: the rows of fp_mul with a second product row before each reduction row --
// 192 wide MACs instead of 256.  Every intermediate total stays below 3p + 3p 2^32 < 2^288 (the nine words of the
// accumulator pair) and the result below (2 p^2 + p 2^256) / 2^256 < 1.38 p: one conditional subtraction.  The helper
// sequence is emulated word by word, with every dropped carry asserted zero, in tools/gen_fp_sqr.py (check2).
template <class P> KZG_HD Fp<P> fp_mul2(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
#if defined(__CUDA_ARCH__) && KZG_FAST_MUL
    uint32_t x[8], y[8];
    asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;\n\t"
        : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
        : "r"(a.l[0]), "r"(a.l[2]), "r"(a.l[4]), "r"(a.l[6]), "r"(b.l[0]));
    asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;\n\t"
        : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]), "=r"(y[4]), "=r"(y[5]), "=r"(y[6]), "=r"(y[7])
        : "r"(a.l[1]), "r"(a.l[3]), "r"(a.l[5]), "r"(a.l[7]), "r"(b.l[0]));
    cmad4(y, c.l[1], c.l[3], c.l[5], c.l[7], d.l[0]);
    cmad4_top(x, c.l[0], c.l[2], c.l[4], c.l[6], d.l[0], y[7]);
    {
        uint32_t m = x[0] * P::INV;
        cmad4(y, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        cmad4_top(x, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, y[7]);
    }
#pragma unroll
    for (int i = 1; i < 8; i++) {
        uint32_t* lo = (i & 1) ? y : x;
        uint32_t* hi = (i & 1) ? x : y;
        shift_mad4(lo[0], hi[1], hi, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
        cmad4_top(lo, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i], hi[7]);
        cmad4(hi, c.l[1], c.l[3], c.l[5], c.l[7], d.l[i]);
        cmad4_top(lo, c.l[0], c.l[2], c.l[4], c.l[6], d.l[i], hi[7]);
        uint32_t m = lo[0] * P::INV;
        cmad4(hi, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        cmad4_top(lo, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m, hi[7]);
    }
    uint32_t t[8];
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, 0;\n\t"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
    return fp_final_sub<P>(t);
#else
    return fp_add(fp_mul(a, b), fp_mul(c, d));
#endif
}
// a b - c d
template <class P> KZG_HD Fp<P> fp_mul2_sub(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    return fp_mul2(a, b, fp_neg(c), d);
}

// Dedicated squaring: 100 wide MACs instead of 128 (28 off-diagonal products doubled + 8 squares, then a Montgomery
// reduction of the 512-bit result with the same even/odd rows as fp_mul).  The instruction list is produced AND emulated
// word by word on the CPU by tools/gen_fp_sqr.py (carry bookkeeping checked there against a^2 / R mod p); kzg_selftest()
// compares the compiled code with fp_mul_portable on the device.  Same canonical result as fp_mul(a, a).
#ifndef KZG_FAST_SQR
#define KZG_FAST_SQR 1
#endif
template <class P> KZG_HD Fp<P> fp_sqr(const Fp<P>& a) {
#if defined(__CUDA_ARCH__) && KZG_FAST_MUL && KZG_FAST_SQR
    // ---- GENERATED by tools/gen_fp_sqr.py (do not edit by hand) ----
    const uint32_t a0 = a.l[0], a1 = a.l[1], a2 = a.l[2], a3 = a.l[3], a4 = a.l[4], a5 = a.l[5], a6 = a.l[6], a7 = a.l[7];
    uint32_t e2, e3, e4, e5, e6, e7, o0, o1, o2, o3, o4, o5, o6, o7, e8, e9, o8, e10, o9, e11, o10, e12, o11, e13,
        o12, o13, d15, d14, d13, d12, d11, d10, d9, d8, d7, d6, d5, d4, d3, d2, d1, T0, T1, T2, T3, T4, T5, T6, T7,
        T8, T9, T10, T11, T12, T13, T14, T15, x0, x1, x2, x3, x4, x5, x6, x7, m, y0, y1, y2, y3, y4, y5, y6, y7, w,
        t0, t1, t2, t3, t4, t5, t6, t7;
    asm("mul.lo.u32 %0, %6, %7;\n\t"
        "mul.hi.u32 %1, %6, %7;\n\t"
        "mul.lo.u32 %2, %8, %7;\n\t"
        "mul.hi.u32 %3, %8, %7;\n\t"
        "mul.lo.u32 %4, %9, %7;\n\t"
        "mul.hi.u32 %5, %9, %7;"
        : "=r"(e2), "=r"(e3), "=r"(e4), "=r"(e5), "=r"(e6), "=r"(e7)
        : "r"(a2), "r"(a0), "r"(a4), "r"(a6));
    asm("mul.lo.u32 %0, %8, %9;\n\t"
        "mul.hi.u32 %1, %8, %9;\n\t"
        "mul.lo.u32 %2, %10, %9;\n\t"
        "mul.hi.u32 %3, %10, %9;\n\t"
        "mul.lo.u32 %4, %11, %9;\n\t"
        "mul.hi.u32 %5, %11, %9;\n\t"
        "mul.lo.u32 %6, %12, %9;\n\t"
        "mul.hi.u32 %7, %12, %9;"
        : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3), "=r"(o4), "=r"(o5), "=r"(o6), "=r"(o7)
        : "r"(a1), "r"(a0), "r"(a3), "r"(a5), "r"(a7));
    asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %7, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %7, 0;\n\t"
        "madc.hi.u32 %5, %9, %7, 0;"
        : "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "=r"(e8), "=r"(e9)
        : "r"(a3), "r"(a1), "r"(a5), "r"(a7));
    asm("mad.lo.cc.u32 %0, %7, %8, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %8, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %8, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %8, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %8, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %8, %5;\n\t"
        "addc.u32 %6, 0, 0;"
        : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7), "=r"(o8)
        : "r"(a2), "r"(a1), "r"(a4), "r"(a6));
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %6, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, %6, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(e6), "+r"(e7), "+r"(e8), "+r"(e9), "=r"(e10)
        : "r"(a4), "r"(a2), "r"(a6));
    asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %7, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %7, %4;\n\t"
        "madc.hi.u32 %5, %9, %7, 0;"
        : "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7), "+r"(o8), "=r"(o9)
        : "r"(a3), "r"(a2), "r"(a5), "r"(a7));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %5, %2;\n\t"
        "madc.hi.u32 %3, %6, %5, 0;"
        : "+r"(e8), "+r"(e9), "+r"(e10), "=r"(e11)
        : "r"(a5), "r"(a3), "r"(a7));
    asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, %6, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, %6, %3;\n\t"
        "addc.u32 %4, 0, 0;"
        : "+r"(o6), "+r"(o7), "+r"(o8), "+r"(o9), "=r"(o10)
        : "r"(a4), "r"(a3), "r"(a6));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, 0, 0;"
        : "+r"(e10), "+r"(e11), "=r"(e12)
        : "r"(a6), "r"(a4));
    asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %5, %2;\n\t"
        "madc.hi.u32 %3, %6, %5, 0;"
        : "+r"(o8), "+r"(o9), "+r"(o10), "=r"(o11)
        : "r"(a5), "r"(a4), "r"(a7));
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
        "madc.hi.u32 %1, %2, %3, 0;"
        : "+r"(e12), "=r"(e13)
        : "r"(a7), "r"(a5));
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, 0, 0;"
        : "+r"(o10), "+r"(o11), "=r"(o12)
        : "r"(a6), "r"(a5));
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
        "madc.hi.u32 %1, %2, %3, 0;"
        : "+r"(o12), "=r"(o13)
        : "r"(a7), "r"(a6));
    asm("add.cc.u32 %0, %0, %13;\n\t"
        "addc.cc.u32 %1, %1, %14;\n\t"
        "addc.cc.u32 %2, %2, %15;\n\t"
        "addc.cc.u32 %3, %3, %16;\n\t"
        "addc.cc.u32 %4, %4, %17;\n\t"
        "addc.cc.u32 %5, %5, %18;\n\t"
        "addc.cc.u32 %6, %6, %19;\n\t"
        "addc.cc.u32 %7, %7, %20;\n\t"
        "addc.cc.u32 %8, %8, %21;\n\t"
        "addc.cc.u32 %9, %9, %22;\n\t"
        "addc.cc.u32 %10, %10, %23;\n\t"
        "addc.cc.u32 %11, %11, %24;\n\t"
        "addc.u32 %12, %12, 0;"
        : "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "+r"(e8), "+r"(e9), "+r"(e10), "+r"(e11), "+r"(e12), "+r"(e13), "+r"(o13)
        : "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6), "r"(o7), "r"(o8), "r"(o9), "r"(o10), "r"(o11), "r"(o12));
    asm("shr.u32 %0, %15, 31;\n\t"
        "shf.l.wrap.b32 %1, %16, %15, 1;\n\t"
        "shf.l.wrap.b32 %2, %17, %16, 1;\n\t"
        "shf.l.wrap.b32 %3, %18, %17, 1;\n\t"
        "shf.l.wrap.b32 %4, %19, %18, 1;\n\t"
        "shf.l.wrap.b32 %5, %20, %19, 1;\n\t"
        "shf.l.wrap.b32 %6, %21, %20, 1;\n\t"
        "shf.l.wrap.b32 %7, %22, %21, 1;\n\t"
        "shf.l.wrap.b32 %8, %23, %22, 1;\n\t"
        "shf.l.wrap.b32 %9, %24, %23, 1;\n\t"
        "shf.l.wrap.b32 %10, %25, %24, 1;\n\t"
        "shf.l.wrap.b32 %11, %26, %25, 1;\n\t"
        "shf.l.wrap.b32 %12, %27, %26, 1;\n\t"
        "shf.l.wrap.b32 %13, %28, %27, 1;\n\t"
        "shl.b32 %14, %28, 1;"
        : "=r"(d15), "=r"(d14), "=r"(d13), "=r"(d12), "=r"(d11), "=r"(d10), "=r"(d9), "=r"(d8), "=r"(d7), "=r"(d6), "=r"(d5), "=r"(d4), "=r"(d3), "=r"(d2), "=r"(d1)
        : "r"(o13), "r"(e13), "r"(e12), "r"(e11), "r"(e10), "r"(e9), "r"(e8), "r"(e7), "r"(e6), "r"(e5), "r"(e4), "r"(e3), "r"(e2), "r"(o0));
    asm("mul.lo.u32 %0, %16, %16;\n\t"
        "mad.hi.cc.u32 %1, %16, %16, %17;\n\t"
        "madc.lo.cc.u32 %2, %18, %18, %19;\n\t"
        "madc.hi.cc.u32 %3, %18, %18, %20;\n\t"
        "madc.lo.cc.u32 %4, %21, %21, %22;\n\t"
        "madc.hi.cc.u32 %5, %21, %21, %23;\n\t"
        "madc.lo.cc.u32 %6, %24, %24, %25;\n\t"
        "madc.hi.cc.u32 %7, %24, %24, %26;\n\t"
        "madc.lo.cc.u32 %8, %27, %27, %28;\n\t"
        "madc.hi.cc.u32 %9, %27, %27, %29;\n\t"
        "madc.lo.cc.u32 %10, %30, %30, %31;\n\t"
        "madc.hi.cc.u32 %11, %30, %30, %32;\n\t"
        "madc.lo.cc.u32 %12, %33, %33, %34;\n\t"
        "madc.hi.cc.u32 %13, %33, %33, %35;\n\t"
        "madc.lo.cc.u32 %14, %36, %36, %37;\n\t"
        "madc.hi.u32 %15, %36, %36, %38;"
        : "=r"(T0), "=r"(T1), "=r"(T2), "=r"(T3), "=r"(T4), "=r"(T5), "=r"(T6), "=r"(T7), "=r"(T8), "=r"(T9), "=r"(T10), "=r"(T11), "=r"(T12), "=r"(T13), "=r"(T14), "=r"(T15)
        : "r"(a0), "r"(d1), "r"(a1), "r"(d2), "r"(d3), "r"(a2), "r"(d4), "r"(d5), "r"(a3), "r"(d6), "r"(d7), "r"(a4), "r"(d8), "r"(d9), "r"(a5), "r"(d10), "r"(d11), "r"(a6), "r"(d12), "r"(d13), "r"(a7), "r"(d14), "r"(d15));
    asm("mov.b32 %0, %9;\n\t"
        "mov.b32 %1, %10;\n\t"
        "mov.b32 %2, %11;\n\t"
        "mov.b32 %3, %12;\n\t"
        "mov.b32 %4, %13;\n\t"
        "mov.b32 %5, %14;\n\t"
        "mov.b32 %6, %15;\n\t"
        "mov.b32 %7, %16;\n\t"
        "mul.lo.u32 %8, %0, %17;"
        : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3), "=r"(x4), "=r"(x5), "=r"(x6), "=r"(x7), "=r"(m)
        : "r"(T0), "r"(T1), "r"(T2), "r"(T3), "r"(T4), "r"(T5), "r"(T6), "r"(T7), "r"(P::INV));
    asm("mul.lo.u32 %0, %8, %9;\n\t"
        "mul.hi.u32 %1, %8, %9;\n\t"
        "mul.lo.u32 %2, %10, %9;\n\t"
        "mul.hi.u32 %3, %10, %9;\n\t"
        "mul.lo.u32 %4, %11, %9;\n\t"
        "mul.hi.u32 %5, %11, %9;\n\t"
        "mul.lo.u32 %6, %12, %9;\n\t"
        "mul.hi.u32 %7, %12, %9;"
        : "=r"(y0), "=r"(y1), "=r"(y2), "=r"(y3), "=r"(y4), "=r"(y5), "=r"(y6), "=r"(y7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(y7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.u32 %0, %2, %3;\n\t"
        "mul.lo.u32 %1, %0, %4;"
        : "=r"(w), "=r"(m)
        : "r"(y0), "r"(x1), "r"(P::INV));
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %10, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %10, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %10, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %10, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %10, 0;\n\t"
        "madc.hi.u32 %8, %13, %10, 0;"
        : "+r"(y0), "=r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7), "+r"(x7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.u32 %0, %2, %3;\n\t"
        "mul.lo.u32 %1, %0, %4;"
        : "=r"(w), "=r"(m)
        : "r"(x0), "r"(y1), "r"(P::INV));
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %10, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %10, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %10, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %10, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %10, 0;\n\t"
        "madc.hi.u32 %8, %13, %10, 0;"
        : "+r"(x0), "=r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(y7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.u32 %0, %2, %3;\n\t"
        "mul.lo.u32 %1, %0, %4;"
        : "=r"(w), "=r"(m)
        : "r"(y0), "r"(x1), "r"(P::INV));
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %10, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %10, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %10, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %10, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %10, 0;\n\t"
        "madc.hi.u32 %8, %13, %10, 0;"
        : "+r"(y0), "=r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7), "+r"(x7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.u32 %0, %2, %3;\n\t"
        "mul.lo.u32 %1, %0, %4;"
        : "=r"(w), "=r"(m)
        : "r"(x0), "r"(y1), "r"(P::INV));
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %10, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %10, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %10, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %10, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %10, 0;\n\t"
        "madc.hi.u32 %8, %13, %10, 0;"
        : "+r"(x0), "=r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(y7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.u32 %0, %2, %3;\n\t"
        "mul.lo.u32 %1, %0, %4;"
        : "=r"(w), "=r"(m)
        : "r"(y0), "r"(x1), "r"(P::INV));
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %10, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %10, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %10, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %10, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %10, 0;\n\t"
        "madc.hi.u32 %8, %13, %10, 0;"
        : "+r"(y0), "=r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7), "+r"(x7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.u32 %0, %2, %3;\n\t"
        "mul.lo.u32 %1, %0, %4;"
        : "=r"(w), "=r"(m)
        : "r"(x0), "r"(y1), "r"(P::INV));
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %10, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %10, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %10, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %10, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %10, 0;\n\t"
        "madc.hi.u32 %8, %13, %10, 0;"
        : "+r"(x0), "=r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(y7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.u32 %0, %2, %3;\n\t"
        "mul.lo.u32 %1, %0, %4;"
        : "=r"(w), "=r"(m)
        : "r"(y0), "r"(x1), "r"(P::INV));
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %10, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %10, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %10, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %10, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %10, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %10, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %10, 0;\n\t"
        "madc.hi.u32 %8, %13, %10, 0;"
        : "+r"(y0), "=r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7)
        : "r"(P::mod(1)), "r"(m), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %10, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7), "+r"(x7)
        : "r"(P::mod(0)), "r"(m), "r"(P::mod(2)), "r"(P::mod(4)), "r"(P::mod(6)));
    asm("add.cc.u32 %0, %8, %9;\n\t"
        "addc.cc.u32 %1, %10, %11;\n\t"
        "addc.cc.u32 %2, %12, %13;\n\t"
        "addc.cc.u32 %3, %14, %15;\n\t"
        "addc.cc.u32 %4, %16, %17;\n\t"
        "addc.cc.u32 %5, %18, %19;\n\t"
        "addc.cc.u32 %6, %20, %21;\n\t"
        "addc.u32 %7, %22, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7)
        : "r"(x0), "r"(y1), "r"(x1), "r"(y2), "r"(x2), "r"(y3), "r"(x3), "r"(y4), "r"(x4), "r"(y5), "r"(x5), "r"(y6), "r"(x6), "r"(y7), "r"(x7));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(t0), "+r"(t1), "+r"(t2), "+r"(t3), "+r"(t4), "+r"(t5), "+r"(t6), "+r"(t7)
        : "r"(T8), "r"(T9), "r"(T10), "r"(T11), "r"(T12), "r"(T13), "r"(T14), "r"(T15));
    uint32_t t[8] = {t0, t1, t2, t3, t4, t5, t6, t7};
    return fp_final_sub<P>(t);
    // ---- end of generated code ----
#else
    return fp_mul(a, a);
#endif
}

template <class P> KZG_HD Fp<P> fp_to_mont(const Fp<P>& a) {
    return fp_mul(a, fp_r2<P>());
}
template <class P> KZG_HD Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> one = fp_zero<P>();
    one.l[0] = 1;
    return fp_mul(a, one);
}

// a^e, e given as 8 little-endian 32-bit limbs (plain integer)
template <class P> KZG_HD Fp<P> fp_pow(const Fp<P>& a, const uint32_t* e) {
    Fp<P> r = fp_one<P>();
    bool started = false;
    for (int i = 7; i >= 0; i--) {
        for (int bit = 31; bit >= 0; bit--) {
            if (started) r = fp_sqr(r);
            if ((e[i] >> bit) & 1) {
                r = started ? fp_mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}
template <class P> KZG_HD Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t e) {
    uint32_t ee[8] = {(uint32_t)e, (uint32_t)(e >> 32), 0, 0, 0, 0, 0, 0};
    return fp_pow(a, ee);
}
// a^-1 = a^(p-2); inv(0) = 0   (Fermat; the specification the fast inversion is checked against)
template <class P> KZG_HD Fp<P> fp_inv_fermat(const Fp<P>& a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = P::mod(i);
    e[0] -= 2;  // mod(0) >= 2 for both fields, no borrow
    return fp_pow(a, e);
}

// ---- binary extended Euclid on plain 256-bit integers -------------------------------------------------
// One inversion costs ~380 cheap add/shift steps instead of 380 Montgomery products: the single-thread
// epilogues (projective -> affine at the end of every MSM, the base case of the batch inversion) are pure
// latency, and this is 4-5 x shorter.
struct U256 {
    uint32_t w[8];
};
KZG_HD bool u256_is_one(const U256& a) {
    uint32_t o = a.w[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) o |= a.w[i];
    return o == 0;
}
KZG_HD bool u256_geq(const U256& a, const U256& b) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (a.w[i] > b.w[i]) return true;
        if (a.w[i] < b.w[i]) return false;
    }
    return true;
}
// a -= b, returns the borrow
KZG_HD uint32_t u256_sub(U256& a, const U256& b) {
    uint64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.w[i] - b.w[i] - bw;
        a.w[i] = (uint32_t)d;
        bw = (d >> 32) & 1;
    }
    return (uint32_t)bw;
}
// a += b, returns the carry
KZG_HD uint32_t u256_add(U256& a, const U256& b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.w[i] + b.w[i];
        a.w[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}
// a = (a + top * 2^256) >> 1
KZG_HD void u256_shr1(U256& a, uint32_t top) {
#pragma unroll
    for (int i = 0; i < 7; i++) a.w[i] = (a.w[i] >> 1) | (a.w[i + 1] << 31);
    a.w[7] = (a.w[7] >> 1) | (top << 31);
}

// a^-1 for a Montgomery residue a (inv(0) = 0): X = (aR)^-1 as an integer by the binary extended Euclidean
// algorithm, then one Montgomery product with R^3 turns a^-1 R^-1 into a^-1 R.
template <class P> KZG_HD Fp<P> fp_inv_euclid(const Fp<P>& a) {
    if (fp_is_zero(a)) return a;
    U256 u, v, x1, x2, p;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u.w[i] = a.l[i];
        p.w[i] = P::mod(i);
        v.w[i] = P::mod(i);
        x1.w[i] = i == 0 ? 1u : 0u;
        x2.w[i] = 0u;
    }
    // invariants: x1 * a = u, x2 * a = v (mod p); u, v odd-or-being-halved, x1, x2 in [0, p)
    while (!u256_is_one(u) && !u256_is_one(v)) {
        while ((u.w[0] & 1u) == 0) {
            u256_shr1(u, 0);
            uint32_t top = 0;
            if (x1.w[0] & 1u) top = u256_add(x1, p);
            u256_shr1(x1, top);
        }
        while ((v.w[0] & 1u) == 0) {
            u256_shr1(v, 0);
            uint32_t top = 0;
            if (x2.w[0] & 1u) top = u256_add(x2, p);
            u256_shr1(x2, top);
        }
        if (u256_geq(u, v)) {
            u256_sub(u, v);
            if (u256_sub(x1, x2)) u256_add(x1, p);
        } else {
            u256_sub(v, u);
            if (u256_sub(x2, x1)) u256_add(x2, p);
        }
    }
    const U256& x = u256_is_one(u) ? x1 : x2;
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = x.w[i];
    // R^3 = mont(R^2, R^2) * R ... : mont_mul(R^2, R^2) = R^3
    const Fp<P> r3 = fp_mul(fp_r2<P>(), fp_r2<P>());
    return fp_mul(r, r3);
}


// ---- binary GCD with 31 steps per multi-limb update (Pornin, "Optimized Binary GCD for Modular Inversion",
// eprint 2020/972, algorithm 2 with k = 32; variable time) -------------------------------------------------
// The Euclid loop above touches all eight limbs of four numbers in every one of its ~380 steps; for the single lane
// that runs it (projective -> affine at the end of every MSM, the bottom of every batch inversion) that is ~100 us
// of dependent instructions.  Here 31 steps at a time run on 64-bit APPROXIMATIONS of a and b (their low 31 bits,
// which decide the parities exactly, and their top 33 bits, which decide the comparisons almost always) and only
// record the 2x2 matrix (f0 g0; f1 g1) of what they did; the matrix is then applied once to the full numbers:
//     (a, b) <- (a f0 + b g0, a f1 + b g1) / 2^31            exact division; a wrong comparison shows up as a
//                                                             negative result and is fixed by negating the row
//     (u, v) <- (u f0 + v g0, u f1 + v g1) / 2^31  mod p     one 31-bit Montgomery reduction step each
// with the invariants a = u y, b = v y (mod p), b odd.  When a reaches 0, b = gcd = 1 and v = 1 / y.
struct BgRow {
    uint32_t f, g;  // magnitudes, <= 2^31
    bool fneg, gneg;
};
// out (9 limbs) = x * f
KZG_HD void bg_mul_small(const uint32_t* x, uint32_t f, uint32_t* out) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)x[i] * f;
        out[i] = (uint32_t)c;
        c >>= 32;
    }
    out[8] = (uint32_t)c;
}
// |x f + y g| >> 31 for signed (f, g) -> out (8 limbs); returns true if x f + y g was negative
KZG_HD bool bg_combine(const uint32_t* x, const uint32_t* y, const BgRow& r, uint32_t* out) {
    uint32_t X[9], Y[9];
    bg_mul_small(x, r.f, X);
    bg_mul_small(y, r.g, Y);
    bool neg = r.fneg;
    if (r.fneg == r.gneg) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            c += (uint64_t)X[i] + Y[i];
            X[i] = (uint32_t)c;
            c >>= 32;
        }
    } else {
        uint64_t bw = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            uint64_t d = (uint64_t)X[i] - Y[i] - bw;
            X[i] = (uint32_t)d;
            bw = (d >> 32) & 1;
        }
        if (bw) {  // |Y| > |X|: the sign is g's, the magnitude the two's complement
            neg = r.gneg;
            uint64_t c = 1;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                c += (uint64_t)(~X[i]);
                X[i] = (uint32_t)c;
                c >>= 32;
            }
        }
    }
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        out[i] = (X[i] >> 31) | (X[i + 1] << 1);
        nz |= out[i];
    }
    return neg && nz != 0;
}
// (u f + v g) / 2^31 mod p for signed (f, g), u, v in [0, p)
template <class P> KZG_HD void bg_combine_mod(const uint32_t* u, const uint32_t* v, const BgRow& r, uint32_t* out) {
    uint32_t us[8], vs[8], X[9], Y[9];
    {  // a negative factor: use p - u (0 stays 0)
        uint32_t nzu = 0, nzv = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            nzu |= u[i];
            nzv |= v[i];
        }
        uint64_t bu = 0, bv = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t du = (uint64_t)P::mod(i) - u[i] - bu;
            uint64_t dv = (uint64_t)P::mod(i) - v[i] - bv;
            bu = (du >> 32) & 1;
            bv = (dv >> 32) & 1;
            us[i] = (r.fneg && nzu) ? (uint32_t)du : u[i];
            vs[i] = (r.gneg && nzv) ? (uint32_t)dv : v[i];
        }
    }
    bg_mul_small(us, r.f, X);
    bg_mul_small(vs, r.g, Y);
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {  // < p 2^31 since f + g <= 2^31
        c += (uint64_t)X[i] + Y[i];
        X[i] = (uint32_t)c;
        c >>= 32;
    }
    const uint32_t q = (X[0] * P::INV) & 0x7fffffffu;  // X + q p = 0 mod 2^31
    c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)q * P::mod(i) + X[i];
        X[i] = (uint32_t)c;
        c >>= 32;
    }
    X[8] += (uint32_t)c;
    uint32_t t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = (X[i] >> 31) | (X[i + 1] << 1);  // < 2 p
    const Fp<P> red = fp_final_sub<P>(t);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = red.l[i];
}
// bit length of an 8-limb number
KZG_HD uint32_t bg_clz32(uint32_t w) {  // w != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)w);
#else
    return (uint32_t)__builtin_clz(w);
#endif
}
KZG_HD uint32_t bg_len(const uint32_t* x) {
    uint32_t len = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (x[i]) len = 32 * i + 32 - bg_clz32(x[i]);
    return len;
}
// low 31 bits + the 33 bits below bit n (n >= 64) of an 8-limb number, as one 64-bit word
KZG_HD uint64_t bg_approx(const uint32_t* x, uint32_t n) {
    const uint32_t s = n - 33, idx = s >> 5, off = s & 31;
    uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {  // (no dynamic indexing: the limbs stay in registers on the device)
        if ((uint32_t)i == idx) w0 = x[i];
        if ((uint32_t)i == idx + 1) w1 = x[i];
        if ((uint32_t)i == idx + 2) w2 = x[i];
    }
    uint64_t top = (((uint64_t)w1 << 32) | w0) >> off;
    if (off) top |= (uint64_t)w2 << (64 - off);
    top &= (1ull << 33) - 1;
    return (top << 31) | (x[0] & 0x7fffffffu);
}

// a^-1 for a Montgomery residue a (inv(0) = 0): X = (aR)^-1 as an integer, then one Montgomery product with R^3
// turns a^-1 R^-1 into a^-1 R.
template <class P> KZG_HD Fp<P> fp_inv(const Fp<P>& y) {
    if (fp_is_zero(y)) return y;
    uint32_t a[8], b[8], u[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = y.l[i];
        b[i] = P::mod(i);
        u[i] = i == 0 ? 1u : 0u;
        v[i] = 0u;
    }
    for (int iter = 0; iter < 40; iter++) {
        uint32_t nza = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) nza |= a[i];
        if (!nza) break;
        uint32_t n = bg_len(a), nb = bg_len(b);
        if (nb > n) n = nb;
        if (n < 64) n = 64;
        uint64_t abar = bg_approx(a, n), bbar = bg_approx(b, n);
        int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
        for (int j = 0; j < 31; j++) {
            if (abar & 1) {
                if (abar < bbar) {
                    uint64_t tt = abar; abar = bbar; bbar = tt;
                    int64_t tf = f0; f0 = f1; f1 = tf;
                    int64_t tg = g0; g0 = g1; g1 = tg;
                }
                abar -= bbar;
                f0 -= f1;
                g0 -= g1;
            }
            abar >>= 1;
            f1 <<= 1;
            g1 <<= 1;
        }
        BgRow r0, r1;
        r0.fneg = f0 < 0; r0.f = (uint32_t)(r0.fneg ? -f0 : f0);
        r0.gneg = g0 < 0; r0.g = (uint32_t)(r0.gneg ? -g0 : g0);
        r1.fneg = f1 < 0; r1.f = (uint32_t)(r1.fneg ? -f1 : f1);
        r1.gneg = g1 < 0; r1.g = (uint32_t)(r1.gneg ? -g1 : g1);
        uint32_t na[8], nb8[8];
        if (bg_combine(a, b, r0, na)) {  // a came out negative: negate the row
            r0.fneg = !r0.fneg;
            r0.gneg = !r0.gneg;
        }
        if (bg_combine(a, b, r1, nb8)) {
            r1.fneg = !r1.fneg;
            r1.gneg = !r1.gneg;
        }
        uint32_t nu[8], nv[8];
        bg_combine_mod<P>(u, v, r0, nu);
        bg_combine_mod<P>(u, v, r1, nv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a[i] = na[i];
            b[i] = nb8[i];
            u[i] = nu[i];
            v[i] = nv[i];
        }
    }
    uint32_t rest = (b[0] ^ 1u);
#pragma unroll
    for (int i = 1; i < 8; i++) rest |= b[i];
#pragma unroll
    for (int i = 0; i < 8; i++) rest |= a[i];
    if (rest) return fp_inv_euclid(y);  // (not reached for a prime modulus: gcd = 1 within 2 * 254 steps)
    Fp<P> x;
#pragma unroll
    for (int i = 0; i < 8; i++) x.l[i] = v[i];
    const Fp<P> r3 = fp_mul(fp_r2<P>(), fp_r2<P>());
    return fp_mul(x, r3);
}

}  // namespace kzg
