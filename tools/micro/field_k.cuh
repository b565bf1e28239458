// Montgomery product with a Karatsuba 8x8-limb multiplication followed by a separated (product-then-reduce)
// Montgomery reduction.  Device only.
//
//   a * b        = z0 + ((al+ah)(bl+bh) - z0 - z2) 2^128 + z2 2^256     three 4x4-limb products: 48 wide MACs
//   reduction    : 8 rows  W += m_i p 2^(32 i),  m_i = W[i] * (-p^-1) mod 2^32                       64 wide MACs + 8 IMAD
// i.e. 112 IMAD.WIDE + 8 IMAD against 128 + 8 for the interleaved schoolbook CIOS of field.cuh; the extra work is
// additions on the ALU pipe, which the field kernels leave mostly idle (27 % busy in the MSM accumulation).
// Every carry chain is one asm statement (the CC flag does not survive between statements).
#pragma once
#include "../../kzg_grandsums_study_b200/csrc/field.cuh"

#if defined(__CUDACC__)
namespace kzg {

// w[0..3] += {x0, x1} * y  (x0*y at words 0-1, x1*y at words 2-3); returns the carry out of word 3
__device__ __forceinline__ uint32_t kmad2(uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t& w3, uint32_t x0, uint32_t x1,
                                         uint32_t y) {
    uint32_t c;
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(w3), "=r"(c)
        : "r"(x0), "r"(x1), "r"(y));
    return c;
}

// z[0..7] = x[0..3] * y[0..3]  (schoolbook, 16 wide MACs).  Row j adds {x0,x2} y_j at words j.., then {x1,x3} y_j at
// words j+1..; the carry of a chain is rippled to the top of z inside a second chain.
__device__ __forceinline__ void kmul4(uint32_t* z, const uint32_t* x, const uint32_t* y) {
#pragma unroll
    for (int i = 0; i < 8; i++) z[i] = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t c = kmad2(z[j], z[j + 1], z[j + 2], z[j + 3], x[0], x[2], y[j]);
        // ripple c into z[j+4 .. 7]
        if (j + 4 <= 7) {
            if (j == 0)
                asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.cc.u32 %2, %2, 0;\n\t addc.u32 %3, %3, 0;\n\t"
                    : "+r"(z[4]), "+r"(z[5]), "+r"(z[6]), "+r"(z[7]) : "r"(c));
            else if (j == 1)
                asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;\n\t"
                    : "+r"(z[5]), "+r"(z[6]), "+r"(z[7]) : "r"(c));
            else if (j == 2)
                asm("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, 0;\n\t" : "+r"(z[6]), "+r"(z[7]) : "r"(c));
            else
                z[7] += c;
        }
        if (j < 3) {
            uint32_t d = kmad2(z[j + 1], z[j + 2], z[j + 3], z[j + 4], x[1], x[3], y[j]);
            if (j == 0)
                asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;\n\t"
                    : "+r"(z[5]), "+r"(z[6]), "+r"(z[7]) : "r"(d));
            else if (j == 1)
                asm("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, %1, 0;\n\t" : "+r"(z[6]), "+r"(z[7]) : "r"(d));
            else
                z[7] += d;
        } else {
            // top row: words 4..7, the product fits (no carry out)
            asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
                "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
                "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
                "madc.hi.u32 %3, %5, %6, %3;\n\t"
                : "+r"(z[4]), "+r"(z[5]), "+r"(z[6]), "+r"(z[7])
                : "r"(x[1]), "r"(x[3]), "r"(y[3]));
        }
    }
}

// r[0..3] = x[0..3] + y[0..3]; returns the carry (0 / 1)
__device__ __forceinline__ uint32_t kadd4(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    uint32_t c;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(c)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
    return c;
}
// x[0..7] -= y[0..7]  (mod 2^256)
__device__ __forceinline__ void ksub8(uint32_t* x, const uint32_t* y) {
    asm("sub.cc.u32 %0, %0, %8;\n\t"
        "subc.cc.u32 %1, %1, %9;\n\t"
        "subc.cc.u32 %2, %2, %10;\n\t"
        "subc.cc.u32 %3, %3, %11;\n\t"
        "subc.cc.u32 %4, %4, %12;\n\t"
        "subc.cc.u32 %5, %5, %13;\n\t"
        "subc.cc.u32 %6, %6, %14;\n\t"
        "subc.u32 %7, %7, %15;\n\t"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
        : "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
}

// T[0..15] = a * b for a, b < 2^254  (Karatsuba over 128-bit halves)
__device__ __forceinline__ void kmul8(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t z0[8], z2[8], zm[8], sa[4], sb[4];
    kmul4(z0, a, b);
    kmul4(z2, a + 4, b + 4);
    const uint32_t ca = kadd4(sa, a, a + 4), cb = kadd4(sb, b, b + 4);
    kmul4(zm, sa, sb);
    // (sa + ca 2^128)(sb + cb 2^128) - z0 - z2 = al bh + ah bl < 2^255: exact mod 2^256
    {
        const uint32_t ma = 0u - ca, mb = 0u - cb;
        asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, %5;\n\t addc.cc.u32 %2, %2, %6;\n\t addc.u32 %3, %3, %7;\n\t"
            : "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7])
            : "r"(sb[0] & ma), "r"(sb[1] & ma), "r"(sb[2] & ma), "r"(sb[3] & ma));
        asm("add.cc.u32 %0, %0, %4;\n\t addc.cc.u32 %1, %1, %5;\n\t addc.cc.u32 %2, %2, %6;\n\t addc.u32 %3, %3, %7;\n\t"
            : "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7])
            : "r"(sa[0] & mb), "r"(sa[1] & mb), "r"(sa[2] & mb), "r"(sa[3] & mb));
    }
    ksub8(zm, z0);
    ksub8(zm, z2);
    // T = z0 + zm 2^128 + z2 2^256
#pragma unroll
    for (int i = 0; i < 4; i++) T[i] = z0[i];
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "=r"(T[4]), "=r"(T[5]), "=r"(T[6]), "=r"(T[7]), "=r"(T[8]), "=r"(T[9]), "=r"(T[10]), "=r"(T[11]), "=r"(c)
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(zm[0]),
          "r"(zm[1]), "r"(zm[2]), "r"(zm[3]), "r"(zm[4]), "r"(zm[5]), "r"(zm[6]), "r"(zm[7]));
    asm("add.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, 0;\n\t"
        "addc.cc.u32 %2, %6, 0;\n\t"
        "addc.u32 %3, %7, 0;\n\t"
        : "=r"(T[12]), "=r"(T[13]), "=r"(T[14]), "=r"(T[15])
        : "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]), "r"(c));
}

// w[0..7] += {x0, x1, x2, x3} * y (x_k * y at words 2k, 2k+1), then the carry goes on into w8; returns the carry out of w8
__device__ __forceinline__ uint32_t kmad4_w8(uint32_t* w, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    uint32_t c;
    asm("mad.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.cc.u32 %7, %13, %14, %7;\n\t"
        "addc.cc.u32 %8, %8, 0;\n\t"
        "addc.u32 %9, 0, 0;\n\t"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "=r"(c)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
    return c;
}
// w[0..7] += {x0..x3} * y; returns the carry out of w7
__device__ __forceinline__ uint32_t kmad4(uint32_t* w, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    uint32_t c;
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "=r"(c)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
    return c;
}

// Montgomery reduction of T[0..15] (< 2^508): returns T / 2^256 mod p, fully reduced.
// Row i adds m_i p at words i..i+8; its two carries into word i+9 are added there together with the carry that
// the previous row's addition into word i+8 produced, so no carry ever ripples further than one word.
template <class P> __device__ __forceinline__ Fp<P> kredc(uint32_t* T) {
    uint32_t W[18];
#pragma unroll
    for (int i = 0; i < 16; i++) W[i] = T[i];
    W[16] = 0;
    W[17] = 0;
    uint32_t pend = 0;  // carry out of the previous row's update of word i+8, to be added at word i+9
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t m = W[i] * P::INV;
        // even limbs of p at words i, i+2, i+4, i+6 (carry runs on into word i+8, then out to word i+9)
        const uint32_t ca = kmad4_w8(W + i, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
        // odd limbs of p at words i+1, i+3, i+5, i+7 (carry out of word i+8 to word i+9)
        const uint32_t cb = kmad4(W + i + 1, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        // word i+9 += ca + cb + pend; the overflow of this addition is next row's pend (it belongs to word i+10)
        uint32_t np;
        const uint32_t s = ca + cb + pend;  // <= 3
        asm("add.cc.u32 %0, %0, %2;\n\t addc.u32 %1, 0, 0;\n\t" : "+r"(W[i + 9]), "=r"(np) : "r"(s));
        pend = np;
    }
    // result = W[8..15] (+ W[16], pend == 0 because the value is < 2p < 2^255)
    return fp_final_sub<P>(W + 8);
}

template <class P> __device__ __forceinline__ Fp<P> fp_mul_k(const Fp<P>& a, const Fp<P>& b) {
    uint32_t T[16];
    kmul8(T, a.l, b.l);
    return kredc<P>(T);
}

}  // namespace kzg
#endif
