/* kzg_oracle.c -- ORACLE, test infrastructure only (CPU restatement in plain C, multi-threaded with OpenMP).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (libkzgb200.so) never does.
 *
 * What it restates
 *   - the arithmetic the reference delegates to the un-vendored npm dependency ffjavascript@0.2.59 /
 *     wasmcurves@0.2.1 (reference package-lock.json:300-309, 905-912), by its published algorithms:
 *       Fr / Fq Montgomery arithmetic (R = 2^256, 4 x 64-bit limbs, fully reduced)
 *       G1 Jacobian add / mixed add / double, toAffine
 *       G1.multiExpAffine: unsigned fixed-window Pippenger, window = pTSizes[floor(log2 n)], one task per
 *         (point chunk, window), windows combined by repeated doubling        (call site polynomial.js:1112)
 *       Fr.fft / Fr.ifft: radix-2, natural order in/out, w = 5^((r-1)/2^k)     (evaluations.js:18, polynomial.js:34)
 *       Fr.batchInverse: Montgomery trick per slice, zero -> zero              (grandsum.js:41)
 *   - the reference's own code on the hot path, function by function (file:line cited at each function):
 *       src/polynomial/polynomial.js, src/polynomial/evaluations.js, src/grandsum/grandsum.js,
 *       src/grandproduct/grandproduct.js, src/grandsum/mset_eq_kzg_prover.js,
 *       src/grandproduct/mset_eq_kzg_prover.js, src/Keccak256Transcript.js
 *
 * PARITY PIN: the reference's tests hold no golden vectors and the reference cannot
 * run on this image (no Node).  This file is pinned against the Python oracle (oracle/py, itself pinned to
 * SURVEY.md Appendix B constants, the EIP-196 vector, keccak256("") and the Appendix F end-to-end vector)
 * in tests/test_oracle_c.py.  "Parity unpinned" at the ffjavascript boundary -- see DESIGN.md.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct { u64 l[4]; } fe;
typedef struct { u64 p[4]; u64 inv; fe r1; fe r2; } field_t;

static const field_t FQ = {
    {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0x87d20782e4866389ull,
    {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
    {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}}};
static const field_t FR = {
    {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
    0xc2e1f593efffffffull,
    {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
    {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}}};

/* ------------------------------------------------------------------------------------------------ field */
static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int geq_p(const u64* a, const field_t* F) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > F->p[i]) return 1;
        if (a[i] < F->p[i]) return 0;
    }
    return 1;
}
static inline void sub_p(u64* a, const field_t* F) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - F->p[i] - (u64)b;
        a[i] = (u64)d;
        b = (d >> 64) & 1;
    }
}
static inline void fe_add(fe* r, const fe* a, const fe* b, const field_t* F) {
    u128 c = 0;
    u64 t[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (u64)c;
        c >>= 64;
    }
    if (geq_p(t, F)) sub_p(t, F);   /* p < 2^254: no carry out */
    memcpy(r->l, t, 32);
}
static inline void fe_sub(fe* r, const fe* a, const fe* b, const field_t* F) {
    u64 t[4];
    u64 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        t[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + F->p[i];
            t[i] = (u64)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, 32);
}
static inline void fe_neg(fe* r, const fe* a, const field_t* F) {
    fe z = {{0, 0, 0, 0}};
    fe_sub(r, &z, a, F);
}
/* Montgomery product, CIOS */
static inline void fe_mul(fe* r, const fe* a, const fe* b, const field_t* F) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (u64)c;
        t[5] = (u64)(c >> 64);
        u64 m = t[0] * F->inv;
        c = (u128)m * F->p[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->p[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (u64)c;
        t[4] = t[5] + (u64)(c >> 64);
    }
    if (t[4] || geq_p(t, F)) sub_p(t, F);
    memcpy(r->l, t, 32);
}
static inline void fe_sqr(fe* r, const fe* a, const field_t* F) { fe_mul(r, a, a, F); }
static void fe_pow(fe* r, const fe* a, const u64 e[4], const field_t* F) {
    fe acc = F->r1, base = *a;
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 64; b++) {
            if ((e[i] >> b) & 1) fe_mul(&acc, &acc, &base, F);
            fe_sqr(&base, &base, F);
        }
    *r = acc;
}
static void fe_inv(fe* r, const fe* a, const field_t* F) {   /* a^(p-2); inv(0) = 0 */
    u64 e[4] = {F->p[0] - 2, F->p[1], F->p[2], F->p[3]};
    fe_pow(r, a, e, F);
}
static inline void fe_to_mont(fe* r, const fe* a, const field_t* F) { fe_mul(r, a, &F->r2, F); }
static inline void fe_from_mont(fe* r, const fe* a, const field_t* F) {
    fe one = {{1, 0, 0, 0}};
    fe_mul(r, a, &one, F);
}
static void fe_from_u64(fe* r, u64 x, const field_t* F) {
    fe t = {{x, 0, 0, 0}};
    fe_to_mont(r, &t, F);
}

/* ------------------------------------------------------------------------------------------------ G1 */
typedef struct { fe x, y; } g1a;          /* affine, Montgomery; infinity = all zero (ffjavascript) */
typedef struct { fe x, y, z; } g1j;       /* Jacobian; infinity: z = 0 */

static inline int g1a_is_inf(const g1a* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void g1j_set_inf(g1j* p) { memset(p, 0, sizeof(*p)); }
static inline int g1j_is_inf(const g1j* p) { return fe_is_zero(&p->z); }

static void g1j_double(g1j* r, const g1j* p) {   /* dbl-2009-l, a = 0 */
    if (g1j_is_inf(p)) { *r = *p; return; }
    const field_t* F = &FQ;
    fe A, B, C, D, E, Fq_, t, x3, y3, z3;
    fe_sqr(&A, &p->x, F);
    fe_sqr(&B, &p->y, F);
    fe_sqr(&C, &B, F);
    fe_add(&t, &p->x, &B, F);
    fe_sqr(&t, &t, F);
    fe_sub(&t, &t, &A, F);
    fe_sub(&t, &t, &C, F);
    fe_add(&D, &t, &t, F);
    fe_add(&E, &A, &A, F);
    fe_add(&E, &E, &A, F);
    fe_sqr(&Fq_, &E, F);
    fe_sub(&x3, &Fq_, &D, F);
    fe_sub(&x3, &x3, &D, F);
    fe_sub(&t, &D, &x3, F);
    fe_mul(&y3, &E, &t, F);
    fe_add(&C, &C, &C, F);
    fe_add(&C, &C, &C, F);
    fe_add(&C, &C, &C, F);
    fe_sub(&y3, &y3, &C, F);
    fe_mul(&z3, &p->y, &p->z, F);
    fe_add(&z3, &z3, &z3, F);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_add_affine(g1j* r, const g1j* p, const g1a* q) {   /* madd-2007-bl with the special cases */
    const field_t* F = &FQ;
    if (g1a_is_inf(q)) { *r = *p; return; }
    if (g1j_is_inf(p)) { r->x = q->x; r->y = q->y; r->z = F->r1; return; }
    fe z1z1, u2, s2, h, hh, i, j, rr, v, t, x3, y3, z3;
    fe_sqr(&z1z1, &p->z, F);
    fe_mul(&u2, &q->x, &z1z1, F);
    fe_mul(&s2, &q->y, &p->z, F);
    fe_mul(&s2, &s2, &z1z1, F);
    fe_sub(&h, &u2, &p->x, F);
    fe_sub(&rr, &s2, &p->y, F);
    if (fe_is_zero(&h)) {
        if (fe_is_zero(&rr)) { g1j_double(r, p); return; }
        g1j_set_inf(r);
        return;
    }
    fe_sqr(&hh, &h, F);
    fe_add(&i, &hh, &hh, F);
    fe_add(&i, &i, &i, F);
    fe_mul(&j, &h, &i, F);
    fe_add(&rr, &rr, &rr, F);
    fe_mul(&v, &p->x, &i, F);
    fe_sqr(&x3, &rr, F);
    fe_sub(&x3, &x3, &j, F);
    fe_sub(&x3, &x3, &v, F);
    fe_sub(&x3, &x3, &v, F);
    fe_sub(&t, &v, &x3, F);
    fe_mul(&y3, &rr, &t, F);
    fe_mul(&t, &p->y, &j, F);
    fe_add(&t, &t, &t, F);
    fe_sub(&y3, &y3, &t, F);
    fe_add(&z3, &p->z, &h, F);
    fe_sqr(&z3, &z3, F);
    fe_sub(&z3, &z3, &z1z1, F);
    fe_sub(&z3, &z3, &hh, F);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_add(g1j* r, const g1j* p, const g1j* q) {   /* add-2007-bl with the special cases */
    const field_t* F = &FQ;
    if (g1j_is_inf(q)) { *r = *p; return; }
    if (g1j_is_inf(p)) { *r = *q; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, y3, z3;
    fe_sqr(&z1z1, &p->z, F);
    fe_sqr(&z2z2, &q->z, F);
    fe_mul(&u1, &p->x, &z2z2, F);
    fe_mul(&u2, &q->x, &z1z1, F);
    fe_mul(&s1, &p->y, &q->z, F);
    fe_mul(&s1, &s1, &z2z2, F);
    fe_mul(&s2, &q->y, &p->z, F);
    fe_mul(&s2, &s2, &z1z1, F);
    fe_sub(&h, &u2, &u1, F);
    fe_sub(&rr, &s2, &s1, F);
    if (fe_is_zero(&h)) {
        if (fe_is_zero(&rr)) { g1j_double(r, p); return; }
        g1j_set_inf(r);
        return;
    }
    fe_add(&i, &h, &h, F);
    fe_sqr(&i, &i, F);
    fe_mul(&j, &h, &i, F);
    fe_add(&rr, &rr, &rr, F);
    fe_mul(&v, &u1, &i, F);
    fe_sqr(&x3, &rr, F);
    fe_sub(&x3, &x3, &j, F);
    fe_sub(&x3, &x3, &v, F);
    fe_sub(&x3, &x3, &v, F);
    fe_sub(&t, &v, &x3, F);
    fe_mul(&y3, &rr, &t, F);
    fe_mul(&t, &s1, &j, F);
    fe_add(&t, &t, &t, F);
    fe_sub(&y3, &y3, &t, F);
    fe_add(&z3, &p->z, &q->z, F);
    fe_sqr(&z3, &z3, F);
    fe_sub(&z3, &z3, &z1z1, F);
    fe_sub(&z3, &z3, &z2z2, F);
    fe_mul(&z3, &z3, &h, F);
    r->x = x3; r->y = y3; r->z = z3;
}
static void g1j_to_affine(g1a* r, const g1j* p) {   /* G1.toAffine: infinity -> zeros */
    const field_t* F = &FQ;
    if (g1j_is_inf(p)) { memset(r, 0, sizeof(*r)); return; }
    fe zi, zi2, zi3;
    fe_inv(&zi, &p->z, F);
    fe_sqr(&zi2, &zi, F);
    fe_mul(&zi3, &zi2, &zi, F);
    fe_mul(&r->x, &p->x, &zi2, F);
    fe_mul(&r->y, &p->y, &zi3, F);
}

/* ------------------------------------------------------------------------------------------------ MSM
 * G1.multiExpAffine of ffjavascript 0.2.59 [dep; SURVEY.md Appendix C]: unsigned windows of
 * pTSizes[floor(log2 n)] bits over the 256-bit little-endian scalar, points split in chunks of
 * clamp(n / (concurrency / nWindows), 2^10, 2^22); one task per (chunk, window): bucket accumulation with
 * mixed adds, then the running-sum bucket reduction; windows combined by repeated doubling, chunks summed. */
static const int PT_SIZES[] = {1, 1, 1, 1, 2, 3, 4, 5, 6, 7, 7, 8, 9, 10, 11, 12, 13, 13, 14, 15, 16, 16, 17, 17, 17, 17,
                               17, 17, 17, 17, 17, 17, 17};

static inline uint32_t scalar_window(const uint8_t* s, int pos, int bits) {
    uint32_t v = 0;
    for (int b = 0; b < bits; b++) {
        int p = pos + b;
        if (p >= 256) break;
        v |= (uint32_t)((s[p >> 3] >> (p & 7)) & 1) << b;
    }
    return v;
}
static void msm_chunk_window(const g1a* bases, const uint8_t* scalars, size_t n, int pos, int bits, g1j* out) {
    size_t nb = (size_t)1 << bits;
    g1j* buckets = (g1j*)calloc(nb, sizeof(g1j));
    for (size_t i = 0; i < n; i++) {
        uint32_t d = scalar_window(scalars + 32 * i, pos, bits);
        if (d) g1j_add_affine(&buckets[d], &buckets[d], &bases[i]);
    }
    g1j run, tot;
    g1j_set_inf(&run);
    g1j_set_inf(&tot);
    for (size_t d = nb - 1; d >= 1; d--) {
        g1j_add(&run, &run, &buckets[d]);
        g1j_add(&tot, &tot, &run);
    }
    free(buckets);
    *out = tot;
}
int ko_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* bases: n x 64 B affine Montgomery-LE; scalars: n x 32 B standard-form LE; out: 64 B affine (toAffine applied) */
void ko_g1_msm(const uint8_t* bases_b, const uint8_t* scalars, size_t n, uint8_t out[64], int threads) {
    const g1a* bases = (const g1a*)bases_b;
    g1j res;
    g1j_set_inf(&res);
    if (n == 0) { memset(out, 0, 64); return; }
    if (threads <= 0) threads = ko_max_threads();
    int lg = 0;
    while (((size_t)2 << lg) <= n) lg++;
    const int c = PT_SIZES[lg];
    const int nwin = (32 * 8 - 1) / c + 1;
    size_t chunk = (size_t)((double)n / ((double)threads / nwin));
    if (chunk > 4194304) chunk = 4194304;
    if (chunk < 1024) chunk = 1024;
    const size_t nchunks = (n + chunk - 1) / chunk;
    g1j* parts = (g1j*)malloc(sizeof(g1j) * nchunks * nwin);
    const long ntasks = (long)(nchunks * nwin);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (long t = 0; t < ntasks; t++) {
        size_t ch = (size_t)t / nwin;
        int w = (int)(t % nwin);
        size_t lo = ch * chunk, hi = lo + chunk < n ? lo + chunk : n;
        int bits = c;
        if (w * c + bits > 256) bits = 256 - w * c;
        msm_chunk_window(bases + lo, scalars + 32 * lo, hi - lo, w * c, bits, &parts[t]);
    }
    for (size_t ch = 0; ch < nchunks; ch++) {
        g1j acc;
        g1j_set_inf(&acc);
        for (int w = nwin - 1; w >= 0; w--) {
            if (!g1j_is_inf(&acc))
                for (int j = 0; j < c; j++) g1j_double(&acc, &acc);
            g1j_add(&acc, &acc, &parts[ch * nwin + w]);
        }
        g1j_add(&res, &res, &acc);
    }
    free(parts);
    g1a a;
    g1j_to_affine(&a, &res);
    memcpy(out, &a, 64);
}

/* synthetic SRS [tau^i]_1, first <= i < first + n (fixed-base 8-bit windows; not part of the reference, which
 * downloads the Hermez file -- .github/workflows/tests.yml:15-19) */
void ko_srs_generate(const uint8_t tau_std[32], u64 first, size_t n, uint8_t* out, int threads) {
    if (threads <= 0) threads = ko_max_threads();
    const field_t* F = &FQ;
    g1a* table = (g1a*)malloc(sizeof(g1a) * 32 * 256);   /* table[w][d] = d * 2^(8w) * G */
    g1j base;
    fe_from_u64(&base.x, 1, F);
    fe_from_u64(&base.y, 2, F);
    base.z = F->r1;
    for (int w = 0; w < 32; w++) {
        g1j acc;
        g1j_set_inf(&acc);
        memset(&table[w * 256], 0, sizeof(g1a));
        for (int d = 1; d < 256; d++) {
            g1j_add(&acc, &acc, &base);
            g1j_to_affine(&table[w * 256 + d], &acc);
        }
        for (int j = 0; j < 8; j++) g1j_double(&base, &base);
    }
    fe tau_raw, tau;
    memcpy(tau_raw.l, tau_std, 32);
    fe_to_mont(&tau, &tau_raw, &FR);
    g1a* pts = (g1a*)out;
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int tid = 0, nt = 1;
#endif
        size_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
        if (lo < hi) {
            u64 e[4] = {first + lo, 0, 0, 0};
            fe cur;
            fe_pow(&cur, &tau, e, &FR);
            for (size_t i = lo; i < hi; i++) {
                fe s;
                fe_from_mont(&s, &cur, &FR);
                g1j acc;
                g1j_set_inf(&acc);
                const uint8_t* sb = (const uint8_t*)s.l;
                for (int w = 0; w < 32; w++)
                    if (sb[w]) g1j_add_affine(&acc, &acc, &table[w * 256 + sb[w]]);
                g1j_to_affine(&pts[i], &acc);
                fe_mul(&cur, &cur, &tau, &FR);
            }
        }
    }
    free(table);
}

/* ------------------------------------------------------------------------------------------------ bulk Fr */
void ko_fr_to_mont(const uint8_t* in, uint8_t* out, size_t n, int threads) {   /* Fr.batchToMontgomery */
    if (threads <= 0) threads = ko_max_threads();
#pragma omp parallel for num_threads(threads)
    for (long i = 0; i < (long)n; i++) fe_to_mont((fe*)out + i, (const fe*)in + i, &FR);
}
void ko_fr_from_mont(const uint8_t* in, uint8_t* out, size_t n, int threads) { /* Fr.batchFromMontgomery */
    if (threads <= 0) threads = ko_max_threads();
#pragma omp parallel for num_threads(threads)
    for (long i = 0; i < (long)n; i++) fe_from_mont((fe*)out + i, (const fe*)in + i, &FR);
}

static void fr_root(fe* w, int k) {   /* Fr.w[k] = 5^((r-1)/2^k), Montgomery */
    u64 e[4];
    u64 rm1[4] = {FR.p[0] - 1, FR.p[1], FR.p[2], FR.p[3]};
    for (int i = 0; i < 4; i++) e[i] = (rm1[i] >> 28) | (i < 3 ? rm1[i + 1] << 36 : 0);
    fe five;
    fe_from_u64(&five, 5, &FR);
    fe_pow(w, &five, e, &FR);
    for (int i = 28; i > k; i--) fe_sqr(w, w, &FR);
}

/* Fr.fft / Fr.ifft: in place, natural order in and out */
static void ntt_inplace(fe* a, int lg, int inverse, int threads) {
    const size_t n = (size_t)1 << lg;
    if (n == 1) return;
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j |= bit;
        if (i < j) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    fe w;
    fr_root(&w, lg);
    if (inverse) fe_inv(&w, &w, &FR);
    fe* tw = (fe*)malloc(sizeof(fe) * (n / 2));
    tw[0] = FR.r1;
    /* tw[k] = w^k, filled in parallel blocks */
    {
        const size_t half = n / 2, blk = 4096;
        const long nblk = (long)((half + blk - 1) / blk);
#pragma omp parallel for num_threads(threads)
        for (long b = 0; b < nblk; b++) {
            size_t lo = (size_t)b * blk, hi = lo + blk < half ? lo + blk : half;
            u64 e[4] = {lo, 0, 0, 0};
            fe cur;
            fe_pow(&cur, &w, e, &FR);
            for (size_t k = lo; k < hi; k++) { tw[k] = cur; fe_mul(&cur, &cur, &w, &FR); }
        }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, step = n / len;
        const long nbf = (long)(n / 2);
#pragma omp parallel for num_threads(threads) schedule(static)
        for (long t = 0; t < nbf; t++) {
            size_t s = ((size_t)t / half) * len, k = (size_t)t % half;
            fe u = a[s + k], v;
            fe_mul(&v, &a[s + k + half], &tw[k * step], &FR);
            fe_add(&a[s + k], &u, &v, &FR);
            fe_sub(&a[s + k + half], &u, &v, &FR);
        }
    }
    free(tw);
    if (inverse) {
        fe ninv;
        fe_from_u64(&ninv, (u64)n, &FR);
        fe_inv(&ninv, &ninv, &FR);
#pragma omp parallel for num_threads(threads)
        for (long i = 0; i < (long)n; i++) fe_mul(&a[i], &a[i], &ninv, &FR);
    }
}
void ko_fr_ntt(uint8_t* data, int log_n, int inverse, int threads) {
    if (threads <= 0) threads = ko_max_threads();
    ntt_inplace((fe*)data, log_n, inverse, threads);
}

/* Fr.batchInverse: the input is sliced across the workers, Montgomery trick per slice, zeros skipped */
static void batch_inverse(fe* v, size_t n, int threads) {
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int tid = 0, nt = 1;
#endif
        size_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
        if (lo < hi) {
            fe* pref = (fe*)malloc(sizeof(fe) * (hi - lo));
            fe acc = FR.r1;
            for (size_t i = lo; i < hi; i++) {
                pref[i - lo] = acc;
                if (!fe_is_zero(&v[i])) fe_mul(&acc, &acc, &v[i], &FR);
            }
            fe inv;
            fe_inv(&inv, &acc, &FR);
            for (size_t i = hi; i-- > lo;) {
                if (fe_is_zero(&v[i])) continue;
                fe t;
                fe_mul(&t, &inv, &pref[i - lo], &FR);
                fe_mul(&inv, &inv, &v[i], &FR);
                v[i] = t;
            }
            free(pref);
        }
    }
}
void ko_fr_batch_inverse(uint8_t* data, size_t n, int threads) {
    if (threads <= 0) threads = ko_max_threads();
    batch_inverse((fe*)data, n, threads);
}

/* ------------------------------------------------------------------------------------------------ Keccak-256 */
static inline u64 rol64(u64 x, unsigned n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static void keccak_f(u64 a[25]) {
    static const u64 RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
        0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    static const unsigned ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int r = 0; r < 24; r++) {
        u64 c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], ROT[x + 5 * y]);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[r];
    }
}
void ko_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {   /* js-sha3 keccak256: pad byte 0x01 */
    u64 a[25];
    memset(a, 0, sizeof(a));
    while (len >= 136) {
        for (int i = 0; i < 17; i++) { u64 w; memcpy(&w, data + 8 * i, 8); a[i] ^= w; }
        keccak_f(a);
        data += 136;
        len -= 136;
    }
    uint8_t blk[136];
    memset(blk, 0, 136);
    if (len) memcpy(blk, data, len);
    blk[len] ^= 0x01;
    blk[135] ^= 0x80;
    for (int i = 0; i < 17; i++) { u64 w; memcpy(&w, blk + 8 * i, 8); a[i] ^= w; }
    keccak_f(a);
    memcpy(out, a, 32);
}

/* ------------------------------------------------------------------------------------------------ transcript
 * src/Keccak256Transcript.js:7-52: every getChallenge() hashes ALL data added so far. */
typedef struct { uint8_t* buf; size_t len, cap; } transcript;
static void tr_reserve(transcript* t, size_t extra) {
    if (t->len + extra > t->cap) {
        t->cap = (t->len + extra) * 2 + 256;
        t->buf = (uint8_t*)realloc(t->buf, t->cap);
    }
}
static void be32(const fe* std, uint8_t out[32]) {
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 8; b++) out[31 - (8 * i + b)] = (uint8_t)(std->l[i] >> (8 * b));
}
static void tr_add_g1(transcript* t, const g1a* p) {          /* addPolCommitment + G1.toRprUncompressed (:42) */
    tr_reserve(t, 64);
    if (g1a_is_inf(p)) {
        memset(t->buf + t->len, 0, 64);
        t->buf[t->len] |= 0x40;
    } else {
        fe x, y;
        fe_from_mont(&x, &p->x, &FQ);
        fe_from_mont(&y, &p->y, &FQ);
        be32(&x, t->buf + t->len);
        be32(&y, t->buf + t->len + 32);
    }
    t->len += 64;
}
static void tr_add_fr(transcript* t, const fe* s) {           /* addFieldElement + Fr.toRprBE (:45) */
    tr_reserve(t, 32);
    fe x;
    fe_from_mont(&x, s, &FR);
    be32(&x, t->buf + t->len);
    t->len += 32;
}
static void tr_challenge(const transcript* t, fe* out) {      /* :50-51: big-endian digest mod r -> Montgomery */
    uint8_t h[32];
    ko_keccak256(t->buf, t->len, h);
    u64 l[4];
    for (int i = 0; i < 4; i++) {
        l[i] = 0;
        for (int b = 0; b < 8; b++) l[i] |= (u64)h[31 - (8 * i + b)] << (8 * b);
    }
    while (geq_p(l, &FR)) sub_p(l, &FR);
    fe raw;
    memcpy(raw.l, l, 32);
    fe_to_mont(out, &raw, &FR);
}

/* ------------------------------------------------------------------------------------------------ Polynomial
 * src/polynomial/polynomial.js (used set).  Coefficients: Montgomery fe, length n. */
typedef struct { fe* c; size_t n; } poly;
static int g_threads = 1;

static poly p_zero(size_t n) {                                 /* :63-66 */
    poly p = {(fe*)calloc(n ? n : 1, sizeof(fe)), n};
    return p;
}
static poly p_clone(const poly* a) {                           /* :80-82 */
    poly p = {(fe*)malloc(sizeof(fe) * (a->n ? a->n : 1)), a->n};
    memcpy(p.c, a->c, sizeof(fe) * a->n);
    return p;
}
static void p_free(poly* p) { free(p->c); p->c = NULL; p->n = 0; }
static size_t p_degree(const poly* p) {                        /* :212-226 */
    for (size_t i = p->n; i-- > 1;)
        if (!fe_is_zero(&p->c[i])) return i;
    return 0;
}
static void p_evaluate(fe* r, const poly* p, const fe* x) {    /* :228-238 Horner */
    fe acc = {{0, 0, 0, 0}};
    if (p->n)
        for (size_t i = p_degree(p) + 1; i-- > 0;) {
            fe_mul(&acc, &acc, x, &FR);
            fe_add(&acc, &acc, &p->c[i], &FR);
        }
    *r = acc;
}
static void p_resize(poly* p, size_t n) {
    if (n > p->n) {
        p->c = (fe*)realloc(p->c, sizeof(fe) * n);
        memset(p->c + p->n, 0, sizeof(fe) * (n - p->n));
        p->n = n;
    }
}
static void p_addsub(poly* a, const poly* b, int sub) {        /* :276-350: the result takes the longer length */
    p_resize(a, b->n);
#pragma omp parallel for num_threads(g_threads)
    for (long i = 0; i < (long)b->n; i++) {
        if (sub) fe_sub(&a->c[i], &a->c[i], &b->c[i], &FR);
        else fe_add(&a->c[i], &a->c[i], &b->c[i], &FR);
    }
}
static void p_mul_scalar(poly* a, const fe* s) {               /* :395-406 */
#pragma omp parallel for num_threads(g_threads)
    for (long i = 0; i < (long)a->n; i++) fe_mul(&a->c[i], &a->c[i], s, &FR);
}
static void p_add_scalar(poly* a, const fe* s) { fe_add(&a->c[0], &a->c[0], s, &FR); }   /* :408-414 */
static void p_sub_scalar(poly* a, const fe* s) { fe_sub(&a->c[0], &a->c[0], s, &FR); }   /* :416-422 */
static int ceil_log2(size_t x) {
    int l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}
/* Evaluations.fromPolynomial(p, extension) (evaluations.js:12-21): zero-pad to nextpow2(len)*ext, fft */
static poly evals_from_poly(const poly* p, size_t ext) {
    size_t size = ((size_t)1 << ceil_log2(p->n)) * ext;
    poly e = p_zero(size);
    memcpy(e.c, p->c, sizeof(fe) * p->n);
    ntt_inplace(e.c, ceil_log2(size), 0, g_threads);
    return e;
}
static void p_multiply(poly* a, const poly* b) {               /* :352-376 (parity domain: full-degree operands) */
    size_t da = p_degree(a), db = p_degree(b);
    int new_power = ceil_log2(da + db + 1);
    size_t new_len = (size_t)1 << new_power;
    poly fa = p_zero(new_len), fb = p_zero(new_len);
    memcpy(fa.c, a->c, sizeof(fe) * (da + 1));
    memcpy(fb.c, b->c, sizeof(fe) * (db + 1));
    ntt_inplace(fa.c, new_power, 0, g_threads);
    ntt_inplace(fb.c, new_power, 0, g_threads);
#pragma omp parallel for num_threads(g_threads)
    for (long i = 0; i < (long)new_len; i++) fe_mul(&fa.c[i], &fa.c[i], &fb.c[i], &FR);
    ntt_inplace(fa.c, new_power, 1, g_threads);
    p_free(&fb);
    free(a->c);
    *a = fa;
}
static void p_shift_omega(poly* a) {                           /* :378-393: fft, rotate left by one, ifft */
    poly e = evals_from_poly(a, 1);
    fe first = e.c[0];
    memmove(e.c, e.c + 1, sizeof(fe) * (e.n - 1));
    e.c[e.n - 1] = first;
    ntt_inplace(e.c, ceil_log2(e.n), 1, g_threads);
    free(a->c);
    *a = e;
}
static int p_div_by_x_sub_value(poly* a, const fe* v) {        /* :814-851 */
    size_t n = a->n;
    poly q = p_zero(n);
    q.c[n - 2] = a->c[n - 1];
    for (size_t i = n - 2; i-- > 0;) {
        fe t;
        fe_mul(&t, v, &q.c[i + 1], &FR);
        fe_add(&q.c[i], &a->c[i + 1], &t, &FR);
    }
    fe nv, chk;
    fe_neg(&nv, v, &FR);
    fe_mul(&chk, &nv, &q.c[0], &FR);
    int ok = fe_eq(&a->c[0], &chk);
    free(a->c);
    *a = q;
    return ok ? 0 : -1;   /* "Polynomial does not divide" */
}
static int p_div_zh(poly* a, size_t n) {                       /* :853-888 */
    size_t deg = p_degree(a);
    size_t ext = a->n / n;
    size_t length = deg < n ? 0 : (size_t)1 << ceil_log2(deg + 1 - n);
    for (size_t i = 0; i < n; i++) fe_neg(&a->c[i], &a->c[i], &FR);
    for (size_t i = n; i < n * ext; i++) {
        fe_sub(&a->c[i], &a->c[i - n], &a->c[i], &FR);
        if (i + ext > n * (ext - 1) && !fe_is_zero(&a->c[i])) return -1;   /* i > n(ext-1) - ext: "not divisible" */
    }
    size_t d = p_degree(a);
    poly out = p_zero(length);
    if (length) memcpy(out.c, a->c, sizeof(fe) * ((d + 1) < length ? (d + 1) : length));
    free(a->c);
    *a = out;
    return 0;
}
static poly p_lagrange1(int power) {                           /* :68-78: ifft of e_0 */
    poly e = p_zero((size_t)1 << power);
    e.c[0] = FR.r1;
    ntt_inplace(e.c, power, 1, g_threads);
    return e;
}

/* commit(pol) = Polynomial.multiExponentiation (polynomial.js:1106-1115) + prover.js:432-434 */
static void commit(const uint8_t* srs, const poly* p, g1a* out) {
    size_t n = p->n ? p_degree(p) + 1 : 0;
    uint8_t* bm = (uint8_t*)malloc(32 * (n ? n : 1));
    ko_fr_from_mont((const uint8_t*)p->c, bm, n, g_threads);   /* Fr.batchFromMontgomery :1109 */
    ko_g1_msm(srs, bm, n, (uint8_t*)out, g_threads);          /* multiExpAffine + toAffine :1112-1113 */
    free(bm);
}

/* ------------------------------------------------------------------------------------------------ argument cores */
/* kind 0: ComputeSGrandSumPolynomial (grandsum.js:6-62); kind 1: ComputeZGrandProductPolynomial (grandproduct.js:6-57) */
static int grand_build(int kind, const fe* ev_f, const fe* ev_t, const fe* sel_f, const fe* sel_t, const fe* gamma,
                       size_t n, poly* out) {
    poly num = p_zero(n), den = p_zero(n);
    const fe one = FR.r1;
#pragma omp parallel for num_threads(g_threads)
    for (long ii = 0; ii < (long)n; ii++) {
        size_t i = (size_t)ii, j = (i + 1) % n;
        fe f, t, a, b;
        fe_add(&f, &ev_f[i], gamma, &FR);
        fe_add(&t, &ev_t[i], gamma, &FR);
        if (kind == 0) {                                       /* grandsum.js:21-38 */
            fe_mul(&a, &t, &sel_f[i], &FR);
            fe_mul(&b, &f, &sel_t[i], &FR);
            fe_sub(&num.c[j], &a, &b, &FR);
            fe_mul(&den.c[j], &f, &t, &FR);
        } else {                                               /* grandproduct.js:21-33 */
            fe_sub(&a, &f, &one, &FR);
            fe_mul(&a, &sel_f[i], &a, &FR);
            fe_add(&num.c[j], &a, &one, &FR);
            fe_sub(&b, &t, &one, &FR);
            fe_mul(&b, &sel_t[i], &b, &FR);
            fe_add(&den.c[j], &b, &one, &FR);
        }
    }
    batch_inverse(den.c, n, g_threads);                        /* grandsum.js:41 / grandproduct.js:36 */
    fe last;
    if (kind == 0) memset(&last, 0, sizeof(last)); else last = one;
    for (size_t i = 0; i < n; i++) {                           /* running sum / product :44-51 / :39-46 */
        size_t j = (i + 1) % n;
        fe t;
        fe_mul(&t, &num.c[j], &den.c[j], &FR);
        if (kind == 0) fe_add(&last, &t, &last, &FR); else fe_mul(&last, &t, &last, &FR);
        num.c[j] = last;
    }
    p_free(&den);
    int ok = kind == 0 ? fe_is_zero(&num.c[0]) : fe_eq(&num.c[0], &one);
    if (!ok) { p_free(&num); return -2; }                      /* "... is not well calculated" */
    ntt_inplace(num.c, ceil_log2(n), 1, g_threads);            /* Polynomial.fromEvaluations */
    *out = num;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ provers
 * kind 0: src/grandsum/mset_eq_kzg_prover.js:144-413; kind 1: src/grandproduct/mset_eq_kzg_prover.js:144-410.
 * cols_f / cols_t: k pointers to n x 32 B standard-form LE; sel_f / sel_t: n x 32 B Montgomery or NULL.
 * proof_out: commitments (64 B each) then evaluations (32 B each) in the reference's key order.
 * challenges_out: beta (zeros if k == 1), gamma, alpha, xi, v -- 5 x 32 B Montgomery.
 * returns 0, or -1 "does not divide", -2 "not well calculated", -3 "not divisible", -4 bad arguments */
int ko_prove(int kind, const uint8_t* srs, size_t srs_points, const uint8_t* const* cols_f, const uint8_t* const* cols_t,
             const uint8_t* sel_f_b, const uint8_t* sel_t_b, int nbits, int k, uint8_t* proof_out, uint8_t* challenges_out,
             int threads) {
    if (threads <= 0) threads = ko_max_threads();
    g_threads = threads;
    const size_t n = (size_t)1 << nbits;
    const int selected = sel_f_b != NULL && sel_t_b != NULL;
    const int gs = kind == 0;
    if (k < 1 || srs_points < 2 * n - 1) return -4;
    const fe one = FR.r1;
    int rc = 0;
    transcript tr = {NULL, 0, 0};
    g1a* cm = (g1a*)proof_out;
    size_t ncm = 0;

    /* ---- round 1 (:144-179) */
    poly* ev_f = (poly*)malloc(sizeof(poly) * k);
    poly* ev_t = (poly*)malloc(sizeof(poly) * k);
    poly* pf = (poly*)malloc(sizeof(poly) * k);
    poly* pt = (poly*)malloc(sizeof(poly) * k);
    for (int i = 0; i < k; i++) {
        ev_f[i] = p_zero(n);
        ev_t[i] = p_zero(n);
        ko_fr_to_mont(cols_f[i], (uint8_t*)ev_f[i].c, n, threads);     /* batchToMontgomery :147-148 */
        ko_fr_to_mont(cols_t[i], (uint8_t*)ev_t[i].c, n, threads);
        pf[i] = p_clone(&ev_f[i]);
        pt[i] = p_clone(&ev_t[i]);
        ntt_inplace(pf[i].c, nbits, 1, threads);                        /* fromEvaluations :151-152 */
        ntt_inplace(pt[i].c, nbits, 1, threads);
        commit(srs, &pf[i], &cm[ncm++]);                                /* :161-162 */
        commit(srs, &pt[i], &cm[ncm++]);
    }
    poly ev_sf = p_zero(n), ev_st = p_zero(n), sf = {NULL, 0}, st = {NULL, 0};
    if (selected) {
        memcpy(ev_sf.c, sel_f_b, 32 * n);
        memcpy(ev_st.c, sel_t_b, 32 * n);
        sf = p_clone(&ev_sf);
        st = p_clone(&ev_st);
        ntt_inplace(sf.c, nbits, 1, threads);                           /* :170-171 */
        ntt_inplace(st.c, nbits, 1, threads);
        commit(srs, &sf, &cm[ncm++]);                                   /* :173-174 */
        commit(srs, &st, &cm[ncm++]);
    } else {
        for (size_t i = 0; i < n; i++) ev_sf.c[i] = ev_st.c[i] = one;   /* getOneEvals :48-53 */
    }

    /* ---- round 2 (:181-231) */
    for (size_t i = 0; i < ncm; i++) tr_add_g1(&tr, &cm[i]);
    fe beta, gamma, alpha, xi, v;
    memset(&beta, 0, sizeof(beta));
    if (k > 1) {
        tr_challenge(&tr, &beta);
        tr_add_fr(&tr, &beta);
    }
    tr_challenge(&tr, &gamma);
    poly F, T, evF, evT;
    if (k > 1) {                                                        /* :207-217 */
        F = p_zero(n);
        T = p_zero(n);
        for (int i = k - 1; i >= 0; i--) {
            p_mul_scalar(&F, &beta);
            p_addsub(&F, &pf[i], 0);
            p_mul_scalar(&T, &beta);
            p_addsub(&T, &pt[i], 0);
        }
        evF = evals_from_poly(&F, 1);
        evT = evals_from_poly(&T, 1);
    } else {
        F = p_clone(&pf[0]);
        T = p_clone(&pt[0]);
        evF = p_clone(&ev_f[0]);
        evT = p_clone(&ev_t[0]);
    }
    poly A;   /* S or Z */
    rc = grand_build(kind, evF.c, evT.c, ev_sf.c, ev_st.c, &gamma, n, &A);
    if (rc) goto done_early;
    g1a cA;
    commit(srs, &A, &cA);                                               /* :229 */
    cm[ncm++] = cA;

    /* ---- round 3 (:233-286) */
    tr_add_fr(&tr, &gamma);
    tr_add_g1(&tr, &cA);
    tr_challenge(&tr, &alpha);
    poly Q = p_zero(n);
    if (selected) {                                                     /* :240-250 */
        poly b1 = p_clone(&st), tmp = p_clone(&st);
        p_multiply(&b1, &st);
        p_addsub(&tmp, &b1, 1);
        p_addsub(&Q, &tmp, 0);
        p_mul_scalar(&Q, &alpha);
        p_free(&b1); p_free(&tmp);
        b1 = p_clone(&sf); tmp = p_clone(&sf);
        p_multiply(&b1, &sf);
        p_addsub(&tmp, &b1, 1);
        p_addsub(&Q, &tmp, 0);
        p_mul_scalar(&Q, &alpha);
        p_free(&b1); p_free(&tmp);
    }
    {
        poly q1 = p_clone(&A);
        p_shift_omega(&q1);                                             /* :252 */
        poly fg = p_clone(&F), tg = p_clone(&T);
        p_add_scalar(&fg, &gamma);
        p_add_scalar(&tg, &gamma);
        if (gs) {
            p_addsub(&q1, &A, 1);                                       /* :253-254 */
            p_multiply(&q1, &fg);                                       /* :259 */
            p_multiply(&q1, &tg);                                       /* :260 */
            if (selected) {                                             /* :262-268 */
                poly a = p_clone(&sf), b = p_clone(&st);
                p_multiply(&a, &tg);
                p_multiply(&b, &fg);
                p_addsub(&q1, &b, 0);
                p_addsub(&q1, &a, 1);
                p_free(&a); p_free(&b);
            } else {                                                    /* :270-272 */
                p_addsub(&q1, &F, 0);
                p_addsub(&q1, &T, 1);
            }
        } else {                                                        /* grandproduct/...prover.js:252-276 */
            poly q2 = p_clone(&A);
            if (selected) {
                p_sub_scalar(&tg, &one);
                p_multiply(&tg, &st);
                p_add_scalar(&tg, &one);
                p_sub_scalar(&fg, &one);
                p_multiply(&fg, &sf);
                p_add_scalar(&fg, &one);
            }
            p_multiply(&q1, &tg);
            p_multiply(&q2, &fg);
            p_addsub(&q1, &q2, 1);
            p_free(&q2);
        }
        p_addsub(&Q, &q1, 0);                                           /* :275 */
        p_mul_scalar(&Q, &alpha);
        p_free(&q1); p_free(&fg); p_free(&tg);
        poly l1 = p_lagrange1(nbits);                                   /* :277-280 */
        poly q3 = p_clone(&A);
        if (!gs) p_sub_scalar(&q3, &one);
        p_multiply(&q3, &l1);
        p_addsub(&Q, &q3, 0);
        p_free(&q3); p_free(&l1);
    }
    {   /* the reference divides a buffer of whole extensions */
        size_t ext = (Q.n + n - 1) / n;
        if (ext < 2) ext = 2;
        p_resize(&Q, ext * n);
    }
    if (p_div_zh(&Q, n)) { rc = -3; goto done_mid; }                    /* :282 */
    g1a cQ;
    commit(srs, &Q, &cQ);                                               /* :284 */
    cm[ncm++] = cQ;

    /* ---- round 4 (:288-318) */
    tr_add_fr(&tr, &alpha);
    tr_add_g1(&tr, &cQ);
    tr_challenge(&tr, &xi);
    fe w, xiw;
    fr_root(&w, nbits);
    fe_mul(&xiw, &xi, &w, &FR);
    fe* ev = (fe*)(proof_out + 64 * (ncm + 2));                         /* evaluations follow the 2 W commitments */
    size_t nev = 0;
    fe* fbar = (fe*)malloc(sizeof(fe) * k);
    fe* tbar = (fe*)malloc(sizeof(fe) * k);
    for (int i = 0; i < k; i++) {
        p_evaluate(&fbar[i], &pf[i], &xi);
        ev[nev++] = fbar[i];
        if (gs) {
            p_evaluate(&tbar[i], &pt[i], &xi);
            ev[nev++] = tbar[i];
        }
    }
    fe sfx = one, stx = one, axw;
    if (selected) {
        p_evaluate(&sfx, &sf, &xi);
        p_evaluate(&stx, &st, &xi);
        ev[nev++] = sfx;
        ev[nev++] = stx;
    }
    p_evaluate(&axw, &A, &xiw);
    ev[nev++] = axw;

    /* ---- round 5 (:320-413) */
    tr_add_fr(&tr, &xi);
    for (size_t i = 0; i < nev; i++) tr_add_fr(&tr, &ev[i]);
    tr_challenge(&tr, &v);
    fe zh = xi, l1x, t0, t1;
    for (int i = 0; i < nbits; i++) fe_sqr(&zh, &zh, &FR);              /* polynomial_utils.js:1-10 */
    fe_sub(&zh, &zh, &one, &FR);
    fe_from_u64(&t0, (u64)n, &FR);                                      /* :12-19 */
    fe_sub(&t1, &xi, &one, &FR);
    fe_mul(&t0, &t0, &t1, &FR);
    fe_inv(&t0, &t0, &FR);
    fe_mul(&l1x, &zh, &t0, &FR);

    poly Rp = p_zero(n);
    if (selected) {                                                     /* :347-356 */
        fe a;
        fe_sqr(&a, &stx, &FR); fe_sub(&a, &stx, &a, &FR);
        p_add_scalar(&Rp, &a); p_mul_scalar(&Rp, &alpha);
        fe_sqr(&a, &sfx, &FR); fe_sub(&a, &sfx, &a, &FR);
        p_add_scalar(&Rp, &a); p_mul_scalar(&Rp, &alpha);
    }
    fe fx, fxg;
    p_evaluate(&fx, &F, &xi);                                           /* :358 */
    fe_add(&fxg, &fx, &gamma, &FR);
    if (gs) {
        fe tx, txg, a, b;
        p_evaluate(&tx, &T, &xi);                                       /* :359 */
        fe_add(&txg, &tx, &gamma, &FR);
        poly r1 = p_clone(&A);                                          /* :361-375 */
        fe m1;
        fe_neg(&m1, &one, &FR);
        p_mul_scalar(&r1, &m1);
        p_add_scalar(&r1, &axw);
        p_mul_scalar(&r1, &fxg);
        p_mul_scalar(&r1, &txg);
        if (selected) {
            fe_mul(&a, &stx, &fxg, &FR);
            fe_mul(&b, &sfx, &txg, &FR);
            p_add_scalar(&r1, &a);
            p_sub_scalar(&r1, &b);
        } else {
            p_add_scalar(&r1, &fx);
            p_sub_scalar(&r1, &tx);
        }
        p_addsub(&Rp, &r1, 0);
        p_mul_scalar(&Rp, &alpha);
        p_free(&r1);
        poly sl = p_clone(&A);                                          /* :378-383 */
        p_mul_scalar(&sl, &l1x);
        p_addsub(&Rp, &sl, 0);
        p_free(&sl);
    } else {                                                            /* grandproduct/...prover.js:341-386 */
        poly r1 = p_zero(n);
        poly txg = p_clone(&T);
        p_add_scalar(&txg, &gamma);                                     /* :354 */
        fe fg = fxg;
        poly zs = p_clone(&A);
        if (selected) {
            fe_sub(&fg, &fg, &one, &FR);
            p_sub_scalar(&txg, &one);
            fe self_g;
            fe_mul(&self_g, &sfx, &fg, &FR);
            fe_add(&self_g, &self_g, &one, &FR);
            p_mul_scalar(&txg, &stx);
            p_add_scalar(&txg, &one);
            p_mul_scalar(&txg, &axw);
            p_addsub(&r1, &txg, 0);
            p_mul_scalar(&zs, &self_g);
            p_addsub(&r1, &zs, 1);
        } else {
            p_mul_scalar(&txg, &axw);
            p_addsub(&r1, &txg, 0);
            p_mul_scalar(&zs, &fg);
            p_addsub(&r1, &zs, 1);
        }
        p_addsub(&Rp, &r1, 0);
        p_mul_scalar(&Rp, &alpha);
        p_free(&r1); p_free(&txg); p_free(&zs);
        poly zl = p_clone(&A);
        p_sub_scalar(&zl, &one);
        p_mul_scalar(&zl, &l1x);
        p_addsub(&Rp, &zl, 0);
        p_free(&zl);
    }
    {
        poly qz = p_clone(&Q);
        p_mul_scalar(&qz, &zh);
        p_addsub(&Rp, &qz, 1);
        p_free(&qz);
    }
    poly W = p_zero(n);                                                 /* :386-403 */
    if (selected) {
        poly a = p_clone(&st);
        p_sub_scalar(&a, &stx);
        p_addsub(&W, &a, 0);
        p_free(&a);
        p_mul_scalar(&W, &v);
        a = p_clone(&sf);
        p_sub_scalar(&a, &sfx);
        p_addsub(&W, &a, 0);
        p_free(&a);
    }
    if (gs)
        for (int i = k - 1; i >= 0; i--) {
            poly a = p_clone(&pt[i]);
            p_sub_scalar(&a, &tbar[i]);
            p_mul_scalar(&W, &v);
            p_addsub(&W, &a, 0);
            p_free(&a);
        }
    for (int i = k - 1; i >= 0; i--) {
        poly a = p_clone(&pf[i]);
        p_sub_scalar(&a, &fbar[i]);
        p_mul_scalar(&W, &v);
        p_addsub(&W, &a, 0);
        p_free(&a);
    }
    p_mul_scalar(&W, &v);
    p_addsub(&W, &Rp, 0);
    if (p_div_by_x_sub_value(&W, &xi)) rc = -1;
    poly Ww = p_clone(&A);                                              /* :406-407 */
    p_sub_scalar(&Ww, &axw);
    if (p_div_by_x_sub_value(&Ww, &xiw)) rc = -1;
    g1a c1, c2;
    commit(srs, &W, &c1);                                               /* :409-410 */
    commit(srs, &Ww, &c2);
    cm[ncm++] = c1;
    cm[ncm++] = c2;
    if (challenges_out) {
        memcpy(challenges_out, &beta, 32);
        memcpy(challenges_out + 32, &gamma, 32);
        memcpy(challenges_out + 64, &alpha, 32);
        memcpy(challenges_out + 96, &xi, 32);
        memcpy(challenges_out + 128, &v, 32);
    }
    p_free(&Rp); p_free(&W); p_free(&Ww);
    free(fbar); free(tbar);
done_mid:
    p_free(&Q);
    p_free(&A);
done_early:
    for (int i = 0; i < k; i++) { p_free(&ev_f[i]); p_free(&ev_t[i]); p_free(&pf[i]); p_free(&pt[i]); }
    free(ev_f); free(ev_t); free(pf); free(pt);
    p_free(&ev_sf); p_free(&ev_st);
    if (sf.c) p_free(&sf);
    if (st.c) p_free(&st);
    p_free(&F); p_free(&T); p_free(&evF); p_free(&evT);
    free(tr.buf);
    return rc;
}
