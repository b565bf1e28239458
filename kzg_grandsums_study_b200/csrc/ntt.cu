// Radix-2^k Fr NTT / iNTT for sm_100a.  Replaces Fr.fft / Fr.ifft of ffjavascript at the reference call
// sites src/polynomial/polynomial.js:34,373,392 and src/polynomial/evaluations.js:18.
//
// Convention: out[k] = sum_j in[j] * w^(jk), w = Fr.w[log2 N] = 5^((r-1)/N), natural order in
// and out; the inverse uses w^-1 and scales by N^-1.
//
// Structure: log2 N is split into P <= 4 passes of at most 8 bits.  With the input index written as
// n = sum_i n_i S_i (S_i = prod_{j>i} R_j) and the output index as k = sum_i k_i prod_{j<i} R_j, pass i
// does R_i-point DFTs over digit n_i for a tile of T consecutive columns, entirely in shared memory
// (the tile's rows -- T * 32 B = 256 contiguous bytes each -- staged by TMA bulk copies; DIF butterflies two stages
// at a time in registers, 128-bit shared-memory accesses), multiplies by the inter-pass twiddle
// w_{N_i}^{k_i * column} and writes back; the last pass works on contiguous rows and scatters T-wide contiguous
// runs into natural order.  Every element therefore crosses HBM once per pass (P reads + P writes); the arithmetic
// (~3 products per element and pass in the butterflies, 1 at every pass boundary; N^-1 of an inverse transform rides
// on the last boundary's table) is what bounds it on B200 -- see DESIGN.md.
//
// Twiddles come from two 8192-entry tables per direction, W = w_{2^26}: lo[i] = W^i, hi[j] = W^(8192 j): every
// twiddle of a transform of size <= 2^26 is hi[.] or hi[.] * lo[.]; plus mid[e] = w_{2^16}^e (2 MB) so that every
// boundary below 2^16 costs one product.  A zero-padded input (Evaluations.fromPolynomial with extension > 1) is
// handled by n_in < N: the loader substitutes zeros.
#include <string.h>

#include "common.cuh"

namespace kzg {

constexpr int NTT_THREADS = 256;
constexpr uint32_t NTT_TILE_COLS = 8;

struct NttPlan {
    uint32_t log_n;
    uint32_t npass;
    uint32_t radbits[4];
};

__device__ __forceinline__ Fr tw_lookup(const Fr* __restrict__ lo, const Fr* __restrict__ hi, uint32_t e) {
    // W^e, e < 2^26
    Fr h = fp_load<FrP>(hi + (e >> TW_BITS));
    uint32_t l = e & (TW_SIZE - 1);
    if (l == 0) return h;
    return fp_mul(h, fp_load<FrP>(lo + l));
}

__device__ __forceinline__ uint32_t bitrev(uint32_t x, uint32_t bits) {
    return bits == 0 ? 0 : (__brev(x) >> (32 - bits));
}

// In-shared-memory R-point DIF DFTs on `cols` independent columns; element (n, c) at sh[n * cols + c].
// Output is left in bit-reversed row order.
// Two radix-2 stages per shared-memory round trip: a thread holds the four elements {j, j + h/2, j + h, j + 3h/2} of a
// block of 2h rows in registers and runs the stage of half-size h and the stage of half-size h/2 on them (a radix-4
// butterfly; in a prime field it costs the same four products as the two radix-2 stages -- w_4 is an ordinary
// element -- but half the loads, stores and barriers).  Butterflies are numbered column-fastest, then group, then j, so
// that from the second round trip on j is uniform across a warp and the trivial twiddles (j = 0) are skipped by whole
// warps instead of being executed under a mask.
// `tw` = the R/2 twiddles w_R^i in shared memory (smem_twiddles): every stage's twiddles are powers of w_R, and a
// warp asks for at most four distinct ones per access (broadcasts), so none of the butterflies waits for the L2.
__device__ __forceinline__ void smem_twiddles(Fr* tw, uint32_t rbits, const Fr* __restrict__ tw_hi) {
    const uint32_t half = (1u << rbits) >> 1;
    const uint32_t stride = TW_SIZE >> rbits;  // w_R^i = W^(i 2^26 / R) = hi[i * 8192 / R]
    for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) fp_store(tw + i, fp_load<FrP>(tw_hi + i * stride));
}
__device__ __forceinline__ void smem_dif(Fr* sh, uint32_t rbits, uint32_t cols, const Fr* tw) {
    const uint32_t R = 1u << rbits;
    uint32_t h = R >> 1;
    while (h >= 2) {
        const uint32_t q = h >> 1;                 // j < q
        const uint32_t G = R / (2 * h);            // groups of 2h rows
        const uint32_t nb = cols * G * q;          // R * cols / 4 butterflies
        const uint32_t step = R / (2 * h);         // w_{2h}^j = w_R^(j R / 2h)
        for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
            const uint32_t c = b % cols;
            const uint32_t rest = b / cols;
            const uint32_t g = rest % G;
            const uint32_t j = rest / G;
            Fr* p0 = sh + (g * 2 * h + j) * cols + c;
            Fr* p1 = p0 + q * cols;
            Fr* p2 = p0 + h * cols;
            Fr* p3 = p2 + q * cols;
            const Fr a0 = fp_load<FrP>(p0), a1 = fp_load<FrP>(p1), a2 = fp_load<FrP>(p2), a3 = fp_load<FrP>(p3);
            const Fr s0 = fp_add(a0, a2), s1 = fp_add(a1, a3);
            Fr t0 = fp_sub(a0, a2), t1 = fp_sub(a1, a3);
            t1 = fp_mul(t1, fp_load<FrP>(tw + ((j + q) * step)));  // w_{2h}^(j + h/2)
            if (j != 0) t0 = fp_mul(t0, fp_load<FrP>(tw + j * step));
            Fr u1 = fp_sub(s0, s1), u3 = fp_sub(t0, t1);
            if (j != 0) {
                const Fr w2 = fp_load<FrP>(tw + 2 * j * step);     // w_h^j
                u1 = fp_mul(u1, w2);
                u3 = fp_mul(u3, w2);
            }
            fp_store(p0, fp_add(s0, s1));
            fp_store(p1, u1);
            fp_store(p2, fp_add(t0, t1));
            fp_store(p3, u3);
        }
        __syncthreads();
        h >>= 2;
    }
    if (h == 1) {  // odd number of stages: one twiddle-free radix-2 stage is left
        const uint32_t nb = cols * (R >> 1);
        for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
            const uint32_t c = b % cols, g = b / cols;
            Fr* p0 = sh + (2 * g) * cols + c;
            Fr* p1 = p0 + cols;
            const Fr a = fp_load<FrP>(p0), d = fp_load<FrP>(p1);
            fp_store(p0, fp_add(a, d));
            fp_store(p1, fp_sub(a, d));
        }
        __syncthreads();
    }
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP): a tile row is 256 contiguous bytes of global memory ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// strided (non-last) pass.  grid = N / (R * T).
// The tile -- R rows of T * 32 = 256 contiguous bytes, S elements apart -- is staged in shared memory by the TMA unit:
// thread n issues ONE bulk copy for row n and the block waits on an mbarrier for the R * 256 bytes; rows beyond the
// valid input (zero-padded transforms, Evaluations.fromPolynomial with an extension) are zero-filled by the threads.
// Inter-pass twiddle w_{N_i}^(k * column), N_i = R * S -- ONE product per element whenever a direct table covers N_i:
//   tw_mode 0: table[e << tw_shift]: `mid` (w_{2^16}^e, 2 MB, L2-resident) for N_i <= 2^16 -- for an inverse transform the
//              copy of `mid` scaled by N^-1 on the LAST strided pass, so that the scaling costs nothing --, `big`
//              (w_{2^24}^e, 512 MB of HBM per direction, built on first use) for 2^17 <= N_i <= 2^24: one random 32-byte
//              gather per element (ld.global.nc.L2::64B: 64 bytes of DRAM traffic each) -- the pass is bound by the
//              integer pipe, the memory system has the room (an L2 prefetch of the tile's twiddles before the butterflies
//              was measured: 128-byte lines fetched, partly twice, and 3.59 instead of 3.53 ms);
//   tw_mode 1: the hi x lo composite (two products), transforms above 2^24 only.
// `always`: the table carries a scale factor, so e = 0 is not a shortcut.
__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_strided_pass_kernel(const Fr* __restrict__ src, uint64_t n_in,
                                                                       Fr* __restrict__ dst, uint32_t rbits, uint32_t T,
                                                                       uint32_t log_stride, uint32_t tw_mode, uint32_t tw_shift,
                                                                       uint32_t always, uint32_t big_table,
                                                                       const Fr* __restrict__ tw_lo,
                                                                       const Fr* __restrict__ tw_hi,
                                                                       const Fr* __restrict__ tw_table) {
    extern __shared__ uint4 smem_raw[];
    Fr* sh = reinterpret_cast<Fr*>(smem_raw);
    __shared__ __align__(8) uint64_t bar;
    const uint32_t R = 1u << rbits;
    const uint64_t S = 1ull << log_stride;
    const uint64_t tiles_per_group = S / T;  // column tiles inside one (hi) group
    const uint64_t grp = blockIdx.x / tiles_per_group;
    const uint64_t c0 = (blockIdx.x % tiles_per_group) * T;
    const uint64_t base = grp * (S << rbits) + c0;
    // rows whose T elements are all valid come through the TMA unit
    uint32_t full_rows = 0;
    if (n_in >= base + T) {
        const uint64_t r = (n_in - base - T) / S + 1;
        full_rows = r < R ? (uint32_t)r : R;
    }
    Fr* tw = sh + R * T;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, full_rows * T * (uint32_t)sizeof(Fr));
    }
    __syncthreads();
    for (uint32_t n = threadIdx.x; n < R; n += blockDim.x) {
        if (n < full_rows) {
            bulk_load(sh + n * T, src + base + (uint64_t)n * S, T * (uint32_t)sizeof(Fr), &bar);
        } else {
            for (uint32_t c = 0; c < T; c++) {
                const uint64_t gi = base + (uint64_t)n * S + c;
                fp_store(sh + n * T + c, gi < n_in ? fp_load<FrP>(src + gi) : fp_zero<FrP>());
            }
        }
    }
    smem_twiddles(tw, rbits, tw_hi);
    const uint32_t log_ni = rbits + log_stride;
    mbar_wait(&bar, 0);
    __syncthreads();
    smem_dif(sh, rbits, T, tw);
    // twiddle w_{N_i}^{k * col}, N_i = R * S
    const uint32_t shift = NTT_MAX_LOG - log_ni;  // composite mode: exponent in units of W = w_{2^26}
#pragma unroll 2
    for (uint32_t idx = threadIdx.x; idx < R * T; idx += blockDim.x) {
        uint32_t c = idx % T, p = idx / T;
        uint32_t k = bitrev(p, rbits);
        Fr v = fp_load<FrP>(sh + idx);
        uint64_t col = c0 + c;
        uint64_t e = ((uint64_t)k * col) & ((1ull << log_ni) - 1);
        if (always || e != 0) {
            if (tw_mode == 0) {
                Fr w;
                if (big_table) {  // the big table: one 32-byte record of a 512 MB array -- ask DRAM for 64 bytes, not the 128-byte line
                    const uint4* q = reinterpret_cast<const uint4*>(tw_table + (e << tw_shift));
                    uint4 a, b;
                    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(q));
                    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(q + 1));
                    w.l[0] = a.x; w.l[1] = a.y; w.l[2] = a.z; w.l[3] = a.w;
                    w.l[4] = b.x; w.l[5] = b.y; w.l[6] = b.z; w.l[7] = b.w;
                } else {
                    w = fp_load<FrP>(tw_table + (e << tw_shift));
                }
                v = fp_mul(v, w);
            }
            else v = fp_mul(v, tw_lookup(tw_lo, tw_hi, (uint32_t)(e << shift)));
        }
        fp_store(dst + base + (uint64_t)k * S + c, v);
    }
}

// last pass: rows of R contiguous elements; a tile is T rows that are consecutive in k_1.
// grid = N / (R * T)   (T = 1 when npass == 1)
__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_last_pass_kernel(const Fr* __restrict__ src, uint64_t n_in,
                                                                    Fr* __restrict__ dst, NttPlan plan, uint32_t T,
                                                                    const Fr* __restrict__ tw_hi, Fr scale,
                                                                    bool do_scale) {
    extern __shared__ uint4 smem_raw[];
    Fr* sh = reinterpret_cast<Fr*>(smem_raw);
    const uint32_t P = plan.npass;
    const uint32_t rbits = plan.radbits[P - 1];
    const uint32_t R = 1u << rbits;
    // mid rows per k_1 value: M = prod_{1<j<P} R_j
    uint32_t mid_bits = 0;
    for (uint32_t j = 1; j + 1 < P; j++) mid_bits += plan.radbits[j];
    const uint64_t M = 1ull << mid_bits;
    const uint64_t mid = blockIdx.x % M;
    const uint64_t k1_0 = (blockIdx.x / M) * T;
    const uint32_t log_s1 = plan.log_n - plan.radbits[0];  // S_1 = N / R_1   (only used when P > 1)
    // (element index fastest: the warp reads 1 KiB of one contiguous row.  Row index fastest -- conflict-free shared
    // stores, 128-byte global segments -- was measured: 47 M instead of 67 M bank conflicts, but 1.086 vs 1.046 ms)
    for (uint32_t idx = threadIdx.x; idx < R * T; idx += blockDim.x) {
        uint32_t n = idx % R, r = idx / R;
        uint64_t gi = (P > 1 ? ((k1_0 + r) << log_s1) : 0) + (mid << rbits) + n;
        Fr v = gi < n_in ? fp_load<FrP>(src + gi) : fp_zero<FrP>();
        fp_store(sh + n * T + r, v);
    }
    Fr* tw = sh + R * T;
    smem_twiddles(tw, rbits, tw_hi);
    __syncthreads();
    smem_dif(sh, rbits, T, tw);
    // output index = k_1 + R_1 * (digit-reversed mid) + k_P * (N / R_P)
    uint64_t out_mid = 0;
    if (P > 2) {
        // storage order of mid: k_2 most significant ... k_{P-1} least; output order: k_2 least significant
        uint64_t rest = mid;
        for (uint32_t j = P - 2; j >= 1; j--) {
            uint32_t kb = plan.radbits[j];
            uint64_t kj = rest & ((1ull << kb) - 1);
            rest >>= kb;
            // digit j has output weight prod_{1<m<j} R_m  (relative to R_1)
            uint32_t wbits = 0;
            for (uint32_t m = 1; m < j; m++) wbits += plan.radbits[m];
            out_mid += kj << wbits;
            if (j == 1) break;
        }
    }
    const uint32_t log_hi = plan.log_n - rbits;  // N / R_P
    for (uint32_t idx = threadIdx.x; idx < R * T; idx += blockDim.x) {
        uint32_t r = idx % T, p = idx / T;
        uint32_t k = bitrev(p, rbits);
        Fr v = fp_load<FrP>(sh + p * T + r);
        if (do_scale) v = fp_mul(v, scale);
        uint64_t oi = ((uint64_t)k << log_hi) + (P > 1 ? (k1_0 + r + (out_mid << plan.radbits[0])) : 0);
        fp_store(dst + oi, v);
    }
}

// ---- twiddle tables -----------------------------------------------------------------------------
__global__ void tw_table_kernel(Fr* __restrict__ lo, Fr* __restrict__ hi, Fr w, Fr w_hi) {
    // lo[i] = w^i, hi[i] = w_hi^i ; one thread per entry (fp_pow_u64: <= 13 squarings)
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TW_SIZE) return;
    fp_store(lo + i, fp_pow_u64(w, i));
    fp_store(hi + i, fp_pow_u64(w_hi, i));
}

Fr fr_from_bytes(const uint8_t b[32]) {
    Fr r;
    memcpy(r.l, b, 32);
    return r;
}
void fr_to_bytes(const Fr& a, uint8_t b[32]) { memcpy(b, a.l, 32); }

static Fr fr_from_u64_host(uint64_t v) {
    Fr r = fp_zero<FrP>();
    r.l[0] = (uint32_t)v;
    r.l[1] = (uint32_t)(v >> 32);
    return fp_to_mont(r);
}

// Fr.w[k] = 5^((r-1)/2^k) in Montgomery form (host)
Fr fr_root_of_unity(uint32_t log_n) {
    // (r - 1) / 2^28 as limbs: r - 1 = 2^28 * t
    static bool init = false;
    static Fr w28;
    if (!init) {
        uint32_t e[8];
        uint32_t rm1[8];
        for (int i = 0; i < 8; i++) rm1[i] = FrP::mod(i);
        rm1[0] -= 1;
        for (int i = 0; i < 8; i++) {
            uint64_t lo = rm1[i] >> 28;
            uint64_t hi = i + 1 < 8 ? ((uint64_t)rm1[i + 1] << 4) : 0;
            e[i] = (uint32_t)(lo | hi);
        }
        w28 = fp_pow(fr_from_u64_host(5), e);
        init = true;
    }
    Fr w = w28;
    for (uint32_t i = 28; i > log_n; i--) w = fp_sqr(w);
    return w;
}

// mid[e] = W^(e 2^10) = w_{2^16}^e: the inter-pass twiddles of every boundary with N_i <= 2^16 in ONE product
__global__ void tw_mid_kernel(Fr* __restrict__ mid, const Fr* __restrict__ lo, const Fr* __restrict__ hi) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (1u << 16)) return;
    const uint32_t x = e << 10;
    Fr v = fp_load<FrP>(hi + (x >> TW_BITS));
    const uint32_t l = x & (TW_SIZE - 1);
    if (l) v = fp_mul(v, fp_load<FrP>(lo + l));
    fp_store(mid + e, v);
}
__global__ void tw_scale_kernel(Fr* __restrict__ out, const Fr* __restrict__ in, Fr scale, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) fp_store(out + i, fp_mul(fp_load<FrP>(in + i), scale));
}
// big[e] = W^(4 e) = w_{2^24}^e, e < 2^24
__global__ void tw_big_kernel(Fr* __restrict__ big, const Fr* __restrict__ lo, const Fr* __restrict__ hi) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t x = e << 2;
    Fr v = fp_load<FrP>(hi + (x >> TW_BITS));
    const uint32_t l = x & (TW_SIZE - 1);
    if (l) v = fp_mul(v, fp_load<FrP>(lo + l));
    fp_store(big + e, v);
}

int ntt_init_tables(kzg_ctx* ctx) {
    for (int dir = 0; dir < 2; dir++) {
        KZG_CUDA(ctx, cudaMalloc((void**)&ctx->tw_lo[dir], sizeof(Fr) * TW_SIZE));
        KZG_CUDA(ctx, cudaMalloc((void**)&ctx->tw_hi[dir], sizeof(Fr) * TW_SIZE));
        KZG_CUDA(ctx, cudaMalloc((void**)&ctx->tw_mid[dir], sizeof(Fr) << 16));
        Fr w = fr_root_of_unity(NTT_MAX_LOG);
        if (dir == 1) w = fp_inv(w);
        Fr w_hi = fp_pow_u64(w, TW_SIZE);
        KZG_LAUNCH(ctx, tw_table_kernel, TW_SIZE / 256, 256, 0, ctx->tw_lo[dir], ctx->tw_hi[dir], w, w_hi);
        KZG_LAUNCH(ctx, tw_mid_kernel, (1u << 16) / 256, 256, 0, ctx->tw_mid[dir], ctx->tw_lo[dir], ctx->tw_hi[dir]);
    }
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

// inverse transforms of 2^log_n points: the `mid` table of the inverse direction times N^-1 (built on first use, 2 MB);
// the last strided pass of a transform always has N_i <= 2^16, so the scaling rides on it for free
static int ntt_scaled_mid(kzg_ctx* ctx, uint32_t log_n, const Fr** out) {
    if (!ctx->tw_mid_scaled[log_n]) {
        Fr* t = nullptr;
        KZG_CUDA(ctx, cudaMalloc((void**)&t, sizeof(Fr) << 16));
        const Fr scale = fp_inv(fr_from_u64_host(1ull << log_n));
        KZG_LAUNCH(ctx, tw_scale_kernel, (1u << 16) / 256, 256, 0, t, ctx->tw_mid[1], scale, 1u << 16);
        KZG_CHECK_LAUNCH(ctx);
        KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // one-off: the other lane's stream may use the table next
        ctx->tw_mid_scaled[log_n] = t;
    }
    *out = ctx->tw_mid_scaled[log_n];
    return KZG_OK;
}
// w_{2^24}^e for every e: the first boundary of a transform of 2^17 .. 2^24 points in ONE product (512 MB per direction:
// HBM capacity spent to delete a product per element, like the MSM's window table)
static int ntt_big_table(kzg_ctx* ctx, int dir, const Fr** out) {
    if (!ctx->tw_big[dir]) {
        Fr* t = nullptr;
        cudaError_t e = cudaMalloc((void**)&t, sizeof(Fr) << 24);
        if (e != cudaSuccess) {  // no room: the composite twiddles still work
            cudaGetLastError();
            *out = nullptr;
            return KZG_OK;
        }
        KZG_LAUNCH(ctx, tw_big_kernel, (1u << 24) / 256, 256, 0, t, ctx->tw_lo[dir], ctx->tw_hi[dir]);
        KZG_CHECK_LAUNCH(ctx);
        KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->tw_big[dir] = t;
    }
    *out = ctx->tw_big[dir];
    return KZG_OK;
}

static NttPlan make_plan(uint32_t log_n) {
    NttPlan p;
    p.log_n = log_n;
    p.npass = log_n <= 10 ? 1 : (log_n + 7) / 8;
    uint32_t base = log_n / p.npass, rem = log_n % p.npass;
    for (uint32_t i = 0; i < 4; i++) p.radbits[i] = 0;
    for (uint32_t i = 0; i < p.npass; i++) p.radbits[i] = base + (i < rem ? 1 : 0);
    return p;
}

// in: n_in valid elements (zero beyond), out: 2^log_n elements.  out may alias in.
int ntt_run(kzg_ctx* ctx, const Fr* in, uint64_t n_in, Fr* out, uint32_t log_n, bool inverse) {
    if (log_n > NTT_MAX_LOG) return set_err(ctx, KZG_ERR_ARG, "ntt: size above 2^26 not supported");
    const uint64_t N = 1ull << log_n;
    if (n_in > N) n_in = N;
    const int dir = inverse ? 1 : 0;
    NttPlan plan = make_plan(log_n);
    Fr scale = fp_one<FrP>();
    if (inverse) scale = fp_inv(fr_from_u64_host(N));
    if (!ctx->ntt_attr_set) {
        KZG_CUDA(ctx, cudaFuncSetAttribute(ntt_strided_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        KZG_CUDA(ctx, cudaFuncSetAttribute(ntt_last_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        ctx->ntt_attr_set = true;
    }
    if (plan.npass == 1) {
        const uint32_t R = 1u << log_n;
        KZG_LAUNCH(ctx, ntt_last_pass_kernel, 1, NTT_THREADS, sizeof(Fr) * (R + R / 2), in, n_in, out, plan, 1u, ctx->tw_hi[dir],
                   scale, inverse);
        KZG_CHECK_LAUNCH(ctx);
        return KZG_OK;
    }
    const Fr* mid_scaled = nullptr;
    if (inverse) KZG_TRY(ntt_scaled_mid(ctx, log_n, &mid_scaled));  // (the last strided pass applies N^-1 with its twiddles)
    const Fr* big = nullptr;
    // (measured: 3.53 vs 3.75 ms at 2^24, 0.428 vs 0.450 at 2^21, 0.240 vs 0.250 at 2^20, 0.146 vs 0.150 at 2^19, level at 2^18)
    if (log_n >= (ctx->tuning.ntt_big_table > 1 ? 17u : 19u) && log_n <= 24 && ctx->tuning.ntt_big_table)
        KZG_TRY(ntt_big_table(ctx, dir, &big));
    Fr* tmp = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&tmp, sizeof(Fr) * N, ctx->stream));
    // tile width T (columns = contiguous elements per row) and threads per block: one thread per two radix-4 butterflies
    const uint32_t T = ctx->tuning.ntt_tile == 4 ? 4u : NTT_TILE_COLS;
    uint32_t log_stride = log_n;
    const Fr* src = in;
    uint64_t src_n = n_in;
    timed_begin(ctx, KZG_TIMED_NTT);
    for (uint32_t i = 0; i + 1 < plan.npass; i++) {
        const uint32_t rb = plan.radbits[i];
        const uint32_t log_ni = log_stride;  // N_i = R_i * S_i = the stride before this pass
        log_stride -= rb;
        const uint32_t R = 1u << rb;
        const uint32_t grid = (uint32_t)(N / ((uint64_t)R * T));
        const bool last_strided = i + 2 == plan.npass;
        const uint32_t always = inverse && last_strided ? 1u : 0u;   // this pass's table carries N^-1
        uint32_t tw_mode = 1, tw_shift = 0, big_table = 0;
        const Fr* table = nullptr;
        if (log_ni <= 16) {
            tw_mode = 0;
            tw_shift = 16 - log_ni;
            table = always ? mid_scaled : ctx->tw_mid[dir];
        } else if (big && log_ni <= 24) {
            tw_mode = 0;
            tw_shift = 24 - log_ni;
            table = big;
            big_table = 1;
        }
        uint32_t threads = R * T / 8;
        if (threads > NTT_THREADS) threads = NTT_THREADS;
        if (threads < 32) threads = 32;
        KZG_LAUNCH(ctx, ntt_strided_pass_kernel, grid, threads, sizeof(Fr) * (R * T + R / 2), src, src_n, tmp, rb, T,
                   log_stride, tw_mode, tw_shift, always, big_table, ctx->tw_lo[dir], ctx->tw_hi[dir], table);
        src = tmp;
        src_n = N;
    }
    {
        const uint32_t rb = plan.radbits[plan.npass - 1];
        const uint32_t R = 1u << rb;
        const uint32_t grid = (uint32_t)(N / ((uint64_t)R * T));
        uint32_t threads = R * T / 8;
        if (threads > NTT_THREADS) threads = NTT_THREADS;
        if (threads < 32) threads = 32;
        KZG_LAUNCH(ctx, ntt_last_pass_kernel, grid, threads, sizeof(Fr) * (R * T + R / 2), tmp, N, out, plan, T, ctx->tw_hi[dir],
                   scale, false);
    }
    timed_end(ctx, KZG_TIMED_NTT);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaFreeAsync(tmp, ctx->stream));
    return KZG_OK;
}

}  // namespace kzg
