// Radix-2^k Fr NTT / iNTT for sm_100a.  Replaces Fr.fft / Fr.ifft of ffjavascript at the reference call
// sites src/polynomial/polynomial.js:34,373,392 and src/polynomial/evaluations.js:18.
//
// Convention: out[k] = sum_j in[j] * w^(jk), w = Fr.w[log2 N] = 5^((r-1)/N), natural order in
// and out; the inverse uses w^-1 and scales by N^-1.
//
// Structure: log2 N is split into P <= 4 passes of at most 8 bits.  With the input index written as
// n = sum_i n_i S_i (S_i = prod_{j>i} R_j) and the output index as k = sum_i k_i prod_{j<i} R_j, pass i
// does R_i-point DFTs over digit n_i for a tile of T consecutive columns, entirely in shared memory
// (radix-2 DIF stages, 128-bit loads/stores, T * 32 B = 256 B contiguous per row), multiplies by the
// inter-pass twiddle w_{N_i}^{k_i * column} and writes back in place; the last pass works on contiguous
// rows and scatters T-wide contiguous runs into natural order.  Every element therefore crosses HBM once
// per pass (P reads + P writes); the arithmetic (log2 N / 2 + 2P modmul per element) is what bounds it
// on B200 -- see DESIGN.md.
//
// Twiddles come from two 8192-entry tables per direction, W = w_{2^26}: lo[i] = W^i, hi[j] = W^(8192 j);
// every twiddle of a transform of size <= 2^26 is hi[.] or hi[.] * lo[.].  A zero-padded input
// (Evaluations.fromPolynomial with extension > 1) is handled by n_in < N: the loader substitutes zeros.
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
__device__ __forceinline__ void smem_dif(Fr* sh, uint32_t rbits, uint32_t cols, const Fr* __restrict__ tw_hi) {
    const uint32_t R = 1u << rbits;
    const uint32_t nbf = (R >> 1) * cols;
    for (uint32_t h = R >> 1; h >= 1; h >>= 1) {
        const uint32_t tw_step = TW_SIZE / (2 * h);  // w_{2h}^j = W^(j * 2^26 / 2h) = hi[j * 8192 / 2h]
        for (uint32_t b = threadIdx.x; b < nbf; b += NTT_THREADS) {
            uint32_t c = b % cols;
            uint32_t jp = b / cols;
            uint32_t j = jp & (h - 1);
            uint32_t i0 = ((jp - j) << 1) + j;
            uint32_t i1 = i0 + h;
            Fr a = fp_load<FrP>(sh + i0 * cols + c);
            Fr d = fp_load<FrP>(sh + i1 * cols + c);
            Fr s = fp_add(a, d);
            Fr t = fp_sub(a, d);
            if (j != 0) t = fp_mul(t, fp_load<FrP>(tw_hi + j * tw_step));
            fp_store(sh + i0 * cols + c, s);
            fp_store(sh + i1 * cols + c, t);
        }
        __syncthreads();
    }
}

// strided (non-last) pass.  grid = N / (R * T).
__global__ void __launch_bounds__(NTT_THREADS) ntt_strided_pass_kernel(const Fr* __restrict__ src, uint64_t n_in,
                                                                       Fr* __restrict__ dst, uint32_t rbits,
                                                                       uint32_t log_stride, uint32_t log_n,
                                                                       const Fr* __restrict__ tw_lo,
                                                                       const Fr* __restrict__ tw_hi) {
    extern __shared__ uint4 smem_raw[];
    Fr* sh = reinterpret_cast<Fr*>(smem_raw);
    const uint32_t R = 1u << rbits;
    const uint32_t T = NTT_TILE_COLS;
    const uint64_t S = 1ull << log_stride;
    const uint64_t tiles_per_group = S / T;  // column tiles inside one (hi) group
    const uint64_t grp = blockIdx.x / tiles_per_group;
    const uint64_t c0 = (blockIdx.x % tiles_per_group) * T;
    const uint64_t base = grp * (S << rbits) + c0;
    for (uint32_t idx = threadIdx.x; idx < R * T; idx += NTT_THREADS) {
        uint32_t c = idx % T, n = idx / T;
        uint64_t gi = base + (uint64_t)n * S + c;
        Fr v = gi < n_in ? fp_load<FrP>(src + gi) : fp_zero<FrP>();
        fp_store(sh + idx, v);
    }
    __syncthreads();
    smem_dif(sh, rbits, T, tw_hi);
    // twiddle w_{N_i}^{k * col}, N_i = R * S : exponent in units of W = w_{2^26}
    const uint32_t log_ni = rbits + log_stride;
    const uint32_t shift = NTT_MAX_LOG - log_ni;
    (void)log_n;
    for (uint32_t idx = threadIdx.x; idx < R * T; idx += NTT_THREADS) {
        uint32_t c = idx % T, p = idx / T;
        uint32_t k = bitrev(p, rbits);
        Fr v = fp_load<FrP>(sh + idx);
        uint64_t col = c0 + c;
        uint64_t e = ((uint64_t)k * col) & ((1ull << log_ni) - 1);
        if (e != 0) v = fp_mul(v, tw_lookup(tw_lo, tw_hi, (uint32_t)(e << shift)));
        fp_store(dst + base + (uint64_t)k * S + c, v);
    }
}

// last pass: rows of R contiguous elements; a tile is T rows that are consecutive in k_1.
// grid = N / (R * T)   (T = 1 when npass == 1)
__global__ void __launch_bounds__(NTT_THREADS) ntt_last_pass_kernel(const Fr* __restrict__ src, uint64_t n_in,
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
    for (uint32_t idx = threadIdx.x; idx < R * T; idx += NTT_THREADS) {
        uint32_t n = idx % R, r = idx / R;
        uint64_t gi = (P > 1 ? ((k1_0 + r) << log_s1) : 0) + (mid << rbits) + n;
        Fr v = gi < n_in ? fp_load<FrP>(src + gi) : fp_zero<FrP>();
        fp_store(sh + n * T + r, v);
    }
    __syncthreads();
    smem_dif(sh, rbits, T, tw_hi);
    // output index = k_1 + R_1 * (digit-reversed mid) + k_P * (N / R_P)
    uint64_t out_mid = 0;
    if (P > 2) {
        // storage order of mid: k_2 most significant ... k_{P-1} least; output order: k_2 least significant
        uint64_t rest = mid;
        uint32_t mult_bits = mid_bits;
        for (uint32_t j = P - 2; j >= 1; j--) {
            uint32_t kb = plan.radbits[j];
            uint64_t kj = rest & ((1ull << kb) - 1);
            rest >>= kb;
            mult_bits -= kb;
            // digit j has output weight prod_{1<m<j} R_m  (relative to R_1)
            uint32_t wbits = 0;
            for (uint32_t m = 1; m < j; m++) wbits += plan.radbits[m];
            out_mid += kj << wbits;
            if (j == 1) break;
        }
        (void)mult_bits;
    }
    const uint32_t log_hi = plan.log_n - rbits;  // N / R_P
    for (uint32_t idx = threadIdx.x; idx < R * T; idx += NTT_THREADS) {
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

int ntt_init_tables(kzg_ctx* ctx) {
    for (int dir = 0; dir < 2; dir++) {
        KZG_CUDA(ctx, cudaMalloc((void**)&ctx->tw_lo[dir], sizeof(Fr) * TW_SIZE));
        KZG_CUDA(ctx, cudaMalloc((void**)&ctx->tw_hi[dir], sizeof(Fr) * TW_SIZE));
        Fr w = fr_root_of_unity(NTT_MAX_LOG);
        if (dir == 1) w = fp_inv(w);
        Fr w_hi = fp_pow_u64(w, TW_SIZE);
        KZG_LAUNCH(ctx, tw_table_kernel, TW_SIZE / 256, 256, 0, ctx->tw_lo[dir], ctx->tw_hi[dir], w, w_hi);
    }
    KZG_CHECK_LAUNCH(ctx);
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
        KZG_LAUNCH(ctx, ntt_last_pass_kernel, 1, NTT_THREADS, sizeof(Fr) * R, in, n_in, out, plan, 1u, ctx->tw_hi[dir],
                   scale, inverse);
        KZG_CHECK_LAUNCH(ctx);
        return KZG_OK;
    }
    Fr* tmp = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&tmp, sizeof(Fr) * N, ctx->stream));
    uint32_t log_stride = log_n;
    const Fr* src = in;
    uint64_t src_n = n_in;
    for (uint32_t i = 0; i + 1 < plan.npass; i++) {
        const uint32_t rb = plan.radbits[i];
        log_stride -= rb;
        const uint32_t R = 1u << rb;
        const uint32_t grid = (uint32_t)(N / ((uint64_t)R * NTT_TILE_COLS));
        KZG_LAUNCH(ctx, ntt_strided_pass_kernel, grid, NTT_THREADS, sizeof(Fr) * R * NTT_TILE_COLS, src, src_n, tmp, rb,
                   log_stride, log_n, ctx->tw_lo[dir], ctx->tw_hi[dir]);
        src = tmp;
        src_n = N;
    }
    {
        const uint32_t rb = plan.radbits[plan.npass - 1];
        const uint32_t R = 1u << rb;
        const uint32_t grid = (uint32_t)(N / ((uint64_t)R * NTT_TILE_COLS));
        KZG_LAUNCH(ctx, ntt_last_pass_kernel, grid, NTT_THREADS, sizeof(Fr) * R * NTT_TILE_COLS, tmp, N, out, plan,
                   NTT_TILE_COLS, ctx->tw_hi[dir], scale, inverse);
    }
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaFreeAsync(tmp, ctx->stream));
    return KZG_OK;
}

}  // namespace kzg
