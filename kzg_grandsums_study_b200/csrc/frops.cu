// Bulk Fr kernels of the prover hot path (sm_100a):
//   fr_convert            Fr.batchToMontgomery / batchFromMontgomery    (prover.js:147-148, polynomial.js:1109)
//   fr_batch_inverse      Fr.batchInverse, 0 -> 0                       (grandsum.js:41, grandproduct.js:36)
//   grand_terms           num/den build loops                           (grandsum.js:21-38, grandproduct.js:21-33)
//   fr_exclusive_scan     running sum / running product                 (grandsum.js:44-51, grandproduct.js:39-46)
//   poly_suffix_*         r_i = a_i + v r_{i+1}: Horner evaluation and division by (X - v)
//                                                                       (polynomial.js:228-238, 814-851)
//   poly_linear_combination   the add/sub/mulScalar/addScalar chains of r(X), W(X)   (prover.js:347-402)
//   poly_degree, all_equal    polynomial.js:212-226, evaluations.js:110-129
// Element layout everywhere: 32-byte Montgomery-LE, array-of-elements; each thread moves whole
// elements with two 128-bit accesses, so a warp touches 1 KiB contiguous per access pair.
#include <string.h>

#include "common.cuh"

namespace kzg {

constexpr int EW_THREADS = 256;

static inline uint32_t grid_for(uint64_t n, uint32_t per_block) { return (uint32_t)((n + per_block - 1) / per_block); }

// ------------------------------------------------------------------------------------------------
// element-wise
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) fr_convert_kernel(const Fr* __restrict__ in, Fr* __restrict__ out, uint64_t n,
                                                                bool to_mont) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_load<FrP>(in + i);
    v = to_mont ? fp_to_mont(v) : fp_from_mont(v);
    fp_store(out + i, v);
}
int fr_convert(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n, bool to_mont) {
    if (n == 0) return KZG_OK;
    KZG_LAUNCH(ctx, fr_convert_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, in, out, n, to_mont);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

__global__ void __launch_bounds__(EW_THREADS) fr_mul_kernel(const Fr* __restrict__ a, const Fr* __restrict__ b,
                                                            Fr* __restrict__ out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_store(out + i, fp_mul(fp_load<FrP>(a + i), fp_load<FrP>(b + i)));
}
int fr_mul_pointwise(kzg_ctx* ctx, const Fr* a, const Fr* b, Fr* out, uint64_t n) {
    if (n == 0) return KZG_OK;
    KZG_LAUNCH(ctx, fr_mul_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, a, b, out, n);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

__global__ void __launch_bounds__(EW_THREADS) fr_fill_kernel(Fr* __restrict__ dst, uint64_t n, Fr v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_store(dst + i, v);
}
int fr_fill(kzg_ctx* ctx, Fr* dst, uint64_t n, const Fr& v) {
    if (n == 0) return KZG_OK;
    KZG_LAUNCH(ctx, fr_fill_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, dst, n, v);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

// out[i] = sum_j c_j * p_j[i] (i < len_j)  + (i == 0 ? constant : 0)
constexpr int LC_MAX = 28;
struct LinCombArgs {
    const Fr* poly[LC_MAX];
    uint64_t len[LC_MAX];
    Fr coeff[LC_MAX];
    Fr constant;
    uint32_t count;
};
__global__ void __launch_bounds__(EW_THREADS) lincomb_kernel(Fr* __restrict__ out, uint64_t n_out, LinCombArgs a) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    Fr acc = i == 0 ? a.constant : fp_zero<FrP>();
    for (uint32_t j = 0; j < a.count; j++) {
        if (i < a.len[j]) acc = fp_add(acc, fp_mul(a.coeff[j], fp_load<FrP>(a.poly[j] + i)));
    }
    fp_store(out + i, acc);
}
static int lincomb_launch(kzg_ctx* ctx, Fr* out, uint64_t n_out, const Fr* const* polys, const uint64_t* lens,
                          const Fr* coeffs, uint32_t count, const Fr& constant) {
    LinCombArgs a;
    memset(&a, 0, sizeof(a));
    for (uint32_t j = 0; j < count; j++) {
        a.poly[j] = polys[j];
        a.len[j] = lens[j];
        a.coeff[j] = coeffs[j];
    }
    a.constant = constant;
    a.count = count;
    KZG_LAUNCH(ctx, lincomb_kernel, grid_for(n_out, EW_THREADS), EW_THREADS, 0, out, n_out, a);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}
// Any number of terms: the kernel takes LC_MAX per launch, further launches add LC_MAX - 1 more to the running result
// (the reference accepts any number of columns, prover.js:34-45).  `out` may alias one of the inputs only when a single
// launch suffices (element i is read before it is written by the same thread).
int poly_linear_combination(kzg_ctx* ctx, Fr* out, uint64_t n_out, const Fr* const* polys, const uint64_t* lens,
                            const Fr* coeffs, uint32_t count, const Fr& constant) {
    if (count <= LC_MAX) return lincomb_launch(ctx, out, n_out, polys, lens, coeffs, count, constant);
    for (uint32_t j = 0; j < count; j++)
        if (polys[j] == out) return set_err(ctx, KZG_ERR_ARG, "linear combination: in-place use with more than 28 terms");
    KZG_TRY(lincomb_launch(ctx, out, n_out, polys, lens, coeffs, LC_MAX, constant));
    for (uint32_t done = LC_MAX; done < count;) {
        const uint32_t take = count - done < LC_MAX - 1 ? count - done : LC_MAX - 1;
        const Fr* ps[LC_MAX];
        uint64_t ls[LC_MAX];
        Fr cs[LC_MAX];
        ps[0] = out;
        ls[0] = n_out;
        cs[0] = fp_one<FrP>();
        for (uint32_t j = 0; j < take; j++) {
            ps[1 + j] = polys[done + j];
            ls[1 + j] = lens[done + j];
            cs[1 + j] = coeffs[done + j];
        }
        KZG_TRY(lincomb_launch(ctx, out, n_out, ps, ls, cs, take + 1, fp_zero<FrP>()));
        done += take;
    }
    return KZG_OK;
}

// ------------------------------------------------------------------------------------------------
// warp helpers over Fr
// ------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ Fp<P> shfl_up_fr(const Fp<P>& v, uint32_t d) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
template <class P>
__device__ __forceinline__ Fp<P> shfl_down_fr(const Fp<P>& v, uint32_t d) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
    return r;
}
template <class P>
__device__ __forceinline__ Fp<P> shfl_fr(const Fp<P>& v, uint32_t src) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}

// ------------------------------------------------------------------------------------------------
// batch inverse, base case (<= 4096 elements): Montgomery trick per thread (E elements), prefix and suffix
// products across the warp by shuffles, ONE inversion per warp (32 * E elements).  Zero maps to zero.
// Larger inputs go through the two-level scheme below (fr_batch_inverse).
// ------------------------------------------------------------------------------------------------
constexpr int BI_E = 8;
// This kernel runs once per call with a handful of warps, so its instructions arrive cold from L2/DRAM and its time
// is latency: the loops are real loops around ONE shared copy of the Montgomery product (~20 KB of code instead of
// ~240 KB with everything unrolled and inlined: 119 -> 90 us came from the faster inversion, the rest from this).
template <class P>
__device__ __noinline__ Fp<P> fp_mul_shared(const Fp<P>& a, const Fp<P>& b) {
    return fp_mul(a, b);
}
template <class P>
__global__ void __launch_bounds__(EW_THREADS) batch_inverse_kernel(const Fp<P>* __restrict__ in, Fp<P>* __restrict__ out,
                                                                   uint64_t n) {
    const uint64_t base = (uint64_t)blockIdx.x * (EW_THREADS * BI_E) + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    Fp<P> v[BI_E], p[BI_E];
    uint32_t zmask = 0;
    const Fp<P> one = fp_one<P>();
#pragma unroll 1
    for (int k = 0; k < BI_E; k++) {
        uint64_t i = base + (uint64_t)k * EW_THREADS;
        v[k] = i < n ? fp_load<P>(in + i) : one;
        if (fp_is_zero(v[k])) {
            zmask |= 1u << k;
            v[k] = one;
        }
        p[k] = k == 0 ? v[0] : fp_mul_shared(p[k - 1], v[k]);
    }
    const Fp<P> total = p[BI_E - 1];
    // inclusive prefix products over lanes
    Fp<P> pre = total;
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {
        Fp<P> o = shfl_up_fr(pre, d);
        if (lane >= d) pre = fp_mul_shared(pre, o);
    }
    // inclusive suffix products over lanes
    Fp<P> suf = total;
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {
        Fp<P> o = shfl_down_fr(suf, d);
        if (lane + d < 32) suf = fp_mul_shared(suf, o);
    }
    Fp<P> inv_all = pre;  // lane 31 holds the warp product
    if (lane == 31) inv_all = fp_inv(pre);
    inv_all = shfl_fr(inv_all, 31);
    // 1 / total_lane = prefix_{lane-1} * suffix_{lane+1} * inv_all
    Fp<P> pre_ex = shfl_up_fr(pre, 1);
    Fp<P> suf_ex = shfl_down_fr(suf, 1);
    Fp<P> inv_t = inv_all;
    if (lane > 0) inv_t = fp_mul_shared(inv_t, pre_ex);
    if (lane < 31) inv_t = fp_mul_shared(inv_t, suf_ex);
#pragma unroll 1
    for (int k = BI_E - 1; k >= 0; k--) {
        Fp<P> r = k == 0 ? inv_t : fp_mul_shared(inv_t, p[k - 1]);
        if (k > 0) inv_t = fp_mul_shared(inv_t, v[k]);
        uint64_t i = base + (uint64_t)k * EW_THREADS;
        if (i < n) fp_store(out + i, (zmask >> k) & 1 ? fp_zero<P>() : r);
    }
}
// Large inputs: two-level Montgomery trick across kernels, so that the number of Fermat inversions (254
// squarings each, executed by a single lane) drops from one per 256 elements to one per 256 * 8^levels:
//   products : thread g multiplies its BI_E elements (zeros skipped) -> P[g]
//   (recurse): P <- 1 / P   (an input 8 times smaller)
//   apply    : thread g re-reads its elements, rebuilds its prefix products and peels the inverses off P[g]
// ~4.3 modmul and 96 bytes per element.  Element k of thread (block, t) is base + k * blockDim + t (coalesced).
template <class P>
__global__ void __launch_bounds__(EW_THREADS) batch_products_kernel(const Fp<P>* __restrict__ in, Fp<P>* __restrict__ prod, uint64_t n) {
    const uint64_t base = (uint64_t)blockIdx.x * (EW_THREADS * BI_E) + threadIdx.x;
    Fp<P> p = fp_one<P>();
#pragma unroll
    for (int k = 0; k < BI_E; k++) {
        uint64_t i = base + (uint64_t)k * EW_THREADS;
        if (i < n) {
            Fp<P> v = fp_load<P>(in + i);
            if (!fp_is_zero(v)) p = fp_mul(p, v);
        }
    }
    fp_store(prod + (uint64_t)blockIdx.x * EW_THREADS + threadIdx.x, p);
}

template <class P>
__global__ void __launch_bounds__(EW_THREADS) batch_apply_kernel(const Fp<P>* __restrict__ in, const Fp<P>* __restrict__ prod_inv,
                                                                 Fp<P>* __restrict__ out, uint64_t n) {
    const uint64_t base = (uint64_t)blockIdx.x * (EW_THREADS * BI_E) + threadIdx.x;
    Fp<P> v[BI_E], p[BI_E];
    uint32_t zmask = 0;
    const Fp<P> one = fp_one<P>();
#pragma unroll
    for (int k = 0; k < BI_E; k++) {
        uint64_t i = base + (uint64_t)k * EW_THREADS;
        v[k] = i < n ? fp_load<P>(in + i) : one;
        if (fp_is_zero(v[k])) {
            zmask |= 1u << k;
            v[k] = one;
        }
        p[k] = k == 0 ? v[0] : fp_mul(p[k - 1], v[k]);
    }
    Fp<P> inv_t = fp_load<P>(prod_inv + (uint64_t)blockIdx.x * EW_THREADS + threadIdx.x);
#pragma unroll
    for (int k = BI_E - 1; k >= 0; k--) {
        Fp<P> r = k == 0 ? inv_t : fp_mul(inv_t, p[k - 1]);
        if (k > 0) inv_t = fp_mul(inv_t, v[k]);
        uint64_t i = base + (uint64_t)k * EW_THREADS;
        if (i < n) fp_store(out + i, (zmask >> k) & 1 ? fp_zero<P>() : r);
    }
}

template <class P>
static int fp_batch_inverse(kzg_ctx* ctx, const Fp<P>* in, Fp<P>* out, uint64_t n) {
    if (n == 0) return KZG_OK;
    const uint32_t blocks = grid_for(n, EW_THREADS * BI_E);
    if (n <= 4096) {
        KZG_LAUNCH(ctx, batch_inverse_kernel<P>, blocks, EW_THREADS, 0, in, out, n);
        KZG_CHECK_LAUNCH(ctx);
        return KZG_OK;
    }
    const uint64_t m = (uint64_t)blocks * EW_THREADS;  // thread products (never zero)
    Fp<P>* prod = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&prod, sizeof(Fp<P>) * m, ctx->stream));
    KZG_LAUNCH(ctx, batch_products_kernel<P>, blocks, EW_THREADS, 0, in, prod, n);
    int r = fp_batch_inverse<P>(ctx, prod, prod, m);
    if (r == KZG_OK) {
        KZG_LAUNCH(ctx, batch_apply_kernel<P>, blocks, EW_THREADS, 0, in, prod, out, n);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) r = set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFreeAsync(prod, ctx->stream);
    return r;
}
int fr_batch_inverse(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n) { return fp_batch_inverse<FrP>(ctx, in, out, n); }
// the same over the base field: the batched-affine bucket rounds of the MSM (msm.cu)
int fq_batch_inverse(kzg_ctx* ctx, const Fq* in, Fq* out, uint64_t n) { return fp_batch_inverse<FqP>(ctx, in, out, n); }

// ------------------------------------------------------------------------------------------------
// grand-sum / grand-product terms.  kind 0: num = t'*selF - f'*selT, den = f'*t'  (f' = F+gamma ...)
//                                   kind 1: num = selF*(f'-1)+1,     den = selT*(t'-1)+1
// Written UN-rotated (index i); the exclusive scan that follows provides the (i+1)%n rotation of
// the reference: S[i+1] = S[i] + term[i], S[0] = 0 and the wrap value is the scan total.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) grand_terms_kernel(int kind, const Fr* __restrict__ ev_f,
                                                                 const Fr* __restrict__ ev_t, const Fr* __restrict__ sel_f,
                                                                 const Fr* __restrict__ sel_t, Fr gamma,
                                                                 Fr* __restrict__ num, Fr* __restrict__ den, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr one = fp_one<FrP>();
    Fr f = fp_add(fp_load<FrP>(ev_f + i), gamma);
    Fr t = fp_add(fp_load<FrP>(ev_t + i), gamma);
    Fr nu, de;
    if (kind == 0) {
        if (sel_f) {
            nu = fp_sub(fp_mul(t, fp_load<FrP>(sel_f + i)), fp_mul(f, fp_load<FrP>(sel_t + i)));
        } else {
            nu = fp_sub(t, f);
        }
        de = fp_mul(f, t);
    } else {
        if (sel_f) {
            nu = fp_add(fp_mul(fp_load<FrP>(sel_f + i), fp_sub(f, one)), one);
            de = fp_add(fp_mul(fp_load<FrP>(sel_t + i), fp_sub(t, one)), one);
        } else {
            nu = f;
            de = t;
        }
    }
    fp_store(num + i, nu);
    fp_store(den + i, de);
}
int grand_terms(kzg_ctx* ctx, int kind, const Fr* ev_f, const Fr* ev_t, const Fr* sel_f, const Fr* sel_t,
                const Fr& gamma, Fr* num, Fr* den, uint64_t n) {
    KZG_LAUNCH(ctx, grand_terms_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, kind, ev_f, ev_t, sel_f, sel_t, gamma,
               num, den, n);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan over Fr (add or mul monoid): single pass, decoupled look-back (below).
// Block tile = SC_THREADS * SC_E contiguous elements; thread t owns SC_E contiguous elements.
// ------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_E = 8;
constexpr int SC_TILE = SC_THREADS * SC_E;

template <int KIND> __device__ __forceinline__ Fr scan_op(const Fr& a, const Fr& b) {
    return KIND == SCAN_ADD ? fp_add(a, b) : fp_mul(a, b);
}
template <int KIND> __device__ __forceinline__ Fr scan_identity() {
    return KIND == SCAN_ADD ? fp_zero<FrP>() : fp_one<FrP>();
}

// inclusive scan of one value per thread across the block; returns inclusive value, writes block total
template <int KIND> __device__ __forceinline__ Fr block_inclusive_scan(Fr v, Fr* sh_warp, Fr& block_total) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        Fr o = shfl_up_fr(v, d);
        if (lane >= d) v = scan_op<KIND>(o, v);
    }
    if (lane == 31) fp_store(sh_warp + wid, v);
    __syncthreads();
    constexpr int NW = SC_THREADS / 32;
    Fr carry = scan_identity<KIND>();
    Fr tot = scan_identity<KIND>();
#pragma unroll
    for (int w = 0; w < NW; w++) {
        Fr x = fp_load<FrP>(sh_warp + w);
        if (w < (int)wid) carry = scan_op<KIND>(carry, x);
        tot = scan_op<KIND>(tot, x);
    }
    block_total = tot;
    __syncthreads();
    return scan_op<KIND>(carry, v);
}

// Single-pass scan with DECOUPLED LOOK-BACK (Merrill & Garland): every tile is read once and written once (64 N bytes),
// in one launch.  A block takes its tile from an atomic ticket (so it only ever waits for tiles whose blocks are already
// running), reduces it, publishes the tile AGGREGATE (flag 1), then its first warp looks back over the predecessors --
// 32 tiles per step, each lane one tile: an INCLUSIVE prefix (flag 2) ends the walk, aggregates are folded in -- and
// publishes its own inclusive prefix (flag 2).  Both monoids (Fr addition for the grand sum, Fr multiplication for the
// grand product) are commutative, so the lanes' values are combined by a plain shuffle tree.  Values are 32 bytes, flags
// separate words: value stores, __threadfence(), flag store on the writer; flag load (volatile), __threadfence(), value
// loads that bypass L1 on the reader.
struct ScanLookback {
    uint32_t* flags;     // per tile: 0 nothing yet, 1 aggregate published, 2 inclusive prefix published
    Fr* aggregate;       // per tile
    Fr* inclusive;       // per tile; inclusive[ntiles - 1] is the grand total
    uint32_t* ticket;
};
__device__ __forceinline__ Fr ld_cg_fr(const Fr* p) {
    Fr r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = __ldcg(q), b = __ldcg(q + 1);
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <int KIND>
__global__ void __launch_bounds__(SC_THREADS) scan_lookback_kernel(const Fr* __restrict__ in, Fr* __restrict__ out, uint64_t n,
                                                                   ScanLookback st) {
    __shared__ Fr sh[SC_THREADS / 32];
    __shared__ Fr sh_incl[SC_THREADS];
    __shared__ Fr sh_prefix;
    __shared__ uint32_t sh_tile;
    if (threadIdx.x == 0) sh_tile = atomicAdd(st.ticket, 1u);
    __syncthreads();
    const uint32_t tile = sh_tile;
    const uint64_t base = (uint64_t)tile * SC_TILE + (uint64_t)threadIdx.x * SC_E;
    Fr v[SC_E];
    Fr acc = scan_identity<KIND>();
#pragma unroll
    for (int k = 0; k < SC_E; k++) {
        const uint64_t i = base + k;
        v[k] = i < n ? fp_load<FrP>(in + i) : scan_identity<KIND>();
        acc = scan_op<KIND>(acc, v[k]);
    }
    Fr tot;
    const Fr incl = block_inclusive_scan<KIND>(acc, sh, tot);
    fp_store(sh_incl + threadIdx.x, incl);
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        Fr prefix = scan_identity<KIND>();
        if (tile == 0) {
            if (lane == 0) {
                fp_store(st.inclusive, tot);
                __threadfence();
                *(volatile uint32_t*)st.flags = 2u;
            }
        } else {
            if (lane == 0) {
                fp_store(st.aggregate + tile, tot);
                __threadfence();
                *(volatile uint32_t*)(st.flags + tile) = 1u;
            }
            int64_t window = (int64_t)tile - 1;  // lane l inspects tile `window - l`
            while (true) {
                const int64_t p = window - (int64_t)lane;
                uint32_t flag = 2u;              // lanes before tile 0 behave like an (identity) inclusive prefix
                if (p >= 0) {
                    do {
                        flag = *(volatile const uint32_t*)(st.flags + p);
                    } while (flag == 0u);
                }
                __threadfence();
                const uint32_t done_mask = __ballot_sync(0xffffffffu, flag == 2u);
                const uint32_t stop = done_mask ? (uint32_t)__ffs(done_mask) - 1u : 32u;  // first lane with an inclusive prefix
                Fr val = scan_identity<KIND>();
                if (p >= 0 && lane <= stop) val = ld_cg_fr(flag == 2u ? st.inclusive + p : st.aggregate + p);
#pragma unroll
                for (uint32_t d = 16; d >= 1; d >>= 1) {
                    const Fr o = shfl_down_fr(val, d);
                    val = scan_op<KIND>(val, o);  // (lanes beyond `stop` hold the identity)
                }
                if (lane == 0) prefix = scan_op<KIND>(val, prefix);
                if (done_mask) break;
                window -= 32;
            }
            if (lane == 0) {
                fp_store(st.inclusive + tile, scan_op<KIND>(prefix, tot));
                __threadfence();
                *(volatile uint32_t*)(st.flags + tile) = 2u;
            }
        }
        if (lane == 0) fp_store(&sh_prefix, prefix);
    }
    __syncthreads();
    Fr run = fp_load<FrP>(&sh_prefix);
    if (threadIdx.x > 0) run = scan_op<KIND>(run, fp_load<FrP>(sh_incl + threadIdx.x - 1));
#pragma unroll
    for (int k = 0; k < SC_E; k++) {
        const uint64_t i = base + k;
        if (i < n) fp_store(out + i, run);
        run = scan_op<KIND>(run, v[k]);
    }
}

int fr_exclusive_scan(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n, ScanKind kind, Fr* total_host) {
    if (n == 0) return KZG_OK;
    const uint32_t nblk = grid_for(n, SC_TILE);
    // state: [flags: nblk + 1 words (the last one is the ticket)] [aggregates: nblk] [inclusive prefixes: nblk]
    const size_t flag_bytes = ((size_t)(nblk + 1) * sizeof(uint32_t) + 31) / 32 * 32;
    uint8_t* state = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&state, flag_bytes + 2 * sizeof(Fr) * nblk, ctx->stream));
    KZG_CUDA(ctx, cudaMemsetAsync(state, 0, flag_bytes, ctx->stream));
    ScanLookback st;
    st.flags = (uint32_t*)state;
    st.ticket = st.flags + nblk;
    st.aggregate = (Fr*)(state + flag_bytes);
    st.inclusive = st.aggregate + nblk;
    if (kind == SCAN_ADD)
        KZG_LAUNCH(ctx, scan_lookback_kernel<SCAN_ADD>, nblk, SC_THREADS, 0, in, out, n, st);
    else
        KZG_LAUNCH(ctx, scan_lookback_kernel<SCAN_MUL>, nblk, SC_THREADS, 0, in, out, n, st);
    KZG_CHECK_LAUNCH(ctx);
    if (total_host) {
        KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, st.inclusive + (nblk - 1), sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
        KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(total_host, ctx->pinned, sizeof(Fr));
    }
    KZG_CUDA(ctx, cudaFreeAsync(state, ctx->stream));
    return KZG_OK;
}

// ------------------------------------------------------------------------------------------------
// weighted suffix recurrence  r_i = a_i + v * r_{i+1},  r_n = 0  (r_0 = P(v); q_i = r_{i+1} = P / (X - v))
// Powers v^(2^j), j = 0..31, are computed on the host (32 squarings) and passed by value.
// Tile = SF_THREADS * SF_E contiguous coefficients = 2^SF_LOG_TILE.
// ------------------------------------------------------------------------------------------------
constexpr int SF_THREADS = 256;
constexpr int SF_LOG_E = 3;
constexpr int SF_E = 1 << SF_LOG_E;
constexpr int SF_LOG_TILE = 8 + SF_LOG_E;  // 2048
constexpr int SF_TILE = 1 << SF_LOG_TILE;

struct PowTable {
    Fr p[32];  // p[j] = v^(2^j)
};

// Inclusive weighted suffix scan across the block: I_t = sum_{j >= t} A_j y^(j-t), with an extra
// element A_T = tail (carry from beyond the block).  y^(2^s) = pw.p[log_unit + s].  sh holds T+1 entries.
__device__ __forceinline__ Fr block_weighted_suffix(Fr a, const Fr& tail, Fr* sh, const PowTable& pw, uint32_t log_unit) {
    const uint32_t t = threadIdx.x;
    fp_store(sh + t, a);
    if (t == 0) fp_store(sh + SF_THREADS, tail);
    __syncthreads();
    uint32_t s = 0;
    for (uint32_t d = 1; d <= SF_THREADS; d <<= 1, s++) {
        Fr o = fp_zero<FrP>();
        bool has = t + d <= SF_THREADS;
        if (has) o = fp_load<FrP>(sh + t + d);
        __syncthreads();
        if (has) {
            a = fp_add(a, fp_mul(o, pw.p[log_unit + s]));
            fp_store(sh + t, a);
        }
        __syncthreads();
    }
    return a;
}

// stage A: aggregate of each tile, A_b = sum_{i in tile} a_i y^(i - start), element weight y = pw.p[log_y]
__global__ void __launch_bounds__(SF_THREADS) suffix_reduce_kernel(const Fr* __restrict__ a, uint64_t n, PowTable pw,
                                                                   uint32_t log_y, Fr* __restrict__ agg) {
    __shared__ Fr sh[SF_THREADS + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SF_TILE + (uint64_t)threadIdx.x * SF_E;
    const Fr y = pw.p[log_y];
    Fr r = fp_zero<FrP>();
#pragma unroll
    for (int k = SF_E - 1; k >= 0; k--) {
        uint64_t i = base + k;
        Fr c = i < n ? fp_load<FrP>(a + i) : fp_zero<FrP>();
        r = fp_add(c, fp_mul(r, y));
    }
    Fr incl = block_weighted_suffix(r, fp_zero<FrP>(), sh, pw, log_y + SF_LOG_E);
    if (threadIdx.x == 0) fp_store(agg + blockIdx.x, incl);
}

// stage C: given the carry of each tile (carry[b] = r at the first index after the tile), write
// q_{i-1} = r_i for every i >= 1 of the tile and r_0 to rem (block 0).
__global__ void __launch_bounds__(SF_THREADS) suffix_apply_kernel(const Fr* __restrict__ a, uint64_t n, PowTable pw,
                                                                  const Fr* __restrict__ carry, Fr* __restrict__ q,
                                                                  Fr* __restrict__ rem) {
    __shared__ Fr sh[SF_THREADS + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SF_TILE + (uint64_t)threadIdx.x * SF_E;
    const Fr v = pw.p[0];
    Fr c[SF_E];
    Fr r = fp_zero<FrP>();
#pragma unroll
    for (int k = SF_E - 1; k >= 0; k--) {
        uint64_t i = base + k;
        c[k] = i < n ? fp_load<FrP>(a + i) : fp_zero<FrP>();
        r = fp_add(c[k], fp_mul(r, v));
    }
    Fr tail = fp_load<FrP>(carry + blockIdx.x);
    block_weighted_suffix(r, tail, sh, pw, SF_LOG_E);
    // carry into this thread = I_{t+1}
    Fr run = fp_load<FrP>(sh + threadIdx.x + 1);
#pragma unroll
    for (int k = SF_E - 1; k >= 0; k--) {
        uint64_t i = base + k;
        run = fp_add(c[k], fp_mul(run, v));  // r_i
        if (i < n) {
            if (i > 0) fp_store(q + i - 1, run);
            else fp_store(rem, run);
        }
    }
}

// stage B (single block): carries over tile aggregates.  carry[b] = sum_{j > b} A_j Y^(j-b-1), Y = v^TILE.
// Each thread owns a run of 2^log_run consecutive aggregates.  total (= r_0 = P(v)) -> carry[nblk].
__global__ void __launch_bounds__(SF_THREADS) suffix_carries_kernel(const Fr* __restrict__ agg, uint32_t nblk, PowTable pw,
                                                                    uint32_t log_tile, uint32_t log_run,
                                                                    Fr* __restrict__ carry) {
    __shared__ Fr sh[SF_THREADS + 1];
    const Fr Y = pw.p[log_tile];
    const uint32_t run_len = 1u << log_run;
    const uint64_t lo = (uint64_t)threadIdx.x * run_len;
    Fr r = fp_zero<FrP>();
    for (int k = (int)run_len - 1; k >= 0; k--) {
        uint64_t b = lo + k;
        Fr cv = b < nblk ? fp_load<FrP>(agg + b) : fp_zero<FrP>();
        r = fp_add(cv, fp_mul(r, Y));
    }
    Fr incl = block_weighted_suffix(r, fp_zero<FrP>(), sh, pw, log_tile + log_run);
    Fr run = fp_load<FrP>(sh + threadIdx.x + 1);  // value just after this thread's run
    for (int k = (int)run_len - 1; k >= 0; k--) {
        uint64_t b = lo + k;
        if (b < nblk) {
            fp_store(carry + b, run);
            run = fp_add(fp_load<FrP>(agg + b), fp_mul(run, Y));
        }
    }
    if (threadIdx.x == 0) fp_store(carry + nblk, incl);
}

static PowTable make_pow_table(const Fr& v) {
    PowTable t;
    t.p[0] = v;
    for (int j = 1; j < 32; j++) t.p[j] = fp_sqr(t.p[j - 1]);
    return t;
}

// shared driver: computes tile carries (and the total r_0) for coefficient vector a.
static int suffix_prepare(kzg_ctx* ctx, const Fr* a, uint64_t n, const PowTable& pw, Fr** carry_out, uint32_t* nblk_out) {
    const uint32_t nblk = grid_for(n, SF_TILE);
    Fr* buf = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&buf, sizeof(Fr) * (2 * (size_t)nblk + 2), ctx->stream));
    Fr* agg = buf;
    Fr* carry = buf + nblk + 1;
    KZG_LAUNCH(ctx, suffix_reduce_kernel, nblk, SF_THREADS, 0, a, n, pw, 0u, agg);
    uint32_t log_run = 0;
    while (((uint64_t)SF_THREADS << log_run) < nblk) log_run++;
    KZG_LAUNCH(ctx, suffix_carries_kernel, 1, SF_THREADS, 0, agg, nblk, pw, (uint32_t)SF_LOG_TILE, log_run, carry);
    KZG_CHECK_LAUNCH(ctx);
    *carry_out = carry;
    *nblk_out = nblk;
    return KZG_OK;
}

// evaluate `count` polynomials, polynomial j at points[j]; results to host (Montgomery)
int poly_evaluate_multi(kzg_ctx* ctx, const Fr* const* polys, const uint64_t* lens, const Fr* points, uint32_t count,
                        Fr* out_host) {
    if (count * sizeof(Fr) > ctx->dev_small_bytes) return set_err(ctx, KZG_ERR_ARG, "too many evaluations in one call");
    Fr* dev_out = (Fr*)ctx->dev_small;
    std::vector<Fr*> to_free;
    for (uint32_t j = 0; j < count; j++) {
        if (lens[j] == 0) {
            KZG_CUDA(ctx, cudaMemsetAsync(dev_out + j, 0, sizeof(Fr), ctx->stream));
            continue;
        }
        PowTable pw = make_pow_table(points[j]);
        Fr* carry = nullptr;
        uint32_t nblk = 0;
        KZG_TRY(suffix_prepare(ctx, polys[j], lens[j], pw, &carry, &nblk));
        KZG_CUDA(ctx, cudaMemcpyAsync(dev_out + j, carry + nblk, sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        to_free.push_back(carry - (nblk + 1));
    }
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, dev_out, sizeof(Fr) * count, cudaMemcpyDeviceToHost, ctx->stream));
    for (Fr* p : to_free) KZG_CUDA(ctx, cudaFreeAsync(p, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out_host, ctx->pinned, sizeof(Fr) * count);
    return KZG_OK;
}

// out (n elements): q_i = r_{i+1} for i < n-1, q_{n-1} = 0; *exact = (r_0 == 0)
int poly_div_x_sub(kzg_ctx* ctx, const Fr* a, uint64_t n, const Fr& v, Fr* out, bool* exact) {
    if (n == 0) {
        *exact = true;
        return KZG_OK;
    }
    PowTable pw = make_pow_table(v);
    Fr* carry = nullptr;
    uint32_t nblk = 0;
    KZG_TRY(suffix_prepare(ctx, a, n, pw, &carry, &nblk));
    Fr* rem = (Fr*)ctx->dev_small;
    KZG_CUDA(ctx, cudaMemsetAsync(out + (n - 1), 0, sizeof(Fr), ctx->stream));
    KZG_LAUNCH(ctx, suffix_apply_kernel, nblk, SF_THREADS, 0, a, n, pw, carry, out, rem);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, rem, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaFreeAsync(carry - (nblk + 1), ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    Fr r0;
    memcpy(&r0, ctx->pinned, sizeof(Fr));
    *exact = fp_is_zero(r0);
    return KZG_OK;
}

// ------------------------------------------------------------------------------------------------
// degree (highest non-zero coefficient index, 0 if none) and broadcast compare
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) degree_kernel(const Fr* __restrict__ a, uint64_t n,
                                                            unsigned long long* __restrict__ best) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cand = 0;
    if (i < n && !fp_is_zero(fp_load<FrP>(a + i))) cand = i;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, cand, d);
        cand = o > cand ? o : cand;
    }
    if ((threadIdx.x & 31) == 0 && cand) atomicMax(best, cand);
}
int poly_degree(kzg_ctx* ctx, const Fr* a, uint64_t n, uint64_t* degree) {
    unsigned long long* slot = (unsigned long long*)ctx->dev_small;
    KZG_CUDA(ctx, cudaMemsetAsync(slot, 0, sizeof(unsigned long long), ctx->stream));
    if (n) KZG_LAUNCH(ctx, degree_kernel, grid_for(n, EW_THREADS), EW_THREADS, 0, a, n, slot);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, slot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long v;
    memcpy(&v, ctx->pinned, sizeof(v));
    *degree = v;
    return KZG_OK;
}

__global__ void __launch_bounds__(EW_THREADS) all_equal_kernel(const Fr* __restrict__ a, uint64_t n, Fr v,
                                                               unsigned int* __restrict__ mismatch) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = i < n && !fp_eq(fp_load<FrP>(a + i), v);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(mismatch, 1u);
}

}  // namespace kzg

using namespace kzg;

extern "C" int kzg_buf_all_equal(kzg_ctx* ctx, kzg_buf* b, const uint8_t value[32], int* out) {
    if (!ctx || !b || !out) return KZG_ERR_ARG;
    unsigned int* slot = (unsigned int*)ctx->dev_small;
    KZG_CUDA(ctx, cudaMemsetAsync(slot, 0, sizeof(unsigned int), ctx->stream));
    if (b->n) KZG_LAUNCH(ctx, all_equal_kernel, grid_for(b->n, EW_THREADS), EW_THREADS, 0, b->d, b->n, fr_from_bytes(value), slot);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, slot, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned int m;
    memcpy(&m, ctx->pinned, sizeof(m));
    *out = m ? 0 : 1;
    return KZG_OK;
}
