// Argument cores and the remaining Polynomial operations:
//   grand_build            ComputeSGrandSumPolynomial / ComputeZGrandProductPolynomial up to the evaluation vector
//                          (reference src/grandsum/grandsum.js:6-62, src/grandproduct/grandproduct.js:6-57)
//   fr_scale_powers        coef_i *= g^i for a root of unity g: Polynomial.shiftOmega (polynomial.js:378-393) in
//                          closed form, and the coset shift of the quotient evaluation
//   poly_div_zh            Polynomial.divZh (polynomial.js:853-888)
#include <string.h>

#include "common.cuh"

namespace kzg {

constexpr int AR_THREADS = 256;
static inline uint32_t grid_for(uint64_t n, uint32_t per_block) { return (uint32_t)((n + per_block - 1) / per_block); }

// ------------------------------------------------------------------------------------------------
// out[i] = in[i] * W^(i * step mod 2^26) * post,  W = w_{2^26} (dir 0) or its inverse (dir 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AR_THREADS) scale_powers_kernel(const Fr* __restrict__ in, Fr* __restrict__ out, uint64_t n,
                                                                  uint32_t step, const Fr* __restrict__ tw_lo,
                                                                  const Fr* __restrict__ tw_hi, Fr post, bool use_post) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_load<FrP>(in + i);
    uint32_t e = (uint32_t)((i * (uint64_t)step) & ((1ull << NTT_MAX_LOG) - 1));
    if (e != 0) {
        Fr h = fp_load<FrP>(tw_hi + (e >> TW_BITS));
        uint32_t l = e & (TW_SIZE - 1);
        if (l) h = fp_mul(h, fp_load<FrP>(tw_lo + l));
        v = fp_mul(v, h);
    }
    if (use_post) v = fp_mul(v, post);
    fp_store(out + i, v);
}

// g = w_{2^log_order}; inverse selects g^-1.  post (optional) is one more factor applied to every element.
int fr_scale_powers(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n, uint32_t log_order, bool inverse, const Fr* post) {
    if (n == 0) return KZG_OK;
    if (log_order > NTT_MAX_LOG) return set_err(ctx, KZG_ERR_ARG, "root of unity order above 2^26");
    const uint32_t step = 1u << (NTT_MAX_LOG - log_order);
    const int dir = inverse ? 1 : 0;
    Fr p = post ? *post : fp_one<FrP>();
    KZG_LAUNCH(ctx, scale_powers_kernel, grid_for(n, AR_THREADS), AR_THREADS, 0, in, out, n, step, ctx->tw_lo[dir],
               ctx->tw_hi[dir], p, post != nullptr);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

// ------------------------------------------------------------------------------------------------
// grand-sum / grand-product evaluation vector.
// acc[0] = identity, acc[i+1] = acc[i] (+|*) num_i / den_i;  *wrap_ok = (value after the last step == identity)
// which is the reference's S[0] == 0 / Z[0] == 1 test after its (i+1)%n rotation.
// ------------------------------------------------------------------------------------------------
int grand_build(kzg_ctx* ctx, int kind, const Fr* ev_f, const Fr* ev_t, const Fr* sel_f, const Fr* sel_t, const Fr& gamma,
                uint64_t n, Fr* acc, bool* wrap_ok) {
    Fr* tmp = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&tmp, sizeof(Fr) * 2 * n, ctx->stream));
    Fr* num = tmp;
    Fr* den = tmp + n;
    int r = grand_terms(ctx, kind, ev_f, ev_t, sel_f, sel_t, gamma, num, den, n);
    if (r == KZG_OK) r = fr_batch_inverse(ctx, den, den, n);
    if (r == KZG_OK) r = fr_mul_pointwise(ctx, num, den, num, n);
    Fr total;
    if (r == KZG_OK) r = fr_exclusive_scan(ctx, num, acc, n, kind == KZG_GRANDSUM ? SCAN_ADD : SCAN_MUL, &total);
    cudaFreeAsync(tmp, ctx->stream);
    if (r != KZG_OK) return r;
    *wrap_ok = kind == KZG_GRANDSUM ? fp_is_zero(total) : fp_eq(total, fp_one<FrP>());
    return KZG_OK;
}

// ------------------------------------------------------------------------------------------------
// divZh: q_i = -a_i (i < n), q_i = q_{i-n} - a_i (i >= n): n independent chains of `ext` steps.
// The reference requires q_i == 0 for i > n*(ext-1) - ext (polynomial.js:875-880).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AR_THREADS) div_zh_kernel(const Fr* __restrict__ a, Fr* __restrict__ q, uint64_t n,
                                                            uint32_t ext, unsigned int* __restrict__ bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t limit = n * (ext - 1) - ext;
    Fr prev = fp_zero<FrP>();
    bool nz = false;
    for (uint32_t e = 0; e < ext; e++) {
        uint64_t idx = i + (uint64_t)e * n;
        prev = fp_sub(prev, fp_load<FrP>(a + idx));
        fp_store(q + idx, prev);
        if (e > 0 && idx > limit && !fp_is_zero(prev)) nz = true;
    }
    if (__any_sync(__activemask(), nz) && nz) atomicOr(bad, 1u);
}

}  // namespace kzg

using namespace kzg;

extern "C" {

static int build_common(kzg_ctx* ctx, int kind, kzg_buf* ev_f, kzg_buf* ev_t, kzg_buf* sel_f, kzg_buf* sel_t,
                        const uint8_t gamma[32], kzg_buf** out) {
    if (!ctx || !ev_f || !ev_t || !gamma || !out) return KZG_ERR_ARG;
    if ((sel_f == nullptr) != (sel_t == nullptr)) return KZG_ERR_ARG;
    const uint64_t n = ev_f->n;
    if (n == 0 || (n & (n - 1))) return set_err(ctx, KZG_ERR_PROTOCOL, "Polynomial length must be a power of two.");
    if (ev_t->n != n || (sel_f && (sel_f->n != n || sel_t->n != n)))
        return set_err(ctx, KZG_ERR_ARG, "grand build: all evaluation vectors must have the same length");
    uint32_t lg = 0;
    while ((1ull << lg) < n) lg++;
    kzg_buf* o = nullptr;
    KZG_TRY(buf_new(ctx, n, false, &o));
    bool ok = false;
    int r = grand_build(ctx, kind, ev_f->d, ev_t->d, sel_f ? sel_f->d : nullptr, sel_t ? sel_t->d : nullptr,
                        fr_from_bytes(gamma), n, o->d, &ok);
    if (r == KZG_OK && !ok)
        r = set_err(ctx, KZG_ERR_PROTOCOL,
                    kind == KZG_GRANDSUM ? "The grand-sum polynomial S is not well calculated"
                                         : "The grand-product polynomial Z is not well calculated");
    if (r == KZG_OK) r = ntt_run(ctx, o->d, n, o->d, lg, true);  // Polynomial.fromEvaluations
    if (r != KZG_OK) {
        kzg_buf_free(ctx, o);
        return r;
    }
    *out = o;
    return KZG_OK;
}

int kzg_grandsum_build(kzg_ctx* ctx, kzg_buf* ev_f, kzg_buf* ev_t, kzg_buf* sel_f, kzg_buf* sel_t, const uint8_t gamma[32],
                       kzg_buf** s_coef) {
    kzg::DeviceGuard _dg(ctx);
    return build_common(ctx, KZG_GRANDSUM, ev_f, ev_t, sel_f, sel_t, gamma, s_coef);
}
int kzg_grandproduct_build(kzg_ctx* ctx, kzg_buf* ev_f, kzg_buf* ev_t, kzg_buf* sel_f, kzg_buf* sel_t,
                           const uint8_t gamma[32], kzg_buf** z_coef) {
    kzg::DeviceGuard _dg(ctx);
    return build_common(ctx, KZG_GRANDPRODUCT, ev_f, ev_t, sel_f, sel_t, gamma, z_coef);
}

// p(X) -> p(wX), w = Fr.w[log2 len]: coefficient i times w^i (what fft / rotate / ifft computes)
int kzg_poly_shift_omega(kzg_ctx* ctx, kzg_buf* a, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !a || !out) return KZG_ERR_ARG;
    const uint64_t n = a->n;
    if (n == 0 || (n & (n - 1))) return set_err(ctx, KZG_ERR_PROTOCOL, "fft must be multiple of 2");
    uint32_t lg = 0;
    while ((1ull << lg) < n) lg++;
    kzg_buf* o = nullptr;
    KZG_TRY(buf_new(ctx, n, false, &o));
    int r = fr_scale_powers(ctx, a->d, o->d, n, lg, false, nullptr);
    if (r != KZG_OK) {
        kzg_buf_free(ctx, o);
        return r;
    }
    *out = o;
    return KZG_OK;
}

int kzg_poly_div_zh(kzg_ctx* ctx, kzg_buf* a, uint64_t domain_size, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !a || !out || domain_size == 0) return KZG_ERR_ARG;
    if (a->n % domain_size) return set_err(ctx, KZG_ERR_ARG, "divZh: length must be a multiple of the domain size");
    const uint64_t n = domain_size;
    const uint32_t ext = (uint32_t)(a->n / n);
    uint64_t deg = 0;
    KZG_TRY(poly_degree(ctx, a->d, a->n, &deg));
    // length = degree < n ? 0 : 2^ceil(log2(degree + 1 - n))   (polynomial.js:855)
    uint64_t len = 0;
    if (deg >= n) {
        len = 1;
        while (len < deg + 1 - n) len <<= 1;
    }
    Fr* q = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&q, sizeof(Fr) * a->n, ctx->stream));
    unsigned int* bad = (unsigned int*)ctx->dev_small;
    KZG_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(unsigned int), ctx->stream));
    KZG_LAUNCH(ctx, div_zh_kernel, grid_for(n, AR_THREADS), AR_THREADS, 0, a->d, q, n, ext, bad);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned int flag;
    memcpy(&flag, ctx->pinned, sizeof(flag));
    if (flag) {
        cudaFreeAsync(q, ctx->stream);
        return set_err(ctx, KZG_ERR_PROTOCOL, "Polynomial is not divisible");
    }
    kzg_buf* o = nullptr;
    int r = buf_new(ctx, len, true, &o);
    if (r == KZG_OK && len) {
        // the reference copies the first degree(q)+1 coefficients; everything above is zero anyway
        uint64_t ncopy = len < a->n ? len : a->n;
        cudaError_t e = cudaMemcpyAsync(o->d, q, sizeof(Fr) * ncopy, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) r = set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFreeAsync(q, ctx->stream);
    if (r != KZG_OK) {
        if (o) kzg_buf_free(ctx, o);
        return r;
    }
    *out = o;
    return KZG_OK;
}

}  // extern "C"
