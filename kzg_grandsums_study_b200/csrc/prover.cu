// Fused five-round provers: grand-sum (reference src/grandsum/mset_eq_kzg_prover.js:144-413) and
// grand-product (src/grandproduct/mset_eq_kzg_prover.js:144-410).  The Keccak transcript stays with the
// caller; each round consumes the challenge the host derived from the previous round's outputs.
//
// The rounds compute the same polynomials as the reference, restructured for the device (every value
// that reaches the proof is a canonical field element / affine point, so the bytes are identical):
//   round 1  columns H2D -> batchToMontgomery -> iNTT(n) -> commit                       (:144-179)
//   round 2  F = sum beta^i F_i in both bases, terms -> batch inverse -> scan -> iNTT -> commit (:181-231)
//   round 3  Q = numerator / Z_H evaluated POINTWISE on the coset g*H_m, m = n or 2n, g = w_{2m}
//            (deg Q < m), instead of the reference's chain of NTT multiplications at sizes 2n and 4n
//            followed by the coefficient-form divZh; divisibility ("Polynomial is not divisible",
//            polynomial.js:876-880) is checked exactly by testing that the numerator vanishes on H. (:233-286)
//   round 4  Horner evaluations as weighted suffix reductions                              (:288-318)
//   round 5  r(X) and the W numerators as ONE fused linear combination each, then the (X - v) division
//            as an affine suffix scan, then two commits                                     (:320-413)
#include <string.h>

#include "common.cuh"

struct kzg_prover {
    kzg_ctx* ctx = nullptr;
    kzg_srs* srs = nullptr;
    int kind = KZG_GRANDSUM;
    uint32_t n_bits = 0, k = 0;
    bool selected = false;
    uint64_t n = 0;
    uint32_t ext = 2;   // quotient coset size m = ext * n
    uint64_t m = 0;
    int round = 0;
    std::vector<kzg::Fr*> ev_f, ev_t, co_f, co_t;  // per column: evaluations on H and coefficients (Montgomery)
    kzg::Fr *ev_self = nullptr, *ev_selt = nullptr, *co_self = nullptr, *co_selt = nullptr;
    kzg::Fr *ev_fc = nullptr, *ev_tc = nullptr, *co_fc = nullptr, *co_tc = nullptr;  // beta-combined (alias column 0 if k == 1)
    kzg::Fr *ev_acc = nullptr, *co_acc = nullptr;  // S or Z
    kzg::Fr* co_q = nullptr;                        // m coefficients
    kzg::Fr* inv_nx = nullptr;                      // 1 / (n (x_i - 1)) on the coset
    kzg::Fr* cos = nullptr;                         // columns evaluated on the coset g H_m (queued in round 2, used in round 3)
    cudaEvent_t cos_ready = nullptr;
    bool cos_queued = false;
    kzg::Fr beta, gamma, alpha, xi, v;
    std::vector<kzg::Fr> evals;  // round-4 outputs in proof order
    std::vector<void*> owned;
};

namespace kzg {

constexpr int PR_THREADS = 256;
static inline uint32_t grid_for(uint64_t n, uint32_t per_block) { return (uint32_t)((n + per_block - 1) / per_block); }

// host-side scalar helpers (O(1) field work per round; field.cuh's portable path)
static Fr h_mul(const Fr& a, const Fr& b) { return fp_mul(a, b); }
static Fr h_add(const Fr& a, const Fr& b) { return fp_add(a, b); }
static Fr h_sub(const Fr& a, const Fr& b) { return fp_sub(a, b); }
static Fr h_from_u64(uint64_t x) {
    Fr r = fp_zero<FrP>();
    r.l[0] = (uint32_t)x;
    r.l[1] = (uint32_t)(x >> 32);
    return fp_to_mont(r);
}

// x_i - 1 scaled by n on the coset x_i = w_{2m}^(2i+1)
__global__ void __launch_bounds__(PR_THREADS) coset_xm1_kernel(Fr* __restrict__ out, uint64_t m, uint32_t step, Fr n_mont,
                                                               const Fr* __restrict__ tw_lo, const Fr* __restrict__ tw_hi) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t e = (uint32_t)(((2 * i + 1) * (uint64_t)step) & ((1ull << NTT_MAX_LOG) - 1));
    Fr x = fp_load<FrP>(tw_hi + (e >> TW_BITS));
    uint32_t l = e & (TW_SIZE - 1);
    if (l) x = fp_mul(x, fp_load<FrP>(tw_lo + l));
    fp_store(out + i, fp_mul(fp_sub(x, fp_one<FrP>()), n_mont));
}

struct QuotArgs {
    const Fr *f, *t, *acc, *self, *selt, *inv_nx;
    Fr gamma, alpha, alpha2, alpha3;
    Fr zh_inv[2];
    uint64_t m;
    uint32_t shift;  // index distance of x -> w x
    uint32_t ext;
    int kind;
};

// numerator of the quotient identity at one point (without the L1 term)
__device__ __forceinline__ Fr quot_numerator(const QuotArgs& a, uint64_t i) {
    const Fr one = fp_one<FrP>();
    uint64_t j = i + a.shift;
    if (j >= a.m) j -= a.m;
    Fr acc = fp_load<FrP>(a.acc + i);
    Fr acc_next = fp_load<FrP>(a.acc + j);
    Fr fg = fp_add(fp_load<FrP>(a.f + i), a.gamma);
    Fr tg = fp_add(fp_load<FrP>(a.t + i), a.gamma);
    Fr A;
    Fr extra = fp_zero<FrP>();
    if (a.self) {
        Fr sf = fp_load<FrP>(a.self + i), st = fp_load<FrP>(a.selt + i);
        if (a.kind == KZG_GRANDSUM) {
            A = fp_mul(fp_mul(fp_sub(acc_next, acc), fg), tg);
            A = fp_add(A, fp_sub(fp_mul(st, fg), fp_mul(sf, tg)));
        } else {
            Fr nf = fp_add(fp_mul(sf, fp_sub(fg, one)), one);
            Fr dt = fp_add(fp_mul(st, fp_sub(tg, one)), one);
            A = fp_sub(fp_mul(acc_next, dt), fp_mul(acc, nf));
        }
        extra = fp_add(fp_mul(a.alpha2, fp_sub(sf, fp_sqr(sf))), fp_mul(a.alpha3, fp_sub(st, fp_sqr(st))));
    } else {
        if (a.kind == KZG_GRANDSUM) {
            A = fp_mul(fp_mul(fp_sub(acc_next, acc), fg), tg);
            A = fp_add(A, fp_sub(fg, tg));  // F - T
        } else {
            A = fp_sub(fp_mul(acc_next, tg), fp_mul(acc, fg));
        }
    }
    return fp_add(fp_mul(a.alpha, A), extra);
}

__global__ void __launch_bounds__(PR_THREADS) quotient_kernel(QuotArgs a, Fr* __restrict__ q) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.m) return;
    Fr num = quot_numerator(a, i);
    Fr acc = fp_load<FrP>(a.acc + i);
    if (a.kind == KZG_GRANDPRODUCT) acc = fp_sub(acc, fp_one<FrP>());
    // L1 * acc / Z_H = acc / (n (x - 1))
    Fr r = fp_add(fp_mul(num, a.zh_inv[i & (a.ext - 1)]), fp_mul(acc, fp_load<FrP>(a.inv_nx + i)));
    fp_store(q + i, r);
}

// numerator must vanish on H (L1 term: acc[0] is the identity by construction)
__global__ void __launch_bounds__(PR_THREADS) vanish_check_kernel(QuotArgs a, unsigned int* __restrict__ bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = false;
    if (i < a.m) nz = !fp_is_zero(quot_numerator(a, i));
    if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}

}  // namespace kzg

using namespace kzg;

static int dev_alloc(kzg_prover* p, uint64_t count, Fr** out) {
    kzg_ctx* ctx = p->ctx;
    void* d = nullptr;
    // stream-ordered pool allocation: after the first proof the pool serves these without touching the driver
    cudaError_t e = cudaMallocAsync(&d, sizeof(Fr) * (count ? count : 1), ctx->stream);
    if (e != cudaSuccess) return set_err(ctx, KZG_ERR_NOMEM, std::string("prover allocation failed: ") + cudaGetErrorString(e));
    p->owned.push_back(d);
    *out = (Fr*)d;
    return KZG_OK;
}

// the commitments of one round: independent MSMs, issued together (two lanes), fetched with one copy
static int commit_many(kzg_prover* p, const Fr* const* coefs, const uint64_t* lens, uint32_t count, uint8_t* out) {
    kzg_ctx* ctx = p->ctx;
    std::vector<MsmJob> jobs(count);
    for (uint32_t i = 0; i < count; i++) {
        jobs[i].bases = srs_bases(ctx, p->srs, 0);
        jobs[i].src = MsmScalarSrc{coefs[i], true};
        jobs[i].n = lens[i] < p->srs->n ? lens[i] : p->srs->n;
    }
    return msm_run_batch(ctx, jobs.data(), count, out);
}

static int commit_dev(kzg_prover* p, const Fr* coef, uint64_t len, uint8_t out[64]) {
    kzg_ctx* ctx = p->ctx;
    uint64_t npts = len < p->srs->n ? len : p->srs->n;
    G1XYZZ* slot = (G1XYZZ*)(ctx->dev_small + 1024);
    MsmScalarSrc src{coef, true};
    KZG_TRY(msm_run_split(ctx, srs_bases(ctx, p->srs, 0), src, npts, slot));
    return msm_result_to_host_affine(ctx, slot, 1, out);
}

static uint32_t log2u(uint64_t n) {
    uint32_t l = 0;
    while ((1ull << l) < n) l++;
    return l;
}

extern "C" {

const char* kzg_prover_last_error(kzg_prover* p) {
    return p ? kzg_last_error(p->ctx) : "null prover";
}
uint32_t kzg_prover_n_evals(kzg_prover* p) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p) return 0;
    return (p->kind == KZG_GRANDSUM ? 2 * p->k : p->k) + (p->selected ? 2 : 0) + 1;
}
uint32_t kzg_prover_n_round1_commitments(kzg_prover* p) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    return p ? 2 * p->k + (p->selected ? 2 : 0) : 0;
}

int kzg_prover_take_evals(kzg_prover* p, uint32_t column, int which, kzg_buf** out) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p || !out || column >= p->k || (which != 0 && which != 1)) return KZG_ERR_ARG;
    kzg_ctx* ctx = p->ctx;
    if (p->round < 5) return set_err(ctx, KZG_ERR_ARG, "evaluations can be taken after round 5 only");
    Fr*& slot = which == 0 ? p->ev_f[column] : p->ev_t[column];
    if (!slot) return set_err(ctx, KZG_ERR_ARG, "evaluations already taken");
    for (size_t i = 0; i < p->owned.size(); i++)
        if (p->owned[i] == (void*)slot) {
            p->owned.erase(p->owned.begin() + i);
            break;
        }
    kzg_buf* b = new kzg_buf();
    b->d = slot;
    b->n = p->n;
    slot = nullptr;
    *out = b;
    return KZG_OK;
}

int kzg_prover_destroy(kzg_prover* p) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p) return KZG_OK;
    if (p->cos_queued) cudaStreamWaitEvent(p->ctx->stream, p->cos_ready, 0);  // lane 1 may still be writing p->cos
    // (an aborted round 1 may have left uploads in flight on the copy stream: the buffers are freed after them)
    cudaEventRecord(p->ctx->ev_join, p->ctx->copy_stream);
    cudaStreamWaitEvent(p->ctx->stream, p->ctx->ev_join, 0);
    if (p->cos) cudaFreeAsync(p->cos, p->ctx->stream);
    for (void* d : p->owned) cudaFreeAsync(d, p->ctx->stream);
    if (p->cos_ready) cudaEventDestroy(p->cos_ready);
    delete p;
    return KZG_OK;
}

int kzg_prover_create(kzg_ctx* ctx, kzg_srs* srs, int kind, uint32_t n_bits, uint32_t n_pols, int selected, kzg_prover** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || !out) return KZG_ERR_ARG;
    if (kind != KZG_GRANDSUM && kind != KZG_GRANDPRODUCT) return set_err(ctx, KZG_ERR_ARG, "unknown argument kind");
    if (n_pols == 0) return set_err(ctx, KZG_ERR_PROTOCOL, "The number of multisets must be greater than 0.");
    if (n_bits < 1 || n_bits > 24) return set_err(ctx, KZG_ERR_ARG, "n_bits must be in [1, 24]");
    const uint64_t n = 1ull << n_bits;
    // the reference needs a ptau of power >= nBits (prover.js:79-81), i.e. 2n - 1 points
    if (srs->n < 2 * n - 1)
        return set_err(ctx, KZG_ERR_PROTOCOL, "The Powers of Tau file is not sufficiently large to commit the polynomials.");
    kzg_prover* p = new kzg_prover();
    p->ctx = ctx;
    p->srs = srs;
    p->kind = kind;
    p->n_bits = n_bits;
    p->k = n_pols;
    p->selected = selected != 0;
    p->n = n;
    // deg Q = 2n - 3 except for the plain grand product, where it is n - 2
    p->ext = (kind == KZG_GRANDPRODUCT && !p->selected) ? 1 : 2;
    p->m = p->ext * n;
    int r = KZG_OK;
    p->ev_f.resize(p->k);
    p->ev_t.resize(p->k);
    p->co_f.resize(p->k);
    p->co_t.resize(p->k);
    for (uint32_t i = 0; i < p->k && r == KZG_OK; i++) {
        r = dev_alloc(p, n, &p->ev_f[i]);
        if (r == KZG_OK) r = dev_alloc(p, n, &p->ev_t[i]);
        if (r == KZG_OK) r = dev_alloc(p, n, &p->co_f[i]);
        if (r == KZG_OK) r = dev_alloc(p, n, &p->co_t[i]);
    }
    if (r == KZG_OK && p->selected) {
        r = dev_alloc(p, n, &p->ev_self);
        if (r == KZG_OK) r = dev_alloc(p, n, &p->ev_selt);
        if (r == KZG_OK) r = dev_alloc(p, n, &p->co_self);
        if (r == KZG_OK) r = dev_alloc(p, n, &p->co_selt);
    }
    if (r == KZG_OK) {
        if (p->k > 1) {
            r = dev_alloc(p, n, &p->ev_fc);
            if (r == KZG_OK) r = dev_alloc(p, n, &p->ev_tc);
            if (r == KZG_OK) r = dev_alloc(p, n, &p->co_fc);
            if (r == KZG_OK) r = dev_alloc(p, n, &p->co_tc);
        } else {
            p->ev_fc = p->ev_f[0];
            p->ev_tc = p->ev_t[0];
            p->co_fc = p->co_f[0];
            p->co_tc = p->co_t[0];
        }
    }
    if (r == KZG_OK) r = dev_alloc(p, n, &p->ev_acc);
    if (r == KZG_OK) r = dev_alloc(p, n, &p->co_acc);
    if (r == KZG_OK) r = dev_alloc(p, p->m, &p->co_q);
    if (r == KZG_OK) {
        // data-independent coset table: 1 / (n (x_i - 1)),  x_i = g w_m^i,  g = w_{2m}; built once per (n, m)
        for (auto& t : ctx->coset_tables)
            if (t.n == n && t.m == p->m) p->inv_nx = t.inv_nx;
        if (!p->inv_nx) {
            Fr* tab = nullptr;
            cudaError_t e = cudaMalloc((void**)&tab, sizeof(Fr) * p->m);
            if (e != cudaSuccess) r = set_err(ctx, KZG_ERR_NOMEM, std::string("coset table allocation failed: ") + cudaGetErrorString(e));
            if (r == KZG_OK) {
                const uint32_t step = 1u << (NTT_MAX_LOG - (log2u(p->m) + 1));
                KZG_LAUNCH(ctx, coset_xm1_kernel, grid_for(p->m, PR_THREADS), PR_THREADS, 0, tab, p->m, step, h_from_u64(n),
                           ctx->tw_lo[0], ctx->tw_hi[0]);
                e = cudaGetLastError();
                if (e != cudaSuccess) r = set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
            }
            if (r == KZG_OK) r = fr_batch_inverse(ctx, tab, tab, p->m);
            if (r == KZG_OK) {
                kzg_ctx::CosetTable t;
                t.n = n;
                t.m = p->m;
                t.inv_nx = tab;
                ctx->coset_tables.push_back(t);
                p->inv_nx = tab;
            } else if (tab) {
                cudaFree(tab);
            }
        }
    }
    if (r != KZG_OK) {
        kzg_prover_destroy(p);
        return r;
    }
    *out = p;
    return KZG_OK;
}

// ---- round 1 -----------------------------------------------------------------------------------------
// Round 3 evaluates the columns F, T, S/Z (and the selectors) on the coset g H_m.  None of that needs the challenge
// alpha, so it is queued on lane 1 (auxiliary stream) as soon as S/Z exists and runs under the commitment of S/Z on
// lane 0, whose sort and bucket-reduction phases leave the integer pipe idle.
static int coset_prefetch(kzg_prover* p) {
    kzg_ctx* ctx = p->ctx;
    const uint64_t n = p->n, m = p->m;
    const uint32_t log_m = log2u(m);
    const uint32_t ncol = p->selected ? 5 : 3;
    if (!p->cos) KZG_CUDA(ctx, cudaMallocAsync((void**)&p->cos, sizeof(Fr) * m * ncol, ctx->stream));
    if (!p->cos_ready) KZG_CUDA(ctx, cudaEventCreateWithFlags(&p->cos_ready, cudaEventDisableTiming));
    const Fr* srcs[5] = {p->co_fc, p->co_tc, p->co_acc, p->co_self, p->co_selt};
    cudaStream_t main_stream = ctx->stream;
    KZG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));  // the coefficients and p->cos exist from here on
    KZG_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    ctx->lane = 1;
    ctx->stream = ctx->aux_stream;
    int r = KZG_OK;
    for (uint32_t c = 0; c < ncol && r == KZG_OK; c++) {
        Fr* dst = p->cos + (uint64_t)c * m;
        r = fr_scale_powers(ctx, srcs[c], dst, n, log_m + 1, false, nullptr);  // coef_j * g^j
        if (r == KZG_OK) r = ntt_run(ctx, dst, n, dst, log_m, false);
    }
    cudaEventRecord(p->cos_ready, ctx->aux_stream);
    ctx->lane = 0;
    ctx->stream = main_stream;
    p->cos_queued = true;
    return r;
}

int kzg_prover_round1(kzg_prover* p, const uint8_t* const* evals_f_std, const uint8_t* const* evals_t_std,
                      const uint8_t* sel_f, const uint8_t* sel_t, uint8_t* commitments_out) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p || !evals_f_std || !evals_t_std || !commitments_out) return KZG_ERR_ARG;
    kzg_ctx* ctx = p->ctx;
    if (p->selected && (!sel_f || !sel_t)) return set_err(ctx, KZG_ERR_ARG, "selected prover needs both selector columns");
    const uint64_t n = p->n;
    const size_t bytes = sizeof(Fr) * n;
    // All uploads are queued at once on the copy stream, in the order the columns are consumed, one event each; the main
    // stream waits for a column right before it converts it, so the conversions and iNTTs of column i run under the
    // uploads of the columns behind it (64 MB at n = 2^20 take 1.2 ms of PCIe time: more than the 0.6 ms of arithmetic).
    // (Splitting the F and T sides over the two lanes was measured and dropped: concurrent uploads only share the link.)
    std::vector<cudaEvent_t> up;
    {
        cudaStream_t cs = ctx->copy_stream;
        KZG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));  // the prover's buffers exist from here on
        KZG_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_fork, 0));
        auto upload = [&](Fr* dst, const uint8_t* src) -> cudaError_t {
            cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cs);
            up.push_back(order_event(ctx));
            if (e == cudaSuccess) e = cudaEventRecord(up.back(), cs);
            return e;
        };
        for (uint32_t i = 0; i < p->k; i++) {
            KZG_CUDA(ctx, upload(p->ev_f[i], evals_f_std[i]));
            KZG_CUDA(ctx, upload(p->ev_t[i], evals_t_std[i]));
        }
        if (p->selected) {
            KZG_CUDA(ctx, upload(p->ev_self, sel_f));
            KZG_CUDA(ctx, upload(p->ev_selt, sel_t));
        }
    }
    size_t next_up = 0;
    auto arrived = [&]() { cudaStreamWaitEvent(ctx->stream, up[next_up++], 0); };
    std::vector<const Fr*> coefs;
    for (uint32_t i = 0; i < p->k; i++) {
        arrived();
        KZG_TRY(fr_convert(ctx, p->ev_f[i], p->ev_f[i], n, true));  // Fr.batchToMontgomery (:147)
        KZG_TRY(ntt_run(ctx, p->ev_f[i], n, p->co_f[i], p->n_bits, true));  // Polynomial.fromEvaluations (:151)
        arrived();
        KZG_TRY(fr_convert(ctx, p->ev_t[i], p->ev_t[i], n, true));  // (:148)
        KZG_TRY(ntt_run(ctx, p->ev_t[i], n, p->co_t[i], p->n_bits, true));  // (:152)
        coefs.push_back(p->co_f[i]);  // commit (:161)
        coefs.push_back(p->co_t[i]);  // commit (:162)
    }
    if (p->selected) {
        arrived();
        KZG_TRY(ntt_run(ctx, p->ev_self, n, p->co_self, p->n_bits, true));  // (:170-171)
        arrived();
        KZG_TRY(ntt_run(ctx, p->ev_selt, n, p->co_selt, p->n_bits, true));
        coefs.push_back(p->co_self);  // commit (:173)
        coefs.push_back(p->co_selt);  // commit (:174)
    }
    // all the round's commitments are independent: issue them together, in the reference's output order
    std::vector<uint64_t> lens(coefs.size(), n);
    KZG_TRY(commit_many(p, coefs.data(), lens.data(), (uint32_t)coefs.size(), commitments_out));
    p->round = 1;
    return KZG_OK;
}

// ---- round 2 -----------------------------------------------------------------------------------------
int kzg_prover_round2(kzg_prover* p, const uint8_t beta[32], const uint8_t gamma[32], uint8_t out_acc[64]) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p || !gamma || !out_acc) return KZG_ERR_ARG;
    kzg_ctx* ctx = p->ctx;
    if (p->round < 1) return set_err(ctx, KZG_ERR_ARG, "prover rounds must run in order");
    const uint64_t n = p->n;
    p->gamma = fr_from_bytes(gamma);
    if (p->k > 1) {
        if (!beta) return set_err(ctx, KZG_ERR_ARG, "vector argument needs beta");
        p->beta = fr_from_bytes(beta);
        // F = sum_i beta^i F_i (:207-214); linear, so it is formed in both bases instead of re-transforming (:216-217)
        std::vector<Fr> pw(p->k);
        pw[0] = fp_one<FrP>();
        for (uint32_t i = 1; i < p->k; i++) pw[i] = h_mul(pw[i - 1], p->beta);
        std::vector<uint64_t> lens(p->k, n);
        std::vector<const Fr*> src(p->k);
        for (uint32_t i = 0; i < p->k; i++) src[i] = p->ev_f[i];
        KZG_TRY(poly_linear_combination(ctx, p->ev_fc, n, src.data(), lens.data(), pw.data(), p->k, fp_zero<FrP>()));
        for (uint32_t i = 0; i < p->k; i++) src[i] = p->ev_t[i];
        KZG_TRY(poly_linear_combination(ctx, p->ev_tc, n, src.data(), lens.data(), pw.data(), p->k, fp_zero<FrP>()));
        for (uint32_t i = 0; i < p->k; i++) src[i] = p->co_f[i];
        KZG_TRY(poly_linear_combination(ctx, p->co_fc, n, src.data(), lens.data(), pw.data(), p->k, fp_zero<FrP>()));
        for (uint32_t i = 0; i < p->k; i++) src[i] = p->co_t[i];
        KZG_TRY(poly_linear_combination(ctx, p->co_tc, n, src.data(), lens.data(), pw.data(), p->k, fp_zero<FrP>()));
    }
    bool ok = false;
    KZG_TRY(grand_build(ctx, p->kind, p->ev_fc, p->ev_tc, p->ev_self, p->ev_selt, p->gamma, n, p->ev_acc, &ok));
    if (!ok)
        return set_err(ctx, KZG_ERR_PROTOCOL,
                       p->kind == KZG_GRANDSUM ? "The grand-sum polynomial S is not well calculated"
                                               : "The grand-product polynomial Z is not well calculated");
    KZG_TRY(ntt_run(ctx, p->ev_acc, n, p->co_acc, p->n_bits, true));
    KZG_TRY(coset_prefetch(p));
    KZG_TRY(commit_dev(p, p->co_acc, n, out_acc));
    p->round = 2;
    return KZG_OK;
}

// ---- round 3 -----------------------------------------------------------------------------------------
int kzg_prover_round3(kzg_prover* p, const uint8_t alpha[32], uint8_t out_q[64]) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p || !alpha || !out_q) return KZG_ERR_ARG;
    kzg_ctx* ctx = p->ctx;
    if (p->round < 2) return set_err(ctx, KZG_ERR_ARG, "prover rounds must run in order");
    const uint64_t n = p->n, m = p->m;
    const uint32_t log_m = log2u(m);
    p->alpha = fr_from_bytes(alpha);

    QuotArgs a;
    memset(&a, 0, sizeof(a));
    a.gamma = p->gamma;
    a.alpha = p->alpha;
    a.alpha2 = h_mul(p->alpha, p->alpha);
    a.alpha3 = h_mul(a.alpha2, p->alpha);
    a.kind = p->kind;
    a.ext = p->ext;

    // (1) exact divisibility test on H
    a.f = p->ev_fc;
    a.t = p->ev_tc;
    a.acc = p->ev_acc;
    a.self = p->ev_self;
    a.selt = p->ev_selt;
    a.m = n;
    a.shift = 1;
    unsigned int* bad = (unsigned int*)ctx->dev_small;
    KZG_CUDA(ctx, cudaMemsetAsync(bad, 0, sizeof(unsigned int), ctx->stream));
    KZG_LAUNCH(ctx, vanish_check_kernel, grid_for(n, PR_THREADS), PR_THREADS, 0, a, bad);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned + 256, bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));

    // (2) the columns on the coset g H_m: queued on lane 1 in round 2 (coset_prefetch)
    if (!p->cos_queued) return set_err(ctx, KZG_ERR_ARG, "prover rounds must run in order");
    KZG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p->cos_ready, 0));
    Fr* cos = p->cos;
    int r = KZG_OK;
    // (3) pointwise quotient, back to coefficients, undo the coset shift
    if (r == KZG_OK) {
        a.f = cos;
        a.t = cos + m;
        a.acc = cos + 2 * m;
        a.self = p->selected ? cos + 3 * m : nullptr;
        a.selt = p->selected ? cos + 4 * m : nullptr;
        a.inv_nx = p->inv_nx;
        a.m = m;
        a.shift = p->ext;
        // Z_H(x_i) = w_{2 ext}^(2 (i mod ext) + 1) - 1
        const Fr one = fp_one<FrP>();
        if (p->ext == 1) {
            a.zh_inv[0] = fp_inv(h_sub(fp_neg(one), one));  // -2
            a.zh_inv[1] = a.zh_inv[0];
        } else {
            Fr w4 = fr_root_of_unity(2);
            a.zh_inv[0] = fp_inv(h_sub(w4, one));
            a.zh_inv[1] = fp_inv(h_sub(fp_neg(w4), one));
        }
        KZG_LAUNCH(ctx, quotient_kernel, grid_for(m, PR_THREADS), PR_THREADS, 0, a, p->co_q);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) r = set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
    }
    if (r == KZG_OK) r = ntt_run(ctx, p->co_q, m, p->co_q, log_m, true);
    if (r == KZG_OK) r = fr_scale_powers(ctx, p->co_q, p->co_q, m, log_m + 1, true, nullptr);
    cudaFreeAsync(p->cos, ctx->stream);  // (the main stream has waited for lane 1's last write)
    p->cos = nullptr;
    p->cos_queued = false;
    KZG_TRY(r);
    KZG_TRY(commit_dev(p, p->co_q, m, out_q));  // synchronises the stream
    unsigned int flag;
    memcpy(&flag, ctx->pinned + 256, sizeof(flag));
    if (flag) return set_err(ctx, KZG_ERR_PROTOCOL, "Polynomial is not divisible");
    p->round = 3;
    return KZG_OK;
}

// ---- round 4 -----------------------------------------------------------------------------------------
int kzg_prover_round4(kzg_prover* p, const uint8_t xi[32], uint8_t* evals_out) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p || !xi || !evals_out) return KZG_ERR_ARG;
    kzg_ctx* ctx = p->ctx;
    if (p->round < 3) return set_err(ctx, KZG_ERR_ARG, "prover rounds must run in order");
    p->xi = fr_from_bytes(xi);
    const Fr xiw = h_mul(p->xi, fr_root_of_unity(p->n_bits));
    std::vector<const Fr*> polys;
    std::vector<uint64_t> lens;
    std::vector<Fr> pts;
    for (uint32_t i = 0; i < p->k; i++) {
        polys.push_back(p->co_f[i]);
        pts.push_back(p->xi);
        if (p->kind == KZG_GRANDSUM) {
            polys.push_back(p->co_t[i]);
            pts.push_back(p->xi);
        }
    }
    if (p->selected) {
        polys.push_back(p->co_self);
        pts.push_back(p->xi);
        polys.push_back(p->co_selt);
        pts.push_back(p->xi);
    }
    polys.push_back(p->co_acc);
    pts.push_back(xiw);
    lens.assign(polys.size(), p->n);
    p->evals.resize(polys.size());
    KZG_TRY(poly_evaluate_multi(ctx, polys.data(), lens.data(), pts.data(), (uint32_t)polys.size(), p->evals.data()));
    for (size_t i = 0; i < p->evals.size(); i++) fr_to_bytes(p->evals[i], evals_out + 32 * i);
    p->round = 4;
    return KZG_OK;
}

// ---- round 5 -----------------------------------------------------------------------------------------
int kzg_prover_round5(kzg_prover* p, const uint8_t v_bytes[32], uint8_t out_w[128]) {
    kzg::DeviceGuard _dg(p ? p->ctx : nullptr);
    if (!p || !v_bytes || !out_w) return KZG_ERR_ARG;
    kzg_ctx* ctx = p->ctx;
    if (p->round < 4) return set_err(ctx, KZG_ERR_ARG, "prover rounds must run in order");
    p->v = fr_from_bytes(v_bytes);
    const uint64_t n = p->n, m = p->m;
    const uint32_t k = p->k;
    const bool gs = p->kind == KZG_GRANDSUM;
    const Fr one = fp_one<FrP>();
    const Fr zero = fp_zero<FrP>();
    const Fr alpha = p->alpha, gamma = p->gamma, xi = p->xi, v = p->v;
    const Fr alpha2 = h_mul(alpha, alpha), alpha3 = h_mul(alpha2, alpha);

    // unpack the evaluations (proof order)
    std::vector<Fr> fbar(k), tbar(k);
    size_t pos = 0;
    for (uint32_t i = 0; i < k; i++) {
        fbar[i] = p->evals[pos++];
        if (gs) tbar[i] = p->evals[pos++];
    }
    Fr self_xi = one, selt_xi = one;
    if (p->selected) {
        self_xi = p->evals[pos++];
        selt_xi = p->evals[pos++];
    }
    const Fr acc_xiw = p->evals[pos++];

    // Z_H(xi), L1(xi)  (polynomial_utils.js:1-19)
    Fr xin = xi;
    for (uint32_t i = 0; i < p->n_bits; i++) xin = fp_sqr(xin);
    const Fr zh = h_sub(xin, one);
    const Fr l1 = h_mul(zh, fp_inv(h_mul(h_from_u64(n), h_sub(xi, one))));

    // combined f(xi), t(xi) by linearity (the reference re-evaluates polF / polT, :358-359)
    std::vector<Fr> bpow(k);
    bpow[0] = one;
    for (uint32_t i = 1; i < k; i++) bpow[i] = h_mul(bpow[i - 1], p->beta);
    Fr fc = zero, tc = zero;
    for (uint32_t i = 0; i < k; i++) {
        fc = h_add(fc, h_mul(bpow[i], fbar[i]));
        if (gs) tc = h_add(tc, h_mul(bpow[i], tbar[i]));
    }

    // W_xi numerator = constant + sum_j coeff_j * poly_j(X)
    std::vector<const Fr*> polys;
    std::vector<uint64_t> lens;
    std::vector<Fr> coeffs;
    Fr constant = zero;
    auto term = [&](const Fr* poly, uint64_t len, const Fr& c) {
        polys.push_back(poly);
        lens.push_back(len);
        coeffs.push_back(c);
    };
    Fr sel_terms = zero;
    if (p->selected) {
        sel_terms = h_add(h_mul(alpha2, h_sub(self_xi, fp_sqr(self_xi))), h_mul(alpha3, h_sub(selt_xi, fp_sqr(selt_xi))));
    }
    if (gs) {
        const Fr fg = h_add(fc, gamma), tg = h_add(tc, gamma);
        const Fr fgtg = h_mul(fg, tg);
        Fr tail = p->selected ? h_sub(h_mul(selt_xi, fg), h_mul(self_xi, tg)) : h_sub(fc, tc);
        constant = h_add(h_mul(alpha, h_add(h_mul(acc_xiw, fgtg), tail)), sel_terms);
        term(p->co_acc, n, h_sub(l1, h_mul(alpha, fgtg)));  // S(X)
    } else {
        const Fr fg1 = p->selected ? h_add(h_mul(self_xi, h_sub(h_add(fc, gamma), one)), one) : h_add(fc, gamma);
        // alpha * zbar * (selT(xi) (T(X) + gamma - 1) + 1)
        const Fr az = h_mul(alpha, acc_xiw);
        const Fr t_coeff = p->selected ? h_mul(az, selt_xi) : az;
        const Fr t_const = p->selected ? h_add(h_mul(selt_xi, h_sub(gamma, one)), one) : gamma;
        constant = h_sub(h_add(h_mul(az, t_const), sel_terms), l1);
        for (uint32_t i = 0; i < k; i++) term(p->co_t[i], n, h_mul(t_coeff, bpow[i]));
        term(p->co_acc, n, h_sub(l1, h_mul(alpha, fg1)));  // Z(X)
    }
    term(p->co_q, m, fp_neg(zh));  // - Z_H(xi) Q(X)
    // opening terms: v^(j+1) (P_j(X) - P_j(xi))
    Fr vp = v;
    auto opening = [&](const Fr* poly, const Fr& value) {
        term(poly, n, vp);
        constant = h_sub(constant, h_mul(vp, value));
        vp = h_mul(vp, v);
    };
    for (uint32_t i = 0; i < k; i++) opening(p->co_f[i], fbar[i]);
    if (gs)
        for (uint32_t i = 0; i < k; i++) opening(p->co_t[i], tbar[i]);
    if (p->selected) {
        opening(p->co_self, self_xi);
        opening(p->co_selt, selt_xi);
    }

    const uint64_t wlen = m > n ? m : n;
    Fr* tmp = nullptr;
    KZG_CUDA(ctx, cudaMallocAsync((void**)&tmp, sizeof(Fr) * (2 * wlen + n), ctx->stream));
    Fr* wnum = tmp;
    Fr* wq = tmp + wlen;
    Fr* wq2 = tmp + 2 * wlen;
    bool exact1 = false, exact2 = false;
    int r = poly_linear_combination(ctx, wnum, wlen, polys.data(), lens.data(), coeffs.data(), (uint32_t)polys.size(), constant);
    if (r == KZG_OK) r = poly_div_x_sub(ctx, wnum, wlen, xi, wq, &exact1);
    if (r == KZG_OK && !exact1) r = set_err(ctx, KZG_ERR_PROTOCOL, "Polynomial does not divide");
    // W_xiw = (acc(X) - acc(xi w)) / (X - xi w)
    if (r == KZG_OK) {
        const Fr* ps[1] = {p->co_acc};
        uint64_t ls[1] = {n};
        Fr cs[1] = {one};
        r = poly_linear_combination(ctx, wnum, n, ps, ls, cs, 1, fp_neg(acc_xiw));
    }
    const Fr xiw = h_mul(xi, fr_root_of_unity(p->n_bits));
    if (r == KZG_OK) r = poly_div_x_sub(ctx, wnum, n, xiw, wq2, &exact2);
    if (r == KZG_OK && !exact2) r = set_err(ctx, KZG_ERR_PROTOCOL, "Polynomial does not divide");
    if (r == KZG_OK) {  // [W_xi], [W_xiw] (:409-410): two independent commitments
        const Fr* coefs[2] = {wq, wq2};
        uint64_t lens[2] = {wlen, n};
        r = commit_many(p, coefs, lens, 2, out_w);
    }
    cudaFreeAsync(tmp, ctx->stream);
    KZG_TRY(r);
    p->round = 5;
    return KZG_OK;
}

}  // extern "C"
