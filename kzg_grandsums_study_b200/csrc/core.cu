// Context, device buffers and the bulk-Fr / Polynomial C ABI of libkzgb200.so.
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace kzg {

int set_err(kzg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int ctx_set_option(kzg_ctx* ctx, const char* name, long long value) {
    MsmTuning& t = ctx->tuning;
    const MsmTuning def;
    const std::string k = name ? name : "";
    const bool reset = value < 0;
    if (k == "aff_rounds") t.aff_rounds = reset ? def.aff_rounds : (int)value;
    else if (k == "aff_m") t.aff_m = reset || value == 0 ? def.aff_m : (uint32_t)(value > 256 ? 256 : value);
    else if (k == "aff_chunks") t.aff_chunks = reset ? def.aff_chunks : (int)value;
    else if (k == "aff_min_entries_log") t.aff_min_entries = reset ? def.aff_min_entries : 1ull << (value > 40 ? 40 : value);
    else if (k == "aff_min_left_log") t.aff_min_left = reset ? def.aff_min_left : 1ull << (value > 40 ? 40 : value);
    else if (k == "aff_min_fill") t.aff_min_fill = reset ? def.aff_min_fill : (double)value;
    else if (k == "part_sort") t.part_sort = reset ? def.part_sort : (int)value;
    else if (k == "red_k0") t.red_k0 = reset ? def.red_k0 : (int)value;
    else if (k == "tail_width") t.tail_width = reset ? def.tail_width : (int)value;
    else if (k == "host_cut_a") t.host_cut_a = reset ? def.host_cut_a : (int)value;
    else if (k == "host_cut_b") t.host_cut_b = reset ? def.host_cut_b : (int)value;
    else if (k == "aff_interleave") t.aff_interleave = reset ? def.aff_interleave : (int)value;
    else if (k == "host_link") t.host_link = reset ? def.host_link : (int)value;
    else if (k == "host_piece_min_log") t.host_piece_min_log = reset ? def.host_piece_min_log : (uint32_t)value;
    else if (k == "split_min_log") t.split_min_log = reset ? def.split_min_log : (int)value;
    else if (k == "split_max_log") t.split_max_log = reset ? def.split_max_log : (int)value;
    else if (k == "msm_merge") t.merge = reset ? def.merge : (int)value;
    else if (k == "timeline") t.timeline = reset ? 0 : (int)value;
    else if (k == "ntt_tile") t.ntt_tile = reset ? def.ntt_tile : (int)value;
    else if (k == "ntt_big_table") t.ntt_big_table = reset ? def.ntt_big_table : (int)value;
    else return set_err(ctx, KZG_ERR_ARG, "unknown option: " + k);
    return KZG_OK;
}

static cudaEvent_t take_event(kzg_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void timed_begin(kzg_ctx* ctx, int tag) {
    if (!ctx->timing) return;
    cudaEvent_t a = take_event(ctx), b = take_event(ctx);
    cudaEventRecord(a, ctx->stream);
    ctx->timed[tag].push_back({a, b});
}
void timed_end(kzg_ctx* ctx, int tag) {
    if (!ctx->timing || ctx->timed[tag].empty()) return;
    cudaEventRecord(ctx->timed[tag].back().second, ctx->stream);
}

// four independent carry chains of wide multiply-accumulates per thread (the cmad4 building block of fp_mul:
// each mad.lo.cc / madc.hi.cc pair is one IMAD.WIDE.U32[.X] in SASS); 16 MACs per inner step
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t iters, uint32_t seed, unsigned long long* sink) {
#if defined(__CUDA_ARCH__)
    uint32_t x[8], y = (seed ^ 0x9e3779b9u) + blockIdx.x * 40503u + threadIdx.x;
    uint32_t a0[8], a1[8], a2[8], a3[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        x[k] = seed * (2 * k + 1) + threadIdx.x * 2654435761u;
        a0[k] = a1[k] = a2[k] = a3[k] = k;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            cmad4(a0, x[0], x[2], x[4], x[6], y);
            cmad4(a1, x[1], x[3], x[5], x[7], y);
            cmad4(a2, x[0], x[3], x[4], x[7], y);
            cmad4(a3, x[1], x[2], x[5], x[6], y);
            y += 0x9e3779b9u;
        }
    }
    unsigned long long z = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) z ^= a0[k] ^ a1[k] ^ a2[k] ^ a3[k];
    if (z == 0x1234567ull) *sink = z;  // keeps the chains live
#endif
}

// the practical ceiling of the field kernels: back-to-back Montgomery products held in registers
// (two independent chains per thread), counted as 136 limb-MACs each
__global__ void __launch_bounds__(256) modmul_peak_kernel(uint32_t iters, uint32_t seed, uint32_t* sink) {
    Fq x = fp_one<FqP>(), y = fp_r2<FqP>();
    x.l[0] += (seed + threadIdx.x) & 0xffff;
    y.l[0] ^= (blockIdx.x * 7u + seed) & 0xffff;
    Fq u = y, w = x;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            x = fp_mul(x, y);
            u = fp_mul(u, w);
        }
    }
    uint32_t z = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) z ^= x.l[k] ^ u.l[k];
    if (z == 0x1234567u) *sink = z;
}

int ctx_scratch(kzg_ctx* ctx, size_t bytes, void** out) {
    const int which = ctx->arena >= 0 ? ctx->arena : ctx->lane;
    void*& arena = which == 0 ? ctx->scratch : which == 1 ? ctx->scratch2 : ctx->scratch3;
    size_t& arena_bytes = which == 0 ? ctx->scratch_bytes : which == 1 ? ctx->scratch2_bytes : ctx->scratch3_bytes;
    if (bytes > arena_bytes) {
        // free the old block once its lane is idle, then a fresh (larger) one; rounded up to limit regrowth
        if (arena) {
            KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            KZG_CUDA(ctx, cudaFree(arena));
            arena = nullptr;
            arena_bytes = 0;
        }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&arena, want);
        if (e != cudaSuccess) {
            want = bytes;
            e = cudaMalloc(&arena, want);
        }
        if (e != cudaSuccess) return set_err(ctx, KZG_ERR_NOMEM, std::string("scratch allocation failed: ") + cudaGetErrorString(e));
        arena_bytes = want;
    }
    *out = arena;
    return KZG_OK;
}

int buf_new(kzg_ctx* ctx, uint64_t n, bool zero, kzg_buf** out) {
    kzg_buf* b = new kzg_buf();
    b->n = n;
    if (n) {
        cudaError_t e = cudaMallocAsync((void**)&b->d, sizeof(Fr) * n, ctx->stream);
        if (e != cudaSuccess) {
            delete b;
            return set_err(ctx, KZG_ERR_NOMEM, std::string("device allocation failed: ") + cudaGetErrorString(e));
        }
        if (zero) {
            e = cudaMemsetAsync(b->d, 0, sizeof(Fr) * n, ctx->stream);
            if (e != cudaSuccess) {
                cudaFreeAsync(b->d, ctx->stream);
                delete b;
                return set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
            }
        }
    }
    *out = b;
    return KZG_OK;
}

// ---- self test kernels --------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <class P> __device__ Fp<P> random_fp(uint64_t& s) {
    Fp<P> r;
    for (int i = 0; i < 8; i += 2) {
        uint64_t z = splitmix(s);
        r.l[i] = (uint32_t)z;
        r.l[i + 1] = (uint32_t)(z >> 32);
    }
    r.l[7] &= 0x1fffffffu;  // < 2^253 < p
    return r;
}
template <class P> __device__ bool selftest_field(uint64_t& s) {
    Fp<P> a = random_fp<P>(s), b = random_fp<P>(s), c = random_fp<P>(s);
    bool ok = true;
    ok &= fp_eq(fp_mul(a, b), fp_mul_portable(a, b));
    ok &= fp_eq(fp_mul(a, a), fp_mul_portable(a, a));
    ok &= fp_eq(fp_sqr(a), fp_mul_portable(a, a));  // the dedicated squaring (100 wide MACs)
    ok &= fp_eq(fp_sqr(fp_sub(b, c)), fp_mul_portable(fp_sub(b, c), fp_sub(b, c)));
    {   // two products under one reduction, incl. the operands that maximise every intermediate total
        const Fp<P> d = fp_sub(a, c), m1 = fp_neg(fp_one<P>());
        ok &= fp_eq(fp_mul2(a, b, c, d), fp_add(fp_mul_portable(a, b), fp_mul_portable(c, d)));
        ok &= fp_eq(fp_mul2_sub(a, b, c, d), fp_sub(fp_mul_portable(a, b), fp_mul_portable(c, d)));
        ok &= fp_eq(fp_mul2(m1, m1, m1, m1), fp_dbl(fp_mul_portable(m1, m1)));
        ok &= fp_eq(fp_mul2_sub(a, b, fp_zero<P>(), d), fp_mul_portable(a, b));
        ok &= fp_is_zero(fp_mul2_sub(a, b, a, b));
    }
    {   // operands made of extreme limbs: every carry of the squaring's chains
        Fp<P> e = a;
        const uint64_t z = splitmix(s);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t pick = (uint32_t)(z >> (4 * i)) & 7u;
            if (pick == 0) e.l[i] = 0xffffffffu;
            else if (pick == 1) e.l[i] = 0;
            else if (pick == 2) e.l[i] = 0x80000000u;
            else if (pick == 3) e.l[i] = 1;
        }
        e.l[7] &= 0x1fffffffu;
        ok &= fp_eq(fp_sqr(e), fp_mul_portable(e, e));
    }
    ok &= fp_eq(fp_mul(fp_add(a, b), c), fp_add(fp_mul(a, c), fp_mul(b, c)));
    ok &= fp_eq(fp_sub(fp_add(a, b), b), a);
    ok &= fp_eq(fp_from_mont(fp_to_mont(a)), a);
    return ok;
}
__global__ void selftest_kernel(uint32_t n, unsigned int* fail) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = 0x1234567ull + 977ull * i;
    bool ok = selftest_field<FqP>(s) && selftest_field<FrP>(s);
    // edge operands: p-1, 0, 1
    Fq m1 = fp_neg(fp_one<FqP>());
    ok &= fp_eq(fp_mul(m1, m1), fp_one<FqP>());
    ok &= fp_eq(fp_mul(m1, m1), fp_mul_portable(m1, m1));
    ok &= fp_eq(fp_sqr(m1), fp_one<FqP>()) && fp_is_zero(fp_sqr(fp_zero<FqP>())) && fp_eq(fp_sqr(fp_one<FqP>()), fp_one<FqP>());
    ok &= fp_eq(fp_sqr(fp_neg(fp_one<FrP>())), fp_one<FrP>()) && fp_eq(fp_sqr(fp_one<FrP>()), fp_one<FrP>());
    ok &= fp_is_zero(fp_mul(fp_zero<FqP>(), m1));
    if ((i & 63) == 0) {
        Fq a = random_fp<FqP>(s);
        if (!fp_is_zero(a)) ok &= fp_eq(fp_mul(a, fp_inv(a)), fp_one<FqP>());
        ok &= fp_eq(fp_inv(a), fp_inv_fermat(a));  // binary GCD (31 steps per update) == a^(p-2)
        ok &= fp_eq(fp_inv_euclid(a), fp_inv_fermat(a));  // plain binary extended Euclid (the fallback)
        Fr b = random_fp<FrP>(s);
        ok &= fp_eq(fp_inv(b), fp_inv_fermat(b));
        ok &= fp_is_zero(fp_inv(fp_zero<FrP>())) && fp_eq(fp_inv(fp_one<FqP>()), fp_one<FqP>());
        ok &= fp_eq(fp_inv(fp_neg(fp_one<FqP>())), fp_neg(fp_one<FqP>()));
        // group law on the generator (1, 2): 2G + G == G + G + G, (G + G) via madd hits the doubling branch
        G1Affine g;
        g.x = fp_one<FqP>();
        g.y = fp_dbl(fp_one<FqP>());
        ok &= g1_affine_on_curve(g);
        G1XYZZ x = xyzz_inf();
        xyzz_madd(x, g);
        xyzz_madd(x, g);          // doubling branch
        G1XYZZ d = xyzz_dbl_affine(g);
        G1Affine xa = xyzz_to_affine(x), da = xyzz_to_affine(d);
        ok &= fp_eq(xa.x, da.x) && fp_eq(xa.y, da.y) && g1_affine_on_curve(xa);
        xyzz_madd(x, g);          // 3G
        G1XYZZ y = d;
        xyzz_add(y, xyzz_from_affine(g));
        G1Affine ya = xyzz_to_affine(y);
        xa = xyzz_to_affine(x);
        ok &= fp_eq(xa.x, ya.x) && fp_eq(xa.y, ya.y) && g1_affine_on_curve(xa);
        xyzz_madd(x, g1_affine_neg(g));  // back to 2G
        xa = xyzz_to_affine(x);
        ok &= fp_eq(xa.x, da.x) && fp_eq(xa.y, da.y);
        G1XYZZ z = xyzz_mul_small(xyzz_from_affine(g), 3);
        G1Affine za = xyzz_to_affine(z);
        ok &= fp_eq(za.x, ya.x) && fp_eq(za.y, ya.y);
        G1XYZZ o = xyzz_from_affine(g);
        xyzz_madd(o, g1_affine_neg(g));  // G - G = inf
        ok &= xyzz_is_inf(o);
    }
    if (!ok) atomicAdd(fail, 1u);
}

}  // namespace kzg

using namespace kzg;

extern "C" {

int kzg_ctx_create(int device, void* stream, kzg_ctx** out) {
    if (!out) return KZG_ERR_ARG;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return KZG_ERR_CUDA;  // no CPU fallback
    kzg_ctx* ctx = new kzg_ctx();
    ctx->device = device;
    struct Restore {  // the caller's current device is left as it was (several contexts per process)
        int prev = -1;
        Restore() { cudaGetDevice(&prev); }
        ~Restore() {
            if (prev >= 0) cudaSetDevice(prev);
        }
    } restore;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return KZG_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return KZG_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return KZG_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    if (cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return KZG_ERR_CUDA;
    }
    {
        int lo = 0, hi = 0;  // (numerically lowest = greatest priority)
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int l = 0; l < 2; l++)
            if (cudaStreamCreateWithPriority(&ctx->inv_stream[l], cudaStreamNonBlocking, hi) != cudaSuccess ||
                cudaStreamCreateWithFlags(&ctx->side_stream[l], cudaStreamNonBlocking) != cudaSuccess) {
                delete ctx;
                return KZG_ERR_CUDA;
            }
    }
    ctx->no_split = getenv("KZGB200_NO_SPLIT") != nullptr;
    // the tuning knobs, read once (kzg_ctx_set_option changes them later)
    static const char* const knobs[] = {"aff_rounds", "aff_m", "aff_chunks", "aff_min_entries_log", "aff_min_left_log",
                                        "aff_min_fill", "part_sort", "red_k0", "tail_width", "host_piece_min_log", "host_link", "aff_interleave",
                                        "split_min_log", "split_max_log", "msm_merge", "timeline", "ntt_tile", "ntt_big_table"};
    for (const char* k : knobs) {
        std::string env = "KZGB200_";
        for (const char* c = k; *c; c++) env += (char)toupper(*c);
        if (const char* ov = getenv(env.c_str())) ctx_set_option(ctx, k, atoll(ov));
    }
    if (const char* ov = getenv("KZGB200_HOST_PIECES")) {  // "a,b" = cuts at a/64 and b/64 of n (b = 64: two pieces)
        int a = 0, b = 64;
        if (sscanf(ov, "%d,%d", &a, &b) >= 1 && a > 0 && a < b && b <= 64) {
            ctx->tuning.host_cut_a = a;
            ctx->tuning.host_cut_b = b;
        }
    }
    if (const char* ov = getenv("KZGB200_L2_FETCH")) {  // experiment: bytes the L2 fetches from DRAM per miss
        size_t before = 0, after = 0;
        cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
        cudaError_t le = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(ov));
        cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "[kzgb200] L2 fetch granularity %zu -> %zu (%s)\n", before, after, cudaGetErrorString(le));
    }
    ctx->pinned_bytes = 1 << 16;
    ctx->dev_small_bytes = 1 << 16;
    if (cudaMallocHost((void**)&ctx->pinned, ctx->pinned_bytes) != cudaSuccess ||
        cudaMalloc((void**)&ctx->dev_small, ctx->dev_small_bytes) != cudaSuccess) {
        delete ctx;
        return KZG_ERR_CUDA;
    }
    // keep freed stream-ordered allocations cached in the pool instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    int r = ntt_init_tables(ctx);
    if (r != KZG_OK) {
        delete ctx;
        return r;
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        delete ctx;
        return KZG_ERR_CUDA;
    }
    *out = ctx;
    return KZG_OK;
}

int kzg_ctx_destroy(kzg_ctx* ctx) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int d = 0; d < 2; d++) {
        cudaFree(ctx->tw_lo[d]);
        cudaFree(ctx->tw_hi[d]);
        cudaFree(ctx->tw_mid[d]);
    }
    for (auto* t : ctx->tw_mid_scaled) cudaFree(t);
    for (auto* t : ctx->tw_big) cudaFree(t);
    for (int tag = 0; tag < KZG_TIMED_TAGS; tag++)
        for (auto& pr : ctx->timed[tag]) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    for (auto& t : ctx->coset_tables) cudaFree(t.inv_nx);
    cudaFree(ctx->scratch);
    cudaFree(ctx->scratch2);
    cudaFree(ctx->scratch3);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->aux_stream) {
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamDestroy(ctx->aux_stream);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    for (int l = 0; l < 2; l++) {
        if (ctx->inv_stream[l]) {
            cudaStreamSynchronize(ctx->inv_stream[l]);
            cudaStreamDestroy(ctx->inv_stream[l]);
        }
        if (ctx->side_stream[l]) {
            cudaStreamSynchronize(ctx->side_stream[l]);
            cudaStreamDestroy(ctx->side_stream[l]);
        }
    }
    for (cudaEvent_t e : ctx->order_events) cudaEventDestroy(e);
    cudaFree(ctx->dev_small);
    cudaFreeHost(ctx->pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return KZG_OK;
}

int kzg_ctx_sync(kzg_ctx* ctx) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_ERR_ARG;
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KZG_OK;
}

int kzg_ctx_set_option(kzg_ctx* ctx, const char* name, int64_t value) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !name) return KZG_ERR_ARG;
    return ctx_set_option(ctx, name, (long long)value);
}

// stream ordering against a caller-owned stream (the multi-GPU MSM exchanges the partial points with a collective
// that runs on the caller's stream): no host synchronisation, one event each way
int kzg_ctx_wait_stream(kzg_ctx* ctx, void* stream) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_ERR_ARG;
    if ((cudaStream_t)stream == ctx->stream) return KZG_OK;
    KZG_CUDA(ctx, cudaEventRecord(ctx->ev_order, (cudaStream_t)stream));
    KZG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_order, 0));
    return KZG_OK;
}
int kzg_stream_wait_ctx(kzg_ctx* ctx, void* stream) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_ERR_ARG;
    if ((cudaStream_t)stream == ctx->stream) return KZG_OK;
    KZG_CUDA(ctx, cudaEventRecord(ctx->ev_order, ctx->stream));
    KZG_CUDA(ctx, cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_order, 0));
    return KZG_OK;
}

const char* kzg_last_error(kzg_ctx* ctx) {
    kzg::DeviceGuard _dg(ctx);
    return ctx ? ctx->err.c_str() : "null context";
}
uint64_t kzg_ctx_launch_count(kzg_ctx* ctx) {
    kzg::DeviceGuard _dg(ctx);
    return ctx ? ctx->launches : 0;
}

// quad-lane group law (ec.cuh) against the scalar formulas: generic sums, doublings through the addition,
// cancellation, infinities, small multiples.  One case per quad.
__global__ void selftest_quad_kernel(uint32_t n, unsigned int* fail) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = t >> 2, j = t & 3;
    if (q >= n) return;
    const uint32_t qm = quad_mask();
    G1Affine g;
    g.x = fp_one<FqP>();
    g.y = fp_dbl(fp_one<FqP>());
    const G1XYZZ base = xyzz_from_affine(g);
    // A = ka G, B = kb G in non-trivial XYZZ representations
    const uint32_t ka = 2 + (q * 7919u) % 1021u, kb = 2 + (q * 104729u) % 509u;
    G1XYZZ A = xyzz_mul_small(base, ka), B = xyzz_mul_small(base, kb);
    const uint32_t kind = q % 8;
    if (kind == 1) B = A;                                   // doubling through the addition
    if (kind == 2) { B = A; B.y = fp_neg(B.y); }            // cancellation
    if (kind == 3) A = xyzz_inf();
    if (kind == 4) B = xyzz_inf();
    if (kind == 5) { B = xyzz_mul_small(base, ka); }        // same point, same representation path
    G1XYZZ want = A;
    xyzz_add(want, B);
    Fq a = quad_coord(A, j);
    quad_add(a, quad_coord(B, j), j, qm);
    bool ok = fp_eq(a, quad_coord(want, j)) || (xyzz_is_inf(want) && j != 2);  // infinity: only ZZ = 0 is specified
    if (xyzz_is_inf(want) && j == 2) ok = fp_is_zero(a);
    G1XYZZ d = xyzz_dbl(A);
    Fq a2 = quad_coord(A, j);
    quad_dbl(a2, j, qm);
    ok &= xyzz_is_inf(d) ? (j != 2 || fp_is_zero(a2)) : fp_eq(a2, quad_coord(d, j));
    const uint32_t k = q % 300u;
    G1XYZZ m = xyzz_mul_small(B, k);
    Fq a3 = quad_coord(B, j);
    quad_mul_small(a3, k, j, qm);
    ok &= xyzz_is_inf(m) ? (j != 2 || fp_is_zero(a3)) : fp_eq(a3, quad_coord(m, j));
    if (!ok) atomicAdd(fail, 1u);
}

int kzg_selftest(kzg_ctx* ctx, uint32_t n_cases) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_ERR_ARG;
    unsigned int* slot = (unsigned int*)ctx->dev_small;
    KZG_CUDA(ctx, cudaMemsetAsync(slot, 0, sizeof(unsigned int), ctx->stream));
    KZG_LAUNCH(ctx, selftest_kernel, (n_cases + 127) / 128, 128, 0, n_cases, slot);
    {
        const uint32_t quads = n_cases < 2048 ? n_cases : 2048;  // ~300 group operations per case
        KZG_LAUNCH(ctx, selftest_quad_kernel, (quads * 4 + 127) / 128, 128, 0, quads, slot);
    }
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, slot, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned int fails;
    memcpy(&fails, ctx->pinned, sizeof(fails));
    if (fails) return set_err(ctx, KZG_ERR_CUDA, "device self-test failed in " + std::to_string(fails) + " cases");
    return KZG_OK;
}

// which = 0: raw IMAD.WIDE.U32 chains; which = 1: fp_mul chains (reported as 136 MACs per product)
static int bench_peak(kzg_ctx* ctx, int which, uint32_t ms, double* macs_per_second) {
    cudaEvent_t a, b;
    KZG_CUDA(ctx, cudaEventCreate(&a));
    KZG_CUDA(ctx, cudaEventCreate(&b));
    const uint32_t blocks = (uint32_t)ctx->sm_count * 8, threads = 256;
    void* sink = (void*)(ctx->dev_small + 2048);
    uint32_t iters = which == 0 ? 1024 : 64;
    const double macs_per_iter = which == 0 ? 64.0 : 8.0 * 136.0;
    double best = 0;
    float total_ms = 0;
    // warm up, then repeat launches until `ms` of kernel time has been spent; report the best launch
    for (int rep = 0; rep < 64 && (rep < 3 || total_ms < (float)ms); rep++) {
        KZG_CUDA(ctx, cudaEventRecord(a, ctx->stream));
        if (which == 0)
            KZG_LAUNCH(ctx, imad_peak_kernel, blocks, threads, 0, iters, 12345u + rep, (unsigned long long*)sink);
        else
            KZG_LAUNCH(ctx, modmul_peak_kernel, blocks, threads, 0, iters, 12345u + rep, (uint32_t*)sink);
        KZG_CUDA(ctx, cudaEventRecord(b, ctx->stream));
        KZG_CUDA(ctx, cudaEventSynchronize(b));
        float t = 0;
        KZG_CUDA(ctx, cudaEventElapsedTime(&t, a, b));
        if (rep >= 1) {
            total_ms += t;
            double rate = (double)blocks * threads * (double)iters * macs_per_iter / (t * 1e-3);
            if (rate > best) best = rate;
        }
        if (t < 2.0f && iters < (1u << 22)) iters *= 4;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *macs_per_second = best;
    return KZG_OK;
}

int kzg_bench_imad_peak(kzg_ctx* ctx, uint32_t ms, double* macs_per_second) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !macs_per_second) return KZG_ERR_ARG;
    return bench_peak(ctx, 0, ms, macs_per_second);
}
int kzg_bench_modmul_peak(kzg_ctx* ctx, uint32_t ms, double* macs_per_second) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !macs_per_second) return KZG_ERR_ARG;
    return bench_peak(ctx, 1, ms, macs_per_second);
}

int kzg_ctx_kernel_time(kzg_ctx* ctx, uint32_t which, int reset, double* ms_out, uint64_t* launches_out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || which >= KZG_TIMED_TAGS) return KZG_ERR_ARG;
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double total = 0;
    for (auto& pr : ctx->timed[which]) {
        float t = 0;
        if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) total += t;
    }
    if (ms_out) *ms_out = total;
    if (launches_out) *launches_out = ctx->timed[which].size();
    if (reset) {
        for (int tag = 0; tag < KZG_TIMED_TAGS; tag++) {
            for (auto& pr : ctx->timed[tag]) {
                ctx->event_pool.push_back(pr.first);
                ctx->event_pool.push_back(pr.second);
            }
            ctx->timed[tag].clear();
        }
        ctx->timing = reset > 0;  // reset = 1: clear and (keep) timing on; reset = -1: clear and switch off
    }
    return KZG_OK;
}

// ---- buffers ------------------------------------------------------------------------------------
int kzg_buf_alloc(kzg_ctx* ctx, uint64_t n, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !out) return KZG_ERR_ARG;
    return buf_new(ctx, n, true, out);
}
int kzg_buf_free(kzg_ctx* ctx, kzg_buf* b) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !b) return KZG_OK;
    if (b->d) cudaFreeAsync(b->d, ctx->stream);
    delete b;
    return KZG_OK;
}
uint64_t kzg_buf_len(kzg_buf* b) { return b ? b->n : 0; }
void* kzg_buf_device_ptr(kzg_buf* b) { return b ? (void*)b->d : nullptr; }

int kzg_buf_upload(kzg_ctx* ctx, kzg_buf* dst, uint64_t dst_off, const uint8_t* host, uint64_t n) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !dst || (!host && n)) return KZG_ERR_ARG;
    if (dst_off + n > dst->n) return set_err(ctx, KZG_ERR_ARG, "upload out of bounds");
    if (n == 0) return KZG_OK;
    KZG_CUDA(ctx, cudaMemcpyAsync(dst->d + dst_off, host, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host buffer may be pageable / reused by the caller
    return KZG_OK;
}
int kzg_buf_download(kzg_ctx* ctx, kzg_buf* src, uint64_t src_off, uint8_t* host, uint64_t n) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !src || (!host && n)) return KZG_ERR_ARG;
    if (src_off + n > src->n) return set_err(ctx, KZG_ERR_ARG, "download out of bounds");
    if (n == 0) return KZG_OK;
    KZG_CUDA(ctx, cudaMemcpyAsync(host, src->d + src_off, sizeof(Fr) * n, cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KZG_OK;
}
int kzg_buf_copy(kzg_ctx* ctx, kzg_buf* dst, uint64_t dst_off, kzg_buf* src, uint64_t src_off, uint64_t n) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !dst || !src) return KZG_ERR_ARG;
    if (dst_off + n > dst->n || src_off + n > src->n) return set_err(ctx, KZG_ERR_ARG, "copy out of bounds");
    if (n == 0) return KZG_OK;
    KZG_CUDA(ctx, cudaMemcpyAsync(dst->d + dst_off, src->d + src_off, sizeof(Fr) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    return KZG_OK;
}
int kzg_buf_fill(kzg_ctx* ctx, kzg_buf* dst, uint64_t off, uint64_t n, const uint8_t value[32]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !dst || !value) return KZG_ERR_ARG;
    if (off + n > dst->n) return set_err(ctx, KZG_ERR_ARG, "fill out of bounds");
    return fr_fill(ctx, dst->d + off, n, fr_from_bytes(value));
}

// ---- bulk Fr ------------------------------------------------------------------------------------
int kzg_fr_to_mont(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !in || !out || in->n != out->n) return KZG_ERR_ARG;
    return fr_convert(ctx, in->d, out->d, in->n, true);
}
int kzg_fr_from_mont(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !in || !out || in->n != out->n) return KZG_ERR_ARG;
    return fr_convert(ctx, in->d, out->d, in->n, false);
}
static int log2_exact(uint64_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1ull << l) < n) l++;
    return l;
}
int kzg_fr_ntt(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out, int inverse) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !in || !out || in->n != out->n) return KZG_ERR_ARG;
    int lg = log2_exact(in->n);
    if (lg < 0) return set_err(ctx, KZG_ERR_PROTOCOL, "fft must be multiple of 2");
    return ntt_run(ctx, in->d, in->n, out->d, (uint32_t)lg, inverse != 0);
}
int kzg_fr_extend_ntt(kzg_ctx* ctx, kzg_buf* coef, uint32_t extension, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !coef || !out || extension == 0 || (extension & (extension - 1))) return KZG_ERR_ARG;
    uint32_t power = 0;
    while ((1ull << power) < coef->n) power++;
    uint64_t len = (1ull << power) * extension;
    int lg = log2_exact(len);
    kzg_buf* o = nullptr;
    KZG_TRY(buf_new(ctx, len, false, &o));
    int r = ntt_run(ctx, coef->d, coef->n, o->d, (uint32_t)lg, false);
    if (r != KZG_OK) {
        kzg_buf_free(ctx, o);
        return r;
    }
    *out = o;
    return KZG_OK;
}
int kzg_fr_batch_inverse(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !in || !out || in->n != out->n) return KZG_ERR_ARG;
    return fr_batch_inverse(ctx, in->d, out->d, in->n);
}

// ---- Polynomial ---------------------------------------------------------------------------------
static int poly_addsub(kzg_ctx* ctx, kzg_buf* a, kzg_buf* b, kzg_buf** out, bool sub) {
    if (!ctx || !a || !b || !out) return KZG_ERR_ARG;
    uint64_t n = a->n > b->n ? a->n : b->n;
    kzg_buf* o = nullptr;
    KZG_TRY(buf_new(ctx, n, false, &o));
    const Fr* polys[2] = {a->d, b->d};
    uint64_t lens[2] = {a->n, b->n};
    Fr coeffs[2] = {fp_one<FrP>(), sub ? fp_neg(fp_one<FrP>()) : fp_one<FrP>()};
    int r = poly_linear_combination(ctx, o->d, n, polys, lens, coeffs, 2, fp_zero<FrP>());
    if (r != KZG_OK) {
        kzg_buf_free(ctx, o);
        return r;
    }
    *out = o;
    return KZG_OK;
}
int kzg_poly_add(kzg_ctx* ctx, kzg_buf* a, kzg_buf* b, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    return poly_addsub(ctx, a, b, out, false);
}
int kzg_poly_sub(kzg_ctx* ctx, kzg_buf* a, kzg_buf* b, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    return poly_addsub(ctx, a, b, out, true);
}

int kzg_poly_mul_scalar(kzg_ctx* ctx, kzg_buf* a, const uint8_t s[32]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !a || !s) return KZG_ERR_ARG;
    const Fr* polys[1] = {a->d};
    uint64_t lens[1] = {a->n};
    Fr coeffs[1] = {fr_from_bytes(s)};
    return poly_linear_combination(ctx, a->d, a->n, polys, lens, coeffs, 1, fp_zero<FrP>());
}
static int poly_addsub_scalar(kzg_ctx* ctx, kzg_buf* a, const uint8_t s[32], bool sub) {
    if (!ctx || !a || !s) return KZG_ERR_ARG;
    if (a->n == 0) return set_err(ctx, KZG_ERR_ARG, "addScalar on an empty polynomial");
    const Fr* polys[1] = {a->d};
    uint64_t lens[1] = {1};
    Fr coeffs[1] = {fp_one<FrP>()};
    Fr c = fr_from_bytes(s);
    if (sub) c = fp_neg(c);
    return poly_linear_combination(ctx, a->d, 1, polys, lens, coeffs, 1, c);
}
int kzg_poly_add_scalar(kzg_ctx* ctx, kzg_buf* a, const uint8_t s[32]) {
    kzg::DeviceGuard _dg(ctx);
    return poly_addsub_scalar(ctx, a, s, false);
}
int kzg_poly_sub_scalar(kzg_ctx* ctx, kzg_buf* a, const uint8_t s[32]) {
    kzg::DeviceGuard _dg(ctx);
    return poly_addsub_scalar(ctx, a, s, true);
}

int kzg_poly_degree(kzg_ctx* ctx, kzg_buf* a, uint64_t* degree) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !a || !degree) return KZG_ERR_ARG;
    return poly_degree(ctx, a->d, a->n, degree);
}
int kzg_poly_evaluate(kzg_ctx* ctx, kzg_buf* a, const uint8_t x[32], uint8_t out[32]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !a || !x || !out) return KZG_ERR_ARG;
    const Fr* polys[1] = {a->d};
    uint64_t lens[1] = {a->n};
    Fr pts[1] = {fr_from_bytes(x)};
    Fr res;
    KZG_TRY(poly_evaluate_multi(ctx, polys, lens, pts, 1, &res));
    fr_to_bytes(res, out);
    return KZG_OK;
}
int kzg_poly_multiply(kzg_ctx* ctx, kzg_buf* a, kzg_buf* b, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !a || !b || !out) return KZG_ERR_ARG;
    uint64_t da = 0, db = 0;
    KZG_TRY(poly_degree(ctx, a->d, a->n, &da));
    KZG_TRY(poly_degree(ctx, b->d, b->n, &db));
    uint32_t lg = 0;
    while ((1ull << lg) < da + db + 1) lg++;
    const uint64_t len = 1ull << lg;
    kzg_buf *fa = nullptr, *fb = nullptr;
    KZG_TRY(buf_new(ctx, len, false, &fa));
    KZG_TRY(buf_new(ctx, len, false, &fb));
    int r = ntt_run(ctx, a->d, da + 1 < a->n ? da + 1 : a->n, fa->d, lg, false);
    if (r == KZG_OK) r = ntt_run(ctx, b->d, db + 1 < b->n ? db + 1 : b->n, fb->d, lg, false);
    if (r == KZG_OK) r = fr_mul_pointwise(ctx, fa->d, fb->d, fa->d, len);
    if (r == KZG_OK) r = ntt_run(ctx, fa->d, len, fb->d, lg, true);
    kzg_buf_free(ctx, fa);
    if (r != KZG_OK) {
        kzg_buf_free(ctx, fb);
        return r;
    }
    *out = fb;
    return KZG_OK;
}
int kzg_poly_lagrange1(kzg_ctx* ctx, uint32_t power, kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !out || power > 26) return KZG_ERR_ARG;
    // iNTT of e_0: every coefficient equals n^-1
    Fr n = fp_zero<FrP>();
    uint64_t N = 1ull << power;
    n.l[0] = (uint32_t)N;
    n.l[1] = (uint32_t)(N >> 32);
    Fr ninv = fp_inv(fp_to_mont(n));
    kzg_buf* o = nullptr;
    KZG_TRY(buf_new(ctx, N, false, &o));
    int r = fr_fill(ctx, o->d, N, ninv);
    if (r != KZG_OK) {
        kzg_buf_free(ctx, o);
        return r;
    }
    *out = o;
    return KZG_OK;
}
int kzg_poly_div_x_sub_value(kzg_ctx* ctx, kzg_buf* a, const uint8_t v[32], kzg_buf** out) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !a || !v || !out) return KZG_ERR_ARG;
    if (a->n < 2) return set_err(ctx, KZG_ERR_ARG, "divByXSubValue needs at least two coefficients");
    kzg_buf* o = nullptr;
    KZG_TRY(buf_new(ctx, a->n, false, &o));
    bool exact = false;
    int r = poly_div_x_sub(ctx, a->d, a->n, fr_from_bytes(v), o->d, &exact);
    if (r == KZG_OK && !exact) r = set_err(ctx, KZG_ERR_PROTOCOL, "Polynomial does not divide");
    if (r != KZG_OK) {
        kzg_buf_free(ctx, o);
        return r;
    }
    *out = o;
    return KZG_OK;
}

}  // extern "C"
