// Internal structures shared by the translation units of libkzgb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/kzgb200.h"
#include "ec.cuh"

// MSM tuning knobs.  Defaults are the measured optima; the KZGB200_* environment variables of the same names are read
// ONCE, when the context is created, and kzg_ctx_set_option changes them afterwards (A/B timing, the forced paths of
// tests/).  None of them changes a result -- every path ends in the same canonical affine point.
struct MsmTuning {
    int aff_rounds = -1;            // batched-affine rounds before the XYZZ walk; -1: the cost rule of msm_affine_rounds
    uint32_t aff_m = 32;            // output points per thread of a round (measured 8 / 12 / 16 / 24 / 32: 26.1 / 25.6 / 25.2 / 25.1 / 25.0 ms for the rounds at 2^24 points)
    int aff_chunks = 0;             // chunks a round is launched in (inversion hidden under the other chunks); 0: auto
    uint64_t aff_min_entries = 20ull << 20;  // no rounds below this many bucket entries
    uint64_t aff_min_left = 13ull << 19;     // a round must leave at least this many points (6.8 M: 2 rounds at 2^21 points, 3 at
                                             // 2^22 / 2^23, 4 at 2^24 -- the sweeps of r02_msm_rounds.md)
    double aff_min_fill = 6.0;      // ... and find at least this many entries per bucket
    int part_sort = -1;             // 0 / 1: direct counting sort / two-level partition sort; -1: by size
    int red_k0 = -1;                // log2 of the level-0 radix of the bucket reduction; -1: by size
    int tail_width = 0;             // threads per tail task (32 / 64 / 128); 0: by size
    int host_cut_a = 0, host_cut_b = 64;  // host-scalar MSM: piece cuts at a/64 and b/64 of the points; 0: default
    int aff_interleave = 2;         // outputs of a round dealt to the lanes of a warp one by one (2: every round, 1: dense rounds only, 0: m consecutive outputs per thread)
    int host_link = 1;              // the pieces of a host-scalar MSM over a window table share one bucket reduction (0: one each)
    uint32_t host_piece_min_log = 21;     // ... pieces from 2^this points on (2^21: the 8-GPU shard, 8.3 -> 7.5 ms e2e)
    int split_min_log = -1, split_max_log = -1;  // two-lane split of ONE msm (off: measured slower since the affine rounds)
    int merge = 1;                  // commitments of one round over one table as one merged pipeline
    int ntt_big_table = 1;          // 512 MB direct twiddle table for the first pass boundary: 0 never, 1 from 2^19 points, 2 from 2^17
    int ntt_tile = 8;               // NTT tile width in elements (8: 256-byte rows, 256 threads; 4: 128-byte rows, 128 threads)
    int timeline = 0;               // debug: print where the time of every affine round goes (events on all three streams)
};

struct kzg_ctx {
    int device = 0;
    MsmTuning tuning;
    cudaStream_t inv_stream[2] = {nullptr, nullptr};  // high priority, one per lane: the inversions of the affine rounds
    cudaStream_t side_stream[2] = {nullptr, nullptr}; // one per lane: every other chunk of an affine round (tails overlap)
    std::vector<cudaEvent_t> order_events;            // ring of ordering events (msm.cu order_event)
    size_t order_next = 0;
    std::vector<std::pair<cudaEvent_t, std::string>> timeline;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::string err;
    uint64_t launches = 0;
    uint32_t msm_window = 0;  // 0 = auto
    // optional per-kernel device timing (bench.py's live roofline): event pairs around tagged launches
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed[6];
    std::vector<cudaEvent_t> event_pool;
    // twiddle tables (device): W = w_{2^26}; lo[i] = W^i, hi[j] = W^(j * 8192); [0] forward, [1] inverse
    kzg::Fr* tw_lo[2] = {nullptr, nullptr};
    kzg::Fr* tw_hi[2] = {nullptr, nullptr};
    kzg::Fr* tw_mid[2] = {nullptr, nullptr};   // mid[e] = w_{2^16}^e: inter-pass twiddles below 2^16 in one product
    kzg::Fr* tw_mid_scaled[27] = {};           // inverse mid table times 2^-log_n (built on first use, ntt.cu)
    kzg::Fr* tw_big[2] = {nullptr, nullptr};   // big[e] = w_{2^24}^e, 512 MB per direction, built on first use (ntt.cu)
    bool ntt_attr_set = false;
    // data-independent coset tables of the fused provers, keyed by (log2 m): 1 / (n (x_i - 1)) on g H_m (prover.cu)
    struct CosetTable {
        uint64_t n = 0, m = 0;
        kzg::Fr* inv_nx = nullptr;
    };
    std::vector<CosetTable> coset_tables;
    // second lane: independent MSMs of one prover round run concurrently (their latency-bound tails overlap the
    // other one's bucket accumulation).  msm.cu's MsmLane swaps `stream` / the scratch arena for the duration of a
    // call, so every launch macro keeps using ctx->stream.
    cudaStream_t aux_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host-scalar MSMs: the piecewise upload, so that both lanes can compute under it
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_order = nullptr;  // kzg_ctx_wait_stream / kzg_stream_wait_ctx
    int lane = 0;
    bool no_split = false;  // KZGB200_NO_SPLIT=1: mid-sized MSMs are not split over the two lanes (A/B timing)
    // persistent scratch per lane (grown on demand)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* scratch2 = nullptr;
    size_t scratch2_bytes = 0;
    void* scratch3 = nullptr;       // third arena: the first of three linked pieces of a host-scalar MSM (its bucket sums must
    size_t scratch3_bytes = 0;      // outlive the third piece, which runs in lane 0's arena)
    int arena = -1;                 // arena of the next ctx_scratch call; -1: the lane's own
    // pinned host staging for small results
    uint8_t* pinned = nullptr;
    size_t pinned_bytes = 0;
    // small device slot for O(1)-sized results (commitment, evaluation values, status words)
    uint8_t* dev_small = nullptr;
    size_t dev_small_bytes = 0;
};

struct kzg_buf {
    kzg::Fr* d = nullptr;
    uint64_t n = 0;
};

struct kzg_srs {
    kzg::G1Affine* d = nullptr;
    uint64_t n = 0;
    uint32_t power = 0;
    // precomputed window table T[w][i] = 2^(tab_c * w) * P_i, w < tab_nwin, row stride n (msm.cu); optional
    kzg::G1Affine* table = nullptr;
    uint32_t tab_c = 0, tab_nwin = 0;
};

namespace kzg {

// twiddle tables: every power of W = w_{2^NTT_MAX_LOG} is hi[e >> TW_BITS] * lo[e & (TW_SIZE - 1)]
constexpr uint32_t TW_BITS = 13;
constexpr uint32_t TW_SIZE = 1u << TW_BITS;
constexpr uint32_t NTT_MAX_LOG = 2 * TW_BITS;  // 26

int set_err(kzg_ctx* ctx, int code, const std::string& msg);

// Every extern "C" entry point that takes a context (or an object that owns one) starts with this guard: the calling
// thread's current CUDA device becomes the context's for the duration of the call and is restored afterwards, so that
// several contexts on different GPUs can be driven from one process (the multi-GPU MSM of mgpu.cu, getCurveFromName(name,
// device) in the host layer).  Allocation, kernel launches and stream operations all go to the current device.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(const kzg_ctx* c) {
        if (!c) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device) changed = cudaSetDevice(c->device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define KZG_CUDA(ctx, expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return kzg::set_err(ctx, KZG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define KZG_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != KZG_OK) return _r; \
    } while (0)

// launch bookkeeping: every kernel launch in the library goes through KZG_LAUNCH so that
// kzg_ctx_launch_count() is an exact count.
#define KZG_LAUNCH(ctx, kernel, grid, block, smem, ...)                         \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);        \
        (ctx)->launches++;                                                      \
    } while (0)

#define KZG_CHECK_LAUNCH(ctx) KZG_CUDA(ctx, cudaGetLastError())

// tags for kzg_ctx_kernel_time
enum { KZG_TIMED_MSM_ACCUMULATE = 0, KZG_TIMED_NTT = 1, KZG_TIMED_MSM_SORT = 2, KZG_TIMED_MSM_REDUCE = 3,
       KZG_TIMED_MSM_FINISH = 4, KZG_TIMED_MSM_AFFINE = 5, KZG_TIMED_TAGS = 6 };
void timed_begin(kzg_ctx* ctx, int tag);
void timed_end(kzg_ctx* ctx, int tag);

int ctx_scratch(kzg_ctx* ctx, size_t bytes, void** out);
int buf_new(kzg_ctx* ctx, uint64_t n, bool zero, kzg_buf** out);

// ---- internal device-level entry points (defined across the .cu files) ----
// ntt.cu
int ntt_run(kzg_ctx* ctx, const Fr* in, uint64_t n_in, Fr* out, uint32_t log_n, bool inverse);
int ntt_init_tables(kzg_ctx* ctx);
// msm.cu
struct MsmScalarSrc {
    const Fr* scalars;  // device
    bool montgomery;    // true: convert from Montgomery on the fly (commit of a polynomial)
};
struct MsmBases {
    const G1Affine* pts;    // plain points (raw flavour)
    const G1Affine* table;  // window table, already offset to the first point of the slice (nullptr: raw flavour)
    uint64_t stride;        // table row stride in points
    uint32_t tab_c, tab_nwin;
};
int msm_run(kzg_ctx* ctx, const MsmBases& bases, MsmScalarSrc src, uint64_t n, G1XYZZ* result_dev);
int msm_run_split(kzg_ctx* ctx, const MsmBases& bases, MsmScalarSrc src, uint64_t n, G1XYZZ* result_dev);
MsmBases srs_bases(kzg_ctx* ctx, const kzg_srs* srs, uint64_t first);
int srs_precompute(kzg_ctx* ctx, kzg_srs* srs, uint32_t c);
int msm_result_to_host_affine(kzg_ctx* ctx, const G1XYZZ* result_dev, uint32_t count, uint8_t out[64]);
// several independent commitments in one go: MSM i runs on lane (i & 1); one D2H of all the affine results
struct MsmJob {
    MsmBases bases;
    MsmScalarSrc src;
    uint64_t n;
};
int msm_run_batch(kzg_ctx* ctx, const MsmJob* jobs, uint32_t count, uint8_t* out_affine);
int ctx_set_option(kzg_ctx* ctx, const char* name, long long value);
cudaEvent_t order_event(kzg_ctx* ctx);  // an event from the context's ring, for stream-to-stream ordering inside one call
// frops.cu
int fr_convert(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n, bool to_mont);
int fr_batch_inverse(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n);
int fq_batch_inverse(kzg_ctx* ctx, const Fq* in, Fq* out, uint64_t n);  // no zeros expected (0 -> 0 all the same)
int poly_degree(kzg_ctx* ctx, const Fr* a, uint64_t n, uint64_t* degree);
int poly_evaluate_multi(kzg_ctx* ctx, const Fr* const* polys, const uint64_t* lens, const Fr* points, uint32_t count,
                        Fr* out_host);
enum ScanKind { SCAN_ADD = 0, SCAN_MUL = 1 };
int fr_exclusive_scan(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n, ScanKind kind, Fr* total_host);
int poly_div_x_sub(kzg_ctx* ctx, const Fr* a, uint64_t n, const Fr& v, Fr* out, bool* exact);
int poly_linear_combination(kzg_ctx* ctx, Fr* out, uint64_t n_out, const Fr* const* polys, const uint64_t* lens,
                            const Fr* coeffs, uint32_t count, const Fr& constant);
int grand_terms(kzg_ctx* ctx, int kind, const Fr* ev_f, const Fr* ev_t, const Fr* sel_f, const Fr* sel_t,
                const Fr& gamma, Fr* num, Fr* den, uint64_t n);
int fr_mul_pointwise(kzg_ctx* ctx, const Fr* a, const Fr* b, Fr* out, uint64_t n);
int fr_fill(kzg_ctx* ctx, Fr* dst, uint64_t n, const Fr& v);
// argument.cu
int fr_scale_powers(kzg_ctx* ctx, const Fr* in, Fr* out, uint64_t n, uint32_t log_order, bool inverse, const Fr* post);
int grand_build(kzg_ctx* ctx, int kind, const Fr* ev_f, const Fr* ev_t, const Fr* sel_f, const Fr* sel_t, const Fr& gamma,
                uint64_t n, Fr* acc, bool* wrap_ok);

// host-side scalar helpers (tiny O(1)/O(log n) field work of the protocol drivers; the same
// field.cuh code compiled for the host)
Fr fr_from_bytes(const uint8_t b[32]);
void fr_to_bytes(const Fr& a, uint8_t b[32]);
Fr fr_root_of_unity(uint32_t log_n);  // Montgomery form, ffjavascript Fr.w[log_n]

}  // namespace kzg
