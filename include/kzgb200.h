/* kzgb200.h -- C ABI of libkzgb200.so: the B200-native (sm_100a) backend for the data-parallel hot path of
 * xavi-pinsach/kzg-grandsums-study (KZG grand-sum / grand-product multiset-equality provers over BN254).
 *
 * The reference is plain JavaScript and has no FFI; its de-facto backend boundary is the set of bulk
 * `curve.Fr.*` / `curve.G1.*` calls it makes into ffjavascript plus the Polynomial / Evaluations classes
 * built on them.  Each entry point below names the reference interface it replaces (file:line under the
 * reference tree).  A thin N-API addon (addon/kzgb200_napi.cc) or the Python ctypes
 * layer (kzg_grandsums_study_b200/_lib.py) binds exactly these symbols.
 *
 * Conventions
 *  - every function returns 0 on success or a negative kzg_status; kzg_last_error(ctx) gives the message
 *    (the host layers rethrow it verbatim -- the prover messages are the reference's own strings).
 *  - field elements are 32-byte little-endian Montgomery residues of BN254 Fr unless a name says `_std`;
 *    G1 points are 64-byte affine (x || y, Montgomery-LE Fq), infinity = 64 zero bytes.  These are the
 *    byte layouts of the reference's `.coef`, `.eval`, ptau section 2 and `proof.*` buffers.
 *  - a context owns one CUDA device + stream and is NOT re-entrant (one in-flight call per context).
 *  - no entry point has a CPU fallback: without a usable CUDA device kzg_ctx_create fails.
 */
#ifndef KZGB200_H
#define KZGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kzg_ctx kzg_ctx;       /* device + stream + scratch; mirrors the `curve` object (getCurveFromQ) */
typedef struct kzg_srs kzg_srs;       /* device-resident [tau^i]_1; mirrors the PTau BigBuffer (prover.js:83-85) */
typedef struct kzg_buf kzg_buf;       /* device Fr vector; mirrors a .coef / .eval buffer */
typedef struct kzg_prover kzg_prover; /* state of one 5-round proof (prover.js:95-106) */

typedef enum {
    KZG_OK = 0,
    KZG_ERR_CUDA = -1,
    KZG_ERR_ARG = -2,
    KZG_ERR_IO = -3,
    KZG_ERR_FORMAT = -4,   /* bad .ptau container / header (ptau_utils.js:4-21) */
    KZG_ERR_PROTOCOL = -5, /* a reference `throw new Error(...)` condition; message is the reference's */
    KZG_ERR_NOMEM = -6
} kzg_status;

enum { KZG_GRANDSUM = 0, KZG_GRANDPRODUCT = 1 };

/* ---- context ------------------------------------------------------------------------------------------ */
/* replaces getCurveFromQ / curve.terminate (ptau_utils.js:13; test/mset_eq_kzg_grandsum.test.js:14-21).
 * `stream` is a cudaStream_t (NULL = create a private non-blocking stream). */
int kzg_ctx_create(int device, void* stream, kzg_ctx** out);
int kzg_ctx_destroy(kzg_ctx* ctx);
int kzg_ctx_sync(kzg_ctx* ctx);
/* Stream ordering against a caller-owned stream without a host synchronisation: the context's stream waits for what is
 * queued on `stream` / `stream` waits for what is queued on the context's stream.  (The multi-rank MSM exchanges its
 * partial points with a collective on the caller's stream.) */
int kzg_ctx_wait_stream(kzg_ctx* ctx, void* stream);
int kzg_stream_wait_ctx(kzg_ctx* ctx, void* stream);
/* Tuning knobs of the MSM (A/B timing, forced paths in tests; none changes a result).  The KZGB200_<NAME> environment
 * variables are read once at kzg_ctx_create; value < 0 restores the default.  Names: aff_rounds, aff_m, aff_chunks,
 * aff_min_entries_log, aff_min_left_log, aff_min_fill, part_sort, red_k0, tail_width, host_piece_min_log,
 * split_min_log, split_max_log, msm_merge. */
int kzg_ctx_set_option(kzg_ctx* ctx, const char* name, int64_t value);
const char* kzg_last_error(kzg_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t kzg_ctx_launch_count(kzg_ctx* ctx);
/* device self-test: fast Montgomery path == portable path, group law identities.  0 = pass */
int kzg_selftest(kzg_ctx* ctx, uint32_t n_cases);
/* roofline denominator for the integer kernels: measured throughput of independent 32x32+64 -> 64 bit
 * multiply-accumulate chains (IMAD.WIDE.U32 in SASS), in MAC/s, over `ms` milliseconds of kernel time */
int kzg_bench_imad_peak(kzg_ctx* ctx, uint32_t ms, double* macs_per_second);
/* same unit, measured on chains of this library's own Montgomery product (136 limb-MACs each): the practical
 * ceiling once the carry handling (IADD3) and the quotient-digit IMADs of a modular product are included */
int kzg_bench_modmul_peak(kzg_ctx* ctx, uint32_t ms, double* macs_per_second);
/* device time in milliseconds the context's kernels tagged `which` took since the last reset (CUDA events
 * on the context's stream; 0 = msm bucket accumulation (XYZZ walk), 5 = its batched-affine rounds, 2/3/4 = msm sort /
 * bucket reduction / finish, 1 = NTT).  Used by bench.py for the live roofline figure. */
int kzg_ctx_kernel_time(kzg_ctx* ctx, uint32_t which, int reset, double* ms_out, uint64_t* launches_out);

/* ---- SRS ---------------------------------------------------------------------------------------------- */
/* replaces readBinFile + readPTauHeader + fd.readToBuffer (prover.js:15-16,83-85; ptau_utils.js:3-24):
 * validates magic "ptau", version <= 1, exactly one header section, n8 == 32, q == BN254 q, header size;
 * uploads the first n_points of section 2 (clamped to the section).  power_out may be NULL. */
int kzg_srs_load_ptau(kzg_ctx* ctx, const char* path, uint64_t n_points, kzg_srs** out, uint32_t* power_out);
/* the same for the points [first, first + n_points) only: one device's shard of a multi-GPU SRS */
int kzg_srs_load_ptau_range(kzg_ctx* ctx, const char* path, uint64_t first, uint64_t n_points, kzg_srs** out,
                            uint32_t* power_out);
/* header only (ptau_utils.js:3-24): power, ceremonyPower */
int kzg_ptau_read_header(kzg_ctx* ctx, const char* path, uint32_t* power, uint32_t* ceremony_power);
/* [tau]_2 : second G2 point of section 3, 128 bytes (verifier.js:18-19) */
int kzg_ptau_read_tau_g2(kzg_ctx* ctx, const char* path, uint8_t out[128]);
int kzg_srs_from_host(kzg_ctx* ctx, const uint8_t* affine, uint64_t n_points, kzg_srs** out);
/* synthetic SRS: [tau^i]_1 for i < n_points computed on the device (the Hermez file of
 * .github/workflows/tests.yml:15-19 cannot be downloaded offline).  tau_std = 32 B standard-form LE. */
int kzg_srs_generate(kzg_ctx* ctx, const uint8_t tau_std[32], uint64_t n_points, kzg_srs** out);
/* the slice [tau^i]_1, first <= i < first + n_points: the shard of one GPU in the multi-GPU MSM */
int kzg_srs_generate_range(kzg_ctx* ctx, const uint8_t tau_std[32], uint64_t first, uint64_t n_points, kzg_srs** out);
/* Lagrange-basis SRS: [L_i(tau)]_1 for i < 2^n_bits, the inverse DFT of the first 2^n_bits monomial points
 * carried out in the group (one-off, ~0.6 s at 2^20).  kzg_commit / kzg_commit_many over it take a polynomial's EVALUATIONS
 * on H (Montgomery) and return the same point as committing iNTT(evaluations) over the monomial SRS -- the commitments of
 * F, T, S / Z (prover.js:151-162; grandsum.js:61) without their iNTTs. */
int kzg_srs_lagrange(kzg_ctx* ctx, kzg_srs* srs, uint32_t n_bits, kzg_srs** out);
/* write a .ptau (sections 1,2,3) from a device SRS; tau_g2 = 128 B [tau]_2 (host computed by the caller) */
int kzg_srs_write_ptau(kzg_ctx* ctx, kzg_srs* srs, uint32_t power, const uint8_t g2_one[128],
                       const uint8_t g2_tau[128], const char* path);
int kzg_srs_download(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, uint64_t count, uint8_t* out);
uint64_t kzg_srs_len(kzg_srs* srs);
void* kzg_srs_device_ptr(kzg_srs* srs); /* device address of the 64-byte affine points (for KZG_BASES_ON_DEVICE) */
int kzg_srs_free(kzg_ctx* ctx, kzg_srs* srs);

/* ---- device Fr vectors (BigBuffer / Uint8Array of n*32 bytes) ---------------------------------------- */
int kzg_buf_alloc(kzg_ctx* ctx, uint64_t n, kzg_buf** out); /* zero-filled, like new Uint8Array */
int kzg_buf_free(kzg_ctx* ctx, kzg_buf* b);
uint64_t kzg_buf_len(kzg_buf* b);
void* kzg_buf_device_ptr(kzg_buf* b);
int kzg_buf_upload(kzg_ctx* ctx, kzg_buf* dst, uint64_t dst_off, const uint8_t* host, uint64_t n);
int kzg_buf_download(kzg_ctx* ctx, kzg_buf* src, uint64_t src_off, uint8_t* host, uint64_t n);
int kzg_buf_copy(kzg_ctx* ctx, kzg_buf* dst, uint64_t dst_off, kzg_buf* src, uint64_t src_off, uint64_t n);
int kzg_buf_fill(kzg_ctx* ctx, kzg_buf* dst, uint64_t off, uint64_t n, const uint8_t value[32]);
/* bytewise compare with a broadcast element (Evaluations.isAllOnes / isAllZeros, evaluations.js:118-129) */
int kzg_buf_all_equal(kzg_ctx* ctx, kzg_buf* b, const uint8_t value[32], int* out);

/* ---- bulk Fr (ffjavascript call sites) ---------------------------------------------------------------- */
int kzg_fr_to_mont(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out);   /* Fr.batchToMontgomery   prover.js:147-148 */
int kzg_fr_from_mont(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out); /* Fr.batchFromMontgomery polynomial.js:1109 */
/* Fr.fft / Fr.ifft (evaluations.js:18; polynomial.js:34,373,392): natural order in and out,
 * w = Fr.w[log2 n], ifft scaled by n^-1.  out may alias in. */
int kzg_fr_ntt(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out, int inverse);
/* Evaluations.fromPolynomial(p, extension) (evaluations.js:12-21): zero-pad to nextpow2(len)*ext, NTT */
int kzg_fr_extend_ntt(kzg_ctx* ctx, kzg_buf* coef, uint32_t extension, kzg_buf** out);
int kzg_fr_batch_inverse(kzg_ctx* ctx, kzg_buf* in, kzg_buf* out); /* Fr.batchInverse grandsum.js:41 (0 -> 0) */

/* ---- Polynomial (used set, polynomial.js) ------------------------------------------------------------- */
int kzg_poly_add(kzg_ctx* ctx, kzg_buf* a, kzg_buf* b, kzg_buf** out);  /* :276-312 result takes the longer length */
int kzg_poly_sub(kzg_ctx* ctx, kzg_buf* a, kzg_buf* b, kzg_buf** out);  /* :314-350 */
int kzg_poly_mul_scalar(kzg_ctx* ctx, kzg_buf* a, const uint8_t s[32]); /* :395-406 in place */
int kzg_poly_add_scalar(kzg_ctx* ctx, kzg_buf* a, const uint8_t s[32]); /* :408-414 */
int kzg_poly_sub_scalar(kzg_ctx* ctx, kzg_buf* a, const uint8_t s[32]); /* :416-422 */
int kzg_poly_degree(kzg_ctx* ctx, kzg_buf* a, uint64_t* degree);        /* :212-226 */
int kzg_poly_evaluate(kzg_ctx* ctx, kzg_buf* a, const uint8_t x[32], uint8_t out[32]); /* :228-238 */
int kzg_poly_multiply(kzg_ctx* ctx, kzg_buf* a, kzg_buf* b, kzg_buf** out);            /* :352-376 */
int kzg_poly_shift_omega(kzg_ctx* ctx, kzg_buf* a, kzg_buf** out);                     /* :378-393 */
int kzg_poly_div_zh(kzg_ctx* ctx, kzg_buf* a, uint64_t domain_size, kzg_buf** out);    /* :853-888 "Polynomial is not divisible" */
int kzg_poly_div_x_sub_value(kzg_ctx* ctx, kzg_buf* a, const uint8_t v[32], kzg_buf** out); /* :814-851 "Polynomial does not divide" */
int kzg_poly_lagrange1(kzg_ctx* ctx, uint32_t power, kzg_buf** out);                   /* :68-78 */

/* ---- argument cores ----------------------------------------------------------------------------------- */
/* ComputeSGrandSumPolynomial (grandsum.js:6-62) / ComputeZGrandProductPolynomial (grandproduct.js:6-57):
 * evaluations in, coefficients of S / Z out.  sel_f / sel_t may be NULL (all ones). */
int kzg_grandsum_build(kzg_ctx* ctx, kzg_buf* ev_f, kzg_buf* ev_t, kzg_buf* sel_f, kzg_buf* sel_t,
                       const uint8_t gamma[32], kzg_buf** s_coef);
int kzg_grandproduct_build(kzg_ctx* ctx, kzg_buf* ev_f, kzg_buf* ev_t, kzg_buf* sel_f, kzg_buf* sel_t,
                           const uint8_t gamma[32], kzg_buf** z_coef);

/* ---- commitments -------------------------------------------------------------------------------------- */
/* commit(pol) = Polynomial.multiExponentiation + G1.toAffine (polynomial.js:1106-1115; prover.js:432-434) */
int kzg_commit(kzg_ctx* ctx, kzg_srs* srs, kzg_buf* coef, uint8_t out_affine[64]);
/* G1.multiExpAffine (polynomial.js:1112): bases 64 B affine Montgomery-LE, scalars 32 B STANDARD-form LE.
 * Pointers are host pointers unless the matching flag says device.  Output is the canonical affine point
 * (a Jacobian triple is not canonical; out_jacobian, if non-NULL, receives x || y || one). */
enum { KZG_BASES_ON_DEVICE = 1, KZG_SCALARS_ON_DEVICE = 2 };
int kzg_g1_msm_affine(kzg_ctx* ctx, const void* bases, const void* scalars_std, uint64_t n, uint32_t flags,
                      uint8_t out_affine[64], uint8_t out_jacobian[96]);
/* MSM over an SRS slice with device-resident standard-form scalars (the standalone sweep of BASELINE C5) */
int kzg_srs_msm(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, kzg_buf* scalars_std, uint64_t n, uint8_t out_affine[64]);
/* multi-GPU MSM: each rank reduces its slice to one partial point in extended-Jacobian
 * form (128 B: X, Y, ZZ, ZZZ) written to DEVICE memory `partial_dev`; after the ranks all-gather the
 * partials, kzg_g1_partials_combine adds `count` of them and returns the affine point. */
int kzg_srs_msm_partial(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, kzg_buf* scalars_std, uint64_t n, void* partial_dev);
int kzg_g1_partials_combine(kzg_ctx* ctx, const void* partials_dev, uint32_t count, uint8_t out_affine[64]);
/* Same, scalars in HOST memory (H2D inside the call): the end-to-end form of kzg_srs_msm / kzg_srs_msm_partial. */
int kzg_srs_msm_host(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, const void* scalars_std_host, uint64_t n,
                     uint8_t out_affine[64]);
int kzg_srs_msm_host_partial(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, const void* scalars_std_host, uint64_t n,
                             void* partial_dev);
/* commit(pol) for `count` polynomials over one SRS as ONE pipeline (multi-MSM over shared bases; the
 * reference's separate commits of one round: prover.js:161-162 [F],[T]; :409-410 [Wxi],[Wxiw]): one sort, one bucket
 * accumulation, one reduction with a bucket set per polynomial.  out_affine: 64 B per polynomial, in order. */
int kzg_commit_many(kzg_ctx* ctx, kzg_srs* srs, kzg_buf* const* coefs, uint32_t count, uint8_t* out_affine);
/* One-off per resident SRS: the window table T[w][i] = 2^(c w) [tau^i]_1 that lets every MSM over this SRS use a
 * single shared bucket set (fewer, larger windows; no doubling chain).  window_bits = 0 picks c from the SRS size.
 * Costs ceil(257/c) x the SRS memory.  Without it the SRS entry points fall back to the per-window MSM. */
int kzg_srs_precompute(kzg_ctx* ctx, kzg_srs* srs, uint32_t window_bits);
/* MSM tuning for the per-window (table-less) path: window bits (0 = auto from n); a non-zero value also makes
 * the SRS entry points ignore their table.  kzg_msm_geometry reports what an n-point MSM will use (srs may be
 * NULL for the raw kzg_g1_msm_affine path; montgomery = 1 for kzg_commit's scalars, 0 for standard form). */
int kzg_msm_geometry(kzg_ctx* ctx, kzg_srs* srs, uint64_t n, int montgomery, uint32_t* window_bits, uint32_t* windows);
/* the same plus the number of batched-affine rounds the bucket accumulation will run before its XYZZ walk (0 below
 * ~2^22 points; each round adds the entries of every bucket pairwise with a shared inversion) */
int kzg_msm_plan(kzg_ctx* ctx, kzg_srs* srs, uint64_t n, int montgomery, uint32_t* window_bits, uint32_t* windows,
                 uint32_t* affine_rounds);
int kzg_msm_set_window(kzg_ctx* ctx, uint32_t c);

/* ---- multi-GPU MSM inside one process (a Node addon cannot run one process per GPU) ----------------------- */
/* G1.multiExpAffine / commit over an SRS sharded across devices (polynomial.js:1106-1115).  Device g keeps the contiguous
 * slice [first_g, first_g + count_g) of the SRS resident with its window table, reduces the matching scalars to ONE
 * 128-byte partial point, device 0 pulls the partials with peer copies and normalises.  devices = NULL / n = 0: all
 * visible devices.  One context per device (kzg_mgpu_ctx), one host thread per device during a call. */
typedef struct kzg_mgpu kzg_mgpu;
int kzg_mgpu_create(const int* devices, uint32_t n_devices, kzg_mgpu** out);
int kzg_mgpu_destroy(kzg_mgpu* m);
uint32_t kzg_mgpu_device_count(kzg_mgpu* m);
kzg_ctx* kzg_mgpu_ctx(kzg_mgpu* m, uint32_t i);
const char* kzg_mgpu_last_error(kzg_mgpu* m);
uint64_t kzg_mgpu_srs_len(kzg_mgpu* m);
int kzg_mgpu_shard(kzg_mgpu* m, uint32_t i, uint64_t* first, uint64_t* count);
int kzg_mgpu_srs_generate(kzg_mgpu* m, const uint8_t tau_std[32], uint64_t n_points);
int kzg_mgpu_srs_from_host(kzg_mgpu* m, const uint8_t* affine, uint64_t n_points);
int kzg_mgpu_srs_load_ptau(kzg_mgpu* m, const char* path, uint64_t n_points);   /* prover.js:15-16,83-85 */
/* scalars in HOST memory (n x 32 B standard form, n <= |SRS|); uploads piecewise under the pieces' MSMs */
int kzg_mgpu_srs_msm_host(kzg_mgpu* m, const void* scalars_std_host, uint64_t n, uint8_t out_affine[64]);
/* scalars made resident once, then any number of MSMs over them */
int kzg_mgpu_scalars_upload(kzg_mgpu* m, const void* scalars_std_host, uint64_t n);
int kzg_mgpu_srs_msm(kzg_mgpu* m, uint8_t out_affine[64]);
/* page-lock / release a caller-owned host buffer for all devices (so that uploads overlap compute) */
int kzg_host_register(void* ptr, uint64_t bytes);
int kzg_host_unregister(void* ptr);

/* ---- fused provers (prover.js:144-413 and the grand-product twin) --------------------------------------- */
/* The Keccak transcript stays with the caller: each round returns the bytes the transcript needs and the
 * next round takes the challenge derived from them. */
int kzg_prover_create(kzg_ctx* ctx, kzg_srs* srs, int kind, uint32_t n_bits, uint32_t n_pols, int selected,
                      kzg_prover** out);
int kzg_prover_destroy(kzg_prover* p);
/* round 1 (:144-179): columns are HOST pointers to n*32 B; F/T standard form, selectors Montgomery (may be
 * NULL when !selected).  commitments_out: 64 B each in the order F0,T0,F1,T1,...[,selF,selT]. */
int kzg_prover_round1(kzg_prover* p, const uint8_t* const* evals_f_std, const uint8_t* const* evals_t_std,
                      const uint8_t* sel_f, const uint8_t* sel_t, uint8_t* commitments_out);
/* round 2 (:181-231): beta ignored unless n_pols > 1.  out: [S] or [Z] */
int kzg_prover_round2(kzg_prover* p, const uint8_t beta[32], const uint8_t gamma[32], uint8_t out_acc[64]);
/* round 3 (:233-286): out: [Q] */
int kzg_prover_round3(kzg_prover* p, const uint8_t alpha[32], uint8_t out_q[64]);
/* round 4 (:288-318): evaluations, 32 B each, in proof order (GS: f0,t0,f1,t1,..[,selF,selT],sxiw;
 * GP: f0,f1,..[,selF,selT],zxiw) */
int kzg_prover_round4(kzg_prover* p, const uint8_t xi[32], uint8_t* evals_out);
/* round 5 (:320-413): out: [Wxi] || [Wxiw] */
int kzg_prover_round5(kzg_prover* p, const uint8_t v[32], uint8_t out_w[128]);
/* After round 5: hand the Montgomery-form evaluations of column `column` (which = 0: F, 1: T) to the caller as a
 * device vector (ownership moves; free with kzg_buf_free).  The reference REPLACES evalsFs[i].eval / evalsTs[i].eval
 * by their Montgomery form as a side effect of proving (prover.js:147-148); the host layer reproduces that with this. */
int kzg_prover_take_evals(kzg_prover* p, uint32_t column, int which, kzg_buf** out);
/* message of the last failed round call (kzg_last_error of the prover's context) */
const char* kzg_prover_last_error(kzg_prover* p);
/* number of evaluations / commitments round 4 / round 1 write */
uint32_t kzg_prover_n_evals(kzg_prover* p);
uint32_t kzg_prover_n_round1_commitments(kzg_prover* p);

/* ---- host helpers ------------------------------------------------------------------------------------- */
/* Keccak-256, original padding (js-sha3 keccak256, Keccak256Transcript.js:50).  Runs on the HOST. */
void kzg_keccak256(const uint8_t* data, size_t len, uint8_t out[32]);
/* Transcript serialisation helpers (Keccak256Transcript.js:42,45), host side:
 * 64 B Montgomery-LE affine -> 64 B big-endian standard form; 32 B Montgomery-LE Fr -> 32 B BE standard. */
void kzg_g1_to_rpr_uncompressed(const uint8_t in[64], uint8_t out[64]);
void kzg_fr_to_rpr_be(const uint8_t in[32], uint8_t out[32]);
/* 32 B big-endian integer (a hash) -> reduced mod r -> 32 B Montgomery-LE (Fr.e, Keccak256Transcript.js:51) */
void kzg_fr_from_hash_be(const uint8_t in[32], uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif /* KZGB200_H */
