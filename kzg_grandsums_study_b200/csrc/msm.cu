// Signed-digit Pippenger G1 MSM for sm_100a.   Replaces G1.multiExpAffine + G1.toAffine
// (reference src/polynomial/polynomial.js:1106-1115, the un-vendored ffjavascript/wasmcurves call).
//
// Two flavours share one pipeline:
//   * SRS MSM (commit, kzg_srs_msm): the SRS carries a precomputed table T[w][i] = 2^(c w) * P_i (affine),
//     built once when the SRS is made resident.  Every (point, window) digit then lands in ONE shared set of
//     2^(c-1) buckets, so there is a single bucket reduction and no doubling chain at the end, and c can be
//     large (22 bits at 2^24 points -> 12 digits per scalar instead of 16).  180 GB of HBM pay for it:
//     12 x 1 GiB at 2^24 points.
//   * raw MSM (kzg_g1_msm_affine on caller-supplied bases): classic per-window bucket sets, windows combined
//     by a Horner chain of doublings.
//
// Pipeline (one stream, no host round trip until the 64-byte result is read):
//   1. msm_digits<count>  scalars -> signed c-bit digits (leaving Montgomery form first when the source is a
//                         polynomial: Fr.batchFromMontgomery fused), histogram of bucket keys (L2 atomics)
//   2. msm_scan_*         exclusive prefix sums of bucket sizes, then of partial-sum slots per bucket (2 x 3 launches)
//   3. msm_digits<scatter> counting-sort scatter of (point index | sign) into bucket order
//   3b. msm_aff_forward / fq_batch_inverse / msm_aff_backward   (from 2^21 points on) batched-affine rounds: the entries
//                         of every bucket are added pairwise with one shared inversion per round (~780 wide MACs per
//                         addition), the list halves each time and ends as a dense array of affine points
//   4. msm_accumulate     one thread per fixed-length SLICE of the sorted list (equal work per lane whatever the
//                         bucket sizes): XYZZ += affine, parking a partial sum at every bucket boundary; random 64 B
//                         gathers (or the dense list the rounds left), next point prefetched during the add.
//                         THE IMAD-bound kernel: 1160 wide MACs per entry (6 products, 2 squarings, 1 double product).
//                         A later piece of a host-scalar MSM opens its buckets with the earlier pieces' sums (MsmCarry).
//   5. msm_collapse       buckets spread over many slices: block-parallel sum of their partials
//   6. msm_reduce_level0  sum_b (b+1) S_b: chunks of 4-8 consecutive buckets -> plain sum U_q and weighted partial t_q
//                         (running-sum trick, the only pass over all buckets)
//   7. msm_tail_tasks     sum_q q U_q as plain column / row sums of the LO x HI arrangement of the chunks (one warp
//      msm_tail_final     each, weight by double-and-add), then T + r0 (C + LO R) -> one XYZZ per bucket set
//   8. msm_horner         raw flavour only: combine the windows (c doublings each)
//   9. g1_finish          sum `count` partial points (count > 1 only for the multi-GPU gather), canonical affine
// The result is a canonical group element, so it is byte-identical to the reference's regardless of window
// size, digit signedness, precomputation or summation order.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace kzg {

// ---------------------------------------------------------------------------------------------
// Debug build (KZGB200_DEBUG_BOUNDS=1 python -m kzg_grandsums_study_b200.build  ->  -DKZG_BOUNDS_CHECK): every
// hand-computed index of the sort, the affine rounds and the walk -- staged shared-memory slots, scratch-arena offsets,
// table gathers -- is checked against the limit the host laid the scratch out with; a violation sets a bit in a device
// word, the access is skipped, and the MSM returns an error naming the site.  (compute-sanitizer is closed on this GPU
// pool; the whole -m gpu suite has been run once under this build: profiles/r02_bounds_check.md.)  In the normal build
// KZG_IDX_OK(...) is the constant `true`.
// ---------------------------------------------------------------------------------------------
#ifdef KZG_BOUNDS_CHECK
struct MsmDebugLimits {
    unsigned long long sorted_entries, mid_entries, partials, prefix_elems, totals, aff_cap[2], table_points, nkeys, dense_in;
};
__device__ MsmDebugLimits g_dbg;
__device__ unsigned int g_dbg_violation;
__device__ __forceinline__ bool kzg_idx_ok(unsigned long long idx, unsigned long long limit, int site) {
    if (idx < limit) return true;
    atomicOr(&g_dbg_violation, 1u << site);
    return false;
}
#define KZG_IDX_OK(idx, limit, site) kzg_idx_ok((unsigned long long)(idx), (unsigned long long)(limit), site)
#define KZG_DBG(field) (g_dbg.field)
#else
#define KZG_IDX_OK(idx, limit, site) true
#define KZG_DBG(field) 0
#endif
enum { DBG_DIGITS_SORTED = 0, DBG_PART_RANK = 1, DBG_PART_STAGE = 2, DBG_PART_MID = 3, DBG_CHUNK_STAGE = 4, DBG_CHUNK_SORTED = 5,
       DBG_CHUNK_COUNTS = 6, DBG_WALK_PARTIALS = 7, DBG_WALK_GATHER = 8, DBG_FWD_PREFIX = 9, DBG_FWD_GATHER = 10,
       DBG_BWD_OUT = 11, DBG_BWD_GATHER = 12, DBG_WALK_DENSE = 13, DBG_FWD_TOTALS = 14 };

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
struct MsmGeom {
    uint32_t c;         // window bits
    uint32_t nwin;      // digits per scalar
    uint32_t nbuckets;  // buckets per set = 2^(c-1)
    uint32_t nsets;     // bucket sets: one per job with a precomputed table, nwin without
    uint32_t table;     // 1: window-table flavour (all windows of a job share one bucket set)
    uint32_t seg;       // slice length: entries per accumulate thread
    uint64_t stride;    // table flavour: entry = w * stride + i
};

// Several MSMs over the SAME bases (the commitments of one prover round share the SRS window table) run as ONE pipeline:
// job j owns the bucket keys [j 2^(c-1), (j+1) 2^(c-1)), so that one sort, one accumulation and one reduction serve all of
// them and each bucket set reduces to its own result.  Table flavour only (one bucket set per job); blockIdx.y = job.
constexpr uint32_t MSM_MAX_JOBS = 8;
struct MsmJobs {
    const Fr* scalars[MSM_MAX_JOBS];
    uint64_t n[MSM_MAX_JOBS];
    uint32_t montgomery;  // bit j: the scalars of job j are Montgomery residues (a polynomial's coefficients)
    uint32_t count;
};

// One piece of an MSM handing its folded bucket sums (one point per non-empty bucket, addressed through pbase) to the
// later pieces over the same bucket geometry (MsmCarry); `ready` is recorded on the producing stream after the fold.
struct MsmReduceLink {
    const G1XYZZ* partials = nullptr;
    const uint32_t* pbase = nullptr;
    uint32_t nkeys = 0;
    cudaEvent_t ready = nullptr;
};

__device__ __forceinline__ uint32_t scalar_bits(const uint32_t* s, uint32_t pos, uint32_t c) {
    // bits [pos, pos+c) of a 256-bit little-endian integer, zero beyond bit 255
    uint32_t word = pos >> 5, off = pos & 31;
    if (word >= 8) return 0;
    uint64_t v = s[word];
    if (word + 1 < 8) v |= (uint64_t)s[word + 1] << 32;
    return (uint32_t)(v >> off) & ((1u << c) - 1u);
}

// Warp-aggregated bucket atomics: lanes of a warp that hit the same key in the same window issue ONE atomic
// (uniform scalars never collide, but skewed inputs -- equal scalars, a short top window -- would otherwise
// serialise millions of atomics on a handful of L2 addresses).
template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_digits_kernel(MsmJobs jobs, MsmGeom g, uint32_t* __restrict__ counts_or_cursor,
                                                         uint32_t* __restrict__ sorted) {
    const uint32_t job = blockIdx.y;
    const Fr* __restrict__ scalars = jobs.scalars[job];
    const uint64_t n = jobs.n[job];
    const bool montgomery = (jobs.montgomery >> job) & 1u;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((uint64_t)blockIdx.x * blockDim.x >= n) return;  // (whole block: the grid is sized for the longest job)
    const uint32_t lane = threadIdx.x & 31;
    Fr s = fp_zero<FrP>();  // lanes past the end run the loop with a zero scalar: the warp stays converged
    if (i < n) {
        // streaming read (evict-first): the scalars are read once and must not push the bucket cursors and the
        // partially written sectors of `sorted` out of L2
        const uint4* q = reinterpret_cast<const uint4*>(scalars + i);
        uint4 a = __ldcs(q), b = __ldcs(q + 1);
        s.l[0] = a.x; s.l[1] = a.y; s.l[2] = a.z; s.l[3] = a.w;
        s.l[4] = b.x; s.l[5] = b.y; s.l[6] = b.z; s.l[7] = b.w;
        if (montgomery) s = fp_from_mont(s);
    }
    uint64_t keep_policy = 0;
    if (SCATTER) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));
    uint32_t carry = 0;
    const uint32_t half = g.nbuckets;  // 2^(c-1)
    const bool table = g.table != 0;
    for (uint32_t w = 0; w < g.nwin; w++) {
        uint32_t raw = scalar_bits(s.l, w * g.c, g.c) + carry;
        uint32_t mag, neg;
        if (raw > half) {
            mag = (1u << g.c) - raw;
            neg = 1;
            carry = 1;
        } else {
            mag = raw;
            neg = 0;
            carry = 0;
        }
        const bool valid = mag != 0;
        const uint32_t key = valid ? (table ? job * g.nbuckets : w * g.nbuckets) + (mag - 1) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const uint32_t leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (valid && lane == leader) base = atomicAdd(&counts_or_cursor[key], (uint32_t)__popc(peers));
        if (SCATTER) {
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid && KZG_IDX_OK(base + rank, KZG_DBG(sorted_entries), DBG_DIGITS_SORTED)) {
                uint32_t entry = table ? (uint32_t)(w * g.stride + i) : (uint32_t)i;
                // every bucket has ONE partially filled 32-byte sector of `sorted` at any time (2^(c-1) x 32 B in
                // total): ask L2 to keep those lines until their other 7 entries arrive
                asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(sorted + base + rank), "r"(entry | (neg << 31)),
                             "l"(keep_policy)
                             : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Exclusive scans over the bucket keys, three phases over tiles of SCAN_TILE keys (tile sums, one-block scan
// of the tile sums, apply).  Run twice:
//   mode 0  f(k) = counts[k]                      -> offsets[] (and the scatter cursors)
//   mode 1  f(k) = slices touched by bucket k     -> pbase[]   (first partial-sum slot of the bucket)
// The sorted entry list is cut into slices of `slice` entries, one accumulate thread each, so a bucket that
// spans slice boundaries produces one partial sum per slice it touches.  Buckets with more than
// HEAVY_PARTS partials are appended to `heavy` and summed by a whole block after the accumulation.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;
constexpr uint32_t HEAVY_PARTS = 8;
constexpr uint32_t HUGE_PARTS = 256;

__device__ __forceinline__ uint32_t parts_of(uint32_t off, uint32_t cnt, uint32_t slice) {
    return cnt == 0 ? 0u : (off + cnt - 1) / slice - off / slice + 1;
}
// `shift` batched-affine rounds (below) leave ceil(count / 2^shift) points in a bucket
__device__ __forceinline__ uint32_t scan_input(int mode, const uint32_t* __restrict__ counts,
                                               const uint32_t* __restrict__ offsets, uint32_t k, uint32_t slice,
                                               uint32_t shift) {
    uint32_t cnt = (counts[k] + ((1u << shift) - 1u)) >> shift;
    return mode == 0 ? cnt : parts_of(offsets[k], cnt, slice);
}

__device__ __forceinline__ uint32_t block_scan_u32(uint32_t v, uint32_t* sh, uint32_t& total) {
    // inclusive scan across the block
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31) sh[wid] = v;
    __syncthreads();
    uint32_t carry = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        uint32_t x = sh[w];
        if (w < (int)wid) carry += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return v + carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan_tiles_kernel(int mode, const uint32_t* __restrict__ counts,
                                                                      const uint32_t* __restrict__ offsets, uint32_t nkeys,
                                                                      uint32_t slice, uint32_t shift,
                                                                      uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        uint32_t i = base + k * SCAN_THREADS;
        if (i < nkeys) acc += scan_input(mode, counts, offsets, i, slice, shift);
    }
    uint32_t tot;
    block_scan_u32(acc, sh, tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of the tile sums in place; grand total -> out[nkeys]
__global__ void __launch_bounds__(SCAN_THREADS) msm_scan_sums_kernel(uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                                                                     uint32_t nkeys, uint32_t* __restrict__ out) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    __shared__ uint32_t carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint32_t start = 0; start < ntiles; start += SCAN_THREADS) {
        uint32_t i = start + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0;
        uint32_t tot;
        uint32_t inc = block_scan_u32(v, sh, tot);
        uint32_t c = carry_sh;
        if (i < ntiles) tile_sums[i] = c + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[nkeys] = carry_sh;
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan_apply_kernel(int mode, const uint32_t* __restrict__ counts,
                                                                      const uint32_t* __restrict__ offsets, uint32_t nkeys,
                                                                      uint32_t slice, uint32_t shift,
                                                                      const uint32_t* __restrict__ tile_sums,
                                                                      uint32_t* __restrict__ out, uint32_t* __restrict__ cursor,
                                                                      uint32_t* __restrict__ heavy,
                                                                      uint32_t* __restrict__ heavy_count,
                                                                      uint32_t* __restrict__ multi,
                                                                      uint32_t* __restrict__ multi_count,
                                                                      uint32_t* __restrict__ huge,
                                                                      uint32_t* __restrict__ huge_count) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    __shared__ uint32_t multi_base;
    // thread owns SCAN_PER_THREAD consecutive keys so that its partial results are a running sum
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t val[SCAN_PER_THREAD];
    uint32_t acc = 0, nmulti = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        uint32_t i = base + k;
        val[k] = i < nkeys ? scan_input(mode, counts, offsets, i, slice, shift) : 0;
        acc += val[k];
        nmulti += (val[k] >= 2 && val[k] <= HEAVY_PARTS) ? 1u : 0u;
    }
    uint32_t tot;
    uint32_t inc = block_scan_u32(acc, sh, tot);
    uint32_t run = tile_sums[blockIdx.x] + inc - acc;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        uint32_t i = base + k;
        if (i < nkeys) {
            out[i] = run;
            if (mode == 0) cursor[i] = run;
            if (mode == 1 && val[k] > HEAVY_PARTS) {
                if (val[k] > HUGE_PARTS) huge[atomicAdd(huge_count, 1u)] = i;
                else heavy[atomicAdd(heavy_count, 1u)] = i;
            }
            run += val[k];
        }
    }
    if (mode == 1) {
        // buckets with 2..HEAVY_PARTS partial sums: compact list (one atomic per block), folded by msm_fold_kernel
        uint32_t mtot;
        uint32_t minc = block_scan_u32(nmulti, sh, mtot);
        if (threadIdx.x == 0) multi_base = mtot ? atomicAdd(multi_count, mtot) : 0u;
        __syncthreads();
        uint32_t pos = multi_base + minc - nmulti;
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; k++)
            if (val[k] >= 2 && val[k] <= HEAVY_PARTS) multi[pos++] = base + k;
    }
}

// The bucket offsets of the sorted list AND of the list every batched-affine round leaves (ceil(count / 2^r) points per
// bucket, r = 1 .. rounds) depend on the counts alone, so one set of three launches emits them all up front:
// out[r * (nkeys + 1) + k], tile_sums[r * ntiles + tile].  Shift 0 also seeds the scatter cursors.
constexpr uint32_t AFF_MAX_ROUNDS = 6;
constexpr uint32_t MAX_SHIFTS = AFF_MAX_ROUNDS + 1;
__global__ void __launch_bounds__(SCAN_THREADS) msm_offsets_tiles_kernel(const uint32_t* __restrict__ counts, uint32_t nkeys,
                                                                         uint32_t nshift, uint32_t ntiles,
                                                                         uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x;
    uint32_t acc[MAX_SHIFTS];
#pragma unroll
    for (uint32_t r = 0; r < MAX_SHIFTS; r++) acc[r] = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        const uint32_t i = base + k * SCAN_THREADS;
        const uint32_t c = i < nkeys ? counts[i] : 0u;
#pragma unroll
        for (uint32_t r = 0; r < MAX_SHIFTS; r++) acc[r] += (c + ((1u << r) - 1u)) >> r;
    }
#pragma unroll
    for (uint32_t r = 0; r < MAX_SHIFTS; r++) {
        if (r >= nshift) break;
        uint32_t tot;
        block_scan_u32(acc[r], sh, tot);
        if (threadIdx.x == 0) tile_sums[r * ntiles + blockIdx.x] = tot;
    }
}
// one block per shift: exclusive scan of its tile sums in place; grand total -> out[r][nkeys]
__global__ void __launch_bounds__(SCAN_THREADS) msm_offsets_sums_kernel(uint32_t* __restrict__ tile_sums, uint32_t ntiles,
                                                                        uint32_t nkeys, uint32_t* __restrict__ out) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    __shared__ uint32_t carry_sh;
    uint32_t* ts = tile_sums + (size_t)blockIdx.x * ntiles;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint32_t start = 0; start < ntiles; start += SCAN_THREADS) {
        const uint32_t i = start + threadIdx.x;
        const uint32_t v = i < ntiles ? ts[i] : 0;
        uint32_t tot;
        const uint32_t inc = block_scan_u32(v, sh, tot);
        const uint32_t c = carry_sh;
        if (i < ntiles) ts[i] = c + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)blockIdx.x * (nkeys + 1) + nkeys] = carry_sh;
}
__global__ void __launch_bounds__(SCAN_THREADS) msm_offsets_apply_kernel(const uint32_t* __restrict__ counts, uint32_t nkeys,
                                                                         uint32_t nshift, uint32_t ntiles,
                                                                         const uint32_t* __restrict__ tile_sums,
                                                                         uint32_t* __restrict__ out, uint32_t* __restrict__ cursor) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    // thread owns SCAN_PER_THREAD consecutive keys so that its partial results are a running sum
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t cnt[SCAN_PER_THREAD];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) cnt[k] = base + k < nkeys ? counts[base + k] : 0u;
    for (uint32_t r = 0; r < nshift; r++) {
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; k++) acc += (cnt[k] + ((1u << r) - 1u)) >> r;
        uint32_t tot;
        const uint32_t inc = block_scan_u32(acc, sh, tot);
        uint32_t run = tile_sums[r * ntiles + blockIdx.x] + inc - acc;
        uint32_t* o = out + (size_t)r * (nkeys + 1);
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; k++) {
            if (base + k < nkeys) {
                o[base + k] = run;
                if (r == 0) cursor[base + k] = run;
                run += (cnt[k] + ((1u << r) - 1u)) >> r;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Partition sort (large inputs): the same counts[] / offsets[] / sorted[] as the kernels above, without one global
// atomic per entry.  218 M entries at 2^24 points cost 2 x 218 M L2 atomics plus as many scattered 4-byte stores in the
// direct scheme (1.7 + 4.0 ms, bound by the L2 atomic rate); here the atomics are shared-memory atomics:
//   key = part * 2^low_bits + lo
//   1. msm_part_hist      per-tile shared histogram over the partitions            -> part_hist[]   (one RED per tile and partition)
//   2. msm_part_scan      one block: partition starts, chunk table (chunks of SORT_CHUNK entries that never straddle
//                         a partition -- a skewed partition simply owns more chunks)
//   3. msm_part_scatter   per tile: entries staged in shared memory grouped by partition, written out in coalesced runs
//                         of (payload, key) pairs                                    -> mid[]
//   4. msm_chunk_hist     per chunk: shared histogram over the 2^low_bits keys of its partition -> counts[] (REDs)
//      (msm_scan_* mode 0: offsets[], cursor[])
//   5. msm_chunk_scatter  per chunk: one returning atomic per (chunk, key) reserves the slots, shared cursors place the
//                         payloads; the chunk is re-read from L2, the stores stay inside the partition's window of sorted[]
// ---------------------------------------------------------------------------------------------
struct SortGeom {
    uint32_t low_bits;      // keys per partition = 2^low_bits
    uint32_t nparts;
    uint32_t tile_scalars;  // scalars per tile of kernels 1 and 3 (<= PART_THREADS * PART_ITEMS)
    uint32_t ntiles;
};
constexpr int PART_THREADS = 512;
constexpr int PART_ITEMS = 2;
constexpr uint32_t PART_TILE_ENTRIES = 8192;   // 64 KB of staged pairs + 16 KB of ranks: two tiles per SM
constexpr uint32_t SORT_MAX_PARTS = 2048;
constexpr uint32_t SORT_MAX_LOW = 11;
constexpr uint32_t SORT_CHUNK = 8192;
constexpr int CHUNK_THREADS = 512;
constexpr size_t CHUNK_SMEM = sizeof(uint32_t) * (SORT_CHUNK + 3 * (1u << SORT_MAX_LOW) + 1) + sizeof(uint16_t) * SORT_CHUNK;

// signed c-bit digits of a scalar, least significant first: a 64-bit bit buffer is refilled limb by limb (static limb
// indices: the scalar stays in registers), the windows beyond bit 255 see zeros plus the carry
template <class F>
__device__ __forceinline__ void msm_for_digits(const Fr& s, const MsmGeom& g, uint32_t key_base, uint64_t i, F&& f) {
    const uint32_t half = g.nbuckets, c = g.c, mask = (1u << g.c) - 1u;
    const bool table = g.table != 0;
    uint64_t buf = 0;
    uint32_t have = 0, w = 0, carry = 0;
    auto emit = [&]() {
        const uint32_t raw = ((uint32_t)buf & mask) + carry;
        buf >>= c;
        uint32_t mag, neg;
        if (raw > half) {
            mag = (1u << c) - raw;
            neg = 1;
            carry = 1;
        } else {
            mag = raw;
            neg = 0;
            carry = 0;
        }
        if (mag != 0) {
            const uint32_t key = (table ? key_base : w * half) + (mag - 1);
            const uint32_t entry = table ? (uint32_t)(w * g.stride + i) : (uint32_t)i;
            f(key, entry | (neg << 31));
        }
        w++;
    };
#pragma unroll
    for (int limb = 0; limb < 8; limb++) {
        buf |= (uint64_t)s.l[limb] << have;
        have += 32;
        while (have >= c && w < g.nwin) {
            emit();
            have -= c;
        }
    }
    while (w < g.nwin) emit();  // the last, partial window(s)
}
__device__ __forceinline__ Fr load_scalar_stream(const Fr* scalars, uint64_t i, bool montgomery) {
    Fr s;
    const uint4* q = reinterpret_cast<const uint4*>(scalars + i);
    uint4 a = __ldcs(q), b = __ldcs(q + 1);
    s.l[0] = a.x; s.l[1] = a.y; s.l[2] = a.z; s.l[3] = a.w;
    s.l[4] = b.x; s.l[5] = b.y; s.l[6] = b.z; s.l[7] = b.w;
    if (montgomery) s = fp_from_mont(s);
    return s;
}

__global__ void __launch_bounds__(PART_THREADS) msm_part_hist_kernel(MsmJobs jobs, MsmGeom g, SortGeom sg,
                                                                     uint32_t* __restrict__ part_hist) {
    __shared__ uint32_t hist[SORT_MAX_PARTS];
    const uint32_t job = blockIdx.y;
    const Fr* __restrict__ scalars = jobs.scalars[job];
    const uint64_t n = jobs.n[job];
    const bool montgomery = (jobs.montgomery >> job) & 1u;
    const uint64_t first = (uint64_t)blockIdx.x * sg.tile_scalars;
    if (first >= n) return;
    const uint32_t key_base = job * g.nbuckets;
    for (uint32_t p = threadIdx.x; p < sg.nparts; p += PART_THREADS) hist[p] = 0;
    __syncthreads();
    const uint64_t last = min(n, first + sg.tile_scalars);
    for (uint64_t i = first + threadIdx.x; i < last; i += PART_THREADS) {
        const Fr s = load_scalar_stream(scalars, i, montgomery);
        msm_for_digits(s, g, key_base, i, [&](uint32_t key, uint32_t) { atomicAdd(&hist[key >> sg.low_bits], 1u); });
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < sg.nparts; p += PART_THREADS)
        if (hist[p]) atomicAdd(&part_hist[p], hist[p]);
}

// one block: pstart[p] = first entry of partition p (pstart[nparts] = total), part_cursor = copy,
// cstart[p] = first chunk of partition p (cstart[nparts] = number of chunks)
__global__ void __launch_bounds__(SCAN_THREADS) msm_part_scan_kernel(const uint32_t* __restrict__ part_hist, uint32_t nparts,
                                                                     uint32_t* __restrict__ pstart, uint32_t* __restrict__ part_cursor,
                                                                     uint32_t* __restrict__ cstart) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    constexpr uint32_t PER = SORT_MAX_PARTS / SCAN_THREADS;
    const uint32_t base = threadIdx.x * PER;
    uint32_t cnt[PER], acc = 0, cacc = 0;
#pragma unroll
    for (uint32_t k = 0; k < PER; k++) {
        cnt[k] = base + k < nparts ? part_hist[base + k] : 0u;
        acc += cnt[k];
        cacc += (cnt[k] + SORT_CHUNK - 1) / SORT_CHUNK;
    }
    uint32_t tot, ctot;
    uint32_t run = block_scan_u32(acc, sh, tot) - acc;
    uint32_t crun = block_scan_u32(cacc, sh, ctot) - cacc;
#pragma unroll
    for (uint32_t k = 0; k < PER; k++) {
        if (base + k < nparts) {
            pstart[base + k] = run;
            part_cursor[base + k] = run;
            cstart[base + k] = crun;
        }
        run += cnt[k];
        crun += (cnt[k] + SORT_CHUNK - 1) / SORT_CHUNK;
    }
    if (threadIdx.x == 0) {
        pstart[nparts] = tot;
        cstart[nparts] = ctot;
    }
}

// exclusive scan of cnt[0 .. n) (n <= THREADS * PER, in shared memory) into off[0 .. n], off[n] = total; every thread
// of the block calls
template <int THREADS, int PER>
__device__ __forceinline__ void block_exclusive_scan_smem(const uint32_t* cnt, uint32_t* off, uint32_t n, uint32_t* warp_sums) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t base = threadIdx.x * PER;
    uint32_t v[PER], acc = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        v[k] = base + k < n ? cnt[base + k] : 0u;
        acc += v[k];
    }
    uint32_t inc = acc;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    uint32_t carry = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < THREADS / 32; k++) {
        const uint32_t x = warp_sums[k];
        if (k < (int)wid) carry += x;
        tot += x;
    }
    uint32_t run = carry + inc - acc;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        if (base + k < n) off[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) off[n] = tot;
    __syncthreads();
}

__global__ void __launch_bounds__(PART_THREADS, 2) msm_part_scatter_kernel(MsmJobs jobs, MsmGeom g, SortGeom sg,
                                                                           uint32_t* __restrict__ part_cursor,
                                                                           uint2* __restrict__ mid) {
    extern __shared__ __align__(16) unsigned char part_smem[];
    const uint32_t job = blockIdx.y;
    const Fr* __restrict__ scalars = jobs.scalars[job];
    const uint64_t n = jobs.n[job];
    const bool montgomery = (jobs.montgomery >> job) & 1u;
    if ((uint64_t)blockIdx.x * sg.tile_scalars >= n) return;
    const uint32_t key_base = job * g.nbuckets;
    uint2* stage = reinterpret_cast<uint2*>(part_smem);                                             // PART_TILE_ENTRIES pairs
    uint16_t* ranks = reinterpret_cast<uint16_t*>(part_smem + sizeof(uint2) * PART_TILE_ENTRIES);   // rank of every digit in its partition
    uint32_t* hist = reinterpret_cast<uint32_t*>(part_smem + (sizeof(uint2) + sizeof(uint16_t)) * PART_TILE_ENTRIES);
    uint32_t* loff = hist + sg.nparts;       // first staged slot of the partition (nparts + 1 values)
    uint32_t* gbase = loff + sg.nparts + 1;  // first slot in mid[]
    __shared__ uint32_t warp_sums[PART_THREADS / 32];
    for (uint32_t p = threadIdx.x; p < sg.nparts; p += PART_THREADS) hist[p] = 0;
    __syncthreads();
    const uint64_t first = (uint64_t)blockIdx.x * sg.tile_scalars;
    const uint64_t last = min(n, first + sg.tile_scalars);
    Fr sc[PART_ITEMS];
#pragma unroll
    for (int k = 0; k < PART_ITEMS; k++) {
        const uint32_t local = threadIdx.x + k * PART_THREADS;
        const uint64_t i = first + local;
        if (i < last) {
            sc[k] = load_scalar_stream(scalars, i, montgomery);
            uint32_t slot = local * g.nwin;  // one slot per (scalar, window)
            msm_for_digits(sc[k], g, key_base, i, [&](uint32_t key, uint32_t) {
                const uint16_t rk = (uint16_t)atomicAdd(&hist[key >> sg.low_bits], 1u);
                if (KZG_IDX_OK(slot, PART_TILE_ENTRIES, DBG_PART_RANK)) ranks[slot] = rk;
                slot++;
            });
        }
    }
    __syncthreads();
    block_exclusive_scan_smem<PART_THREADS, SORT_MAX_PARTS / PART_THREADS>(hist, loff, sg.nparts, warp_sums);
    // reserve the runs in mid[]
    for (uint32_t p = threadIdx.x; p < sg.nparts; p += PART_THREADS) {
        const uint32_t cnt = hist[p];
        gbase[p] = cnt ? atomicAdd(&part_cursor[p], cnt) : 0u;
    }
#pragma unroll
    for (int k = 0; k < PART_ITEMS; k++) {
        const uint32_t local = threadIdx.x + k * PART_THREADS;
        const uint64_t i = first + local;
        if (i < last) {
            uint32_t slot = local * g.nwin;
            msm_for_digits(sc[k], g, key_base, i, [&](uint32_t key, uint32_t entry) {
                const uint32_t at = loff[key >> sg.low_bits] + ranks[slot++];
                if (KZG_IDX_OK(at, PART_TILE_ENTRIES, DBG_PART_STAGE)) stage[at] = make_uint2(entry, key);
            });
        }
    }
    __syncthreads();
    const uint32_t total = loff[sg.nparts];
    for (uint32_t sidx = threadIdx.x; sidx < total; sidx += PART_THREADS) {
        const uint2 e = stage[sidx];
        const uint32_t p = e.y >> sg.low_bits;
        if (KZG_IDX_OK(gbase[p] + (sidx - loff[p]), KZG_DBG(mid_entries), DBG_PART_MID)) mid[gbase[p] + (sidx - loff[p])] = e;
    }
}

// chunk -> (partition, entry range): the largest p with cstart[p] <= chunk
__device__ __forceinline__ bool chunk_range(const uint32_t* __restrict__ pstart, const uint32_t* __restrict__ cstart,
                                            uint32_t nparts, uint32_t chunk, uint32_t& part, uint32_t& begin, uint32_t& end) {
    if (chunk >= cstart[nparts]) return false;
    uint32_t lo = 0, hi = nparts;  // cstart[lo] <= chunk < cstart[hi]
    while (hi - lo > 1) {
        const uint32_t mid_ = (lo + hi) >> 1;
        if (cstart[mid_] <= chunk) lo = mid_; else hi = mid_;
    }
    part = lo;
    begin = pstart[lo] + (chunk - cstart[lo]) * SORT_CHUNK;
    end = min(pstart[lo + 1], begin + SORT_CHUNK);
    return true;
}

__global__ void __launch_bounds__(CHUNK_THREADS) msm_chunk_hist_kernel(const uint2* __restrict__ mid,
                                                                       const uint32_t* __restrict__ pstart,
                                                                       const uint32_t* __restrict__ cstart, SortGeom sg,
                                                                       uint32_t* __restrict__ counts,
                                                                       uint16_t* __restrict__ chunk_hist) {
    __shared__ uint32_t h[1u << SORT_MAX_LOW];
    uint32_t part, begin, end;
    if (!chunk_range(pstart, cstart, sg.nparts, blockIdx.x, part, begin, end)) return;
    const uint32_t nlow = 1u << sg.low_bits, mask = nlow - 1;
    for (uint32_t k = threadIdx.x; k < nlow; k += CHUNK_THREADS) h[k] = 0;
    __syncthreads();
#pragma unroll 4
    for (uint32_t e = begin + threadIdx.x; e < end; e += CHUNK_THREADS) atomicAdd(&h[__ldg(&mid[e]).y & mask], 1u);
    __syncthreads();
    uint32_t* dst = counts + ((size_t)part << sg.low_bits);
    uint16_t* keep = chunk_hist + ((size_t)blockIdx.x << sg.low_bits);  // (a chunk holds SORT_CHUNK <= 65535 entries)
    for (uint32_t k = threadIdx.x; k < nlow; k += CHUNK_THREADS) {
        keep[k] = (uint16_t)h[k];
        if (h[k] && KZG_IDX_OK(((size_t)part << sg.low_bits) + k, KZG_DBG(nkeys), DBG_CHUNK_COUNTS)) atomicAdd(&dst[k], h[k]);
    }
}

__global__ void __launch_bounds__(CHUNK_THREADS) msm_chunk_scatter_kernel(const uint2* __restrict__ mid,
                                                                          const uint32_t* __restrict__ pstart,
                                                                          const uint32_t* __restrict__ cstart, SortGeom sg,
                                                                          const uint16_t* __restrict__ chunk_hist,
                                                                          uint32_t* __restrict__ cursor,
                                                                          uint32_t* __restrict__ sorted) {
    // the chunk's payloads are staged in shared memory in key order and leave in runs (one run per key: a full
    // 32-byte sector on average) instead of one scattered 4-byte store per entry
    extern __shared__ __align__(16) unsigned char chunk_smem[];
    uint32_t* stage = reinterpret_cast<uint32_t*>(chunk_smem);                 // SORT_CHUNK payloads
    uint32_t* cnt = stage + SORT_CHUNK;                                        // counts, then local cursors
    uint32_t* loff = cnt + (1u << SORT_MAX_LOW);                               // first staged slot of the key (+1)
    uint32_t* base = loff + (1u << SORT_MAX_LOW) + 1;                          // first slot in sorted[] for this chunk's entries of the key
    uint16_t* skey = reinterpret_cast<uint16_t*>(base + (1u << SORT_MAX_LOW)); // key of the staged payload
    __shared__ uint32_t warp_sums[CHUNK_THREADS / 32];
    uint32_t part, begin, end;
    if (!chunk_range(pstart, cstart, sg.nparts, blockIdx.x, part, begin, end)) return;
    const uint32_t nlow = 1u << sg.low_bits, mask = nlow - 1;
    uint32_t* cur = cursor + ((size_t)part << sg.low_bits);
    const uint16_t* keep = chunk_hist + ((size_t)blockIdx.x << sg.low_bits);
    for (uint32_t k = threadIdx.x; k < nlow; k += CHUNK_THREADS) {
        const uint32_t c = keep[k];
        cnt[k] = c;
        base[k] = c ? atomicAdd(&cur[k], c) : 0u;
    }
    __syncthreads();
    block_exclusive_scan_smem<CHUNK_THREADS, (1 << SORT_MAX_LOW) / CHUNK_THREADS>(cnt, loff, nlow, warp_sums);
    for (uint32_t k = threadIdx.x; k < nlow; k += CHUNK_THREADS) cnt[k] = 0;
    __syncthreads();
#pragma unroll 4
    for (uint32_t e = begin + threadIdx.x; e < end; e += CHUNK_THREADS) {
        const uint2 v = __ldg(&mid[e]);
        const uint32_t k = v.y & mask;
        const uint32_t slot = loff[k] + atomicAdd(&cnt[k], 1u);
        if (KZG_IDX_OK(slot, SORT_CHUNK, DBG_CHUNK_STAGE)) {
            stage[slot] = v.x;
            skey[slot] = (uint16_t)k;
        }
    }
    __syncthreads();
    const uint32_t total = end - begin;
    for (uint32_t sidx = threadIdx.x; sidx < total; sidx += CHUNK_THREADS) {
        const uint32_t k = skey[sidx];
        if (KZG_IDX_OK(base[k] + (sidx - loff[k]), KZG_DBG(sorted_entries), DBG_CHUNK_SORTED)) sorted[base[k] + (sidx - loff[k])] = stage[sidx];
    }
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation
// ---------------------------------------------------------------------------------------------
// Random gathers from the multi-GiB window table (tools/micro/gather.cu, B200): a plain load that misses makes the L2
// fetch the whole 128-byte line (127 B of DRAM traffic per 32- or 64-byte gather); with the .L2::64B qualifier it
// fetches 64 B.  The RATE is the same either way -- ~44 G random 32-byte sectors per second, a 64-byte point costs two
// -- but half the DRAM bandwidth stays free for the streaming traffic of the same kernels.
__device__ __forceinline__ uint4 ldg_gather16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// A coordinate is ONE 256-bit access (sm_100: LDG.E.256 / STG.E.256), not two 128-bit ones: half the instructions in the
// load / store queues of the accumulation kernels (ncu --set full of the backward pass: lg_throttle 0.7-0.8 stalls per
// issue with 128-bit accesses).  Every Fq these helpers touch sits at a multiple of 32 bytes: the table, the dense lists,
// the prefix / total arrays are 256-byte-aligned arrays of 32- or 64-byte elements.  (The XYZZ partial sums stay on
// 128-bit accesses: ptxas 12.9 crashes on 256-bit accesses inside the shared group-law functions.)
__device__ __forceinline__ Fq load_fq_gather(const Fq* p) {
    Fq r;
    asm volatile("ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ Fq load_fq_global(const Fq* p) {  // read-only data, streamed
    Fq r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void store_fq_global(Fq* p, const Fq& v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.l[0]), "r"(v.l[1]), "r"(v.l[2]), "r"(v.l[3]),
                 "r"(v.l[4]), "r"(v.l[5]), "r"(v.l[6]), "r"(v.l[7])
                 : "memory");
}
__device__ __forceinline__ G1Affine load_affine_gather(const G1Affine* p) {
    G1Affine r;
    r.x = load_fq_gather(&p->x);
    r.y = load_fq_gather(&p->y);
    return r;
}
__device__ __forceinline__ G1Affine load_affine(const G1Affine* p) {
    G1Affine r;
    r.x = load_fq_global(&p->x);
    r.y = load_fq_global(&p->y);
    return r;
}
__device__ __forceinline__ void store_xyzz(G1XYZZ* p, const G1XYZZ& v) {
    fp_store(&p->x, v.x);
    fp_store(&p->y, v.y);
    fp_store(&p->zz, v.zz);
    fp_store(&p->zzz, v.zzz);
}
__device__ __forceinline__ G1XYZZ load_xyzz(const G1XYZZ* p) {
    G1XYZZ v;
    v.x = fp_load<FqP>(&p->x);
    v.y = fp_load<FqP>(&p->y);
    v.zz = fp_load<FqP>(&p->zz);
    v.zzz = fp_load<FqP>(&p->zzz);
    return v;
}

// Bucket sums of EARLIER pieces of the same MSM over the same bucket geometry (a host-scalar MSM is cut into pieces so
// that its uploads hide behind compute), latest piece first.  A piece starts every bucket of its walk from the sum the
// latest earlier piece left for it -- which already contains the pieces before that one -- so merging the pieces costs
// no group operation at all; a bucket that is empty in this piece is picked up by the reduction the same way.
struct MsmCarry {
    const G1XYZZ* in[2];
    const uint32_t* pbase[2];
};
__device__ __forceinline__ G1XYZZ load_xyzz(const G1XYZZ* p);
__device__ __noinline__ G1XYZZ carry_point(const MsmCarry& cy, uint32_t key) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (cy.in[k]) {
            const uint32_t a = cy.pbase[k][key];
            if (cy.pbase[k][key + 1] != a) return load_xyzz(cy.in[k] + a);
        }
    }
    return xyzz_inf();
}

__device__ __forceinline__ G1Affine load_dense_point(const Fq* xs, const Fq* ys, uint32_t i) {
    G1Affine r;
    r.x = load_fq_global(xs + i);
    r.y = load_fq_global(ys + i);
    return r;
}

// One thread per slice of `slice` consecutive entries of the sorted list: equal work per lane no matter how
// the bucket sizes fluctuate.  Whenever the walk crosses a bucket boundary the running sum is parked in the
// partial-sum slot of (bucket, slice):  pbase[key] + (slice index - first slice of the bucket).
// DIRECT: the list is a dense array of points in bucket order (what the batched-affine rounds below leave), not indices.
// CARRY: the walk opens every bucket with the sum earlier pieces of the same MSM left for it (MsmCarry).
template <bool DIRECT, bool CARRY>
__global__ void __launch_bounds__(128, 4) msm_accumulate_kernel(const G1Affine* __restrict__ bases,
                                                             const Fq* __restrict__ dense_x, const Fq* __restrict__ dense_y,
                                                             const uint32_t* __restrict__ sorted,
                                                             const uint32_t* __restrict__ offsets,
                                                             const uint32_t* __restrict__ pbase, uint32_t nkeys,
                                                             uint32_t slice, G1XYZZ* __restrict__ partials, MsmCarry cy) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = offsets[nkeys];
    const uint64_t begin64 = (uint64_t)t * slice;
    if (begin64 >= total) return;
    const uint32_t begin = (uint32_t)begin64;
    const uint32_t end = (uint32_t)min((uint64_t)total, begin64 + slice);
    // bucket of the first entry: largest key with offsets[key] <= begin (empty buckets repeat an offset: the
    // largest such key is the non-empty one)
    uint32_t lo = 0, hi = nkeys;  // invariant: offsets[lo] <= begin < offsets[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= begin) lo = mid; else hi = mid;
    }
    uint32_t key = lo;
    uint32_t key_end = offsets[key + 1];

    G1XYZZ acc = xyzz_inf();
    if (CARRY && begin == offsets[key]) acc = carry_point(cy, key);  // this slice opens its bucket
    uint32_t e = DIRECT ? begin : sorted[begin];
#ifdef KZG_BOUNDS_CHECK
    for (uint32_t j = begin; j < end; j++) {  // every operand of the slice, before anything is fetched
        if (DIRECT ? !KZG_IDX_OK(j, KZG_DBG(dense_in), DBG_WALK_DENSE)
                   : !KZG_IDX_OK(sorted[j] & 0x7fffffffu, KZG_DBG(table_points), DBG_WALK_GATHER)) return;
    }
#endif
    G1Affine p = DIRECT ? load_dense_point(dense_x, dense_y, e) : load_affine_gather(bases + (e & 0x7fffffffu));
    for (uint32_t j = begin; j < end; j++) {
        if (j == key_end) {  // bucket boundary inside the slice
            if (KZG_IDX_OK(pbase[key] + (t - offsets[key] / slice), KZG_DBG(partials), DBG_WALK_PARTIALS))
                store_xyzz(partials + pbase[key] + (t - offsets[key] / slice), acc);
            acc = xyzz_inf();
            do {
                key++;
                key_end = offsets[key + 1];
            } while (key_end <= j);
            if (CARRY) acc = carry_point(cy, key);
        }
        // prefetch the next entry while this one is being added
        uint32_t e_next = e;
        G1Affine p_next = p;
        if (j + 1 < end) {
            e_next = DIRECT ? j + 1 : sorted[j + 1];
            p_next = DIRECT ? load_dense_point(dense_x, dense_y, e_next) : load_affine_gather(bases + (e_next & 0x7fffffffu));
        }
        if (!DIRECT && (e >> 31)) p.y = fp_neg(p.y);
        xyzz_madd(acc, p);
        e = e_next;
        p = p_next;
    }
    if (KZG_IDX_OK(pbase[key] + (t - offsets[key] / slice), KZG_DBG(partials), DBG_WALK_PARTIALS))
        store_xyzz(partials + pbase[key] + (t - offsets[key] / slice), acc);
}

// ---------------------------------------------------------------------------------------------
// Batched-affine bucket rounds.  The XYZZ walk above pays 10 modmul per entry because it never inverts.  With many
// independent additions in flight the inversions can be shared (Montgomery's trick), and an AFFINE addition
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// then costs 3 modmul for its share of the batch inversion + 3 for the formulas = 6.  One round adds the points of
// every bucket PAIRWISE by position (entries 2i and 2i+1 of the bucket -> point i of the bucket in the next list; an odd
// last entry is copied), so a round halves the list and all of its additions are independent.  After a few rounds the
// dense list that is left (ceil(count / 2^rounds) points per bucket, still in bucket order) goes through the XYZZ
// walk and the bucket reduction as before.  Per round:
//   msm_aff_forward   thread t owns AFF_M consecutive OUTPUT points: denominators d_j (x only: 32 B per operand),
//                     exclusive prefix products -> prefix[j][t], operand handles -> desc[j][t], thread product -> totals[t]
//   fq_batch_inverse  totals <- 1 / totals   (frops.cu, ~4.3 / AFF_M modmul per addition)
//   msm_aff_backward  walks its outputs backwards peeling 1 / d_j off the inverse (2 modmul), finishes the additions,
//                     writes the affine results as two planes (all x, all y: the next forward pass reads x only)
// Exceptional pairs are decided from the operands alone, identically in both kernels: an operand at infinity, P + P
// (tangent: d = 2 y, numerator 3 x^2), P + (-P) = infinity; they put d = 1 into the batch where no quotient is needed.
// ---------------------------------------------------------------------------------------------
constexpr int AFF_THREADS = 128;
struct AffRound {
    const G1Affine* bases;    // round 1: window table / points, addressed through `sorted` (index | sign << 31)
    const uint32_t* sorted;   // round 1 only, nullptr afterwards
    const Fq* in_x;           // later rounds: the dense list left by the previous round, as two planes (all x, all y):
    const Fq* in_y;           // the forward pass needs the x coordinates only and then streams half the bytes
    const uint32_t* off_in;   // bucket offsets of the input list  (nkeys + 1)
    const uint32_t* off_out;  // bucket offsets of the output list (nkeys + 1): scan of ceil(count / 2)
    uint32_t nkeys;
    uint32_t nthreads;        // threads of the round = row stride of prefix[][]
};
enum { AFF_ADD = 0, AFF_TAKE_1 = 1, AFF_TAKE_2 = 2, AFF_DOUBLE = 3, AFF_INF = 4 };

__device__ __forceinline__ Fq load_fq_ldg(const Fq* p) { return load_fq_global(p); }
// operand handles: position in the dense input list, or (round 1) the sorted entry itself: index | sign << 31
// (the sign is applied by the consumer: a prefetch must not touch what it loads -- the first use of a loaded
// register is where the warp waits)
template <bool INDEXED>
__device__ __forceinline__ G1Affine aff_request_point(const AffRound& a, uint32_t h) {
    if (!(INDEXED ? KZG_IDX_OK(h & 0x7fffffffu, KZG_DBG(table_points), DBG_BWD_GATHER) : KZG_IDX_OK(h, KZG_DBG(dense_in), DBG_BWD_GATHER))) {
        G1Affine z;
        z.x = fp_zero<FqP>();
        z.y = fp_zero<FqP>();
        return z;
    }
    if (INDEXED) return load_affine_gather(a.bases + (h & 0x7fffffffu));
    G1Affine p;
    p.x = load_fq_ldg(a.in_x + h);
    p.y = load_fq_ldg(a.in_y + h);
    return p;
}
template <bool INDEXED>
__device__ __forceinline__ G1Affine aff_load_point(const AffRound& a, uint32_t h) {
    G1Affine p = aff_request_point<INDEXED>(a, h);
    if (INDEXED && (h >> 31)) p.y = fp_neg(p.y);
    return p;
}
// both operands of output o's pair
template <bool INDEXED>
__device__ __forceinline__ void aff_request_pair(const AffRound& a, uint32_t o, uint32_t h1, uint32_t h2, G1Affine& p1, G1Affine& p2) {
    // (Measured: letting the forward pass of round 1 save the x coordinates so that this one gathers only the y halves
    // costs more in the forward pass -- 6.4 GB of extra stores at 2^24 points, 5.0 -> 6.3 ms -- than it saves here,
    // 9.7 -> 9.6 ms: the backward pass is not bound by its gathers.)
    p1 = aff_request_point<INDEXED>(a, h1);
    p2 = aff_request_point<INDEXED>(a, h2);
}
// kind of the pair and the denominator it contributes to the batch
__device__ __forceinline__ int aff_classify(const G1Affine& p1, const G1Affine& p2, Fq& d) {
    d = fp_one<FqP>();
    if (g1_affine_is_inf(p1)) return AFF_TAKE_2;
    if (g1_affine_is_inf(p2)) return AFF_TAKE_1;
    if (fp_eq(p1.x, p2.x)) {
        if (fp_eq(p1.y, p2.y) && !fp_is_zero(p1.y)) {
            d = fp_dbl(p1.y);
            return AFF_DOUBLE;
        }
        return AFF_INF;
    }
    d = fp_sub(p2.x, p1.x);
    return AFF_ADD;
}
// Walk over the output points of one thread, one bucket at a time: output o of bucket `key` takes the entries
// 2 (o - out_begin) and 2 (o - out_begin) + 1 of the bucket's input run (the second one only if it exists: `pair`).
// The loops below are real loops, not unrolled ones: a body of five inlined Montgomery products is ~17 KB of code, and
// the unrolled version (8 bodies, every warp running through 150 KB of straight-line code once) was bound by
// instruction fetch (issue-active 15 %).
struct AffCursor {
    uint32_t key, out_begin, out_end, in_begin, in_end;
};
__device__ __forceinline__ AffCursor aff_seek(const AffRound& a, uint32_t o) {
    uint32_t lo = 0, hi = a.nkeys;  // invariant: off_out[lo] <= o < off_out[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.off_out[mid] <= o) lo = mid; else hi = mid;
    }
    AffCursor c;
    c.key = lo;
    c.out_begin = a.off_out[lo];
    c.out_end = a.off_out[lo + 1];
    c.in_begin = a.off_in[lo];
    c.in_end = a.off_in[lo + 1];
    return c;
}
template <bool INDEXED>
__device__ __forceinline__ void aff_handles(const AffRound& a, const AffCursor& c, uint32_t o, uint32_t& h1, uint32_t& h2, bool& pair) {
    const uint32_t i0 = c.in_begin + 2 * (o - c.out_begin);
    pair = i0 + 1 < c.in_end;
    h1 = INDEXED ? a.sorted[i0] : i0;
    h2 = INDEXED ? (pair ? a.sorted[i0 + 1] : 0u) : i0 + 1;
}
__device__ __forceinline__ void aff_step_up(const AffRound& a, AffCursor& c, uint32_t o) {  // o = previous output + 1
    if (o >= c.out_end) {
        do {
            c.key++;
            c.out_end = a.off_out[c.key + 1];
        } while (c.out_end <= o);
        c.out_begin = a.off_out[c.key];
        c.in_begin = a.off_in[c.key];
        c.in_end = a.off_in[c.key + 1];
    }
}

template <bool INDEXED>
__device__ __forceinline__ Fq aff_load_x(const AffRound& a, uint32_t h) {
    if (!(INDEXED ? KZG_IDX_OK(h & 0x7fffffffu, KZG_DBG(table_points), DBG_FWD_GATHER) : KZG_IDX_OK(h, KZG_DBG(dense_in), DBG_FWD_GATHER)))
        return fp_one<FqP>();
    return INDEXED ? load_fq_gather(&(a.bases + (h & 0x7fffffffu))->x) : load_fq_ldg(a.in_x + h);
}

// Outputs of thread t: `count` of them, first, first + step, ...  Plain: m consecutive outputs.  Interleaved (il): the 32
// lanes of a warp share a tile of 32 m consecutive outputs, lane l takes l, l + 32, ... -- at every step a warp then reads
// 32 CONSECUTIVE pairs of the dense list and writes 32 consecutive results (coalesced) instead of 32 runs 2 m points apart.
__device__ __forceinline__ void aff_thread_outputs(uint32_t t, uint32_t m, uint32_t total, bool il, uint32_t& first,
                                                   uint32_t& count, uint32_t& step) {
    const uint64_t f = il ? (uint64_t)(t >> 5) * 32 * m + (t & 31) : (uint64_t)t * m;
    step = il ? 32u : 1u;
    first = (uint32_t)f;
    count = 0;
    if (f < total) {
        const uint64_t left = (total - f + step - 1) / step;
        count = (uint32_t)(left < m ? left : m);
    }
}

// `desc`: the operand handles of every output, desc[j][t] = (h1, h2), h2 == h1 when the output has a single operand
// (handles of one bucket are distinct: an entry is one (window, point) pair) -- the backward pass then needs no bucket
// cursor of its own and its gathers are ONE load away from their addresses instead of three (offsets -> sorted[] -> table).
template <bool INDEXED>
__global__ void __launch_bounds__(AFF_THREADS, 6) msm_aff_forward_kernel(AffRound a, uint32_t m, uint32_t t_first, uint32_t t_end,
                                                                         Fq* __restrict__ prefix, Fq* __restrict__ totals,
                                                                         uint2* __restrict__ desc, bool il) {
    const uint32_t t = t_first + blockIdx.x * blockDim.x + threadIdx.x;  // (a round is launched in chunks of threads)
    if (t >= t_end) return;
    Fq acc = fp_one<FqP>();
    const uint32_t total = a.off_out[a.nkeys];
    uint32_t first, count, step;
    aff_thread_outputs(t, m, total, il, first, count, step);
    if (count) {
        AffCursor c = aff_seek(a, first);
        uint32_t h1, h2;
        bool pair;
        Fq x1, x2;  // x coordinates of the pair in hand, fetched one iteration ahead
        aff_handles<INDEXED>(a, c, first, h1, h2, pair);
        if (pair) {
            x1 = aff_load_x<INDEXED>(a, h1);
            x2 = aff_load_x<INDEXED>(a, h2);
        }
        Fq* pre = prefix + t;
        uint2* dsc = desc + t;
#pragma unroll 1
        for (uint32_t j = 0, o = first; j < count; j++, o += step, pre += a.nthreads) {
            if (KZG_IDX_OK((uint64_t)j * a.nthreads + t, KZG_DBG(prefix_elems), DBG_FWD_PREFIX))
                *dsc = make_uint2(h1, pair ? h2 : h1);
            dsc += a.nthreads;
            uint32_t nh1 = 0, nh2 = 0;
            bool npair = false;
            Fq nx1, nx2;
            if (j + 1 < count) {
                aff_step_up(a, c, o + step);
                aff_handles<INDEXED>(a, c, o + step, nh1, nh2, npair);
                if (npair) {
                    nx1 = aff_load_x<INDEXED>(a, nh1);
                    nx2 = aff_load_x<INDEXED>(a, nh2);
                }
            }
            if (pair && KZG_IDX_OK((uint64_t)j * a.nthreads + t, KZG_DBG(prefix_elems), DBG_FWD_PREFIX)) {
                Fq d;
                if (fp_eq(x1, x2) || fp_is_zero(x1) || fp_is_zero(x2)) {  // exceptional: decide on the full points
                    const G1Affine p1 = aff_load_point<INDEXED>(a, h1), p2 = aff_load_point<INDEXED>(a, h2);
                    aff_classify(p1, p2, d);
                } else {
                    d = fp_sub(x2, x1);
                }
                store_fq_global(pre, acc);
                acc = fp_mul(acc, d);
            }
            h1 = nh1;
            h2 = nh2;
            pair = npair;
            x1 = nx1;
            x2 = nx2;
        }
    }
    if (KZG_IDX_OK(t, KZG_DBG(totals), DBG_FWD_TOTALS)) store_fq_global(totals + t, acc);
}

// Backward pass, driven by the forward pass's descriptors: no bucket cursor of its own (measured against a version that
// walked the offsets again and fetched its handles from sorted[]: 26.5 -> 25.4 ms for the rounds at 2^24 points, 5.96 ->
// 5.69 at 2^22 -- the gathers are one load away from their addresses instead of three, and the loop is shorter).
// Software pipeline: the handles of output o - 2 (8 bytes, coalesced) and the operands and prefix product of output o - 1
// are requested before the addition of output o is computed, and not touched until then (the signs of round 1 are
// applied at use: with fp_neg inside the prefetch the warp waited for the gather right there, 40 % of all stall samples).
template <bool INDEXED>
__global__ void __launch_bounds__(AFF_THREADS, 4) msm_aff_backward_kernel(AffRound a, uint32_t m, uint32_t t_first, uint32_t t_end,
                                                                               const Fq* __restrict__ prefix,
                                                                               const Fq* __restrict__ inv_totals,
                                                                               const uint2* __restrict__ desc,
                                                                               Fq* __restrict__ out_x, Fq* __restrict__ out_y, bool il) {
    const uint32_t t = t_first + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_end) return;
    const uint32_t total = a.off_out[a.nkeys];
    uint32_t first, count, step;
    aff_thread_outputs(t, m, total, il, first, count, step);
    if (count == 0) return;
    Fq s = load_fq_ldg(inv_totals + t);
    const uint64_t top = (uint64_t)(count - 1) * a.nthreads + t;
    const Fq* pre_ptr = prefix + top;
    const uint2* dsc = desc + top;
    uint2 h = __ldg(dsc);                                                  // output j
    uint2 hn = count > 1 ? __ldg(dsc - a.nthreads) : make_uint2(0u, 0u);   // output j - 1
    G1Affine p1, p2;
    Fq pre;
    if (h.y != h.x) {
        aff_request_pair<INDEXED>(a, 0, h.x, h.y, p1, p2);
        pre = load_fq_ldg(pre_ptr);
    } else {
        p1 = aff_request_point<INDEXED>(a, h.x);
    }
#pragma unroll 1
    for (uint32_t j = count - 1, o = first + (count - 1) * step;; j--, o -= step, pre_ptr -= a.nthreads, dsc -= a.nthreads) {
        const bool more = j > 0;
        G1Affine n1, n2;
        Fq npre;
        uint2 hnn = make_uint2(0u, 0u);
        if (more) {
            if (j > 1) hnn = __ldg(dsc - 2 * (size_t)a.nthreads);   // output j - 2
            if (hn.y != hn.x) {
                aff_request_pair<INDEXED>(a, 0, hn.x, hn.y, n1, n2);
                npre = load_fq_ldg(pre_ptr - a.nthreads);
            } else {
                n1 = aff_request_point<INDEXED>(a, hn.x);
            }
        }
        G1Affine r = p1;
        if (INDEXED && (h.x >> 31)) r.y = fp_neg(r.y);
        if (h.y != h.x) {
            if (INDEXED && (h.y >> 31)) p2.y = fp_neg(p2.y);
            Fq d;
            const int kind = aff_classify(r, p2, d);
            const Fq inv = fp_mul(s, pre);
            s = fp_mul(s, d);
            if (kind == AFF_ADD || kind == AFF_DOUBLE) {
                Fq num = fp_sub(p2.y, r.y);
                if (kind == AFF_DOUBLE) {
                    const Fq xx = fp_sqr(r.x);
                    num = fp_add(fp_dbl(xx), xx);
                }
                const Fq lam = fp_mul(num, inv);
                const Fq x3 = fp_sub(fp_sub(fp_sqr(lam), r.x), p2.x);
                r.y = fp_sub(fp_mul(lam, fp_sub(r.x, x3)), r.y);
                r.x = x3;
            } else if (kind == AFF_TAKE_2) {
                r = p2;
            } else if (kind == AFF_INF) {
                r.x = fp_zero<FqP>();
                r.y = fp_zero<FqP>();
            }
        }
        if (KZG_IDX_OK(o, KZG_DBG(aff_cap[0]), DBG_BWD_OUT)) {
            store_fq_global(out_x + o, r.x);
            store_fq_global(out_y + o, r.y);
        }
        if (!more) break;
        p1 = n1;
        p2 = n2;
        pre = npre;
        h = hn;
        hn = hnn;
    }
}

// The kernels after the accumulation are latency-bound (few warps, long dependent chains) and run once per MSM, so
// their instructions come cold out of L2/DRAM: they all CALL one shared copy of the 14-product addition and of the
// doubling instead of inlining ~60 KB of straight-line code at every use.
__device__ __noinline__ void xyzz_add_fn(G1XYZZ& acc, const G1XYZZ& b) { xyzz_add(acc, b); }
__device__ __noinline__ void xyzz_dbl_fn(G1XYZZ& p) { p = xyzz_dbl(p); }
__device__ G1XYZZ xyzz_mul_small_fn(const G1XYZZ& p, uint32_t k) {
    G1XYZZ r = xyzz_inf();
    int top = 31;
    while (top >= 0 && !((k >> top) & 1)) top--;
#pragma unroll 1
    for (int bit = top; bit >= 0; bit--) {
        xyzz_dbl_fn(r);
        if ((k >> bit) & 1) xyzz_add_fn(r, p);
    }
    return r;
}

// Buckets cut by slice boundaries hold 2..HEAVY_PARTS partial sums (about one bucket in two): one thread per such
// bucket folds them into the first slot, so that the reduction reads exactly one point per bucket, divergence-free.
__global__ void __launch_bounds__(128) msm_fold_kernel(G1XYZZ* __restrict__ partials, const uint32_t* __restrict__ pbase,
                                                       const uint32_t* __restrict__ multi,
                                                       const uint32_t* __restrict__ multi_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *multi_count) return;
    const uint32_t key = multi[i];
    const uint32_t a = pbase[key], b = pbase[key + 1];
    G1XYZZ v = load_xyzz(partials + a);
#pragma unroll 1
    for (uint32_t j = a + 1; j < b; j++) {
        G1XYZZ o = load_xyzz(partials + j);
        xyzz_add(v, o);
    }
    store_xyzz(partials + a, v);
}

// A bucket that spans many slices holds many partial sums: one block per such bucket tree-sums them in shared
// memory and leaves the total in the bucket's first slot.  (Buckets with <= HEAVY_PARTS partials are summed by
// the reader, load_bucket.)
constexpr int RED_THREADS = 128;

__device__ __forceinline__ Fq shfl_xor_fq(const Fq& v, uint32_t d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], d);
    return r;
}
// Points in quad form (ec.cuh): lane j of every group of four lanes holds coordinate j.  All-reduce over the
// quads whose lane indices differ in the bits [4, top]: afterwards each of them holds the sum.
__device__ __forceinline__ void quad_butterfly(Fq& acc, uint32_t top, uint32_t j, uint32_t qm) {
    __syncwarp();
#pragma unroll 1
    for (uint32_t d = top; d >= 4; d >>= 1) {
        const Fq o = shfl_xor_fq(acc, d);
        quad_add(acc, o, j, qm);
    }
}

// Heavy buckets (more than HEAVY_PARTS partial sums: skewed scalars, or the short top window whose few digit values
// each collect n / 2^bits points).  Their partial sums are contiguous, so this is a plain sum, done in quad-lane
// arithmetic: one WARP per bucket up to HUGE_PARTS partials (8 quads stride over them, 3-level butterfly), one
// 512-thread BLOCK (128 quads) per bucket beyond.  The total lands in the bucket's first slot.
__global__ void __launch_bounds__(128) msm_collapse_kernel(G1XYZZ* __restrict__ partials, const uint32_t* __restrict__ pbase,
                                                           const uint32_t* __restrict__ heavy,
                                                           const uint32_t* __restrict__ heavy_count) {
    const uint32_t nheavy = *heavy_count;
    const uint32_t lane = threadIdx.x & 31, j = threadIdx.x & 3;
    const uint32_t qm = quad_mask();
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t h = gwarp; h < nheavy; h += nwarps) {
        const uint32_t key = heavy[h];
        const uint32_t a = pbase[key], b = pbase[key + 1];
        Fq acc = fp_zero<FqP>();
#pragma unroll 1
        for (uint32_t k = a + (lane >> 2); k < b; k += 8) {
            const Fq o = quad_load(partials + k, j);
            quad_add(acc, o, j, qm);
        }
        quad_butterfly(acc, 16, j, qm);  // (also orders every read of slot a before the write below)
        if (lane < 4) quad_store(partials + a, j, acc);
        __syncwarp();
    }
}
__global__ void __launch_bounds__(512) msm_collapse_huge_kernel(G1XYZZ* __restrict__ partials, const uint32_t* __restrict__ pbase,
                                                                 const uint32_t* __restrict__ huge,
                                                                 const uint32_t* __restrict__ huge_count) {
    __shared__ G1XYZZ sh[16];
    const uint32_t nhuge = *huge_count;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = threadIdx.x & 3;
    const uint32_t qm = quad_mask();
    for (uint32_t h = blockIdx.x; h < nhuge; h += gridDim.x) {
        const uint32_t key = huge[h];
        const uint32_t a = pbase[key], b = pbase[key + 1];
        Fq acc = fp_zero<FqP>();
#pragma unroll 1
        for (uint32_t k = a + (threadIdx.x >> 2); k < b; k += 128) {
            const Fq o = quad_load(partials + k, j);
            quad_add(acc, o, j, qm);
        }
        quad_butterfly(acc, 16, j, qm);
        if (lane < 4) quad_store(sh + warp, j, acc);
        __syncthreads();
        if (warp == 0) {
            const uint32_t q = lane >> 2;  // quad q adds the sums of warps q and q + 8
            Fq v = quad_load(sh + q, j);
            {
                const Fq o = quad_load(sh + q + 8, j);
                quad_add(v, o, j, qm);
            }
            quad_butterfly(v, 16, j, qm);
            if (lane < 4) quad_store(partials + a, j, v);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// bucket reduction: per set  sum_{b < B} (b+1) S_b  in three launches whose serial depth does not grow with B.
//   level 0   one thread per chunk q of r0 = 2^k0 consecutive buckets (the only pass over all the buckets, pipe-bound):
//               U_q = sum_j S_{r0 q + j}                 (plain)
//               t_q = sum_j (j+1) S_{r0 q + j}           (running-sum trick: 2 additions per bucket)
//             so that  sum_b (b+1) S_b = sum_q t_q + r0 sum_q q U_q.
//   tail      q = hi LO + lo with LO = 2^a ~ sqrt(n1):   sum_q q U_q = sum_lo lo C_lo + LO sum_hi hi R_hi
//             where the column sums C_lo and the row sums R_hi are PLAIN sums -- no serial weighted recurrence left.
//             msm_tail_tasks: one warp per column / row / run of LO t-values: lane-serial partial sums, a shuffle
//             tree, then the weight (< 2^a) by double-and-add on lane 0.
//             msm_tail_final: one block per set adds the three groups and forms  T + r0 (C + LO R).
// Every step of the tail is latency-bound (one XYZZ addition is ~5 us for a lone warp), so what counts is the number
// of dependent additions: ~20 per tail launch, against ~65 for a radix-4 hierarchy of running sums.
// ---------------------------------------------------------------------------------------------

// level 0 reads the accumulate output through pbase: one point per bucket (msm_fold / msm_collapse have folded the
// partial sums of a bucket into its first slot), none for an empty bucket
// (cy: the bucket sums of earlier pieces of the same MSM, MsmCarry -- the pieces share THIS reduction instead of paying
// one each)
__global__ void __launch_bounds__(RED_THREADS, 4) msm_reduce_level0_kernel(const G1XYZZ* __restrict__ in,
                                                                           const uint32_t* __restrict__ pbase,
                                                                           MsmCarry cy, uint32_t nbuckets, uint32_t radix,
                                                                           G1XYZZ* __restrict__ out_u, uint32_t n_out,
                                                                           G1XYZZ* __restrict__ out_t) {
    const uint32_t set = blockIdx.y;
    const uint32_t q = blockIdx.x * RED_THREADS + threadIdx.x;
    if (q >= n_out) return;
    const uint32_t lo = q * radix;
    const uint32_t hi = min(lo + radix, nbuckets);
    G1XYZZ run = xyzz_inf(), tot = xyzz_inf();
#pragma unroll 1
    for (uint32_t b = hi; b-- > lo;) {
        const uint32_t key = set * nbuckets + b;
        const uint32_t a = pbase[key];
        if (pbase[key + 1] != a) {
            G1XYZZ o = load_xyzz(in + a);
            xyzz_add_fn(run, o);
        } else if (cy.in[0]) {  // empty in this piece: what the earlier pieces left (a walk that opens a bucket takes it along)
            G1XYZZ o = carry_point(cy, key);
            xyzz_add_fn(run, o);
        }
        xyzz_add_fn(tot, run);
    }
    store_xyzz(out_u + (size_t)set * n_out + q, run);
    store_xyzz(out_t + (size_t)set * n_out + q, tot);
}

struct TailGeom {
    uint32_t n1;      // chunks per set after level 0
    uint32_t k0;      // log2 of the level-0 radix
    uint32_t log_lo;  // a
    uint32_t lo, hi;  // LO = 2^a columns, HI = ceil(n1 / LO) rows
    uint32_t ntask;   // LO column tasks, HI row tasks, HI runs of t-values
};

// one block (32 quads) per task: a column, a row or a run of t-values
template <int TAIL_THREADS>
__global__ void __launch_bounds__(TAIL_THREADS) msm_tail_tasks_kernel(const G1XYZZ* __restrict__ u,
                                                                      const G1XYZZ* __restrict__ t, TailGeom tg,
                                                                      G1XYZZ* __restrict__ w) {
    __shared__ G1XYZZ sh[TAIL_THREADS / 32];
    const uint32_t set = blockIdx.y, task = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t j = threadIdx.x & 3, quad = threadIdx.x >> 2;
    const uint32_t qm = quad_mask();
    const G1XYZZ* src;
    uint32_t first, step, count, weight;
    if (task < tg.lo) {  // column `task`: q = hi LO + task
        src = u + (size_t)set * tg.n1;
        first = task;
        step = tg.lo;
        count = tg.n1 > task ? (tg.n1 - task + tg.lo - 1) >> tg.log_lo : 0u;
        weight = task;
    } else if (task < tg.lo + tg.hi) {  // row
        const uint32_t h = task - tg.lo;
        src = u + (size_t)set * tg.n1;
        first = h << tg.log_lo;
        step = 1;
        count = min(tg.lo, tg.n1 - first);
        weight = h;
    } else {  // run of t-values
        const uint32_t r = task - tg.lo - tg.hi;
        src = t + (size_t)set * tg.n1;
        first = r << tg.log_lo;
        step = 1;
        count = min(tg.lo, tg.n1 - first);
        weight = 1;
    }
    G1XYZZ* dst = w + (size_t)set * tg.ntask + task;
    if (weight == 0) {  // (block-uniform) column 0 and row 0 carry weight 0
        if (threadIdx.x < 4) quad_store(dst, j, fp_zero<FqP>());
        return;
    }
    Fq acc = fp_zero<FqP>();
#pragma unroll 1
    for (uint32_t k = quad; k < count; k += TAIL_THREADS / 4) {
        const Fq b = quad_load(src + first + (size_t)k * step, j);
        quad_add(acc, b, j, qm);
    }
    quad_butterfly(acc, 16, j, qm);
    if (lane < 4) quad_store(sh + warp, j, acc);
    __syncthreads();
    if (warp == 0) {
        Fq v = quad < TAIL_THREADS / 32 ? quad_load(sh + quad, j) : fp_zero<FqP>();
        if (TAIL_THREADS > 32) quad_butterfly(v, TAIL_THREADS / 32 * 2, j, qm);  // quads 0..W-1: lane distances 2W .. 4
        if (quad == 0) {
            if (weight > 1) quad_mul_small(v, weight, j, qm);
            quad_store(dst, j, v);
        }
    }
}

// one block per set: warps 0-2 add the column results, 3-5 the row results, 6-7 the t results
__global__ void __launch_bounds__(256) msm_tail_final_kernel(const G1XYZZ* __restrict__ w, TailGeom tg,
                                                             G1XYZZ* __restrict__ set_sums) {
    __shared__ G1XYZZ sh[8];
    const uint32_t set = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = threadIdx.x & 3;
    const uint32_t qm = quad_mask();
    uint32_t first, count, nw, rank;
    if (warp < 3) {
        first = 0; count = tg.lo; nw = 3; rank = warp;
    } else if (warp < 6) {
        first = tg.lo; count = tg.hi; nw = 3; rank = warp - 3;
    } else {
        first = tg.lo + tg.hi; count = tg.hi; nw = 2; rank = warp - 6;
    }
    const G1XYZZ* src = w + (size_t)set * tg.ntask + first;
    Fq acc = fp_zero<FqP>();
#pragma unroll 1
    for (uint32_t k = rank * 8 + (lane >> 2); k < count; k += nw * 8) {
        const Fq b = quad_load(src + k, j);
        quad_add(acc, b, j, qm);
    }
    quad_butterfly(acc, 16, j, qm);
    if (lane < 4) quad_store(sh + warp, j, acc);
    __syncthreads();
    // the three group totals on the first quad of three different warps: C, 2^a R, T
    if (lane < 4 && warp < 3) {
        const uint32_t base = warp * 3, n = warp == 2 ? 2u : 3u;
        Fq v = quad_load(sh + base, j);
#pragma unroll 1
        for (uint32_t k = 1; k < n; k++) {
            const Fq o = quad_load(sh + base + k, j);
            quad_add(v, o, j, qm);
        }
        if (warp == 1) {
#pragma unroll 1
            for (uint32_t k = 0; k < tg.log_lo; k++) quad_dbl(v, j, qm);
        }
        quad_store(sh + base, j, v);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        Fq v = quad_load(sh + 3, j);
        Fq o = quad_load(sh + 0, j);
        quad_add(v, o, j, qm);
#pragma unroll 1
        for (uint32_t k = 0; k < tg.k0; k++) quad_dbl(v, j, qm);
        o = quad_load(sh + 6, j);
        quad_add(v, o, j, qm);
        quad_store(set_sums + set, j, v);
    }
}

// result = sum_w 2^(c*w) * windows[w]   (raw flavour)
__global__ void msm_horner_kernel(const G1XYZZ* __restrict__ windows, uint32_t nwin, uint32_t c, G1XYZZ* __restrict__ result) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1XYZZ acc = load_xyzz(windows + (nwin - 1));
    for (int w = (int)nwin - 2; w >= 0; w--) {
#pragma unroll 1
        for (uint32_t k = 0; k < c; k++) acc = xyzz_dbl(acc);
        G1XYZZ o = load_xyzz(windows + w);
        xyzz_add(acc, o);
    }
    store_xyzz(result, acc);
}

// sum `count` XYZZ points into one XYZZ point (the pieces of a host-scalar shard before the all-gather): one warp,
// quad-lane arithmetic as in g1_finish
__global__ void __launch_bounds__(32) g1_sum_kernel(const G1XYZZ* __restrict__ parts, uint32_t count, G1XYZZ* __restrict__ out) {
    if (blockIdx.x != 0) return;
    const uint32_t lane = threadIdx.x & 31, j = lane & 3;
    if (count == 1) {
        if (lane == 0) store_xyzz(out, load_xyzz(parts));
        return;
    }
    const uint32_t qm = quad_mask();
    Fq a = fp_zero<FqP>();
#pragma unroll 1
    for (uint32_t k = lane >> 2; k < count; k += 8) {
        const Fq o = quad_load(parts + k, j);
        quad_add(a, o, j, qm);
    }
    quad_butterfly(a, 16, j, qm);
    if (lane < 4) quad_store(out, j, a);
}

// sum `count` XYZZ points and write the canonical affine encoding.  One warp: the partial points (the shards of a
// multi-GPU MSM, the pieces of a host-scalar MSM) are summed in quad-lane arithmetic -- 8 quads stride over them, then a
// 3-level butterfly -- and lane 0 does the one inversion.
__global__ void __launch_bounds__(32) g1_finish_kernel(const G1XYZZ* __restrict__ parts, uint32_t count, G1Affine* __restrict__ out) {
    if (blockIdx.x != 0) return;
    const uint32_t lane = threadIdx.x & 31, j = lane & 3;
    G1XYZZ acc;
    if (count == 1) {
        acc = load_xyzz(parts);
    } else {
        const uint32_t qm = quad_mask();
        Fq a = fp_zero<FqP>();
#pragma unroll 1
        for (uint32_t k = lane >> 2; k < count; k += 8) {
            const Fq o = quad_load(parts + k, j);
            quad_add(a, o, j, qm);
        }
        quad_butterfly(a, 16, j, qm);
        acc = quad_gather(qm, a);
    }
    if (lane != 0) return;
    G1Affine r = xyzz_to_affine(acc);
    fp_store(&out->x, r.x);
    fp_store(&out->y, r.y);
}

// ---------------------------------------------------------------------------------------------
// SRS window table:  T[w][i] = 2^(c w) P_i.   Stage 1 walks the doubling chain of each point in XYZZ and
// parks every window's value; stage 2 normalises the nwin values of a point with one shared inversion
// (Montgomery trick over ZZ*ZZZ) and writes the affine table.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) srs_table_chain_kernel(const G1Affine* __restrict__ pts, uint64_t first, uint64_t count,
                                                              uint32_t c, uint32_t nwin, G1XYZZ* __restrict__ tmp) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    G1Affine p = load_affine(pts + first + t);
    G1XYZZ acc = xyzz_from_affine(p);
    for (uint32_t w = 1; w < nwin; w++) {
#pragma unroll 1
        for (uint32_t k = 0; k < c; k++) acc = xyzz_dbl(acc);
        store_xyzz(tmp + (size_t)(w - 1) * count + t, acc);
    }
}

__global__ void __launch_bounds__(128) srs_table_normalise_kernel(const G1Affine* __restrict__ pts, uint64_t first,
                                                                  uint64_t count, uint64_t stride, uint32_t nwin,
                                                                  const G1XYZZ* __restrict__ tmp, Fq* __restrict__ prefix,
                                                                  G1Affine* __restrict__ table) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    G1Affine p = load_affine(pts + first + t);
    fp_store(&table[first + t].x, p.x);
    fp_store(&table[first + t].y, p.y);
    if (g1_affine_is_inf(p)) {  // every multiple of the point at infinity is the point at infinity
        for (uint32_t w = 1; w < nwin; w++) {
            fp_store(&table[(size_t)w * stride + first + t].x, fp_zero<FqP>());
            fp_store(&table[(size_t)w * stride + first + t].y, fp_zero<FqP>());
        }
        return;
    }
    // a point of odd prime order never doubles to infinity: every ZZ*ZZZ below is non-zero
    Fq run = fp_one<FqP>();
    for (uint32_t w = 1; w < nwin; w++) {
        const G1XYZZ* v = tmp + (size_t)(w - 1) * count + t;
        fp_store(prefix + (size_t)(w - 1) * count + t, run);
        run = fp_mul(run, fp_mul(fp_load<FqP>(&v->zz), fp_load<FqP>(&v->zzz)));
    }
    Fq inv = fp_inv(run);
    for (uint32_t w = nwin - 1; w >= 1; w--) {
        const G1XYZZ* v = tmp + (size_t)(w - 1) * count + t;
        Fq zz = fp_load<FqP>(&v->zz), zzz = fp_load<FqP>(&v->zzz);
        Fq i_w = fp_mul(inv, fp_load<FqP>(prefix + (size_t)(w - 1) * count + t));  // 1 / (zz zzz)
        inv = fp_mul(inv, fp_mul(zz, zzz));
        G1Affine* dst = table + (size_t)w * stride + first + t;
        fp_store(&dst->x, fp_mul(fp_load<FqP>(&v->x), fp_mul(i_w, zzz)));  // X / ZZ
        fp_store(&dst->y, fp_mul(fp_load<FqP>(&v->y), fp_mul(i_w, zz)));   // Y / ZZZ
    }
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static uint32_t windows_for(uint32_t c, bool montgomery) {
    // Montgomery sources are reduced (< r < 2^254): ceil(255/c) digits leave the top one carry-free.
    // Raw standard-form scalars may use all 256 bits: ceil(257/c).
    const uint32_t bits = montgomery ? 255 : 257;
    return (bits + c - 1) / c;
}

// raw flavour: per-window bucket sets, keep the sets small
static uint32_t auto_window_raw(uint64_t n) {
    uint32_t lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    int c = (int)lg - 4;
    if (c < 4) c = 4;
    if (c > 17) c = 17;
    return (uint32_t)c;
}

// table flavour: one bucket set.  Cost model in units of one mixed addition (0.16 ns at full pipe rate):
//   n * digits(c)                 bucket accumulation
//   x 1.12 if 2^(c-1) > 2^19      direct counting sort only: it keeps one partially written 128 B line per bucket; beyond
//                                 ~64 MB of such lines they no longer fit the 126 MB L2 and every 4-byte store becomes a
//                                 DRAM read-modify-write (measured: 6.3 ms instead of 2.5 ms at 2^24 points).  The
//                                 partition sort (>= 2^27 entries) confines its stores to one partition's window.
//   + 4.6 * 2^(c-1)               bucket reduction (measured: 0.64 ms at 2^19 buckets, 1.96 ms at 2^21, ~0.25 ms of it
//                                 independent of the bucket count)
constexpr uint64_t PART_SORT_MIN_ENTRIES = 1ull << 22;
uint32_t msm_table_window(uint64_t n) {
    uint32_t best = 4;
    double best_cost = 1e300;
    for (uint32_t c = 4; c <= 23; c++) {
        const double entries = (double)n * windows_for(c, false);
        double cost = entries;
        if (c > 20 && entries < (double)PART_SORT_MIN_ENTRIES) cost *= 1.12;
        cost += 4.6 * (double)(1ull << (c - 1));
        // a short top window (scalars are < 2^254) funnels n / 2^bits points into each of its few buckets: thousands of
        // partial sums per bucket for the collapse kernels, hot addresses for the sort (measured: +0.3 ms at 2^18, c = 18)
        const uint32_t top_bits = 254 - c * ((254 + c - 1) / c - 1);
        double slice = 4.0 * (entries / (double)(1ull << (c - 1)) + 1.0);  // (as msm_run picks it)
        if (slice > entries / 303104.0) slice = entries / 303104.0;
        if (slice < 16) slice = 16;
        if (slice > 512) slice = 512;
        const double hot_parts = (double)(n >> top_bits) / slice;  // partial sums per bucket of the top window
        if (hot_parts > 8.0 && entries < (double)PART_SORT_MIN_ENTRIES) cost += 0.1 * entries;  // hot addresses in the direct sort
        if (hot_parts > 256.0) cost += 3.0e6;                      // block-tier collapse: ~50 dependent quad additions
        else if (hot_parts > 8.0) cost += (hot_parts / 8.0 + 3.0) * 15.6e3;  // warp tier: 2.5 us per dependent quad addition
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static MsmGeom msm_geometry(kzg_ctx* ctx, const MsmBases& b, uint64_t n, bool montgomery, uint32_t njobs = 1) {
    MsmGeom g;
    if (b.table) {
        g.c = b.tab_c;
        g.nwin = windows_for(g.c, montgomery);
        if (g.nwin > b.tab_nwin) g.nwin = b.tab_nwin;  // (never: the table is built for raw scalars)
        g.nsets = njobs;
        g.table = 1;
        g.stride = b.stride;
    } else {
        g.c = ctx->msm_window ? ctx->msm_window : auto_window_raw(n);
        if (g.c < 2) g.c = 2;
        if (g.c > 22) g.c = 22;
        g.nwin = windows_for(g.c, montgomery);
        g.nsets = g.nwin;
        g.table = 0;
        g.stride = 0;
    }
    g.nbuckets = 1u << (g.c - 1);
    g.seg = 0;
    return g;
}

// batched-affine rounds before the XYZZ walk.  Measured on B200 (profiles/r01_msm_affine.md, r02_msm_small.md): a round
// over DENSE points runs at 0.10-0.12 ns per addition against 0.16-0.17 ns for the XYZZ walk; round 1 gathers its operands
// from the window table twice (x for the denominators, then the points) and is bound by the rate of random 32-byte
// sectors (44 G/s whatever the table size beyond the L2, tools/micro/gather.cu sweep): 0.15 ns.  The inversion of a round
// is a chain of ~8 small launches (0.13-0.2 ms): it runs on a high-priority side stream under the forward pass of the
// round's later chunks, so what a round still pays is its scan-free launch overhead.  A round is worth it while its
// buckets hold >= aff_min_fill entries on average and the list it leaves keeps >= aff_min_left points.
static uint32_t msm_affine_rounds(kzg_ctx* ctx, uint64_t max_entries, uint32_t nkeys) {
    const MsmTuning& tn = ctx->tuning;
    uint32_t aff_rounds = 0;
    if (max_entries >= tn.aff_min_entries) {
        double fill = (double)max_entries / nkeys;
        uint64_t left = max_entries / 2;
        while (aff_rounds < AFF_MAX_ROUNDS && fill >= tn.aff_min_fill && left >= tn.aff_min_left) {
            aff_rounds++;
            fill *= 0.5;
            left /= 2;
        }
    }
    if (tn.aff_rounds >= 0) aff_rounds = tn.aff_rounds > (int)AFF_MAX_ROUNDS ? AFF_MAX_ROUNDS : (uint32_t)tn.aff_rounds;
    return aff_rounds;
}

// an event for stream-to-stream ordering inside one call, from a per-context ring (no timing, never destroyed before
// the context is): a wait captures the state of the event when it is enqueued, so a slot can be re-recorded later
cudaEvent_t order_event(kzg_ctx* ctx) {
    if (ctx->order_events.size() < 256) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        ctx->order_events.push_back(e);
        return e;
    }
    ctx->order_next = (ctx->order_next + 1) % ctx->order_events.size();
    return ctx->order_events[ctx->order_next];
}

// `jobs.count` MSMs over the same bases as one pipeline; results[j] receives the XYZZ sum of job j (device memory).
// More than one job needs the window-table flavour.
int msm_run_multi(kzg_ctx* ctx, const MsmBases& bases, const MsmJobs& jobs, G1XYZZ* results, const MsmReduceLink* add_in,
                  uint32_t n_add_in, MsmReduceLink* defer_out) {
    const uint32_t njobs = jobs.count;
    if (njobs == 0) return KZG_OK;
    if (njobs > MSM_MAX_JOBS) return set_err(ctx, KZG_ERR_ARG, "msm: too many jobs in one pipeline");
    if (njobs > 1 && !bases.table) return set_err(ctx, KZG_ERR_ARG, "msm: merged jobs need the SRS window table");
    uint64_t n = 0, total_n = 0;  // longest job, all jobs
    bool any_std = false;
    for (uint32_t j = 0; j < njobs; j++) {
        if (jobs.n[j] > n) n = jobs.n[j];
        total_n += jobs.n[j];
        if (!((jobs.montgomery >> j) & 1u)) any_std = true;
    }
    if (total_n == 0) {
        KZG_CUDA(ctx, cudaMemsetAsync(results, 0, sizeof(G1XYZZ) * njobs, ctx->stream));
        return KZG_OK;
    }
    if (n >= (1ull << 27)) return set_err(ctx, KZG_ERR_ARG, "msm: at most 2^27 - 1 points per call");
    const MsmTuning& tn = ctx->tuning;
    // (standard-form scalars may use all 256 bits: one digit more than Montgomery residues at some window sizes; a
    // merged pipeline uses the larger count for all its jobs -- the extra digit of a reduced scalar is zero)
    MsmGeom g = msm_geometry(ctx, bases, n, !any_std, njobs);
    const uint32_t nkeys = g.nsets * g.nbuckets;
    const uint64_t max_entries = total_n * g.nwin;
    if (max_entries >= (1ull << 32)) return set_err(ctx, KZG_ERR_ARG, "msm: n * windows exceeds 2^32 entries");
    if ((uint64_t)nkeys >= (1ull << 31)) return set_err(ctx, KZG_ERR_ARG, "msm: too many bucket keys");
    const uint32_t aff_rounds = msm_affine_rounds(ctx, max_entries, nkeys);
    const uint32_t aff_m = tn.aff_m;  // output points per thread of a round
    // upper bounds of the list lengths: sum_b ceil(k_b / 2) <= (entries + buckets) / 2
    uint64_t aff_entries[AFF_MAX_ROUNDS + 1];
    aff_entries[0] = max_entries;
    // (and a list never grows: with more buckets than entries the first bound alone would exceed the previous length,
    // and the prefix / thread-product arrays are sized by round 1)
    for (uint32_t r = 1; r <= aff_rounds; r++) {
        aff_entries[r] = (aff_entries[r - 1] + nkeys) / 2 + 1;
        if (aff_entries[r] > aff_entries[r - 1]) aff_entries[r] = aff_entries[r - 1];
    }
    const uint64_t walk_entries = aff_entries[aff_rounds];  // what the XYZZ walk sees
    // slice length: a few average buckets (every slice start costs one extra partial sum), but short enough to
    // give every SM several waves of equal-sized tasks
    uint64_t slice = 4 * (walk_entries / nkeys + 1);
    const uint64_t waves = (uint64_t)ctx->sm_count * 512 * 4;
    if (slice > walk_entries / waves) slice = walk_entries / waves;
    if (slice < 16) slice = 16;
    if (slice > 512) slice = 512;
    g.seg = (uint32_t)slice;
    const uint64_t max_tasks = (walk_entries + slice - 1) / slice;
    const uint64_t max_parts = max_tasks + nkeys + 1;

    // bucket reduction geometry: level-0 radix (enough chunks to keep every SM sub-partition busy), then the
    // LO x HI shape of the tail
    TailGeom tg;
    // (measured: 2^19 buckets 0.59 ms with chunks of 16 against 0.62 with 8, 2^21 buckets 1.54 against 1.63)
    tg.k0 = (uint64_t)g.nbuckets * g.nsets >= (1u << 19) ? 4 : 2;
    if (tn.red_k0 >= 0) tg.k0 = (uint32_t)tn.red_k0;
    if (tg.k0 > 6) tg.k0 = 6;
    while (tg.k0 > 0 && (1u << tg.k0) > g.nbuckets) tg.k0--;
    tg.n1 = (g.nbuckets + (1u << tg.k0) - 1) >> tg.k0;
    tg.log_lo = 0;
    while ((1ull << (2 * tg.log_lo)) < tg.n1) tg.log_lo++;
    tg.lo = 1u << tg.log_lo;
    tg.hi = (tg.n1 + tg.lo - 1) >> tg.log_lo;
    tg.ntask = tg.lo + 2 * tg.hi;

    // scratch layout
    const uint32_t nshift = aff_rounds + 1;
    const uint32_t ntiles = (nkeys + SCAN_TILE - 1) / SCAN_TILE;
    size_t off = 0;
    const size_t o_counts = off;   off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_offsets = off;  off = align_up(off + sizeof(uint32_t) * (nkeys + 1) * nshift, 256);  // [shift][nkeys + 1]
    const size_t o_cursor = off;   off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_segoff = off;   off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_sorted = off;   off = align_up(off + sizeof(uint32_t) * max_entries, 256);
    const size_t o_partials = off; off = align_up(off + sizeof(G1XYZZ) * max_parts, 256);
    const size_t o_u = off;        off = align_up(off + sizeof(G1XYZZ) * tg.n1 * g.nsets, 256);
    const size_t o_tv = off;       off = align_up(off + sizeof(G1XYZZ) * tg.n1 * g.nsets, 256);
    const size_t o_lp = off;       off = align_up(off + sizeof(G1XYZZ) * tg.ntask * g.nsets, 256);
    const size_t o_sets = off;     off = align_up(off + sizeof(G1XYZZ) * g.nsets, 256);
    const size_t o_tiles = off;    off = align_up(off + sizeof(uint32_t) * ntiles * nshift, 256);
    const size_t o_heavy = off;    off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_multi = off;    off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_huge = off;     off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    // batched-affine rounds: two point lists (ping-pong), prefix products, thread products
    size_t o_aff_pts[2] = {0, 0}, o_aff_prefix = 0, o_aff_totals = 0, o_aff_desc = 0;
    if (aff_rounds) {
        // (whole warps: with interleaved outputs the 32 lanes of a warp share one tile of 32 m outputs)
        const uint64_t threads1 = ((aff_entries[1] + aff_m - 1) / aff_m + 31) / 32 * 32;
        o_aff_pts[0] = off;   off = align_up(off + sizeof(G1Affine) * aff_entries[1], 256);
        o_aff_pts[1] = off;   off = align_up(off + sizeof(G1Affine) * (aff_rounds > 1 ? aff_entries[2] : 0), 256);
        o_aff_prefix = off;   off = align_up(off + sizeof(Fq) * threads1 * aff_m, 256);
        o_aff_totals = off;   off = align_up(off + sizeof(Fq) * threads1, 256);
        o_aff_desc = off;     off = align_up(off + sizeof(uint2) * threads1 * aff_m, 256);
    }
    // partition sort (large inputs)
    SortGeom sg;
    memset(&sg, 0, sizeof(sg));
    // (measured on B200: 0.24 vs 0.32 ms at 2^20 points, 0.74 vs 1.11 ms at 2^22, 2.9 vs 5.7 ms at 2^24; level below 2^18)
    bool use_part_sort = max_entries >= PART_SORT_MIN_ENTRIES;
    if (tn.part_sort >= 0) use_part_sort = tn.part_sort != 0;
    uint32_t max_chunks = 0;
    size_t o_mid = 0, o_phist = 0, o_pstart = 0, o_pcursor = 0, o_cstart = 0, o_chist = 0;
    if (use_part_sort) {
        uint32_t key_bits = 0;
        while ((1ull << key_bits) < nkeys) key_bits++;
        sg.low_bits = key_bits > 9 ? key_bits - 9 : 0;
        if (sg.low_bits < 4) sg.low_bits = 4;
        if (sg.low_bits > SORT_MAX_LOW) sg.low_bits = SORT_MAX_LOW;
        sg.nparts = (uint32_t)(((uint64_t)nkeys + (1u << sg.low_bits) - 1) >> sg.low_bits);
        if (sg.nparts > SORT_MAX_PARTS) {
            use_part_sort = false;
        } else {
            sg.tile_scalars = PART_TILE_ENTRIES / g.nwin;
            if (sg.tile_scalars > PART_THREADS * PART_ITEMS) sg.tile_scalars = PART_THREADS * PART_ITEMS;
            if (sg.tile_scalars > PART_THREADS) sg.tile_scalars = PART_THREADS;  // equal work per thread
            if (sg.tile_scalars == 0) use_part_sort = false;
        }
    }
    if (use_part_sort) {
        sg.ntiles = (uint32_t)((n + sg.tile_scalars - 1) / sg.tile_scalars);  // per job (grid.x; grid.y = job)
        max_chunks = sg.nparts + (uint32_t)(max_entries / SORT_CHUNK) + 1;
        o_mid = off;     off = align_up(off + sizeof(uint2) * max_entries, 256);
        o_phist = off;   off = align_up(off + sizeof(uint32_t) * (sg.nparts + 1), 256);
        o_pstart = off;  off = align_up(off + sizeof(uint32_t) * (sg.nparts + 1), 256);
        o_pcursor = off; off = align_up(off + sizeof(uint32_t) * (sg.nparts + 1), 256);
        o_cstart = off;  off = align_up(off + sizeof(uint32_t) * (sg.nparts + 1), 256);
        o_chist = off;   off = align_up(off + (sizeof(uint16_t) * max_chunks << sg.low_bits), 256);
    }
    void* base = nullptr;
    KZG_TRY(ctx_scratch(ctx, off, &base));
    uint8_t* sc = (uint8_t*)base;
    uint32_t* counts = (uint32_t*)(sc + o_counts);
    uint32_t* offsets = (uint32_t*)(sc + o_offsets);
    uint32_t* cursor = (uint32_t*)(sc + o_cursor);
    uint32_t* segoff = (uint32_t*)(sc + o_segoff);
    uint32_t* sorted = (uint32_t*)(sc + o_sorted);
    G1XYZZ* partials = (G1XYZZ*)(sc + o_partials);
    G1XYZZ* u_arrays = (G1XYZZ*)(sc + o_u);
    G1XYZZ* tvals = (G1XYZZ*)(sc + o_tv);
    G1XYZZ* tail_parts = (G1XYZZ*)(sc + o_lp);
    G1XYZZ* set_sums = (G1XYZZ*)(sc + o_sets);
    uint32_t* tile_sums = (uint32_t*)(sc + o_tiles);
    uint32_t* heavy = (uint32_t*)(sc + o_heavy);  // [0] = count, [1..] = keys
    uint32_t* multi = (uint32_t*)(sc + o_multi);  // same layout: buckets with 2..HEAVY_PARTS partial sums
    uint32_t* huge = (uint32_t*)(sc + o_huge);    // same layout: buckets with more than HUGE_PARTS partial sums
    uint2* mid = (uint2*)(sc + o_mid);
    uint32_t* part_hist = (uint32_t*)(sc + o_phist);
    uint32_t* pstart = (uint32_t*)(sc + o_pstart);
    uint32_t* part_cursor = (uint32_t*)(sc + o_pcursor);
    uint32_t* cstart = (uint32_t*)(sc + o_cstart);
    uint16_t* chunk_hist = (uint16_t*)(sc + o_chist);

    const G1Affine* pts = bases.table ? bases.table : bases.pts;
#ifdef KZG_BOUNDS_CHECK
    MsmDebugLimits lim;
    memset(&lim, 0, sizeof(lim));
    {
        KZG_CUDA(ctx, cudaDeviceSynchronize());  // the limits are one device global: debug builds run the lanes one after the other
        lim.sorted_entries = max_entries;
        lim.mid_entries = max_entries;
        lim.partials = max_parts;
        lim.prefix_elems = aff_rounds ? (((aff_entries[1] + aff_m - 1) / aff_m + 31) / 32 * 32) * aff_m : 0;
        lim.totals = aff_rounds ? ((aff_entries[1] + aff_m - 1) / aff_m + 31) / 32 * 32 : 0;
        lim.table_points = bases.table ? (uint64_t)(g.nwin - 1) * g.stride + n : n;
        lim.nkeys = nkeys;
        const unsigned int zero = 0;
        KZG_CUDA(ctx, cudaMemcpyToSymbol(g_dbg, &lim, sizeof(lim)));
        KZG_CUDA(ctx, cudaMemcpyToSymbol(g_dbg_violation, &zero, sizeof(zero)));
    }
#endif
    KZG_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (nkeys + 1), ctx->stream));
    KZG_CUDA(ctx, cudaMemsetAsync(heavy, 0, sizeof(uint32_t), ctx->stream));
    KZG_CUDA(ctx, cudaMemsetAsync(multi, 0, sizeof(uint32_t), ctx->stream));
    KZG_CUDA(ctx, cudaMemsetAsync(huge, 0, sizeof(uint32_t), ctx->stream));
    const dim3 dgrid((uint32_t)((n + 255) / 256), njobs);
    timed_begin(ctx, KZG_TIMED_MSM_SORT);
    if (use_part_sort) {
        static const size_t part_smem_max = (sizeof(uint2) + sizeof(uint16_t)) * PART_TILE_ENTRIES + sizeof(uint32_t) * (3 * SORT_MAX_PARTS + 2);
        const size_t part_smem = (sizeof(uint2) + sizeof(uint16_t)) * PART_TILE_ENTRIES + sizeof(uint32_t) * (3 * sg.nparts + 2);
        static bool attr_set[64] = {};
        if (!attr_set[ctx->device & 63]) {
            KZG_CUDA(ctx, cudaFuncSetAttribute(msm_part_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_max));
            KZG_CUDA(ctx, cudaFuncSetAttribute(msm_chunk_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHUNK_SMEM));
            attr_set[ctx->device & 63] = true;
        }
        const dim3 tgrid(sg.ntiles, njobs);
        KZG_CUDA(ctx, cudaMemsetAsync(part_hist, 0, sizeof(uint32_t) * (sg.nparts + 1), ctx->stream));
        KZG_LAUNCH(ctx, msm_part_hist_kernel, tgrid, PART_THREADS, 0, jobs, g, sg, part_hist);
        KZG_LAUNCH(ctx, msm_part_scan_kernel, 1, SCAN_THREADS, 0, part_hist, sg.nparts, pstart, part_cursor, cstart);
        KZG_LAUNCH(ctx, msm_part_scatter_kernel, tgrid, PART_THREADS, part_smem, jobs, g, sg, part_cursor, mid);
        KZG_LAUNCH(ctx, msm_chunk_hist_kernel, max_chunks, CHUNK_THREADS, 0, mid, pstart, cstart, sg, counts, chunk_hist);
    } else {
        KZG_LAUNCH(ctx, msm_digits_kernel<false>, dgrid, 256, 0, jobs, g, counts, nullptr);
    }
    // bucket offsets of the sorted list and of every round's output list (one scan set for all of them)
    KZG_LAUNCH(ctx, msm_offsets_tiles_kernel, ntiles, SCAN_THREADS, 0, counts, nkeys, nshift, ntiles, tile_sums);
    KZG_LAUNCH(ctx, msm_offsets_sums_kernel, nshift, SCAN_THREADS, 0, tile_sums, ntiles, nkeys, offsets);
    KZG_LAUNCH(ctx, msm_offsets_apply_kernel, ntiles, SCAN_THREADS, 0, counts, nkeys, nshift, ntiles, tile_sums, offsets, cursor);
    if (use_part_sort)
        KZG_LAUNCH(ctx, msm_chunk_scatter_kernel, max_chunks, CHUNK_THREADS, CHUNK_SMEM, mid, pstart, cstart, sg, chunk_hist, cursor,
                   sorted);
    else
        KZG_LAUNCH(ctx, msm_digits_kernel<true>, dgrid, 256, 0, jobs, g, cursor, sorted);
    timed_end(ctx, KZG_TIMED_MSM_SORT);

    // batched-affine rounds: (sorted, offsets) -> dense point lists, half as long each time.  A round is launched in
    // chunks of threads: all forward chunks first, each followed by the inversion of ITS thread products on the lane's
    // high-priority side stream, then the backward chunks, each waiting for its own inversion only -- the latency-bound
    // inversion chains (~8 small launches) run under the remaining forward chunks and the earlier backward chunks.
    const uint32_t* walk_offsets = offsets;
    const Fq *walk_x = nullptr, *walk_y = nullptr;  // the dense list the rounds leave: a plane of x and a plane of y
    if (aff_rounds) timed_begin(ctx, KZG_TIMED_MSM_AFFINE);
    cudaStream_t main_stream = ctx->stream;
    cudaStream_t side_stream = ctx->side_stream[ctx->lane];
    cudaStream_t inv_stream = ctx->inv_stream[ctx->lane];
    int rr = KZG_OK;
    for (uint32_t r = 1; r <= aff_rounds && rr == KZG_OK; r++) {
        uint32_t* off_out = offsets + (size_t)r * (nkeys + 1);
        Fq* out_x = (Fq*)(sc + o_aff_pts[(r & 1) ^ 1]);
        Fq* out_y = out_x + aff_entries[(r & 1) ? 1 : 2];  // (capacity of that buffer in points)
        Fq* prefix = (Fq*)(sc + o_aff_prefix);
        Fq* totals = (Fq*)(sc + o_aff_totals);
        uint2* desc = (uint2*)(sc + o_aff_desc);
        const bool aff_il = tn.aff_interleave == 2 || (tn.aff_interleave == 1 && r > 1);
        AffRound ar;
        ar.bases = pts;
        ar.sorted = r == 1 ? sorted : nullptr;
        ar.in_x = walk_x;
        ar.in_y = walk_y;
        ar.off_in = walk_offsets;
        ar.off_out = off_out;
        ar.nkeys = nkeys;
        ar.nthreads = (uint32_t)(((aff_entries[r] + aff_m - 1) / aff_m + 31) / 32 * 32);
#ifdef KZG_BOUNDS_CHECK
        KZG_CUDA(ctx, cudaDeviceSynchronize());
        lim.aff_cap[0] = aff_entries[r];                       // what this round may write
        lim.dense_in = r == 1 ? 0 : aff_entries[r - 1];        // what it may read from the previous round's list
        KZG_CUDA(ctx, cudaMemcpyToSymbol(g_dbg, &lim, sizeof(lim)));
#endif
        // Chunks of threads, a multiple of the block size each, alternating between the lane's main stream and its side
        // stream: consecutive kernels of ONE stream do not overlap, so with all chunks on one stream every chunk would
        // expose its own tail (a block lives ~90 us: measured +1.6 ms at 2^24 points with 4 chunks per round); on two
        // streams the next chunk's blocks fill the previous chunk's tail.
        uint32_t nchunks = tn.aff_chunks > 0 ? (uint32_t)tn.aff_chunks : 2;
        const uint32_t min_chunk_threads = (uint32_t)ctx->sm_count * AFF_THREADS * 8;
        while (nchunks > 1 && ar.nthreads / nchunks < min_chunk_threads) nchunks--;
        if (nchunks > 16) nchunks = 16;
        uint32_t chunk_threads = (ar.nthreads + nchunks - 1) / nchunks;
        chunk_threads = (chunk_threads + AFF_THREADS - 1) / AFF_THREADS * AFF_THREADS;
        nchunks = (ar.nthreads + chunk_threads - 1) / chunk_threads;
        const bool piped = nchunks > 1;
        cudaEvent_t inv_done[16];
        auto mark = [&](cudaStream_t st, const char* what, uint32_t k) {  // KZGB200_TIMELINE=1: where does a round's time go
            if (!tn.timeline) return;
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            char label[64];
            snprintf(label, sizeof(label), "r%u %s%u", r, what, k);
            ctx->timeline.push_back({e, label});
        };
        mark(main_stream, "round_begin", 0);
        if (piped) {  // the side stream joins in: everything queued so far (the previous round) is its input too
            cudaEvent_t fork = order_event(ctx);
            cudaEventRecord(fork, main_stream);
            cudaStreamWaitEvent(side_stream, fork, 0);
        }
        for (uint32_t k = 0; k < nchunks; k++) {
            const uint32_t t0 = k * chunk_threads;
            const uint32_t t1 = t0 + chunk_threads < ar.nthreads ? t0 + chunk_threads : ar.nthreads;
            const uint32_t blocks = (t1 - t0 + AFF_THREADS - 1) / AFF_THREADS;
            ctx->stream = (k & 1) ? side_stream : main_stream;
            if (r == 1)
                KZG_LAUNCH(ctx, msm_aff_forward_kernel<true>, blocks, AFF_THREADS, 0, ar, aff_m, t0, t1, prefix, totals, desc, aff_il);
            else
                KZG_LAUNCH(ctx, msm_aff_forward_kernel<false>, blocks, AFF_THREADS, 0, ar, aff_m, t0, t1, prefix, totals, desc, aff_il);
            mark(ctx->stream, "F_end", k);
            if (piped) {
                cudaEvent_t fwd_done = order_event(ctx);
                inv_done[k] = order_event(ctx);
                cudaEventRecord(fwd_done, ctx->stream);
                cudaStreamWaitEvent(inv_stream, fwd_done, 0);
                ctx->stream = inv_stream;
                mark(inv_stream, "I_begin", k);
                rr = fq_batch_inverse(ctx, totals + t0, totals + t0, t1 - t0);
                mark(inv_stream, "I_end", k);
                cudaEventRecord(inv_done[k], inv_stream);
                if (rr != KZG_OK) {
                    nchunks = k + 1;
                    break;
                }
            } else {
                rr = fq_batch_inverse(ctx, totals, totals, ar.nthreads);
            }
        }
        for (uint32_t k = 0; k < nchunks; k++) {
            const uint32_t t0 = k * chunk_threads;
            const uint32_t t1 = t0 + chunk_threads < ar.nthreads ? t0 + chunk_threads : ar.nthreads;
            const uint32_t blocks = (t1 - t0 + AFF_THREADS - 1) / AFF_THREADS;
            ctx->stream = (k & 1) ? side_stream : main_stream;
            if (piped) cudaStreamWaitEvent(ctx->stream, inv_done[k], 0);  // (also on the error path: the inversion stream drains)
            if (rr != KZG_OK) continue;
            mark(ctx->stream, "B_begin", k);
            if (r == 1)
                KZG_LAUNCH(ctx, msm_aff_backward_kernel<true>, blocks, AFF_THREADS, 0, ar, aff_m, t0, t1, prefix, totals, desc, out_x, out_y, aff_il);
            else
                KZG_LAUNCH(ctx, msm_aff_backward_kernel<false>, blocks, AFF_THREADS, 0, ar, aff_m, t0, t1, prefix, totals, desc, out_x, out_y, aff_il);
            mark(ctx->stream, "B_end", k);
        }
        ctx->stream = main_stream;
        if (piped) {  // join: the main stream continues after the side stream's chunks
            cudaEvent_t join = order_event(ctx);
            cudaEventRecord(join, side_stream);
            cudaStreamWaitEvent(main_stream, join, 0);
        }
        walk_offsets = off_out;
        walk_x = out_x;
        walk_y = out_y;
    }
    KZG_TRY(rr);
    if (aff_rounds) timed_end(ctx, KZG_TIMED_MSM_AFFINE);
    if (tn.timeline && !ctx->timeline.empty()) {
        cudaStreamSynchronize(main_stream);
        cudaStreamSynchronize(inv_stream);
        cudaStreamSynchronize(side_stream);
        for (auto& m : ctx->timeline) {
            float t = 0;
            cudaEventElapsedTime(&t, ctx->timeline[0].first, m.first);
            fprintf(stderr, "[timeline] %-16s %8.3f ms\n", m.second.c_str(), t);
            if (&m != &ctx->timeline[0]) cudaEventDestroy(m.first);
        }
        cudaEventDestroy(ctx->timeline[0].first);
        ctx->timeline.clear();
    }

    // partial-sum slots of the XYZZ walk over what is left
    KZG_LAUNCH(ctx, msm_scan_tiles_kernel, ntiles, SCAN_THREADS, 0, 1, counts, walk_offsets, nkeys, g.seg, aff_rounds, tile_sums);
    KZG_LAUNCH(ctx, msm_scan_sums_kernel, 1, SCAN_THREADS, 0, tile_sums, ntiles, nkeys, segoff);
    KZG_LAUNCH(ctx, msm_scan_apply_kernel, ntiles, SCAN_THREADS, 0, 1, counts, walk_offsets, nkeys, g.seg, aff_rounds, tile_sums,
               segoff, cursor, heavy + 1, heavy, multi + 1, multi, huge + 1, huge);
    const uint32_t ablocks = (uint32_t)((max_tasks + 127) / 128);
#ifdef KZG_BOUNDS_CHECK
    KZG_CUDA(ctx, cudaDeviceSynchronize());
    lim.dense_in = aff_rounds ? aff_entries[aff_rounds] : 0;
    KZG_CUDA(ctx, cudaMemcpyToSymbol(g_dbg, &lim, sizeof(lim)));
#endif
    // earlier pieces of the same MSM (add_in[0] the oldest): the walk opens every bucket with what they left, latest first
    MsmCarry cy;
    memset(&cy, 0, sizeof(cy));
    if (n_add_in > 2) return set_err(ctx, KZG_ERR_ARG, "msm: at most two earlier pieces");
    for (uint32_t k = 0; k < n_add_in; k++) {
        if (add_in[k].nkeys != nkeys) return set_err(ctx, KZG_ERR_ARG, "msm: linked pieces must share the bucket geometry");
        KZG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, add_in[k].ready, 0));
        cy.in[n_add_in - 1 - k] = add_in[k].partials;
        cy.pbase[n_add_in - 1 - k] = add_in[k].pbase;
    }
    timed_begin(ctx, KZG_TIMED_MSM_ACCUMULATE);
    if (aff_rounds && n_add_in)
        KZG_LAUNCH(ctx, (msm_accumulate_kernel<true, true>), ablocks, 128, 0, (const G1Affine*)nullptr, walk_x, walk_y,
                   (const uint32_t*)nullptr, walk_offsets, segoff, nkeys, g.seg, partials, cy);
    else if (aff_rounds)
        KZG_LAUNCH(ctx, (msm_accumulate_kernel<true, false>), ablocks, 128, 0, (const G1Affine*)nullptr, walk_x, walk_y,
                   (const uint32_t*)nullptr, walk_offsets, segoff, nkeys, g.seg, partials, cy);
    else if (n_add_in)
        KZG_LAUNCH(ctx, (msm_accumulate_kernel<false, true>), ablocks, 128, 0, pts, (const Fq*)nullptr, (const Fq*)nullptr, sorted,
                   offsets, segoff, nkeys, g.seg, partials, cy);
    else
        KZG_LAUNCH(ctx, (msm_accumulate_kernel<false, false>), ablocks, 128, 0, pts, (const Fq*)nullptr, (const Fq*)nullptr, sorted,
                   offsets, segoff, nkeys, g.seg, partials, cy);
    timed_end(ctx, KZG_TIMED_MSM_ACCUMULATE);
    timed_begin(ctx, KZG_TIMED_MSM_REDUCE);
    KZG_LAUNCH(ctx, msm_collapse_huge_kernel, 128, 512, 0, partials, segoff, huge + 1, huge);
    KZG_LAUNCH(ctx, msm_collapse_kernel, (uint32_t)ctx->sm_count * 4, 128, 0, partials, segoff, heavy + 1, heavy);
    {
        // every multi-part bucket contains a slice boundary: at most max_tasks of them
        const uint64_t max_multi = max_tasks < nkeys ? max_tasks : nkeys;
        KZG_LAUNCH(ctx, msm_fold_kernel, (uint32_t)((max_multi + 127) / 128), 128, 0, partials, segoff, multi + 1, multi);
    }
    if (defer_out) {  // this piece's bucket sums go into a later piece's reduction
        defer_out->partials = partials;
        defer_out->pbase = segoff;
        defer_out->nkeys = nkeys;
        defer_out->ready = order_event(ctx);
        KZG_CUDA(ctx, cudaEventRecord(defer_out->ready, ctx->stream));
        timed_end(ctx, KZG_TIMED_MSM_REDUCE);
        KZG_CHECK_LAUNCH(ctx);
        return KZG_OK;
    }
    // table flavour: every bucket set IS a result (one per job); raw flavour: the sets are the windows of one result
    G1XYZZ* sums_out = g.table ? results : set_sums;
    {
        const dim3 grid((tg.n1 + RED_THREADS - 1) / RED_THREADS, g.nsets);
        // (measured: calling the shared addition at 126 registers / 16 warps per SM beats the inlined 168-register
        // version with its spills: 0.60 vs 0.66 ms for the whole reduction at 2^19 buckets)
        KZG_LAUNCH(ctx, msm_reduce_level0_kernel, grid, RED_THREADS, 0, partials, segoff, cy, g.nbuckets, 1u << tg.k0, u_arrays, tg.n1,
                   tvals);
    }
    {
        // task width: all the tasks should be resident at once (one wave)
        int width = (uint64_t)tg.ntask * g.nsets > (uint64_t)ctx->sm_count * 3 ? 64 : 128;
        if (tn.tail_width > 0) width = tn.tail_width;
        if (width == 32)
            KZG_LAUNCH(ctx, msm_tail_tasks_kernel<32>, dim3(tg.ntask, g.nsets), 32, 0, u_arrays, tvals, tg, tail_parts);
        else if (width == 64)
            KZG_LAUNCH(ctx, msm_tail_tasks_kernel<64>, dim3(tg.ntask, g.nsets), 64, 0, u_arrays, tvals, tg, tail_parts);
        else
            KZG_LAUNCH(ctx, msm_tail_tasks_kernel<128>, dim3(tg.ntask, g.nsets), 128, 0, u_arrays, tvals, tg, tail_parts);
    }
    KZG_LAUNCH(ctx, msm_tail_final_kernel, g.nsets, 256, 0, tail_parts, tg, sums_out);
    if (!g.table) KZG_LAUNCH(ctx, msm_horner_kernel, 1, 32, 0, set_sums, g.nwin, g.c, results);
    timed_end(ctx, KZG_TIMED_MSM_REDUCE);
    KZG_CHECK_LAUNCH(ctx);
#ifdef KZG_BOUNDS_CHECK
    {
        KZG_CUDA(ctx, cudaDeviceSynchronize());
        unsigned int bad = 0;
        KZG_CUDA(ctx, cudaMemcpyFromSymbol(&bad, g_dbg_violation, sizeof(bad)));
        if (bad) {
            char msg[128];
            snprintf(msg, sizeof(msg), "msm: out-of-bounds index caught by the debug build, site mask 0x%x (enum DBG_* in msm.cu)", bad);
            return set_err(ctx, KZG_ERR_CUDA, msg);
        }
    }
#endif
    return KZG_OK;
}

int msm_run(kzg_ctx* ctx, const MsmBases& bases, MsmScalarSrc src, uint64_t n, G1XYZZ* result_dev) {
    MsmJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.scalars[0] = src.scalars;
    jobs.n[0] = n;
    jobs.montgomery = src.montgomery ? 1u : 0u;
    jobs.count = 1;
    return msm_run_multi(ctx, bases, jobs, result_dev, nullptr, 0, nullptr);
}
// one piece of a linked pair (see MsmReduceLink)
static int msm_run_linked(kzg_ctx* ctx, const MsmBases& bases, MsmScalarSrc src, uint64_t n, G1XYZZ* result_dev,
                          const MsmReduceLink* add_in, uint32_t n_add_in, MsmReduceLink* defer_out) {
    MsmJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.scalars[0] = src.scalars;
    jobs.n[0] = n;
    jobs.montgomery = src.montgomery ? 1u : 0u;
    jobs.count = 1;
    return msm_run_multi(ctx, bases, jobs, result_dev, add_in, n_add_in, defer_out);
}

int msm_result_to_host_affine(kzg_ctx* ctx, const G1XYZZ* result_dev, uint32_t count, uint8_t out[64]) {
    static_assert(sizeof(G1Affine) == 64, "affine layout");
    G1Affine* slot = (G1Affine*)ctx->dev_small;
    timed_begin(ctx, KZG_TIMED_MSM_FINISH);
    KZG_LAUNCH(ctx, g1_finish_kernel, 1, 32, 0, result_dev, count, slot);
    timed_end(ctx, KZG_TIMED_MSM_FINISH);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, slot, 64, cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->pinned, 64);
    return KZG_OK;
}

// sum of two XYZZ points (the halves of a split MSM)
__global__ void g1_add2_kernel(const G1XYZZ* __restrict__ parts, G1XYZZ* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1XYZZ a = load_xyzz(parts);
    G1XYZZ b = load_xyzz(parts + 1);
    xyzz_add(a, b);
    store_xyzz(out, a);
}

// The two halves of the input are issued on the two lanes, so that one half's latency-bound tail (bucket reduction,
// gather) overlaps the other half's pipe-bound accumulation; the two partial points are added at the end.
// Measured on B200: pays only around 2^23 points (21.2 vs 21.8 ms) -- below that the second bucket reduction (its
// cost is per bucket, not per point) eats the overlap (2^20: 3.97 vs 3.71 ms), above it the tail is negligible.
int msm_run_split(kzg_ctx* ctx, const MsmBases& bases, MsmScalarSrc src, uint64_t n, G1XYZZ* result_dev) {
    // Two halves on the two lanes (the latency-bound tail of one under the accumulation of the other) won 3 % around
    // 2^23 points before the batched-affine rounds; with them one undivided MSM is faster everywhere (2^23: 18.5 vs
    // 20.0 ms, 2^24: 33.7 vs 35.7 ms, 2^22: 9.8 vs 11.1 ms), so the range is empty unless set for an experiment.
    uint64_t split_min = 0, split_max = 0;  // split when split_min < n <= split_max
    if (ctx->tuning.split_min_log >= 0) split_min = 1ull << ctx->tuning.split_min_log;
    if (ctx->tuning.split_max_log >= 0) split_max = 1ull << ctx->tuning.split_max_log;
    if (n <= split_min || n > split_max || ctx->lane != 0 || ctx->no_split) return msm_run(ctx, bases, src, n, result_dev);
    const uint64_t h = n / 2;
    G1XYZZ* halves = (G1XYZZ*)(ctx->dev_small + 12288);
    cudaStream_t main_stream = ctx->stream;
    KZG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
    KZG_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    int r = msm_run(ctx, bases, src, h, halves);
    if (r == KZG_OK) {
        MsmBases b2 = bases;
        b2.pts = bases.pts + h;
        if (bases.table) b2.table = bases.table + h;
        ctx->lane = 1;
        ctx->stream = ctx->aux_stream;
        r = msm_run(ctx, b2, MsmScalarSrc{src.scalars + h, src.montgomery}, n - h, halves + 1);
        ctx->lane = 0;
        ctx->stream = main_stream;
    }
    cudaEventRecord(ctx->ev_join, ctx->aux_stream);
    cudaStreamWaitEvent(main_stream, ctx->ev_join, 0);
    KZG_TRY(r);
    KZG_LAUNCH(ctx, g1_add2_kernel, 1, 32, 0, halves, result_dev);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

// count affine results out of count XYZZ sums, one warp each (lane 0 of the warp does the inversion)
__global__ void __launch_bounds__(32) g1_finish_many_kernel(const G1XYZZ* __restrict__ sums, G1Affine* __restrict__ out) {
    if (threadIdx.x != 0) return;
    const G1XYZZ acc = load_xyzz(sums + blockIdx.x);
    const G1Affine r = xyzz_to_affine(acc);
    fp_store(&out[blockIdx.x].x, r.x);
    fp_store(&out[blockIdx.x].y, r.y);
}

// Independent MSMs (the commitments of one prover round).  Jobs that share one window table are MERGED into one
// pipeline (msm_run_multi: one sort, one accumulation, one reduction with a bucket set per job -- the fixed ~0.85 ms
// tail of an MSM is paid once, and the merged list is long enough for the batched-affine rounds); what cannot be merged
// (no table, different bases, more than MSM_MAX_JOBS) runs as before, group i on lane i & 1 -- lane 1 is the context's
// auxiliary stream with its own scratch arena -- so that the latency-bound tail of one group overlaps the pipe-bound
// accumulation of the other.  Every job leaves its affine result in its own 64-byte slot; one D2H copy and one
// synchronisation fetch them all.
int msm_run_batch(kzg_ctx* ctx, const MsmJob* jobs, uint32_t count, uint8_t* out_affine) {
    if (count == 0) return KZG_OK;
    if (count > 30) {  // (the result slots hold 30: longer batches go in segments, one synchronisation each)
        for (uint32_t i = 0; i < count; i += 30) KZG_TRY(msm_run_batch(ctx, jobs + i, count - i < 30 ? count - i : 30, out_affine + 64 * i));
        return KZG_OK;
    }
    G1Affine* affine_slots = (G1Affine*)(ctx->dev_small + 4096);    // 30 x 64 B
    G1XYZZ* xyzz_slots = (G1XYZZ*)(ctx->dev_small + 8192);          // 30 x 128 B
    // groups of consecutive jobs over the same table
    struct Group { uint32_t first, count; };
    std::vector<Group> groups;
    for (uint32_t i = 0; i < count;) {
        uint32_t len = 1;
        if (ctx->tuning.merge && jobs[i].bases.table)
            while (i + len < count && len < MSM_MAX_JOBS && jobs[i + len].bases.table == jobs[i].bases.table &&
                   jobs[i + len].bases.stride == jobs[i].bases.stride && jobs[i + len].bases.tab_c == jobs[i].bases.tab_c)
                len++;
        groups.push_back({i, len});
        i += len;
    }
    cudaStream_t main_stream = ctx->stream;
    int r = KZG_OK;
    const bool two_lanes = groups.size() > 1;
    if (two_lanes) {
        // lane 1 may start once everything queued so far on the main stream (the scalars' producers) is done
        KZG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
        KZG_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    }
    for (size_t gi = 0; gi < groups.size() && r == KZG_OK; gi++) {
        const Group& gr = groups[gi];
        const int lane = two_lanes ? (int)(gi & 1) : 0;
        ctx->lane = lane;
        ctx->stream = lane ? ctx->aux_stream : main_stream;
        MsmJobs mj;
        memset(&mj, 0, sizeof(mj));
        mj.count = gr.count;
        for (uint32_t j = 0; j < gr.count; j++) {
            mj.scalars[j] = jobs[gr.first + j].src.scalars;
            mj.n[j] = jobs[gr.first + j].n;
            if (jobs[gr.first + j].src.montgomery) mj.montgomery |= 1u << j;
        }
        r = msm_run_multi(ctx, jobs[gr.first].bases, mj, xyzz_slots + gr.first, nullptr, 0, nullptr);
        if (r == KZG_OK) {
            KZG_LAUNCH(ctx, g1_finish_many_kernel, gr.count, 32, 0, xyzz_slots + gr.first, affine_slots + gr.first);
            if (cudaGetLastError() != cudaSuccess) r = set_err(ctx, KZG_ERR_CUDA, "g1_finish launch failed");
        }
    }
    ctx->lane = 0;
    ctx->stream = main_stream;
    if (two_lanes) {
        // join: the main stream continues only after lane 1 has drained (also on the error path)
        cudaEventRecord(ctx->ev_join, ctx->aux_stream);
        cudaStreamWaitEvent(main_stream, ctx->ev_join, 0);
    }
    KZG_TRY(r);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned + 1024, affine_slots, 64 * count, cudaMemcpyDeviceToHost, main_stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(main_stream));
    memcpy(out_affine, ctx->pinned + 1024, 64 * count);
    return KZG_OK;
}

// build (or rebuild) the window table of an SRS; c = 0 picks the window from the SRS size
int srs_precompute(kzg_ctx* ctx, kzg_srs* srs, uint32_t c) {
    if (srs->table) {
        KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(srs->table);
        srs->table = nullptr;
        srs->tab_c = srs->tab_nwin = 0;
    }
    if (srs->n == 0) return KZG_OK;
    if (c == 0) c = msm_table_window(srs->n);
    if (c < 2 || c > 23) return set_err(ctx, KZG_ERR_ARG, "srs table window must be in [2, 23]");
    const uint32_t nwin = windows_for(c, false);
    if ((uint64_t)nwin * srs->n >= (1ull << 31)) return set_err(ctx, KZG_ERR_ARG, "srs table would exceed 2^31 entries");
    G1Affine* table = nullptr;
    cudaError_t e = cudaMalloc((void**)&table, sizeof(G1Affine) * srs->n * nwin);
    if (e != cudaSuccess)
        return set_err(ctx, KZG_ERR_NOMEM, std::string("SRS window table allocation failed: ") + cudaGetErrorString(e));
    const uint64_t chunk = 1ull << 20;  // points per pass (bounds the XYZZ staging buffer)
    const uint64_t cmax = srs->n < chunk ? srs->n : chunk;
    G1XYZZ* tmp = nullptr;
    Fq* prefix = nullptr;
    if (nwin > 1) {
        e = cudaMalloc((void**)&tmp, sizeof(G1XYZZ) * cmax * (nwin - 1));
        if (e == cudaSuccess) e = cudaMalloc((void**)&prefix, sizeof(Fq) * cmax * (nwin - 1));
        if (e != cudaSuccess) {
            cudaFree(table);
            cudaFree(tmp);
            return set_err(ctx, KZG_ERR_NOMEM, std::string("SRS window table staging allocation failed: ") + cudaGetErrorString(e));
        }
    }
    for (uint64_t first = 0; first < srs->n; first += chunk) {
        const uint64_t count = srs->n - first < chunk ? srs->n - first : chunk;
        const uint32_t blocks = (uint32_t)((count + 127) / 128);
        if (nwin > 1) KZG_LAUNCH(ctx, srs_table_chain_kernel, blocks, 128, 0, srs->d, first, count, c, nwin, tmp);
        KZG_LAUNCH(ctx, srs_table_normalise_kernel, blocks, 128, 0, srs->d, first, count, srs->n, nwin, tmp, prefix, table);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    cudaFree(prefix);
    if (e != cudaSuccess) {
        cudaFree(table);
        return set_err(ctx, KZG_ERR_CUDA, cudaGetErrorString(e));
    }
    srs->table = table;
    srs->tab_c = c;
    srs->tab_nwin = nwin;
    return KZG_OK;
}

MsmBases srs_bases(kzg_ctx* ctx, const kzg_srs* srs, uint64_t first) {
    MsmBases b;
    b.pts = srs->d + first;
    b.table = nullptr;
    b.stride = 0;
    b.tab_c = b.tab_nwin = 0;
    // an explicit kzg_msm_set_window() asks for the classic per-window path (tuning / cross-checks)
    if (srs->table && ctx->msm_window == 0) {
        b.table = srs->table + first;
        b.stride = srs->n;
        b.tab_c = srs->tab_c;
        b.tab_nwin = srs->tab_nwin;
    }
    return b;
}

}  // namespace kzg

using namespace kzg;

// device slot for the XYZZ result of the MSM in flight (inside ctx->dev_small, past the 64-byte affine slot)
static inline G1XYZZ* result_slot(kzg_ctx* ctx) { return (G1XYZZ*)(ctx->dev_small + 1024); }

static MsmBases raw_bases(const G1Affine* pts) {
    MsmBases b;
    b.pts = pts;
    b.table = nullptr;
    b.stride = 0;
    b.tab_c = b.tab_nwin = 0;
    return b;
}

extern "C" {

int kzg_msm_geometry(kzg_ctx* ctx, kzg_srs* srs, uint64_t n, int montgomery, uint32_t* window_bits, uint32_t* windows) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_ERR_ARG;
    MsmBases b = srs ? srs_bases(ctx, srs, 0) : raw_bases(nullptr);
    MsmGeom g = msm_geometry(ctx, b, n ? n : 1, montgomery != 0);
    if (window_bits) *window_bits = g.c;
    if (windows) *windows = g.nwin;
    return KZG_OK;
}

int kzg_msm_plan(kzg_ctx* ctx, kzg_srs* srs, uint64_t n, int montgomery, uint32_t* window_bits, uint32_t* windows,
                 uint32_t* affine_rounds) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_ERR_ARG;
    MsmBases b = srs ? srs_bases(ctx, srs, 0) : raw_bases(nullptr);
    MsmGeom g = msm_geometry(ctx, b, n ? n : 1, montgomery != 0);
    if (window_bits) *window_bits = g.c;
    if (windows) *windows = g.nwin;
    if (affine_rounds) *affine_rounds = msm_affine_rounds(ctx, (n ? n : 1) * g.nwin, g.nsets * g.nbuckets);
    return KZG_OK;
}
int kzg_msm_set_window(kzg_ctx* ctx, uint32_t c) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx) return KZG_ERR_ARG;
    if (c != 0 && (c < 2 || c > 22)) return set_err(ctx, KZG_ERR_ARG, "msm window must be 0 (auto) or in [2, 22]");
    ctx->msm_window = c;
    return KZG_OK;
}

int kzg_srs_precompute(kzg_ctx* ctx, kzg_srs* srs, uint32_t window_bits) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs) return KZG_ERR_ARG;
    return srs_precompute(ctx, srs, window_bits);
}

// commit(pol): MSM length = min(len, |SRS|); coefficients beyond the SRS must be zero (the reference
// slices PTau to degree+1 points, polynomial.js:1107-1108 -- trailing zero coefficients never matter).
int kzg_commit(kzg_ctx* ctx, kzg_srs* srs, kzg_buf* coef, uint8_t out_affine[64]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || !coef || !out_affine) return KZG_ERR_ARG;
    uint64_t n = coef->n;
    if (n > srs->n) {
        uint64_t deg = 0;
        KZG_TRY(poly_degree(ctx, coef->d, coef->n, &deg));
        if (deg + 1 > srs->n)
            return set_err(ctx, KZG_ERR_PROTOCOL, "The Powers of Tau file is not sufficiently large to commit the polynomials.");
        n = srs->n;
    }
    MsmScalarSrc src{coef->d, true};
    KZG_TRY(msm_run_split(ctx, srs_bases(ctx, srs, 0), src, n, result_slot(ctx)));
    return msm_result_to_host_affine(ctx, result_slot(ctx), 1, out_affine);
}

int kzg_srs_msm(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, kzg_buf* scalars_std, uint64_t n, uint8_t out_affine[64]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || !scalars_std || !out_affine) return KZG_ERR_ARG;
    if (first + n > srs->n || n > scalars_std->n) return set_err(ctx, KZG_ERR_ARG, "msm: slice out of bounds");
    MsmScalarSrc src{scalars_std->d, false};
    KZG_TRY(msm_run_split(ctx, srs_bases(ctx, srs, first), src, n, result_slot(ctx)));
    return msm_result_to_host_affine(ctx, result_slot(ctx), 1, out_affine);
}

int kzg_srs_msm_partial(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, kzg_buf* scalars_std, uint64_t n, void* partial_dev) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || !scalars_std || !partial_dev) return KZG_ERR_ARG;
    if (first + n > srs->n || n > scalars_std->n) return set_err(ctx, KZG_ERR_ARG, "msm: slice out of bounds");
    MsmScalarSrc src{scalars_std->d, false};
    return msm_run_split(ctx, srs_bases(ctx, srs, first), src, n, (G1XYZZ*)partial_dev);
}

int kzg_g1_partials_combine(kzg_ctx* ctx, const void* partials_dev, uint32_t count, uint8_t out_affine[64]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !partials_dev || !out_affine || count == 0) return KZG_ERR_ARG;
    return msm_result_to_host_affine(ctx, (const G1XYZZ*)partials_dev, count, out_affine);
}

// scalars from host memory against a resident SRS (the e2e path of bench.py: H2D of the scalars is inside).
// Large inputs are cut into growing pieces (1/8, 2/8, 5/8 of the points): all the uploads are queued on the
// auxiliary stream (copy engine) at once, and the MSM of piece k on the main stream waits only for ITS upload, so that
// only the first, small upload is exposed and every later one hides behind the previous piece's MSM (each piece's MSM
// takes longer than the next piece's upload at PCIe 5 x16 rates).  Leaves `*parts_out` partial points in `slots`.
// (The overlap needs pinned host memory; with pageable memory the call is still correct, just serial.)
static int srs_msm_host_pieces(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, const void* scalars_std_host, uint64_t n,
                               G1XYZZ* slots, uint32_t* parts_out) {
    Fr* tmp = nullptr;
    if (n) KZG_CUDA(ctx, cudaMallocAsync((void**)&tmp, sizeof(Fr) * n, ctx->stream));
    const Fr* host = (const Fr*)scalars_std_host;
    int r = KZG_OK;
    uint32_t parts = 1;
    if (n >= (1ull << ctx->tuning.host_piece_min_log)) {
        // every piece pays the fixed part of an MSM again (~1.3 ms with 2^19 buckets, ~2.2 ms with the 2^21 buckets of a
        // c = 22 table): three pieces of 1/8, 2/8, 5/8 for the former, two of 3/16, 13/16 for the latter
        // Two pieces (3/16, 13/16) over a window table: they SHARE one bucket reduction (the first piece hands its folded
        // bucket sums to the second one's level 0, MsmReduceLink), so a piece costs its sort and accumulation only.
        // Without a table (per-window bucket sets whose geometry depends on the piece size) every piece reduces on its
        // own: three pieces of 1/8, 2/8, 5/8.
        const bool table_flavour = srs->table && ctx->msm_window == 0;
        const bool lanes = ctx->lane == 0 && !ctx->no_split;
        uint64_t cut[4] = {0, n / 8, n / 8 + n / 4, n};
        parts = 3;
        if (table_flavour) {
            parts = 2;
            cut[1] = n / 16 * 3;
            cut[2] = n;
            // With two pieces the step is (upload of everything) + (MSM of the last 13/16): the last piece cannot start
            // before the whole upload has arrived (10 ms for 512 MiB), and the GPU idles once the first piece is done.
            // Three linked pieces (3/64, 13/64, 48/64) keep it busy during the upload and leave a smaller rest: e2e 37.2 ->
            // 35.5 ms at 2^24 points, 20.7 -> 19.7 at 2^23, 10.75 -> 10.4 at 2^22.  What remains above the resident-scalar
            // MSM (32.8 ms) is the first upload (0.5 ms) and the pieces' own inefficiency: two more sorts, fewer affine
            // rounds in the small pieces.
            if (n >= (1ull << 22) && lanes && ctx->tuning.host_link) {
                parts = 3;
                cut[1] = n / 64 * 3;
                cut[2] = n / 4;
            }
        }
        if (ctx->tuning.host_cut_a > 0) {  // tuning: cuts at a/64 and b/64 of n (b = 64: two pieces)
            const int a = ctx->tuning.host_cut_a, b = ctx->tuning.host_cut_b;
            cut[1] = n / 64 * a;
            cut[2] = b < 64 ? n / 64 * b : n;
            cut[3] = n;
            parts = b < 64 ? 3 : 2;
        }
        // All uploads are queued at once on the copy stream.  Piece k runs on lane k & 1 (own stream, own scratch arena)
        // and waits only for its own upload, so the next piece starts sorting while the latency-bound tail of the
        // previous one (bucket reduction, the inversions of the affine rounds) is still running.
        cudaEvent_t up[3] = {order_event(ctx), order_event(ctx), order_event(ctx)};
        cudaStream_t main_stream = ctx->stream;
        cudaError_t e = cudaEventRecord(ctx->ev_fork, main_stream);  // tmp exists from here on
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0);
        for (uint32_t k = 0; k < parts && e == cudaSuccess; k++) {
            e = cudaMemcpyAsync(tmp + cut[k], host + cut[k], sizeof(Fr) * (cut[k + 1] - cut[k]), cudaMemcpyHostToDevice,
                                ctx->copy_stream);
            if (e == cudaSuccess) e = cudaEventRecord(up[k], ctx->copy_stream);
        }
        if (e != cudaSuccess) r = set_err(ctx, KZG_ERR_CUDA, std::string("msm upload: ") + cudaGetErrorString(e));
        // (measured, e2e ms linked / each piece reducing on its own: 6.05 / 6.57 at 2^21 points, 11.03 / 12.37 at 2^22,
        // 20.71 / 21.06 at 2^23 with two pieces; host_link = 0: never)
        bool link = lanes && table_flavour && ctx->tuning.host_link != 0;
        for (uint32_t k = 0; k + 1 < parts; k++) link = link && cut[k + 1] > cut[k] && cut[k + 1] < n;
        MsmReduceLink lk[2];
        for (uint32_t k = 0; k < parts && r == KZG_OK; k++) {
            const int lane = lanes ? (int)(k & 1) : 0;
            ctx->lane = lane;
            ctx->stream = lane ? ctx->aux_stream : main_stream;
            cudaStreamWaitEvent(ctx->stream, up[k], 0);
            const MsmBases pb = srs_bases(ctx, srs, first + cut[k]);
            const MsmScalarSrc ps{tmp + cut[k], false};
            if (link) {
                // every piece but the last stops after its fold; a later piece opens each bucket of its walk with the sum
                // the earlier ones left for it (MsmCarry) and the last one reduces.  The bucket sums of a piece live in
                // its scratch arena until then: of three pieces the first and the last share lane 0 (one stream), so the
                // first one works in the third arena.
                const bool last = k + 1 == parts;
                if (parts == 3 && k == 0) ctx->arena = 2;
                r = msm_run_linked(ctx, pb, ps, cut[k + 1] - cut[k], slots, lk, k, last ? nullptr : &lk[k]);
                ctx->arena = -1;
            } else {
                r = msm_run(ctx, pb, ps, cut[k + 1] - cut[k], slots + k);
            }
        }
        if (link) parts = 1;
        ctx->lane = 0;
        ctx->stream = main_stream;
        // (also on the error path: tmp is freed on the main stream, after every upload and every piece that was queued)
        cudaEventRecord(ctx->ev_join, ctx->copy_stream);
        cudaStreamWaitEvent(main_stream, ctx->ev_join, 0);
        cudaEventRecord(ctx->ev_join, ctx->aux_stream);
        cudaStreamWaitEvent(main_stream, ctx->ev_join, 0);
        if (r != KZG_OK) cudaStreamSynchronize(main_stream);
    } else {
        if (n) KZG_CUDA(ctx, cudaMemcpyAsync(tmp, host, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
        r = msm_run(ctx, srs_bases(ctx, srs, first), MsmScalarSrc{tmp, false}, n, slots);
    }
    if (tmp) cudaFreeAsync(tmp, ctx->stream);
    *parts_out = parts;
    return r;
}

int kzg_srs_msm_host(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, const void* scalars_std_host, uint64_t n,
                     uint8_t out_affine[64]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || (!scalars_std_host && n) || !out_affine) return KZG_ERR_ARG;
    if (first + n > srs->n) return set_err(ctx, KZG_ERR_ARG, "msm: slice out of bounds");
    G1XYZZ* slots = (G1XYZZ*)(ctx->dev_small + 8192);
    uint32_t parts = 1;
    KZG_TRY(srs_msm_host_pieces(ctx, srs, first, scalars_std_host, n, slots, &parts));
    return msm_result_to_host_affine(ctx, slots, parts, out_affine);
}

// the multi-GPU form of the same: the rank's shard of the scalars comes from HOST memory (piecewise upload hidden
// behind the pieces' MSMs) and the shard's sum is left as ONE XYZZ partial in device memory for the all-gather
int kzg_srs_msm_host_partial(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, const void* scalars_std_host, uint64_t n,
                             void* partial_dev) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || (!scalars_std_host && n) || !partial_dev) return KZG_ERR_ARG;
    if (first + n > srs->n) return set_err(ctx, KZG_ERR_ARG, "msm: slice out of bounds");
    G1XYZZ* slots = (G1XYZZ*)(ctx->dev_small + 8192);
    uint32_t parts = 1;
    KZG_TRY(srs_msm_host_pieces(ctx, srs, first, scalars_std_host, n, slots, &parts));
    KZG_LAUNCH(ctx, g1_sum_kernel, 1, 32, 0, slots, parts, (G1XYZZ*)partial_dev);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

// commit(pol) for several polynomials over one SRS in one pipeline (multi-MSM over shared bases; the
// prover's [F],[T] of round 1 and [W_xi],[W_xiw] of round 5, prover.js:161-162,409-410).  out_affine: 64 B per polynomial.
int kzg_commit_many(kzg_ctx* ctx, kzg_srs* srs, kzg_buf* const* coefs, uint32_t count, uint8_t* out_affine) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || !srs || (!coefs && count) || (!out_affine && count)) return KZG_ERR_ARG;
    std::vector<MsmJob> jobs(count);
    for (uint32_t i = 0; i < count; i++) {
        if (!coefs[i]) return KZG_ERR_ARG;
        uint64_t n = coefs[i]->n;
        if (n > srs->n) {
            uint64_t deg = 0;
            KZG_TRY(poly_degree(ctx, coefs[i]->d, coefs[i]->n, &deg));
            if (deg + 1 > srs->n)
                return set_err(ctx, KZG_ERR_PROTOCOL, "The Powers of Tau file is not sufficiently large to commit the polynomials.");
            n = srs->n;
        }
        jobs[i].bases = srs_bases(ctx, srs, 0);
        jobs[i].src = MsmScalarSrc{coefs[i]->d, true};
        jobs[i].n = n;
    }
    return msm_run_batch(ctx, jobs.data(), count, out_affine);
}

int kzg_g1_msm_affine(kzg_ctx* ctx, const void* bases, const void* scalars_std, uint64_t n, uint32_t flags,
                      uint8_t out_affine[64], uint8_t out_jacobian[96]) {
    kzg::DeviceGuard _dg(ctx);
    if (!ctx || (!bases && n) || (!scalars_std && n) || !out_affine) return KZG_ERR_ARG;
    const G1Affine* d_bases = (const G1Affine*)bases;
    const Fr* d_scalars = (const Fr*)scalars_std;
    G1Affine* tmp_bases = nullptr;
    Fr* tmp_scalars = nullptr;
    int r = KZG_OK;
    if (n && !(flags & KZG_BASES_ON_DEVICE)) {
        KZG_CUDA(ctx, cudaMallocAsync((void**)&tmp_bases, sizeof(G1Affine) * n, ctx->stream));
        KZG_CUDA(ctx, cudaMemcpyAsync(tmp_bases, bases, sizeof(G1Affine) * n, cudaMemcpyHostToDevice, ctx->stream));
        d_bases = tmp_bases;
    }
    if (n && !(flags & KZG_SCALARS_ON_DEVICE)) {
        KZG_CUDA(ctx, cudaMallocAsync((void**)&tmp_scalars, sizeof(Fr) * n, ctx->stream));
        KZG_CUDA(ctx, cudaMemcpyAsync(tmp_scalars, scalars_std, sizeof(Fr) * n, cudaMemcpyHostToDevice, ctx->stream));
        d_scalars = tmp_scalars;
    }
    MsmScalarSrc src{d_scalars, false};
    r = msm_run(ctx, raw_bases(d_bases), src, n, result_slot(ctx));
    if (r == KZG_OK) r = msm_result_to_host_affine(ctx, result_slot(ctx), 1, out_affine);
    if (tmp_bases) cudaFreeAsync(tmp_bases, ctx->stream);
    if (tmp_scalars) cudaFreeAsync(tmp_scalars, ctx->stream);
    if (r == KZG_OK && out_jacobian) {
        // G1.multiExpAffine returns a Jacobian triple (polynomial.js:1112); a triple is not canonical, so the
        // normalised representative (x, y, 1) -- the all-zero triple for infinity -- is returned.
        memcpy(out_jacobian, out_affine, 64);
        bool inf = true;
        for (int i = 0; i < 64; i++) inf &= out_affine[i] == 0;
        Fq one = inf ? fp_zero<FqP>() : fp_one<FqP>();
        memcpy(out_jacobian + 64, one.l, 32);
    }
    return r;
}

}  // extern "C"
