// Signed-digit Pippenger G1 MSM for sm_100a.   Replaces G1.multiExpAffine + G1.toAffine
// (reference src/polynomial/polynomial.js:1106-1115, the un-vendored ffjavascript/wasmcurves call).
//
// Pipeline (all on one stream, no host round trip until the 64-byte result is read):
//   1. msm_count      : scalars -> signed c-bit digits (optionally leaving Montgomery form first,
//                       Fr.batchFromMontgomery fused), histogram of bucket keys (L2 atomics).
//   2. msm_scan       : exclusive prefix sums of bucket sizes and of per-bucket segment counts.
//   3. msm_scatter    : counting-sort scatter of (point index | sign) into bucket order.
//   4. msm_accumulate : one thread per bucket segment, XYZZ += affine over its slice of the sorted
//                       list (random 64 B gathers from the resident SRS, next point prefetched while the
//                       current one is being added).  This is the IMAD-bound kernel: 10 modmul per entry.
//   5. msm_reduce_1/2 : weighted bucket sums  sum_b (b+1) * S_b  per window: chunked running sums,
//                       shared-memory tree over the chunk partials.
//   6. msm_horner     : combine the windows (c doublings each) -> one XYZZ point in device memory.
//   7. g1_finish      : sum `count` partial points (count > 1 only for the multi-GPU gather) and convert
//                       to the canonical affine encoding.
// The result is a canonical group element, so it is byte-identical to the reference's regardless of
// window size, digit signedness or summation order.
#include <string.h>

#include "common.cuh"

namespace kzg {

// ---------------------------------------------------------------------------------------------
// digits
// ---------------------------------------------------------------------------------------------
struct MsmGeom {
    uint32_t c;        // window bits
    uint32_t nwin;     // number of windows
    uint32_t nbuckets; // buckets per window = 2^(c-1)
    uint32_t seg;      // max entries per accumulate task
};

__device__ __forceinline__ uint32_t scalar_bits(const uint32_t* s, uint32_t pos, uint32_t c) {
    // bits [pos, pos+c) of a 256-bit little-endian integer, zero beyond bit 255
    uint32_t word = pos >> 5, off = pos & 31;
    if (word >= 8) return 0;
    uint64_t v = s[word];
    if (word + 1 < 8) v |= (uint64_t)s[word + 1] << 32;
    return (uint32_t)(v >> off) & ((1u << c) - 1u);
}

template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_digits_kernel(const Fr* __restrict__ scalars, uint64_t n, bool montgomery,
                                                         MsmGeom g, uint32_t* __restrict__ counts_or_cursor,
                                                         uint32_t* __restrict__ sorted) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = fp_load<FrP>(scalars + i);
    if (montgomery) s = fp_from_mont(s);
    uint32_t carry = 0;
    const uint32_t half = g.nbuckets;  // 2^(c-1)
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
        if (mag != 0) {
            uint32_t key = w * g.nbuckets + (mag - 1);
            if (SCATTER) {
                uint32_t pos = atomicAdd(&counts_or_cursor[key], 1u);
                sorted[pos] = (uint32_t)i | (neg << 31);
            } else {
                atomicAdd(&counts_or_cursor[key], 1u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// scan of bucket sizes: offsets[k] = sum_{j<k} counts[j], segoff[k] = sum_{j<k} ceil(counts[j]/seg)
// Three phases over tiles of SCAN_TILE keys: tile sums, one-block scan of the tile sums, apply.
// Buckets that need more than one accumulate task are appended to `heavy` (collapsed after accumulation).
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ uint2 block_scan_pair(uint2 v, uint2* sh, uint2& total) {
    // inclusive scan of (a, b) pairs across the block
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        uint32_t oa = __shfl_up_sync(0xffffffffu, v.x, d), ob = __shfl_up_sync(0xffffffffu, v.y, d);
        if (lane >= d) {
            v.x += oa;
            v.y += ob;
        }
    }
    if (lane == 31) sh[wid] = v;
    __syncthreads();
    uint2 carry = make_uint2(0, 0), tot = make_uint2(0, 0);
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        uint2 x = sh[w];
        if (w < (int)wid) {
            carry.x += x.x;
            carry.y += x.y;
        }
        tot.x += x.x;
        tot.y += x.y;
    }
    __syncthreads();
    total = tot;
    return make_uint2(v.x + carry.x, v.y + carry.y);
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan_tiles_kernel(const uint32_t* __restrict__ counts, uint32_t nkeys,
                                                                      uint32_t seg, uint2* __restrict__ tile_sums) {
    __shared__ uint2 sh[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x;
    uint2 acc = make_uint2(0, 0);
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        uint32_t i = base + k * SCAN_THREADS;
        if (i < nkeys) {
            uint32_t cnt = counts[i];
            acc.x += cnt;
            acc.y += (cnt + seg - 1) / seg;
        }
    }
    uint2 tot;
    block_scan_pair(acc, sh, tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of the tile sums in place; totals -> offsets[nkeys], segoff[nkeys]
__global__ void __launch_bounds__(SCAN_THREADS) msm_scan_sums_kernel(uint2* __restrict__ tile_sums, uint32_t ntiles,
                                                                     uint32_t nkeys, uint32_t* __restrict__ offsets,
                                                                     uint32_t* __restrict__ segoff) {
    __shared__ uint2 sh[SCAN_THREADS / 32];
    __shared__ uint2 carry_sh;
    if (threadIdx.x == 0) carry_sh = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t start = 0; start < ntiles; start += SCAN_THREADS) {
        uint32_t i = start + threadIdx.x;
        uint2 v = i < ntiles ? tile_sums[i] : make_uint2(0, 0);
        uint2 tot;
        uint2 inc = block_scan_pair(v, sh, tot);
        uint2 c = carry_sh;
        if (i < ntiles) tile_sums[i] = make_uint2(c.x + inc.x - v.x, c.y + inc.y - v.y);
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = make_uint2(c.x + tot.x, c.y + tot.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        offsets[nkeys] = carry_sh.x;
        segoff[nkeys] = carry_sh.y;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_scan_apply_kernel(const uint32_t* __restrict__ counts, uint32_t nkeys,
                                                                      uint32_t seg, const uint2* __restrict__ tile_sums,
                                                                      uint32_t* __restrict__ offsets,
                                                                      uint32_t* __restrict__ cursor,
                                                                      uint32_t* __restrict__ segoff,
                                                                      uint32_t* __restrict__ heavy,
                                                                      uint32_t* __restrict__ heavy_count) {
    __shared__ uint2 sh[SCAN_THREADS / 32];
    // thread owns SCAN_PER_THREAD consecutive keys so that its partial results are a running sum
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t cnt[SCAN_PER_THREAD];
    uint2 acc = make_uint2(0, 0);
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        uint32_t i = base + k;
        cnt[k] = i < nkeys ? counts[i] : 0;
        acc.x += cnt[k];
        acc.y += (cnt[k] + seg - 1) / seg;
    }
    uint2 tot;
    uint2 inc = block_scan_pair(acc, sh, tot);
    uint2 t0 = tile_sums[blockIdx.x];
    uint32_t ra = t0.x + inc.x - acc.x, rb = t0.y + inc.y - acc.y;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        uint32_t i = base + k;
        if (i < nkeys) {
            offsets[i] = ra;
            cursor[i] = ra;
            segoff[i] = rb;
            uint32_t nseg = (cnt[k] + seg - 1) / seg;
            if (nseg > 1) heavy[atomicAdd(heavy_count, 1u)] = i;
            ra += cnt[k];
            rb += nseg;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ G1Affine load_affine(const G1Affine* p) {
    G1Affine r;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    r.x.l[0] = a.x; r.x.l[1] = a.y; r.x.l[2] = a.z; r.x.l[3] = a.w;
    r.x.l[4] = b.x; r.x.l[5] = b.y; r.x.l[6] = b.z; r.x.l[7] = b.w;
    r.y.l[0] = c.x; r.y.l[1] = c.y; r.y.l[2] = c.z; r.y.l[3] = c.w;
    r.y.l[4] = d.x; r.y.l[5] = d.y; r.y.l[6] = d.z; r.y.l[7] = d.w;
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

__global__ void __launch_bounds__(128) msm_accumulate_kernel(const G1Affine* __restrict__ bases,
                                                             const uint32_t* __restrict__ sorted,
                                                             const uint32_t* __restrict__ offsets,
                                                             const uint32_t* __restrict__ segoff, uint32_t nkeys,
                                                             uint32_t seg, G1XYZZ* __restrict__ partials) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ntasks = segoff[nkeys];
    if (t >= ntasks) return;
    // largest key with segoff[key] <= t  (segoff is non-decreasing; empty buckets repeat a value)
    uint32_t lo = 0, hi = nkeys;  // invariant: segoff[lo] <= t < segoff[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (segoff[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t key = lo;
    const uint32_t s = t - segoff[key];
    uint32_t begin = offsets[key] + s * seg;
    uint32_t end = min(begin + seg, offsets[key + 1]);

    G1XYZZ acc = xyzz_inf();
    uint32_t e = sorted[begin];
    G1Affine p = load_affine(bases + (e & 0x7fffffffu));
    for (uint32_t j = begin; j < end; j++) {
        // prefetch the next entry while this one is being added
        uint32_t e_next = e;
        G1Affine p_next = p;
        if (j + 1 < end) {
            e_next = sorted[j + 1];
            p_next = load_affine(bases + (e_next & 0x7fffffffu));
        }
        if (e >> 31) p.y = fp_neg(p.y);
        xyzz_madd(acc, p);
        e = e_next;
        p = p_next;
    }
    store_xyzz(partials + t, acc);
}

// After accumulation a bucket that was split over several tasks holds several partial sums: one block per
// such bucket tree-sums them in shared memory and leaves the total in the bucket's first slot.
constexpr int COLLAPSE_THREADS = 128;
__global__ void __launch_bounds__(COLLAPSE_THREADS) msm_collapse_kernel(G1XYZZ* __restrict__ partials,
                                                                        const uint32_t* __restrict__ segoff,
                                                                        const uint32_t* __restrict__ heavy,
                                                                        const uint32_t* __restrict__ heavy_count) {
    __shared__ G1XYZZ sh[COLLAPSE_THREADS];
    const uint32_t nheavy = *heavy_count;
    for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const uint32_t key = heavy[h];
        const uint32_t a = segoff[key], b = segoff[key + 1];
        G1XYZZ v = xyzz_inf();
        for (uint32_t j = a + threadIdx.x; j < b; j += COLLAPSE_THREADS) {
            G1XYZZ o = load_xyzz(partials + j);
            xyzz_add(v, o);
        }
        store_xyzz(sh + threadIdx.x, v);
        __syncthreads();
        for (uint32_t s = COLLAPSE_THREADS / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                G1XYZZ x = load_xyzz(sh + threadIdx.x);
                G1XYZZ y = load_xyzz(sh + threadIdx.x + s);
                xyzz_add(x, y);
                store_xyzz(sh + threadIdx.x, x);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) store_xyzz(partials + a, load_xyzz(sh));
        __syncthreads();
    }
}

// the (collapsed) sum of one bucket
__device__ __forceinline__ G1XYZZ load_bucket(const G1XYZZ* partials, const uint32_t* segoff, uint32_t key) {
    uint32_t a = segoff[key], b = segoff[key + 1];
    if (a == b) return xyzz_inf();
    return load_xyzz(partials + a);
}

// ---------------------------------------------------------------------------------------------
// bucket reduction: per window  sum_{b<B} (b+1) * S_b
// ---------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 128;

__device__ __forceinline__ void block_tree_sum(G1XYZZ& v, G1XYZZ* sh) {
    const uint32_t tid = threadIdx.x;
    store_xyzz(sh + tid, v);
    __syncthreads();
    for (uint32_t s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
            G1XYZZ a = load_xyzz(sh + tid);
            G1XYZZ b = load_xyzz(sh + tid + s);
            xyzz_add(a, b);
            store_xyzz(sh + tid, a);
        }
        __syncthreads();
    }
    v = load_xyzz(sh);
}

// grid = (blocks_per_window, nwin); thread -> chunk of `chunk` consecutive buckets
__global__ void __launch_bounds__(RED_THREADS) msm_reduce1_kernel(const G1XYZZ* __restrict__ partials,
                                                                  const uint32_t* __restrict__ segoff, MsmGeom g,
                                                                  uint32_t chunk, G1XYZZ* __restrict__ out) {
    __shared__ G1XYZZ sh[RED_THREADS];
    const uint32_t w = blockIdx.y;
    const uint32_t ci = blockIdx.x * RED_THREADS + threadIdx.x;
    const uint32_t nch = (g.nbuckets + chunk - 1) / chunk;
    G1XYZZ total = xyzz_inf();
    if (ci < nch) {
        const uint32_t lo = ci * chunk;
        const uint32_t hi = min(lo + chunk, g.nbuckets);
        G1XYZZ run = xyzz_inf();
        for (uint32_t b = hi; b-- > lo;) {
            G1XYZZ s = load_bucket(partials, segoff, w * g.nbuckets + b);
            xyzz_add(run, s);
            xyzz_add(total, run);
        }
        // total = sum (b - lo + 1) S_b ; add lo * run
        if (lo != 0) {
            G1XYZZ scaled = xyzz_mul_small(run, lo);
            xyzz_add(total, scaled);
        }
    }
    block_tree_sum(total, sh);
    if (threadIdx.x == 0) store_xyzz(out + (size_t)w * gridDim.x + blockIdx.x, total);
}

// grid = nwin; sums the per-block partials of a window
__global__ void __launch_bounds__(RED_THREADS) msm_reduce2_kernel(const G1XYZZ* __restrict__ in, uint32_t per_window,
                                                                  G1XYZZ* __restrict__ windows) {
    __shared__ G1XYZZ sh[RED_THREADS];
    const uint32_t w = blockIdx.x;
    G1XYZZ total = xyzz_inf();
    for (uint32_t j = threadIdx.x; j < per_window; j += RED_THREADS) {
        G1XYZZ o = load_xyzz(in + (size_t)w * per_window + j);
        xyzz_add(total, o);
    }
    block_tree_sum(total, sh);
    if (threadIdx.x == 0) store_xyzz(windows + w, total);
}

// result = sum_w 2^(c*w) * windows[w]
__global__ void msm_horner_kernel(const G1XYZZ* __restrict__ windows, MsmGeom g, G1XYZZ* __restrict__ result) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1XYZZ acc = load_xyzz(windows + (g.nwin - 1));
    for (int w = (int)g.nwin - 2; w >= 0; w--) {
        for (uint32_t k = 0; k < g.c; k++) acc = xyzz_dbl(acc);
        G1XYZZ o = load_xyzz(windows + w);
        xyzz_add(acc, o);
    }
    store_xyzz(result, acc);
}

// sum `count` XYZZ points and write the canonical affine encoding
__global__ void g1_finish_kernel(const G1XYZZ* __restrict__ parts, uint32_t count, G1Affine* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1XYZZ acc = xyzz_inf();
    for (uint32_t i = 0; i < count; i++) {
        G1XYZZ o = load_xyzz(parts + i);
        xyzz_add(acc, o);
    }
    G1Affine a = xyzz_to_affine(acc);
    fp_store(&out->x, a.x);
    fp_store(&out->y, a.y);
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static uint32_t auto_window(uint64_t n) {
    uint32_t lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    int c = (int)lg - 4;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    return (uint32_t)c;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static MsmGeom msm_geometry(kzg_ctx* ctx, uint64_t n, bool montgomery) {
    MsmGeom g;
    g.c = ctx->msm_window ? ctx->msm_window : auto_window(n);
    if (g.c < 2) g.c = 2;
    if (g.c > 22) g.c = 22;
    // Montgomery sources are reduced (< r < 2^254): ceil(255/c) windows leave the top digit carry-free.
    // Raw standard-form scalars may use all 256 bits: ceil(257/c).
    const uint32_t bits = montgomery ? 255 : 257;
    g.nwin = (bits + g.c - 1) / g.c;
    g.nbuckets = 1u << (g.c - 1);
    g.seg = 0;
    return g;
}

int msm_run(kzg_ctx* ctx, const G1Affine* bases, MsmScalarSrc src, uint64_t n, G1XYZZ* result_dev) {
    if (n == 0) {
        KZG_CUDA(ctx, cudaMemsetAsync(result_dev, 0, sizeof(G1XYZZ), ctx->stream));
        return KZG_OK;
    }
    if (n >= (1ull << 27)) return set_err(ctx, KZG_ERR_ARG, "msm: at most 2^27 - 1 points per call");
    MsmGeom g = msm_geometry(ctx, n, src.montgomery);
    const uint64_t avg = n / g.nbuckets + 1;
    uint64_t seg = 4 * avg;
    if (seg < 256) seg = 256;
    g.seg = (uint32_t)seg;
    const uint32_t nkeys = g.nwin * g.nbuckets;
    const uint64_t max_entries = n * g.nwin;
    const uint64_t max_tasks = (uint64_t)nkeys + max_entries / g.seg + 1;
    if (max_entries >= (1ull << 32)) return set_err(ctx, KZG_ERR_ARG, "msm: n * windows exceeds 2^32 entries");

    const uint32_t red_chunk = g.nbuckets >= 2048 ? 16 : (g.nbuckets >= 128 ? 4 : 1);
    const uint32_t nch = (g.nbuckets + red_chunk - 1) / red_chunk;
    const uint32_t red_blocks = (nch + RED_THREADS - 1) / RED_THREADS;

    // scratch layout
    size_t off = 0;
    const size_t o_counts = off;   off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_offsets = off;  off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_cursor = off;   off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_segoff = off;   off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    const size_t o_sorted = off;   off = align_up(off + sizeof(uint32_t) * max_entries, 256);
    const size_t o_partials = off; off = align_up(off + sizeof(G1XYZZ) * max_tasks, 256);
    const size_t o_red = off;      off = align_up(off + sizeof(G1XYZZ) * (size_t)g.nwin * red_blocks, 256);
    const size_t o_windows = off;  off = align_up(off + sizeof(G1XYZZ) * g.nwin, 256);
    const uint32_t ntiles = (nkeys + SCAN_TILE - 1) / SCAN_TILE;
    const size_t o_tiles = off;    off = align_up(off + sizeof(uint2) * ntiles, 256);
    const size_t o_heavy = off;    off = align_up(off + sizeof(uint32_t) * (nkeys + 1), 256);
    void* base = nullptr;
    KZG_TRY(ctx_scratch(ctx, off, &base));
    uint8_t* sc = (uint8_t*)base;
    uint32_t* counts = (uint32_t*)(sc + o_counts);
    uint32_t* offsets = (uint32_t*)(sc + o_offsets);
    uint32_t* cursor = (uint32_t*)(sc + o_cursor);
    uint32_t* segoff = (uint32_t*)(sc + o_segoff);
    uint32_t* sorted = (uint32_t*)(sc + o_sorted);
    G1XYZZ* partials = (G1XYZZ*)(sc + o_partials);
    G1XYZZ* red = (G1XYZZ*)(sc + o_red);
    G1XYZZ* windows = (G1XYZZ*)(sc + o_windows);
    uint2* tile_sums = (uint2*)(sc + o_tiles);
    uint32_t* heavy = (uint32_t*)(sc + o_heavy);       // [0] = count, [1..] = keys

    KZG_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (nkeys + 1), ctx->stream));
    KZG_CUDA(ctx, cudaMemsetAsync(heavy, 0, sizeof(uint32_t), ctx->stream));
    const uint32_t dblocks = (uint32_t)((n + 255) / 256);
    KZG_LAUNCH(ctx, msm_digits_kernel<false>, dblocks, 256, 0, src.scalars, n, src.montgomery, g, counts, nullptr);
    KZG_LAUNCH(ctx, msm_scan_tiles_kernel, ntiles, SCAN_THREADS, 0, counts, nkeys, g.seg, tile_sums);
    KZG_LAUNCH(ctx, msm_scan_sums_kernel, 1, SCAN_THREADS, 0, tile_sums, ntiles, nkeys, offsets, segoff);
    KZG_LAUNCH(ctx, msm_scan_apply_kernel, ntiles, SCAN_THREADS, 0, counts, nkeys, g.seg, tile_sums, offsets, cursor, segoff,
               heavy + 1, heavy);
    KZG_LAUNCH(ctx, msm_digits_kernel<true>, dblocks, 256, 0, src.scalars, n, src.montgomery, g, cursor, sorted);
    const uint32_t ablocks = (uint32_t)((max_tasks + 127) / 128);
    timed_begin(ctx, KZG_TIMED_MSM_ACCUMULATE);
    KZG_LAUNCH(ctx, msm_accumulate_kernel, ablocks, 128, 0, bases, sorted, offsets, segoff, nkeys, g.seg, partials);
    timed_end(ctx, KZG_TIMED_MSM_ACCUMULATE);
    KZG_LAUNCH(ctx, msm_collapse_kernel, (uint32_t)ctx->sm_count * 2, COLLAPSE_THREADS, 0, partials, segoff, heavy + 1, heavy);
    KZG_LAUNCH(ctx, msm_reduce1_kernel, dim3(red_blocks, g.nwin), RED_THREADS, 0, partials, segoff, g, red_chunk, red);
    KZG_LAUNCH(ctx, msm_reduce2_kernel, g.nwin, RED_THREADS, 0, red, red_blocks, windows);
    KZG_LAUNCH(ctx, msm_horner_kernel, 1, 32, 0, windows, g, result_dev);
    KZG_CHECK_LAUNCH(ctx);
    return KZG_OK;
}

int msm_result_to_host_affine(kzg_ctx* ctx, const G1XYZZ* result_dev, uint32_t count, uint8_t out[64]) {
    static_assert(sizeof(G1Affine) == 64, "affine layout");
    G1Affine* slot = (G1Affine*)ctx->dev_small;
    KZG_LAUNCH(ctx, g1_finish_kernel, 1, 32, 0, result_dev, count, slot);
    KZG_CHECK_LAUNCH(ctx);
    KZG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, slot, 64, cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->pinned, 64);
    return KZG_OK;
}

}  // namespace kzg

using namespace kzg;

// device slot for the XYZZ result of the MSM in flight (inside ctx->dev_small, past the 64-byte affine slot)
static inline G1XYZZ* result_slot(kzg_ctx* ctx) { return (G1XYZZ*)(ctx->dev_small + 1024); }

extern "C" {

int kzg_msm_geometry(kzg_ctx* ctx, uint64_t n, int montgomery, uint32_t* window_bits, uint32_t* windows) {
    if (!ctx) return KZG_ERR_ARG;
    MsmGeom g = msm_geometry(ctx, n ? n : 1, montgomery != 0);
    if (window_bits) *window_bits = g.c;
    if (windows) *windows = g.nwin;
    return KZG_OK;
}

int kzg_msm_set_window(kzg_ctx* ctx, uint32_t c) {
    if (!ctx) return KZG_ERR_ARG;
    if (c != 0 && (c < 2 || c > 22)) return set_err(ctx, KZG_ERR_ARG, "msm window must be 0 (auto) or in [2, 22]");
    ctx->msm_window = c;
    return KZG_OK;
}

// commit(pol): MSM length = min(len, |SRS|); coefficients beyond the SRS must be zero (the reference
// slices PTau to degree+1 points, polynomial.js:1107-1108 -- trailing zero coefficients never matter).
int kzg_commit(kzg_ctx* ctx, kzg_srs* srs, kzg_buf* coef, uint8_t out_affine[64]) {
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
    KZG_TRY(msm_run(ctx, srs->d, src, n, result_slot(ctx)));
    return msm_result_to_host_affine(ctx, result_slot(ctx), 1, out_affine);
}

int kzg_srs_msm(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, kzg_buf* scalars_std, uint64_t n, uint8_t out_affine[64]) {
    if (!ctx || !srs || !scalars_std || !out_affine) return KZG_ERR_ARG;
    if (first + n > srs->n || n > scalars_std->n) return set_err(ctx, KZG_ERR_ARG, "msm: slice out of bounds");
    MsmScalarSrc src{scalars_std->d, false};
    KZG_TRY(msm_run(ctx, srs->d + first, src, n, result_slot(ctx)));
    return msm_result_to_host_affine(ctx, result_slot(ctx), 1, out_affine);
}

int kzg_srs_msm_partial(kzg_ctx* ctx, kzg_srs* srs, uint64_t first, kzg_buf* scalars_std, uint64_t n, void* partial_dev) {
    if (!ctx || !srs || !scalars_std || !partial_dev) return KZG_ERR_ARG;
    if (first + n > srs->n || n > scalars_std->n) return set_err(ctx, KZG_ERR_ARG, "msm: slice out of bounds");
    MsmScalarSrc src{scalars_std->d, false};
    return msm_run(ctx, srs->d + first, src, n, (G1XYZZ*)partial_dev);
}

int kzg_g1_partials_combine(kzg_ctx* ctx, const void* partials_dev, uint32_t count, uint8_t out_affine[64]) {
    if (!ctx || !partials_dev || !out_affine || count == 0) return KZG_ERR_ARG;
    return msm_result_to_host_affine(ctx, (const G1XYZZ*)partials_dev, count, out_affine);
}

int kzg_g1_msm_affine(kzg_ctx* ctx, const void* bases, const void* scalars_std, uint64_t n, uint32_t flags,
                      uint8_t out_affine[64], uint8_t out_jacobian[96]) {
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
    r = msm_run(ctx, d_bases, src, n, result_slot(ctx));
    if (r == KZG_OK) r = msm_result_to_host_affine(ctx, result_slot(ctx), 1, out_affine);
    if (tmp_bases) cudaFreeAsync(tmp_bases, ctx->stream);
    if (tmp_scalars) cudaFreeAsync(tmp_scalars, ctx->stream);
    if (r == KZG_OK && out_jacobian) {
        // G1.multiExpAffine returns a Jacobian triple (polynomial.js:1112); a triple is not canonical, so the
        // normalised representative (x, y, 1) -- or (0, 0, 0)... the all-zero triple for infinity -- is returned.
        memcpy(out_jacobian, out_affine, 64);
        bool inf = true;
        for (int i = 0; i < 64; i++) inf &= out_affine[i] == 0;
        Fq one = inf ? fp_zero<FqP>() : fp_one<FqP>();
        memcpy(out_jacobian + 64, one.l, 32);
    }
    return r;
}

}  // extern "C"
