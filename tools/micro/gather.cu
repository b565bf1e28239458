// How many DRAM bytes does a random 32-byte / 64-byte gather from a multi-GiB table cost on B200, and does the load
// flavour change it?  (The batched-affine round 1 of the MSM is bound by exactly this: profiles/r01_msm_affine.md.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather tools/micro/gather.cu && ./gather [GiB]
// Prints, per variant, the time of 2^27 gathers and the implied rate; run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum
// to see the bytes per gather.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { V_LDG = 0, V_CG, V_CS, V_LU, V_CV, V_NOALLOC, V_EVICT_FIRST, V_L2_64, V_L2_128, V_L2_256, V_V8, V_V8_EF, V_COUNT };
static const char* NAMES[V_COUNT] = {"ld.global.nc (__ldg)",      "ld.global.cg",          "ld.global.cs",
                                     "ld.global.lu",              "ld.global.cv",          "ld.global.nc.L1::no_allocate",
                                     "ld.global.L2::evict_first", "ld.global.nc.L2::64B",  "ld.global.nc.L2::128B",
                                     "ld.global.nc.L2::256B",     "ld.global.v8.b32 (256-bit)", "ld.global.L2::evict_first.v8.b32"};

template <int V>
__device__ __forceinline__ uint4 load16(const uint4* p) {
    uint4 r;
    if (V == V_LDG) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_CG) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_CS) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_LU) asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_CV) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_NOALLOC)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_EVICT_FIRST) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    }
    if (V == V_L2_64) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_L2_128) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == V_L2_256) asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// each thread: `per` gathers of BYTES bytes at pseudo-random 64-byte-aligned records (two independent ones in flight)
template <int V, int BYTES>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ table, uint64_t records, uint32_t per, uint32_t* out) {
    uint64_t s = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per; k++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const uint64_t rec = (s >> 20) % records;
        const uint4* p = table + rec * 4;  // 64-byte records
        if (V == V_V8 || V == V_V8_EF) {
#pragma unroll
            for (int j = 0; j < BYTES / 32; j++) {
                uint32_t a0, a1, a2, a3, a4, a5, a6, a7;
                if (V == V_V8)
                    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(p + 2 * j));
                else
                    asm volatile("ld.global.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(a4), "=r"(a5), "=r"(a6), "=r"(a7) : "l"(p + 2 * j));
                acc += a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
            }
        } else {
#pragma unroll
            for (int j = 0; j < BYTES / 16; j++) {
                uint4 v = load16<V>(p + j);
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// the access pattern of the dense affine rounds: every thread walks its OWN contiguous run of 64-byte records (lanes
// 32 records = 2 KB apart), so a warp-wide load touches 32 different lines although the kernel as a whole streams
__global__ void __launch_bounds__(256) strided_kernel(const uint4* __restrict__ table, uint64_t records, uint32_t per, int bytes,
                                                     uint32_t* out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per; k++) {
        const uint64_t rec = (t * per + k) % records;
        const uint4* p = table + rec * 4;
        for (int j = 0; j < bytes / 16; j++) {
            uint4 v = load16<V_LDG>(p + j);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// and the fully coalesced version of the same stream (a warp reads 512 consecutive bytes per instruction)
__global__ void __launch_bounds__(256) coalesced_kernel(const uint4* __restrict__ table, uint64_t records, uint32_t per, uint32_t* out) {
    const uint64_t base = ((uint64_t)blockIdx.x * blockDim.x) * per * 4;  // uint4 units: the block's records
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per * 4; k++) {
        uint4 v = load16<V_LDG>(table + (base + (uint64_t)k * blockDim.x + threadIdx.x) % (records * 4));
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
static void run_patterns(const uint4* table, uint64_t records, uint32_t* out) {
    const uint32_t per = 32, threads = 256;
    const uint64_t gathers = 1ull << 27;
    const uint32_t blocks = (uint32_t)(gathers / per / threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms = 0;
    for (int bytes = 32; bytes <= 64; bytes += 32) {
        strided_kernel<<<blocks, threads>>>(table, records, per, bytes, out);
        cudaEventRecord(a);
        strided_kernel<<<blocks, threads>>>(table, records, per, bytes, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("per-thread contiguous runs       %2d B of every 64-byte record: %7.3f ms  %6.2f G records/s\n", bytes, ms, gathers / ms / 1e6);
    }
    coalesced_kernel<<<blocks, threads>>>(table, records, per, out);
    cudaEventRecord(a);
    coalesced_kernel<<<blocks, threads>>>(table, records, per, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("warp-coalesced stream            64 B records: %7.3f ms  %6.2f G records/s  (%5.2f TB/s)\n", ms, gathers / ms / 1e6,
           gathers * 64.0 / ms / 1e9);
}

template <int V, int BYTES>
static void run(const uint4* table, uint64_t records, uint32_t* out) {
    const uint32_t per = 32, threads = 256;
    const uint64_t gathers = 1ull << 27;
    const uint32_t blocks = (uint32_t)(gathers / per / threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    gather_kernel<V, BYTES><<<blocks, threads>>>(table, records, per, out);
    cudaEventRecord(a);
    gather_kernel<V, BYTES><<<blocks, threads>>>(table, records, per, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    printf("%-32s %2d B: %7.3f ms  %6.2f G gathers/s  (%5.2f TB/s if 128 B each, %5.2f if 64, %5.2f if 32)%s\n", NAMES[V], BYTES, ms,
           gathers / ms / 1e6, gathers * 128.0 / ms / 1e9, gathers * 64.0 / ms / 1e9, gathers * 32.0 / ms / 1e9,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int V>
static void both(const uint4* table, uint64_t records, uint32_t* out) {
    run<V, 32>(table, records, out);
    run<V, 64>(table, records, out);
}

// `gather sweep`: the working-set curve -- is the ~44 G sectors/s of the 12 GiB table a TLB / page-locality limit (then it
// rises for small tables) or a sector-rate limit of the L2 / DRAM path (then it is flat once the table exceeds the L2)?
static int sweep() {
    const double sizes_gib[] = {0.0625, 0.125, 0.25, 0.5, 1, 2, 4, 8, 12, 16, 24, 32};
    uint32_t* out = nullptr;
    cudaMalloc(&out, 4);
    for (double g : sizes_gib) {
        const uint64_t bytes = (uint64_t)(g * (double)(1ull << 30)), records = bytes / 64;
        uint4* table = nullptr;
        if (cudaMalloc(&table, bytes) != cudaSuccess) {
            printf("allocation of %.3f GiB failed\n", g);
            return 1;
        }
        cudaMemset(table, 1, bytes);
        printf("== table %.4f GiB\n", g);
        both<V_LDG>(table, records, out);
        both<V_L2_64>(table, records, out);
        cudaFree(table);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "sweep")) return sweep();
    const uint64_t gib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 8;
    const uint64_t bytes = gib << 30, records = bytes / 64;
    uint4* table = nullptr;
    uint32_t* out = nullptr;
    if (cudaMalloc(&table, bytes) != cudaSuccess || cudaMalloc(&out, 4) != cudaSuccess) {
        printf("allocation failed\n");
        return 1;
    }
    cudaMemset(table, 1, bytes);
    size_t gran = 0;
    cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
    printf("table %llu GiB, cudaLimitMaxL2FetchGranularity = %zu\n", (unsigned long long)gib, gran);
    run_patterns(table, records, out);
    both<V_LDG>(table, records, out);
    both<V_CG>(table, records, out);
    both<V_CS>(table, records, out);
    both<V_LU>(table, records, out);
    both<V_CV>(table, records, out);
    both<V_NOALLOC>(table, records, out);
    both<V_EVICT_FIRST>(table, records, out);
    both<V_L2_64>(table, records, out);
    both<V_L2_128>(table, records, out);
    both<V_L2_256>(table, records, out);
    both<V_V8>(table, records, out);
    both<V_V8_EF>(table, records, out);
    if (argc > 2) {
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[2]));
        cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity = %zu\n", gran);
        both<V_LDG>(table, records, out);
    }
    return 0;
}
