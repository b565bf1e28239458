// Pipe throughput microbenchmarks for B200 (sm_100a): DFMA, IMAD.WIDE, IADD3, and co-issue of DFMA + IMAD.WIDE
// from different warps of the same SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipes.cu -o pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void dfma_body(double& a0, double& a1, double& a2, double& a3, double& a4, double& a5, double& a6,
                                          double& a7, double x, double y) {
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a0) : "d"(x), "d"(y));
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a1) : "d"(x), "d"(y));
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a2) : "d"(x), "d"(y));
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a3) : "d"(x), "d"(y));
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a4) : "d"(x), "d"(y));
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a5) : "d"(x), "d"(y));
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a6) : "d"(x), "d"(y));
    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(a7) : "d"(x), "d"(y));
}
__device__ __forceinline__ void imad_body(uint32_t* a, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                 "madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                 "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                 "madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.u32 %7, %11, %12, %7;\n\t"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                 : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
__device__ __forceinline__ void iadd_body(uint32_t* a, uint32_t x, uint32_t y) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("add.u32 %0, %0, %1;\n\t add.u32 %0, %0, %2;" : "+r"(a[k]) : "r"(x + k), "r"(y));
}

// mode: 0 = all warps DFMA; 1 = all warps IMAD.WIDE; 2 = even warps DFMA, odd warps IMAD.WIDE; 3 = IADD3;
//       4 = each warp interleaves DFMA and IMAD.WIDE; 5 = every warp: DFMA + IMAD + IADD interleaved
__global__ void __launch_bounds__(256) pipes(int mode, uint32_t iters, double* sink) {
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    double x = 1.0000001 + blockIdx.x * 1e-9, y = 0.9999999;
    uint32_t u[8], v[8], w[8];
    for (int k = 0; k < 8; k++) { u[k] = k + threadIdx.x; v[k] = 2 * k + 1; w[k] = k; }
    uint32_t yy = blockIdx.x * 77 + threadIdx.x;
    const int warp = threadIdx.x >> 5;
    bool do_f = mode == 0 || (mode == 2 && (warp & 1) == 0) || mode == 4 || mode == 5;
    bool do_i = mode == 1 || (mode == 2 && (warp & 1) == 1) || mode == 4 || mode == 5;
    bool do_a = mode == 3 || mode == 5;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (do_f) dfma_body(a0, a1, a2, a3, a4, a5, a6, a7, x, y);
            if (do_i) { imad_body(u, v[0], v[1], v[2], v[3], yy); imad_body(w, v[4], v[5], v[6], v[7], yy); yy += 0x9e3779b9u; }
            if (do_a) iadd_body(u, yy, v[3]);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    for (int k = 0; k < 8; k++) s += u[k] + w[k];
    if (s == 1.2345) *sink = s;
}

int main() {
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, clock %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    double* sink;
    CHECK(cudaMalloc(&sink, 8));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    const uint32_t iters = 20000;
    const char* names[] = {"DFMA only", "IMAD.WIDE only", "DFMA (even warps) + IMAD.WIDE (odd warps)", "IADD only",
                           "DFMA + IMAD.WIDE interleaved in every warp", "DFMA + IMAD.WIDE + IADD interleaved"};
    for (int mode = 0; mode < 6; mode++) {
        pipes<<<blocks, threads>>>(mode, 100, sink);
        CHECK(cudaDeviceSynchronize());
        cudaEventRecord(a);
        pipes<<<blocks, threads>>>(mode, iters, sink);
        cudaEventRecord(b);
        CHECK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        double thr = (double)blocks * threads * iters * 4;   // per-thread body repetitions
        double nf = 0, ni = 0, na = 0;
        if (mode == 0) nf = thr * 8;
        if (mode == 1) ni = thr * 8;
        if (mode == 2) { nf = thr * 8 / 2; ni = thr * 8 / 2; }
        if (mode == 3) na = thr * 16;
        if (mode == 4) { nf = thr * 8; ni = thr * 8; }
        if (mode == 5) { nf = thr * 8; ni = thr * 8; na = thr * 16; }
        double per_clk_sm = 1.0 / (ms * 1e-3) / prop.multiProcessorCount / (clk_khz * 1e3);
        printf("%-50s %8.3f ms  DFMA %6.1f /clk/SM  IMAD.WIDE %6.1f /clk/SM  IADD %6.1f /clk/SM\n", names[mode], ms,
               nf * per_clk_sm, ni * per_clk_sm, na * per_clk_sm);
    }
    return 0;
}
