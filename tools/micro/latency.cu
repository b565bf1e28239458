// Latency of the field / group primitives for a LONE warp (the regime of the MSM's bucket-reduction tail):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../kzg_grandsums_study_b200/csrc latency.cu -o latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ec.cuh"

using namespace kzg;

template <int CH> __global__ void mul_chain(uint32_t iters, uint32_t* sink, long long* cyc) {
    Fq x[CH], y = fp_r2<FqP>();
    for (int c = 0; c < CH; c++) { x[c] = fp_one<FqP>(); x[c].l[0] += threadIdx.x + c; }
    y.l[0] ^= threadIdx.x;
    long long t0 = clock64();
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = fp_mul(x[c], y);
    }
    long long t1 = clock64();
    uint32_t z = 0;
    for (int c = 0; c < CH; c++) for (int k = 0; k < 8; k++) z ^= x[c].l[k];
    if (z == 0x1234567u) *sink = z;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void add_chain(const G1XYZZ* p, uint32_t iters, G1XYZZ* out, long long* cyc) {
    G1XYZZ acc = p[0], b = p[1];
    long long t0 = clock64();
    for (uint32_t i = 0; i < iters; i++) xyzz_add(acc, b);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dbl_chain(const G1XYZZ* p, uint32_t iters, G1XYZZ* out, long long* cyc) {
    G1XYZZ acc = p[0];
    long long t0 = clock64();
    for (uint32_t i = 0; i < iters; i++) acc = xyzz_dbl(acc);
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void qadd_chain(const G1XYZZ* p, uint32_t iters, G1XYZZ* out, long long* cyc) {
    const uint32_t j = threadIdx.x & 3, qm = quad_mask();
    Fq a = quad_coord(p[0], j);
    const Fq b = quad_coord(p[1], j);
    long long t0 = clock64();
    for (uint32_t i = 0; i < iters; i++) quad_add(a, b, j, qm);
    long long t1 = clock64();
    quad_store(out + (blockIdx.x * blockDim.x + threadIdx.x) / 4, j, a);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void qdbl_chain(const G1XYZZ* p, uint32_t iters, G1XYZZ* out, long long* cyc) {
    const uint32_t j = threadIdx.x & 3, qm = quad_mask();
    Fq a = quad_coord(p[0], j);
    long long t0 = clock64();
    for (uint32_t i = 0; i < iters; i++) quad_dbl(a, j, qm);
    long long t1 = clock64();
    quad_store(out + (blockIdx.x * blockDim.x + threadIdx.x) / 4, j, a);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void inv_chain(uint32_t iters, Fq* out, long long* cyc) {
    Fq x = fp_r2<FqP>();
    x.l[0] += threadIdx.x;
    long long t0 = clock64();
    for (uint32_t i = 0; i < iters; i++) { x = fp_inv(x); x.l[0] ^= 1; }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    uint32_t* sink; long long* cyc; G1XYZZ *pts, *out; Fq* fo;
    cudaMalloc(&sink, 4); cudaMallocManaged(&cyc, 8); cudaMallocManaged(&pts, 2 * sizeof(G1XYZZ)); cudaMalloc(&out, 64 << 20);
    cudaMalloc(&fo, 4096);
    // two distinct points: G = (1, 2) and 2G, in Montgomery form, ZZ = ZZZ = 1
    G1Affine g; g.x = fp_one<FqP>(); g.y = fp_dbl(fp_one<FqP>());
    pts[0] = xyzz_dbl_affine(g);
    pts[1] = xyzz_from_affine(g);
    const uint32_t it = 2000;
    const int grids[3] = {1, 148, 592};
    for (int gi = 0; gi < 3; gi++) {
        int grid = grids[gi];
        for (int warps = 1; warps <= 4; warps *= 2) {
            int th = 32 * warps;
            mul_chain<1><<<grid, th>>>(it, sink, cyc); cudaDeviceSynchronize(); double m1 = (double)*cyc / it;
            mul_chain<2><<<grid, th>>>(it, sink, cyc); cudaDeviceSynchronize(); double m2 = (double)*cyc / it / 2;
            mul_chain<4><<<grid, th>>>(it, sink, cyc); cudaDeviceSynchronize(); double m4 = (double)*cyc / it / 4;
            add_chain<<<grid, th>>>(pts, it, out, cyc); cudaDeviceSynchronize(); double a = (double)*cyc / it;
            dbl_chain<<<grid, th>>>(pts, it, out, cyc); cudaDeviceSynchronize(); double d = (double)*cyc / it;
            inv_chain<<<1, 32>>>(20, fo, cyc); cudaDeviceSynchronize(); double iv = (double)*cyc / 20;
            qadd_chain<<<grid, th>>>(pts, it, out, cyc); cudaDeviceSynchronize(); double qa = (double)*cyc / it;
            qdbl_chain<<<grid, th>>>(pts, it, out, cyc); cudaDeviceSynchronize(); double qd = (double)*cyc / it;
            printf("grid %3d x %d warps/block: modmul cycles 1 chain %.0f | 2 chains %.0f | 4 chains %.0f per product; xyzz_add %.0f; xyzz_dbl %.0f; quad_add %.0f; quad_dbl %.0f; fp_inv (32 divergent lanes) %.0f cycles\n",
                   grid, warps, m1, m2, m4, a, d, qa, qd, iv);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
