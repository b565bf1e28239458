// Correctness + throughput of fp_mul_k (Karatsuba + separated reduction) against field.cuh's fp_mul / fp_mul_portable.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 fieldk_test.cu -o fieldk_test
// Result on B200 (round 1): bit-exact, but 49 G modmul/s against 68 G for the CIOS product -- the 16 wide MACs saved
// are paid for with ~175 extra moves / selects / adds per product (382 instructions instead of 207), so the
// experiment stays out of the library.
#include <cstdio>
#include <cuda_runtime.h>
#include "field_k.cuh"
using namespace kzg;

__device__ uint64_t splitmix(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <class P> __device__ Fp<P> rnd(uint64_t& s, int mode) {
    Fp<P> r;
    for (int i = 0; i < 8; i += 2) {
        uint64_t z = splitmix(s);
        r.l[i] = (uint32_t)z;
        r.l[i + 1] = (uint32_t)(z >> 32);
    }
    r.l[7] &= 0x1fffffffu;
    if (mode == 1) for (int i = 0; i < 8; i++) r.l[i] = P::mod(i);           // p - 1
    if (mode == 1) r.l[0] -= 1;
    if (mode == 2) for (int i = 0; i < 8; i++) r.l[i] = (i < 4) ? 0xffffffffu : 0;   // low half all ones
    if (mode == 3) for (int i = 0; i < 8; i++) r.l[i] = (i >= 4 && i < 7) ? 0xffffffffu : (i == 7 ? 0x1fffffffu : 0);
    if (mode == 4) for (int i = 0; i < 8; i++) r.l[i] = 0;
    if (mode == 5) { for (int i = 0; i < 8; i++) r.l[i] = 0xffffffffu; r.l[7] = 0x1fffffffu; }
    return r;
}
template <class P> __device__ bool check(uint64_t& s, int ma, int mb) {
    Fp<P> a = rnd<P>(s, ma), b = rnd<P>(s, mb);
    return fp_eq(fp_mul_k(a, b), fp_mul_portable(a, b));
}
__global__ void test_kernel(unsigned int n, unsigned int* fail) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = 0xabcdef12345ull + 7919ull * i;
    bool ok = true;
    ok &= check<FqP>(s, 0, 0) && check<FrP>(s, 0, 0);
    if (i < 36) ok &= check<FqP>(s, i / 6, i % 6) && check<FrP>(s, i / 6, i % 6);
    if (!ok) atomicAdd(fail, 1u);
}
template <int WHICH> __global__ void __launch_bounds__(256) thr_kernel(uint32_t iters, uint32_t seed, uint32_t* sink) {
    Fq x = fp_one<FqP>(), y = fp_r2<FqP>();
    x.l[0] += (seed + threadIdx.x) & 0xffff;
    y.l[0] ^= (blockIdx.x * 7u + seed) & 0xffff;
    Fq u = y, w = x;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (WHICH == 0) { x = fp_mul(x, y); u = fp_mul(u, w); }
            else { x = fp_mul_k(x, y); u = fp_mul_k(u, w); }
        }
    }
    uint32_t z = 0;
    for (int k = 0; k < 8; k++) z ^= x.l[k] ^ u.l[k];
    if (z == 0x1234567u) *sink = z;
}
int main() {
    unsigned int* fail;
    cudaMalloc(&fail, 8);
    cudaMemset(fail, 0, 8);
    test_kernel<<<4096, 128>>>(4096 * 128, fail);
    unsigned int h = 123;
    cudaMemcpy(&h, fail, 4, cudaMemcpyDeviceToHost);
    printf("correctness: %u failures of %u cases (%s)\n", h, 4096 * 128 * 2, cudaGetErrorString(cudaGetLastError()));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    uint32_t* sink = (uint32_t*)fail + 1;
    for (int which = 0; which < 2; which++) {
        for (int occ = 0; occ < 2; occ++) {
            int blocks = prop.multiProcessorCount * (occ ? 8 : 2), threads = 256;
            uint32_t iters = 4096;
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(a);
                if (which == 0) thr_kernel<0><<<blocks, threads>>>(iters, rep, sink);
                else thr_kernel<1><<<blocks, threads>>>(iters, rep, sink);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (rep && ms < best) best = ms;
            }
            double muls = (double)blocks * threads * iters * 8;
            printf("%s, %d blocks/SM: %.3f ms, %.2f G modmul/s\n", which ? "fp_mul_k (karatsuba)" : "fp_mul   (cios)     ", occ ? 8 : 2, best,
                   muls / best / 1e6);
        }
    }
    return 0;
}
