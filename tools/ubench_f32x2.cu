// Microbenchmark: do packed f32x2 instructions (FADD2/FMUL2/FFMA2, sm_100+) retire two fp32
// operations per issue slot?  Prints Gop/s (one "op" = one scalar add/mul/fma).
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void k_scalar_fma(float* o, float a, float b) {
    float x[8]; for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i]; o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_scalar_add(float* o, float a, float b) {
    float x[8]; for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = x[i] + a;
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i]; o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__global__ void k_packed_fma(float* o, float a, float b) {
    unsigned long long x[8], A = pk(a, a), B = pk(b, b);
    for (int i = 0; i < 8; ++i) x[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x[i]) : "l"(x[i]), "l"(A), "l"(B));
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_add(float* o, float a, float b) {
    unsigned long long x[8], A = pk(a, a);
    for (int i = 0; i < 8; ++i) x[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(x[i]) : "l"(x[i]), "l"(A));
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K> static float run(K k, float* o, int grid, int block) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<grid, block>>>(o, 1.0001f, 0.5f); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int r = 0; r < 5; ++r) k<<<grid, block>>>(o, 1.0001f, 0.5f); cudaEventRecord(b);
    cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    int grid = 148 * 8, block = 256; float* o; cudaMalloc(&o, grid * block * 4);
    double n = (double)grid * block * ITER * 8;
    float t;
    t = run(k_scalar_fma, o, grid, block); printf("scalar FFMA : %.3f ms  %.1f Gop/s\n", t, n / t / 1e6);
    t = run(k_packed_fma, o, grid, block); printf("packed FFMA2: %.3f ms  %.1f Gop/s\n", t, 2 * n / t / 1e6);
    t = run(k_scalar_add, o, grid, block); printf("scalar FADD : %.3f ms  %.1f Gop/s\n", t, n / t / 1e6);
    t = run(k_packed_add, o, grid, block); printf("packed FADD2: %.3f ms  %.1f Gop/s\n", t, 2 * n / t / 1e6);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
