// Microbenchmark: issue throughput of scalar FADD/FFMA vs packed FADD2/FFMA2 (add.f32x2 / fma.rn.f32x2) on sm_100a.
// Decides whether the neighbor-list distance test should process two candidates per lane with packed math.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_scalar(float* out, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; ++i) {
        x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
        x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long x, unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long x, unsigned long long a) {
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(a));
    return r;
}
__global__ void k_packed(float* out, float a, float b) {
    unsigned long long A, B, x[8];
    asm("mov.b64 %0, {%1,%1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1,%1};" : "=l"(B) : "f"(b));
    for (int k = 0; k < 8; ++k) { float v = threadIdx.x + k; asm("mov.b64 %0, {%1,%1};" : "=l"(x[k]) : "f"(v)); }
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fma2(x[k], A, B);
    }
    float s = 0;
    for (int k = 0; k < 8; ++k) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_add(float* out, float a) {
    unsigned long long A, x[8];
    asm("mov.b64 %0, {%1,%1};" : "=l"(A) : "f"(a));
    for (int k = 0; k < 8; ++k) { float v = threadIdx.x + k; asm("mov.b64 %0, {%1,%1};" : "=l"(x[k]) : "f"(v)); }
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = add2(x[k], A);
    }
    float s = 0;
    for (int k = 0; k < 8; ++k) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_scalar_add(float* out, float a) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; ++i) {
        x0 = __fadd_rn(x0, a); x1 = __fadd_rn(x1, a); x2 = __fadd_rn(x2, a); x3 = __fadd_rn(x3, a);
        x4 = __fadd_rn(x4, a); x5 = __fadd_rn(x5, a); x6 = __fadd_rn(x6, a); x7 = __fadd_rn(x7, a);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;   // 64 warps/SM
    auto run = [&](const char* name, auto launch, double ops_per_thread_iter) {
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); for (int r = 0; r < 5; ++r) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        double winstr = (double)blocks * threads / 32 * ITERS * 8;
        printf("%-12s %.3f ms  %.1f G warp-instr/s  %.2f T lane-ops/s\n", name, ms, winstr / ms / 1e6, winstr * 32 * ops_per_thread_iter / ms / 1e9);
    };
    run("FFMA", [&] { k_scalar<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 1);
    run("FFMA2", [&] { k_packed<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 2);
    run("FADD", [&] { k_scalar_add<<<blocks, threads>>>(out, 0.5f); }, 1);
    run("FADD2", [&] { k_packed_add<<<blocks, threads>>>(out, 0.5f); }, 2);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
