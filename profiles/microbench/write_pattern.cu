// Microbenchmark: floor of the COO row-write pattern.  One warp per atom writes its row
// (src[p0..p0+cnt) = i, dst[...] = k, shifts[3*p0 .. 3*(p0+cnt)) = 0) with coalesced 4-byte stores.
//   order=0: atoms visited in index order (rows land sequentially in memory)
//   order=1: atoms visited in a random permutation (what a cell-ordered sweep does to randomly indexed atoms)
// N = 1M atoms, 90 neighbors each -> 1.81 GB written per launch.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
template <int ST>
__device__ __forceinline__ void st(int* p, int v) {
    if (ST == 0) *p = v;
    else if (ST == 1) __stcs(p, v);   // streaming (evict-first)
    else if (ST == 2) __stwt(p, v);   // write-through
    else __stcg(p, v);                // cache-global (L2 only)
}
template <int ST>
__global__ void k_rows_st(const int* __restrict__ order, const int* __restrict__ ptr, int n, int* __restrict__ src,
                          int* __restrict__ dst, int* __restrict__ sh) {
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nw) {
        const int i = order[w];
        const size_t p0 = (size_t)ptr[i];
        const int cnt = ptr[i + 1] - ptr[i];
        for (int k = lane; k < cnt; k += 32) { st<ST>(src + p0 + k, i); st<ST>(dst + p0 + k, w); }
        int* s = sh + 3 * p0;
        for (int e = lane; e < 3 * cnt; e += 32) st<ST>(s + e, 0);
    }
}
__global__ void k_rows(const int* __restrict__ order, const int* __restrict__ ptr, int n, int* __restrict__ src,
                       int* __restrict__ dst, int* __restrict__ sh, int vec) {
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nw) {
        const int i = order[w];
        const size_t p0 = (size_t)ptr[i];
        const int cnt = ptr[i + 1] - ptr[i];
        for (int k = lane; k < cnt; k += 32) { src[p0 + k] = i; dst[p0 + k] = w; }
        int* s = sh + 3 * p0;
        if (vec) {
            // 16-byte stores on the aligned middle part
            const size_t a0 = (size_t)s;
            int head = (int)(((16 - (a0 & 15)) & 15) >> 2);
            if (head > 3 * cnt) head = 3 * cnt;
            if (lane < head) s[lane] = 0;
            const int nv = (3 * cnt - head) >> 2;
            int4* v = reinterpret_cast<int4*>(s + head);
            for (int e = lane; e < nv; e += 32) v[e] = make_int4(0, 0, 0, 0);
            for (int e = head + 4 * nv + lane; e < 3 * cnt; e += 32) s[e] = 0;
        } else {
            for (int e = lane; e < 3 * cnt; e += 32) s[e] = 0;
        }
    }
}
int main() {
    const int n = 1000000, per = 90;
    std::vector<int> ptr(n + 1), seq(n), rnd(n);
    std::mt19937 g(1);
    long long P = 0;
    for (int i = 0; i < n; ++i) { ptr[i] = (int)P; P += per + (int)(g() % 21) - 10; seq[i] = i; rnd[i] = i; }
    ptr[n] = (int)P;
    std::shuffle(rnd.begin(), rnd.end(), g);
    int *dptr, *dseq, *drnd, *src, *dst, *sh;
    cudaMalloc(&dptr, (n + 1) * 4); cudaMalloc(&dseq, n * 4); cudaMalloc(&drnd, n * 4);
    cudaMalloc(&src, P * 4); cudaMalloc(&dst, P * 4); cudaMalloc(&sh, P * 12);
    cudaMemcpy(dptr, ptr.data(), (n + 1) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dseq, seq.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(drnd, rnd.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int vec = 0; vec < 2; ++vec)
        for (int ord = 0; ord < 2; ++ord) {
            const int* o = ord ? drnd : dseq;
            for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
                k_rows<<<blocks, 256>>>(o, dptr, n, src, dst, sh, vec);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                for (int r = 0; r < 5; ++r) k_rows<<<blocks, 256>>>(o, dptr, n, src, dst, sh, vec);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
                printf("vec=%d order=%s blocks=%5d  %.3f ms  %.0f GB/s\n", vec, ord ? "random" : "sequential", blocks, ms,
                       P * 20.0 / ms / 1e6);
            }
        }
    {
        auto run = [&](const char* name, auto kern) {
            kern<<<148 * 8, 256>>>(drnd, dptr, n, src, dst, sh);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            for (int r = 0; r < 5; ++r) kern<<<148 * 8, 256>>>(drnd, dptr, n, src, dst, sh);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            printf("random order, store variant %-22s %.3f ms  %.0f GB/s\n", name, ms, P * 20.0 / ms / 1e6);
        };
        run("default (st.global)", k_rows_st<0>);
        run("st.global.cs", k_rows_st<1>);
        run("st.global.wt", k_rows_st<2>);
        run("st.global.cg", k_rows_st<3>);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
