// microbench.cu -- throughput probes that drive kernel design decisions (not product code).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 256;

__global__ void k_atoms_ret(int *out, int spread)
{
    __shared__ int cnt[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    int acc = 0;
    unsigned a = threadIdx.x * 7u;
    for (int it = 0; it < ITERS; it++) {
        acc += atomicAdd(&cnt[(a * spread) & 2047], 1);
        a += 13;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_atoms_noret(int *out, int spread)
{
    __shared__ int cnt[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    unsigned a = threadIdx.x * 7u;
    for (int it = 0; it < ITERS; it++) {
        atomicAdd(&cnt[(a * spread) & 2047], 1);
        a += 13;
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = cnt[threadIdx.x];
}

__global__ void k_sts_lds(int *out)
{
    __shared__ int cnt[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    int acc = 0;
    unsigned a = threadIdx.x * 7u;
    for (int it = 0; it < ITERS; it++) {
        const int v = cnt[a & 2047];
        cnt[(a + 1) & 2047] = v + 1;
        acc += v;
        a += 13;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_match(int *out, int groups)
{
    int acc = 0;
    int v = (threadIdx.x & 31) % groups;
    for (int it = 0; it < ITERS; it++) {
        const unsigned m = __match_any_sync(0xffffffffu, v + it);
        acc += __popc(m);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_gatom(int *g, int *out, int naddr, int ret)
{
    int acc = 0;
    unsigned a = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    for (int it = 0; it < 64; it++) {
        const unsigned idx = (a >> 8) % (unsigned)naddr;
        if (ret) acc += atomicAdd(&g[idx], 1); else atomicAdd(&g[idx], 1);
        a = a * 1664525u + 1013904223u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
float timeit(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main()
{
    int *out, *g;
    const int blocks = 148 * 4, threads = 512;
    CK(cudaMalloc(&out, blocks * threads * 4 * 4));
    CK(cudaMalloc(&g, 64 << 20));
    CK(cudaMemset(g, 0, 64 << 20));
    const double lane_ops = (double)blocks * threads * ITERS;
    for (int spread : {1, 0}) {
        float ms = timeit([&] { k_atoms_ret<<<blocks, threads>>>(out, spread); });
        printf("smem atomicAdd with return, %s: %.3f ms -> %.1f G lane-ops/s (%.2f cyc/lane/SM @1.9GHz)\n",
               spread ? "spread" : "one address", ms, lane_ops / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 / lane_ops);
        ms = timeit([&] { k_atoms_noret<<<blocks, threads>>>(out, spread); });
        printf("smem atomicAdd no return,   %s: %.3f ms -> %.1f G lane-ops/s (%.2f cyc/lane/SM)\n",
               spread ? "spread" : "one address", ms, lane_ops / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 / lane_ops);
    }
    {
        float ms = timeit([&] { k_sts_lds<<<blocks, threads>>>(out); });
        printf("smem LDS+STS pair:            %.3f ms -> %.1f G lane-pairs/s (%.2f cyc/lane/SM)\n", ms,
               lane_ops / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 / lane_ops);
    }
    for (int groups : {1, 4, 8, 32}) {
        float ms = timeit([&] { k_match<<<blocks, threads>>>(out, groups); });
        printf("match_any, %2d groups/warp:   %.3f ms -> %.1f G lane-ops/s (%.1f cyc/warp-instr/SMSP)\n", groups, ms,
               lane_ops / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 * 4 / (lane_ops / 32));
    }
    const double gops = (double)blocks * threads * 64;
    for (int naddr : {64, 4096, 1 << 20, 16 << 20}) {
        for (int ret : {0, 1}) {
            float ms = timeit([&] { k_gatom<<<blocks, threads>>>(g, out, naddr, ret); });
            printf("global atomicAdd %s, %8d addresses: %.3f ms -> %.1f G atomics/s\n", ret ? "ret  " : "noret", naddr, ms,
                   gops / ms / 1e6);
        }
    }
    return 0;
}
