// Development probe (not part of the library): one-way latency of a CTA-to-CTA hand-over through global memory (st.relaxed.gpu by the
// producer, ld.relaxed.gpu polling by the consumer, as the sweep's hand-over rings do), idle and while the other SMs stream memory at
// full bandwidth; and the same through distributed shared memory inside a cluster of two CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/handover_probe scripts/handover_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void st_relaxed(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// blocks 0 and 1 play ping-pong on two words; all other blocks copy `src` to `dst` until the game is over (load = 1)
__global__ void pingpong(uint32_t *flags, int iters, long long *cyc, const uint4 *src, uint4 *dst, size_t n, int load)
{
    uint32_t *A = flags, *B = flags + 64, *done = flags + 128; // (separate 128-byte lines)
    if (blockIdx.x >= 2) {
        if (!load) return;
        size_t i = (size_t)(blockIdx.x - 2) * blockDim.x + threadIdx.x, stride = (size_t)(gridDim.x - 2) * blockDim.x;
        while (ld_relaxed(done) == 0) {
            for (int k = 0; k < 64; k++, i += stride) {
                if (i >= n) i -= n;
                dst[i] = src[i];
            }
        }
        return;
    }
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
        if (blockIdx.x == 0) {
            st_relaxed(A, i);
            while (ld_relaxed(B) != (uint32_t)i) {}
        } else {
            while (ld_relaxed(A) != (uint32_t)i) {}
            st_relaxed(B, i);
        }
    }
    long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0) st_relaxed(done, 1);
}

// the same game through distributed shared memory: cluster of 2 CTAs, remote store + local polling
__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(int iters, long long *cyc)
{
    __shared__ volatile uint32_t box;
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x == 0) box = 0;
    cl.sync();
    const unsigned rank = cl.block_rank();
    volatile uint32_t *peer = cl.map_shared_rank((uint32_t *)&box, rank ^ 1);
    if (threadIdx.x == 0 && blockIdx.x < 2) {
        long long t0 = clock64();
        for (int i = 1; i <= iters; i++) {
            if (rank == 0) {
                *peer = i;
                while (box != (uint32_t)i) {}
            } else {
                while (box != (uint32_t)i) {}
                *peer = i;
            }
        }
        cyc[rank] = clock64() - t0;
    }
    cl.sync();
}

int main()
{
    uint32_t *flags; long long *cyc; uint4 *src, *dst;
    const size_t n = (size_t)1 << 26; // 1 GiB each
    cudaMalloc(&flags, 1024); cudaMalloc(&cyc, 64); cudaMalloc(&src, n * 16); cudaMalloc(&dst, n * 16);
    cudaMemset(src, 1, n * 16);
    const int iters = 2000;
    long long h[2];
    for (int load = 0; load < 2; load++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(flags, 0, 1024);
            pingpong<<<148, 1024>>>(flags, iters, cyc, src, dst, n, load);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("global-memory hand-over, %s: %.0f cycles one way (round trip / 2; store + polling load)  (%s)\n", load ? "146 SMs streaming a 1 GiB copy" : "idle GPU",
               (double)h[0] / iters / 2, cudaGetErrorString(cudaGetLastError()));
    }
    pingpong_dsmem<<<2, 32>>>(iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("distributed-shared-memory hand-over inside a cluster of 2: %.0f cycles one way  (%s)\n", (double)h[0] / iters / 2, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
