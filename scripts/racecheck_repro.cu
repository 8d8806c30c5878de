// Minimal reproducer for the racecheck finding of round 1: compute-sanitizer --tool racecheck reports
// shared-memory hazards between the slot stores and loads of the lock-step sweep (sgm_common.cuh: sts_s / lds_s), which are
// ordered by an mbarrier (producer: st.shared, __syncwarp, mbarrier.arrive [release]; consumer: mbarrier.try_wait.parity
// [acquire], ld.shared).  This file is that pattern and nothing else: warp 0 writes a slot and arrives, warp 1 waits and reads,
// 64 rounds with two alternating mbarriers.  If racecheck flags THIS kernel, it does not model inline-PTX mbarrier ordering and
// the finding in the sweep is a tool limitation; the checksum proves the consumer always saw the producer's data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o racecheck_repro scripts/racecheck_repro.cu
//   compute-sanitizer --tool racecheck ./racecheck_repro
#include <cstdio>
#include <cstdint>
__global__ void repro(uint32_t *out)
{
    __shared__ __align__(8) unsigned long long mb[2];
    __shared__ uint32_t slot[2][32];
    const uint32_t mba = (uint32_t)__cvta_generic_to_shared(mb);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mba + threadIdx.x * 8) : "memory");
    __syncthreads();
    uint32_t sum = 0;
    for (int t = 0; t < 64; t++) {
        const uint32_t b = mba + (t & 1) * 8, par = (t >> 1) & 1;
        if (warp == 0) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot[t & 1][lane])), "r"((uint32_t)(t * 32 + lane)) : "memory");
            __syncwarp();
            if (lane == 0) asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(b) : "memory");
        } else {
            uint32_t ok;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(par) : "memory");
            } while (!ok);
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(&slot[t & 1][lane])) : "memory");
            sum += v == (uint32_t)(t * 32 + lane);
        }
        // (the producer may run at most one round ahead, like a column of the sweep: the consumer's read of round t-2 is ordered
        // before the producer's write of round t by the CTA barrier every second round)
        if (t & 1) __syncthreads();
    }
    if (warp == 1) out[lane] = sum;
}
int main()
{
    uint32_t *d, h[32];
    cudaMalloc(&d, sizeof h);
    repro<<<1, 64>>>(d);
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    int good = 0;
    for (int i = 0; i < 32; i++) good += h[i] == 64;
    printf("cuda %s, lanes that saw every value: %d / 32\n", cudaGetErrorString(e), good);
    return good == 32 ? 0 : 1;
}
