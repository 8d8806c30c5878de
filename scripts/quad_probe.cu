// Development probe (not part of the library): how fast can ONE warp run the SGM step in the block layout when a lane holds 8
// words (8 lanes per pixel, 4 pixels per warp), with W warps per SM and nothing else going on?  Prints cycles per "row" (three
// path steps + the saturating sums + shared-memory traffic shaped like the sweep's) for several warp counts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I calibrating_b200/csrc -o scripts/quad_probe scripts/quad_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "sgm_common.cuh"

template <int MODE> __global__ void __launch_bounds__(1024, 1) probe(uint32_t *out, long long *cyc, int iters, const int16_t *C = nullptr, int16_t *S = nullptr, int width1 = 0, int16_t *S2 = nullptr)
{
    extern __shared__ __align__(16) uint32_t sm[];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + wi * 8192 + lane * 32;
    for (int i = threadIdx.x; i < (int)blockDim.x * 64; i += blockDim.x) sm[i] = (i * 2654435761u) & 0x0FFF0FFFu;
    __syncthreads();
    const uint32_t P1v = 600u * 0x10001u, P2mP1v = 1800u * 0x10001u, BIG = 0x7FFF7FFFu;
    long long t0 = clock64();
    if (MODE == 0) { // quad: 8 words per lane, groups of 8 lanes
        const int li = lane & 7;
        const uint32_t ku = li != 0, au = li == 0 ? BIG : 0u, kd = li != 7, ad = li == 7 ? BIG : 0u;
        uint32_t T0[8], T1[8], Td[8], c[8], s[8], L[8], pm[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { T0[i] = T1[i] = Td[i] = 0; s[i] = 0; pm[i] = 0; }
        for (int it = 0; it < iters; it++) {
            const uint32_t a = sbase + (it & 3) * 1024;
            uint32_t x[4], y[4];
            lds_s<4>(a, x); lds_s<4>(a + 16, y);
#pragma unroll
            for (int i = 0; i < 4; i++) { c[i] = x[i]; c[4 + i] = y[i]; }
            lds_s<4>(a + 2048, x); lds_s<4>(a + 2064, y);
#pragma unroll
            for (int i = 0; i < 4; i++) { T0[i] = x[i] & 0x0FFF0FFFu; T0[4 + i] = y[i] & 0x0FFF0FFFu; }
            sgm_step_blk<8, 8, false>(T0, c, L, pm, P1v, P2mP1v, ku, au, kd, ad);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = __viaddmin_u16x2(s[i] & 0x0FFF0FFFu, L[i], BIG);
            sgm_step_blk<8, 8, false>(T1, c, L, pm, P1v, P2mP1v, ku, au, kd, ad);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = __viaddmin_u16x2(s[i], L[i], BIG);
#pragma unroll
            for (int i = 0; i < 4; i++) { x[i] = T0[i]; y[i] = T0[4 + i]; }
            sts_s<4>(a + 2048, x); sts_s<4>(a + 2064, y);
            sgm_step_blk<8, 8, false>(Td, c, L, pm, P1v, P2mP1v, ku, au, kd, ad);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = __viaddmin_u16x2(s[i], L[i], BIG);
#pragma unroll
            for (int i = 0; i < 4; i++) { x[i] = s[i]; y[i] = s[4 + i]; }
            sts_s<4>(a + 1024 * 3 - 1024, x); // (keeps the sums alive)
            __syncwarp();
        }
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) r ^= s[i] ^ T1[i] ^ Td[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 4) { // quad + the sweep's global traffic: 4 columns x (C | S) = 2 KB per row and warp through an 8-stage cp.async ring
        constexpr int R = 8;
        const int li = lane & 7, g = lane >> 3;
        const uint32_t ku = li != 0, au = li == 0 ? BIG : 0u, kd = li != 7, ad = li == 7 ? BIG : 0u;
        uint32_t T0[8], T1[8], Td[8], c[8], s[8], L[8], pm[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { T0[i] = T1[i] = Td[i] = 0; pm[i] = 0; }
        // per warp: ring 8 stages x 2 KB = 16 KB, slots 4 KB  (20 KB per warp)
        const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sm) + wi * 20480, slot = ring + 16384 + g * 256 + li * 32;
        const int x = (blockIdx.x * (blockDim.x >> 5) + wi) * 4;
        const long long rs = (long long)width1 * 128;
        // 128 segments of 16 B per row: lane handles segments lane, lane+32 (C: 4 columns x 16) and lane+64, lane+96 (S)
        const int16_t *srcC = C + (long long)x * 128 + lane * 8, *srcS = S + (long long)x * 128 + lane * 8;
        int16_t *sp = S + (long long)(x + g) * 128 + li * 16;
        const uint32_t dst = ring + lane * 16;
        uint32_t o_iss = 0, o_cur = 0;
        auto issue = [&](uint32_t sz) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + o_iss), "l"(srcC), "r"(sz) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + o_iss + 512), "l"(srcC + 256), "r"(sz) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + o_iss + 1024), "l"(srcS), "r"(sz) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + o_iss + 1536), "l"(srcS + 256), "r"(sz) : "memory");
            srcC += rs; srcS += rs; o_iss = (o_iss + 2048) & 16383;
        };
        for (int p = 0; p < R - 1; p++) { issue(16u); cp_async_commit(); }
        cp_async_wait<R - 2>();
        __syncwarp();
        const uint32_t cbase = ring + g * 256 + li * 32;
        uint32_t x4[4], y4[4];
        lds_s<4>(cbase, x4); lds_s<4>(cbase + 16, y4);
#pragma unroll
        for (int i = 0; i < 4; i++) { c[i] = x4[i]; c[4 + i] = y4[i]; }
        lds_s<4>(cbase + 1024, x4); lds_s<4>(cbase + 1040, y4);
#pragma unroll
        for (int i = 0; i < 4; i++) { s[i] = x4[i]; s[4 + i] = y4[i]; }
        for (int it = 0; it < iters; it++) {
            const uint32_t si = slot + (it & 1) * 1024, so = slot + ((it + 1) & 1) * 1024;
            lds_s<4>(si, x4); lds_s<4>(si + 16, y4);
#pragma unroll
            for (int i = 0; i < 4; i++) { T0[i] = x4[i] & 0x0FFF0FFFu; T0[4 + i] = y4[i] & 0x0FFF0FFFu; c[i] &= 0x0FFF0FFFu; c[4 + i] &= 0x0FFF0FFFu; }
            lds_s<4>(si + 2048, x4); lds_s<4>(si + 2064, y4);
#pragma unroll
            for (int i = 0; i < 4; i++) { T1[i] = x4[i] & 0x0FFF0FFFu; T1[4 + i] = y4[i] & 0x0FFF0FFFu; }
            sgm_step_blk<8, 8, false>(T0, c, L, pm, P1v, P2mP1v, ku, au, kd, ad);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = __viaddmin_u16x2(s[i] & 0x0FFF0FFFu, L[i], BIG);
            sgm_step_blk<8, 8, false>(T1, c, L, pm, P1v, P2mP1v, ku, au, kd, ad);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = __viaddmin_u16x2(s[i], L[i], BIG);
#pragma unroll
            for (int i = 0; i < 4; i++) { x4[i] = T0[i]; y4[i] = T0[4 + i]; }
            sts_s<4>(so, x4); sts_s<4>(so + 16, y4);
#pragma unroll
            for (int i = 0; i < 4; i++) { x4[i] = T1[i]; y4[i] = T1[4 + i]; }
            sts_s<4>(so + 2048, x4); sts_s<4>(so + 2064, y4);
            sgm_step_blk<8, 8, false>(Td, c, L, pm, P1v, P2mP1v, ku, au, kd, ad);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = __viaddmin_u16x2(s[i], L[i], BIG);
            __stcg((uint4 *)sp, make_uint4(s[0], s[1], s[2], s[3]));
            __stcg((uint4 *)sp + 1, make_uint4(s[4], s[5], s[6], s[7]));
            sp += rs;
            __syncwarp();
            issue(it + R - 1 < iters ? 16u : 0u);
            cp_async_commit();
            cp_async_wait<R - 2>();
            __syncwarp();
            o_cur = (o_cur + 2048) & 16383;
            lds_s<4>(cbase + o_cur, x4); lds_s<4>(cbase + o_cur + 16, y4);
#pragma unroll
            for (int i = 0; i < 4; i++) { c[i] = x4[i]; c[4 + i] = y4[i]; }
            lds_s<4>(cbase + o_cur + 1024, x4); lds_s<4>(cbase + o_cur + 1040, y4);
#pragma unroll
            for (int i = 0; i < 4; i++) { s[i] = x4[i]; s[4 + i] = y4[i]; }
        }
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) r ^= s[i] ^ T1[i] ^ Td[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE >= 2) { // pair layout + the sweep's global traffic: C (and S) rows through an 8-stage cp.async ring, S written back
        constexpr int R = 8;
        uint32_t T0[2], T1[2], Td[2], c[2], s[2], L0[2], L1[2], L2[2], pm[2] = {0, 0};
        T0[0] = T0[1] = T1[0] = T1[1] = Td[0] = Td[1] = 0;
        const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sm) + wi * 8192, slot = ring + 4096 + lane * 8;
        const int nw = blockDim.x >> 5, half = nw / 2;
        const bool upw = MODE == 5 && wi >= half;               // MODE 5: second half of the warps sweeps bottom-up and writes S2
        const bool readS = MODE == 2 || (MODE == 5 && !upw);
        const int x = MODE == 5 ? blockIdx.x * half + (upw ? wi - half : wi) : blockIdx.x * nw + wi;
        const long long rs = (long long)width1 * 128 * (upw ? -1 : 1);
        const long long o0 = (long long)x * 128 + (upw ? (long long)(iters - 1) * width1 * 128 : 0);
        const int isS = lane >> 4, r = lane & 15;
        const int16_t *src = (isS ? S : C) + o0 + r * 8;
        int16_t *sp = (upw ? S2 : S) + o0 + lane * 4;
        const uint32_t dst = ring + lane * 16;
        uint32_t o_iss = 0, o_cur = 0;
        for (int p = 0; p < R - 1; p++) {
            if (readS || !isS) cp_async16_s(dst + o_iss, src);
            src += rs; o_iss = (o_iss + 512) & 4095;
            cp_async_commit();
        }
        cp_async_wait<R - 2>();
        __syncwarp();
        lds_s<2>(ring + lane * 8, c);
        lds_s<2>(ring + 256 + lane * 8, s);
        for (int it = 0; it < iters; it++) {
            lds_s<2>(slot + (it & 1) * 256, T0);
            lds_s<2>(slot + 512 + (it & 1) * 256, T1);
            T0[0] &= 0x0FFF0FFFu; T0[1] &= 0x0FFF0FFFu; T1[0] &= 0x0FFF0FFFu; T1[1] &= 0x0FFF0FFFu;
            c[0] &= 0x0FFF0FFFu; c[1] &= 0x0FFF0FFFu;
            sgm_step<2, false>(T0, c, L0, pm, P1v, P2mP1v, lane);
            sgm_step<2, false>(T1, c, L1, pm, P1v, P2mP1v, lane);
            sts_s<2>(slot + ((it + 1) & 1) * 256, T0);
            sts_s<2>(slot + 512 + ((it + 1) & 1) * 256, T1);
            sgm_step<2, false>(Td, c, L2, pm, P1v, P2mP1v, lane);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint32_t v = __viaddmin_u16x2(L0[i], L1[i], BIG);
                v = __viaddmin_u16x2(v, L2[i], BIG);
                s[i] = __viaddmin_u16x2(s[i] & 0x0FFF0FFFu, v, BIG);
            }
            stcg_regs<2>(sp, s);
            sp += rs;
            __syncwarp();
            if (readS || !isS) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + o_iss), "l"(src), "r"(it + R - 1 < iters ? 16u : 0u) : "memory");
            src += rs; o_iss = (o_iss + 512) & 4095;
            cp_async_commit();
            cp_async_wait<R - 2>();
            __syncwarp();
            o_cur = (o_cur + 512) & 4095;
            lds_s<2>(ring + o_cur + lane * 8, c);
            if (readS) lds_s<2>(ring + o_cur + 256 + lane * 8, s);
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s[0] ^ s[1] ^ T1[0] ^ Td[1];
    } else { // the pair layout of the shipped sweep: 2 words per lane, one pixel per warp
        uint32_t T0[2], T1[2], Td[2], c[2], s[2], L0[2], L1[2], L2[2], pm[2] = {0, 0};
        T0[0] = T0[1] = T1[0] = T1[1] = Td[0] = Td[1] = s[0] = s[1] = 0;
        const uint32_t sb2 = (uint32_t)__cvta_generic_to_shared(sm) + wi * 8192 + lane * 8;
        for (int it = 0; it < iters; it++) {
            const uint32_t a = sb2 + (it & 3) * 256;
            lds_s<2>(a, c);
            lds_s<2>(a + 2048, T0);
            T0[0] &= 0x0FFF0FFFu; T0[1] &= 0x0FFF0FFFu;
            sgm_step<2, false>(T0, c, L0, pm, P1v, P2mP1v, lane);
            sgm_step<2, false>(T1, c, L1, pm, P1v, P2mP1v, lane);
            sts_s<2>(a + 2048, T0);
            sgm_step<2, false>(Td, c, L2, pm, P1v, P2mP1v, lane);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint32_t v = __viaddmin_u16x2(L0[i], L1[i], BIG);
                v = __viaddmin_u16x2(v, L2[i], BIG);
                s[i] = __viaddmin_u16x2(s[i] & 0x0FFF0FFFu, v, BIG);
            }
            sts_s<2>(a + 1024, s);
            __syncwarp();
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s[0] ^ s[1] ^ T1[0] ^ Td[1];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 8192);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 8192);
    const int iters = 1080;
    long long h[148];
    const int W1 = 148 * 24;
    int16_t *C, *S;
    cudaMalloc(&C, (size_t)iters * W1 * 256); cudaMalloc(&S, (size_t)iters * W1 * 256);
    cudaMemset(C, 1, (size_t)iters * W1 * 256); cudaMemset(S, 0, (size_t)iters * W1 * 256);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 8192);
    cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 8192);
    cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 10 * 20480);
    cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 26 * 8192);
    int16_t *S2;
    cudaMalloc(&S2, (size_t)iters * W1 * 256);
    for (int rep = 0; rep < 3; rep++) {
        probe<5><<<138, 26 * 32, 26 * 8192>>>(out, cyc, iters, C, S, 138 * 13, S2);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    { double avg = 0; for (int i = 0; i < 138; i++) avg += h[i]; avg /= 138;
      printf("pair, 13 columns x 2 sweeps per CTA on 138 CTAs (the shipped kernel's shape, no synchronisation): %7.1f cycles per row (%s)\n", avg / iters, cudaGetErrorString(cudaGetLastError())); }
    for (int mode = 0; mode < 5; mode++)
        for (int warps : {1, 2, 4, 6, 7, 8, 10, 12, 16, 24}) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) probe<0><<<148, warps * 32, 24 * 8192>>>(out, cyc, iters);
                else if (mode == 1) probe<1><<<148, warps * 32, 24 * 8192>>>(out, cyc, iters);
                else if (mode == 2) probe<2><<<148, warps * 32, 24 * 8192>>>(out, cyc, iters, C, S, W1);
                else if (mode == 3) probe<3><<<148, warps * 32, 24 * 8192>>>(out, cyc, iters, C, S, W1);
                else if (warps <= 10) probe<4><<<148, warps * 32, 10 * 20480>>>(out, cyc, iters, C, S, W1);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
            const int cols = (mode == 0 || mode == 4) ? 4 * warps : warps;
            if (mode == 4 && warps > 10) continue;
            printf("%s warps/SM %2d: %7.1f cycles per row of %3d columns  -> %6.1f cycles per column-row  (%s)\n", mode == 0 ? "quad" : mode == 1 ? "pair" : mode == 2 ? "pair+C,S stream" : mode == 3 ? "pair+C stream,S2 write" : "quad+C,S stream", warps,
                   avg / iters, cols, avg / iters / cols, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
