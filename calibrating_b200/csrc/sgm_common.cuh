// sgm_common.cuh -- device helpers shared by the aggregation kernels (sgbm_agg.cu, sgbm_wave.cu): the packed int16x2
// SGM step, cp.async / bulk-copy / mbarrier wrappers and the bounded spin-wait.  Not part of the public ABI.
#pragma once
#include <stdint.h>

#include "b2s_internal.h"

namespace {


constexpr int WARPS = 8;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
template <int NP> struct Stages { static constexpr int value = (NP <= 2) ? 16 : (NP == 3 ? 10 : 8); }; // x 3 sources x 8 warps <= 98 KB


template <int NP> __device__ __forceinline__ void store_regs(int16_t *dst, const uint32_t (&v)[NP])
{
    if constexpr (NP == 1) *(uint32_t *)dst = v[0];
    else if constexpr (NP == 2) *(uint2 *)dst = make_uint2(v[0], v[1]);
    else if constexpr (NP == 4) *(uint4 *)dst = make_uint4(v[0], v[1], v[2], v[3]);
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) ((uint32_t *)dst)[i] = v[i];
    }
}

template <int NP> __device__ __forceinline__ void ldcg_regs(const int16_t *src, uint32_t (&v)[NP])
{
    if constexpr (NP == 1) v[0] = __ldcg((const uint32_t *)src);
    else if constexpr (NP == 2) { uint2 t = __ldcg((const uint2 *)src); v[0] = t.x; v[1] = t.y; }
    else if constexpr (NP == 4) { uint4 t = __ldcg((const uint4 *)src); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) v[i] = __ldcg((const uint32_t *)src + i);
    }
}
template <int NP> __device__ __forceinline__ void stcg_regs(int16_t *dst, const uint32_t (&v)[NP])
{
    if constexpr (NP == 1) __stcg((uint32_t *)dst, v[0]);
    else if constexpr (NP == 2) __stcg((uint2 *)dst, make_uint2(v[0], v[1]));
    else if constexpr (NP == 4) __stcg((uint4 *)dst, make_uint4(v[0], v[1], v[2], v[3]));
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) __stcg((uint32_t *)dst + i, v[i]);
    }
}
// shared memory through 32-bit shared-window addresses (keeps generic->shared conversions out of the loop)
template <int NP> __device__ __forceinline__ void lds_s(uint32_t addr, uint32_t (&v)[NP])
{
    if constexpr (NP == 2) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr) : "memory");
    else if constexpr (NP == 4)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr) : "memory");
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[i]) : "r"(addr + 4 * i) : "memory");
    }
}
template <int NP> __device__ __forceinline__ void sts_s(uint32_t addr, const uint32_t (&v)[NP])
{
    if constexpr (NP == 2) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(v[0]), "r"(v[1]) : "memory");
    else if constexpr (NP == 4)
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr + 4 * i), "r"(v[i]) : "memory");
    }
}
__device__ __forceinline__ void cp_async16_s(uint32_t saddr, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}

// one step of one path: T = normalised state of the predecessor pixel (in/out), c = C of this pixel, L = L_r of this pixel
template <int NP, bool PAD>
__device__ __forceinline__ void sgm_step(uint32_t (&T)[NP], const uint32_t (&c)[NP], uint32_t (&L)[NP], const uint32_t (&padmask)[NP],
                                         uint32_t P1v, uint32_t P2mP1v, int lane)
{
    const uint32_t BIG = 0x7FFF7FFFu;
    uint32_t up = __shfl_up_sync(0xffffffffu, T[NP - 1], 1);
    uint32_t dn = __shfl_down_sync(0xffffffffu, T[0], 1);
    // the d = -1 / d = D sentinels of the edge lanes as a multiply-add (FMA pipe) instead of a select: the kernels that
    // use this step are bound by the ALU pipe, which the packed min / max cannot leave
    // (the factors go through an empty asm so that the compiler does not turn the multiply-add back into a select)
    uint32_t ku = lane != 0 ? 1u : 0u, kd = lane != 31 ? 1u : 0u;
    asm("" : "+r"(ku));
    asm("" : "+r"(kd));
    up = up * ku + (lane == 0 ? BIG : 0u);
    dn = dn * kd + (lane == 31 ? BIG : 0u);
    uint32_t m = BIG;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        uint32_t lft = __byte_perm(i == 0 ? up : T[i - 1], T[i], 0x5432);      // L(d-1)
        uint32_t rgt = __byte_perm(T[i], i == NP - 1 ? dn : T[i + 1], 0x5432); // L(d+1)
        uint32_t t = __vimin3_s16x2(lft, rgt, P2mP1v);
        t = __viaddmin_s16x2(t, P1v, T[i]);
        L[i] = c[i] + t; // packed add as a 32-bit add (FMA pipe instead of the busier ALU pipe): 0 <= t <= P2 and C >= 0, no carry between halves
        if (PAD) L[i] |= padmask[i];
        m = __vmins2(m, L[i]);
    }
    m = __vmins2(m, __byte_perm(m, m, 0x1032));                 // both halves = min over this lane's disparities
    m = (uint32_t)__reduce_min_sync(0xffffffffu, (int)m);       // signed 32-bit min of (v,v) pairs = (min,min)
#pragma unroll
    for (int i = 0; i < NP; i++) {
        T[i] = L[i] - m; // both halves of L are >= their half of m: the 32-bit difference has no borrow = packed difference
        if (PAD) T[i] |= padmask[i];
    }
}

// ---- spin-wait guard and mbarrier helpers (fused vertical sweep and the bulk-copy pipeline of the horizontal scans) ----
constexpr int HO_SLOTS = 4;
constexpr unsigned long long WAIT_TIMEOUT_NS = 2000000000ull; // a hand-over / neighbour wait longer than 2 s is a lost strip: flag it, do not hang
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// slow path of the spin loops: true when the wait should be abandoned (timeout, or another warp already flagged an error)
__device__ __forceinline__ bool wait_expired(int spins, unsigned long long &t0, int *err)
{
    if ((spins & 255) != 0) return false;
    if (*(volatile int *)err != 0) return true;
    const unsigned long long now = global_ns();
    if (t0 == 0) t0 = now;
    return now - t0 > WAIT_TIMEOUT_NS;
}

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(addr) : "memory");
}
// wait until the phase of parity `parity` of the mbarrier has completed (acquire)
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity, int *err)
{
    uint32_t ok;
    int spins = 0;
    unsigned long long t0 = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && wait_expired(++spins, t0, err)) {
            *(volatile int *)err = 1;
            break;
        }
    } while (!ok);
}

// the same with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint expires) instead
// of coming back to spin; for waits that normally last a pipeline stage rather than a few cycles (sgbm_wave.cu)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t addr, uint32_t parity, int *err)
{
    uint32_t ok;
    int spins = 0;
    unsigned long long t0 = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(100000u) : "memory");
        if (!ok && wait_expired(++spins, t0, err)) {
            *(volatile int *)err = 1;
            break;
        }
    } while (!ok);
}

// 1-D bulk copy global -> shared (UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t sdst, const void *gsrc, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t addr, uint32_t bytes)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(addr), "r"(bytes) : "memory");
}

} // namespace
