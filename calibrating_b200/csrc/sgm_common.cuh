// sgm_common.cuh -- device helpers shared by the aggregation kernels (sgbm_agg.cu, sgbm_wave.cu): the packed int16x2
// SGM step, cp.async / bulk-copy / mbarrier wrappers and the bounded spin-wait.  Not part of the public ABI.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include <mutex>

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

// One SGM step of an LPP-lane group (LPP = 4 or 8) in the block layout (SgbmGeom::layout 1): T = normalised state of the predecessor pixel (in/out), c = C of this
// pixel, L = L_r of this pixel.  Word i of a lane = word li*N + i of the pixel; word w holds disparities (16b+j, 16b+8+j),
// b = w/8, j = w%8.  The d-1 neighbour of word j > 0 is word j-1, of word 0 it is (hi of the previous block's word 7, lo of
// this block's word 7); the d+1 neighbour of word j < 7 is word j+1, of word 7 it is (hi of this block's word 0, lo of the next
// block's word 0).  Same arithmetic as sgm_step (sgm_common.cuh).
template <int N, int LPP, bool PAD>
__device__ __forceinline__ void sgm_step_blk(uint32_t (&T)[N], const uint32_t (&c)[N], uint32_t (&L)[N], const uint32_t (&padmask)[N],
                                           uint32_t P1v, uint32_t P2mP1v, uint32_t ku, uint32_t au, uint32_t kd, uint32_t ad)
{
    static_assert(N % 8 == 0, "whole blocks of 16 disparities per lane");
    uint32_t up = __shfl_up_sync(0xffffffffu, T[N - 1], 1, LPP);
    uint32_t dn = __shfl_down_sync(0xffffffffu, T[0], 1, LPP);
    up = up * ku + au; // first / last lane of the group: the d = -1 / d = D sentinels (multiply-add: FMA pipe, the ALU pipe is the busy one)
    dn = dn * kd + ad;
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t lft, rgt;
        if (i % 8 == 0) lft = __byte_perm(i == 0 ? up : T[i - 1], T[i + 7], 0x5432);
        else lft = T[i - 1];
        if (i % 8 == 7) rgt = __byte_perm(T[i - 7], i == N - 1 ? dn : T[i + 1], 0x5432);
        else rgt = T[i + 1];
        uint32_t t = __vimin3_s16x2(lft, rgt, P2mP1v);
        t = __viaddmin_s16x2(t, P1v, T[i]);
        L[i] = c[i] + t; // 0 <= t <= P2 and C >= 0: no carry between the halves
        if (PAD) L[i] |= padmask[i];
    }
    uint32_t m = __vimin3_s16x2(L[0], L[1], L[2]);
#pragma unroll
    for (int i = 3; i + 1 < N; i += 2) m = __vimin3_s16x2(m, L[i], L[i + 1]);
    m = __vmins2(m, L[N - 1]);
    m = __vmins2(m, __byte_perm(m, m, 0x1032));
#pragma unroll
    for (int o = LPP / 2; o > 0; o >>= 1) m = __vmins2(m, __shfl_xor_sync(0xffffffffu, m, o, LPP));
#pragma unroll
    for (int i = 0; i < N; i++) {
        T[i] = L[i] - m; // both halves of L are >= their half of m: no borrow
        if (PAD) T[i] |= padmask[i];
    }
}

// The same step for a whole warp per pixel (the horizontal scans) in the block layout: a lane holds NP = 2 or 4 consecutive
// words, so a block of 8 words spans LB = 8 / NP lanes.  Word 0 of a block (register 0 of the lanes with lane % LB == 0) takes its
// d-1 neighbour from (hi of the previous word, lo of the block's word 7 = last register of lane + LB - 1); word 7 takes its d+1
// neighbour from (hi of the block's word 0, lo of the next word): two extra shuffles and two PRMTs with a per-lane selector
// (identity for the lanes in the middle of a block) instead of the 2 PRMTs per register of the pair layout.
template <int NP, bool PAD>
__device__ __forceinline__ void sgm_step_b32(uint32_t (&T)[NP], const uint32_t (&c)[NP], uint32_t (&L)[NP], const uint32_t (&padmask)[NP],
                                             uint32_t P1v, uint32_t P2mP1v, int lane)
{
    static_assert(NP == 2 || NP == 4, "block layout: 2 or 4 words per lane");
    constexpr int LB = 8 / NP;
    const uint32_t BIG = 0x7FFF7FFFu;
    uint32_t up = __shfl_up_sync(0xffffffffu, T[NP - 1], 1);       // word w-1 of register 0
    uint32_t dn = __shfl_down_sync(0xffffffffu, T[0], 1);          // word w+1 of the last register
    const uint32_t w7 = __shfl_down_sync(0xffffffffu, T[NP - 1], LB - 1); // the block's word 7 (for the lanes that hold word 0)
    const uint32_t w0 = __shfl_up_sync(0xffffffffu, T[0], LB - 1);        // the block's word 0 (for the lanes that hold word 7)
    uint32_t ku = lane != 0 ? 1u : 0u, kd = lane != 31 ? 1u : 0u;
    asm("" : "+r"(ku));
    asm("" : "+r"(kd));
    up = up * ku + (lane == 0 ? BIG : 0u);  // d = -1 / d = Dp sentinels (multiply-add: FMA pipe)
    dn = dn * kd + (lane == 31 ? BIG : 0u);
    const int p = lane % LB;
    const uint32_t lft0 = __byte_perm(up, w7, p == 0 ? 0x5432u : 0x3210u);          // (hi of previous word, lo of word 7) or the previous word
    const uint32_t rgtN = __byte_perm(w0, dn, p == LB - 1 ? 0x5432u : 0x7654u);     // (hi of word 0, lo of next word) or the next word
    uint32_t m = BIG;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const uint32_t lft = i == 0 ? lft0 : T[i - 1], rgt = i == NP - 1 ? rgtN : T[i + 1];
        uint32_t t = __vimin3_s16x2(lft, rgt, P2mP1v);
        t = __viaddmin_s16x2(t, P1v, T[i]);
        L[i] = c[i] + t;
        if (PAD) L[i] |= padmask[i];
        m = __vmins2(m, L[i]);
    }
    m = __vmins2(m, __byte_perm(m, m, 0x1032));
    m = (uint32_t)__reduce_min_sync(0xffffffffu, (int)m);
#pragma unroll
    for (int i = 0; i < NP; i++) {
        T[i] = L[i] - m;
        if (PAD) T[i] |= padmask[i];
    }
}

// ---- spin-wait guard and mbarrier helpers (fused vertical sweep and the bulk-copy pipeline of the horizontal scans) ----
constexpr int HO_SLOTS = 4;
// a hand-over / neighbour wait longer than this is a lost strip: flag it, do not hang.  2 s by default; B2S_WAIT_TIMEOUT_S=<seconds>
// raises it (compute-sanitizer slows the kernels down a thousandfold).  One copy per translation unit, set by wait_timeout_init().
__device__ unsigned long long b2s_wait_timeout_ns = 2000000000ull;
inline cudaError_t wait_timeout_init()
{
    static std::once_flag once;
    static cudaError_t e = cudaSuccess;
    std::call_once(once, [] {
        const char *s = getenv("B2S_WAIT_TIMEOUT_S");
        if (s && atof(s) > 0) {
            const unsigned long long ns = (unsigned long long)(atof(s) * 1e9);
            int ndev = 0;
            cudaGetDeviceCount(&ndev);
            int cur = 0;
            cudaGetDevice(&cur);
            for (int d = 0; d < ndev && e == cudaSuccess; d++) {
                cudaSetDevice(d);
                e = cudaMemcpyToSymbol(b2s_wait_timeout_ns, &ns, sizeof ns);
            }
            cudaSetDevice(cur);
        }
    });
    return e;
}
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
    return now - t0 > b2s_wait_timeout_ns;
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

// The same wait with the retry loop written out: a warp that waits for a neighbour re-issues TRYWAIT about every 66 cycles (the
// hardware's own suspend limit; ptxas drops the suspend-time hint on sm_100a), and every instruction of the retry path takes an
// issue slot from the warps that are not waiting.  The compiler's loop costs 8 instructions per retry (it rebuilds the operands),
// this one 2.75: four attempts per trip, the time-out check (wait_expired) only every 1024 attempts.
__device__ __noinline__ void mbar_wait_retry(uint32_t addr, uint32_t parity, int *err)
{
    uint32_t ok;
    unsigned long long t0 = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .u32 n;\n"
            "mov.u32 n, 256;\n"
            "B2S_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "@p bra B2S_DONE;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "@p bra B2S_DONE;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "@p bra B2S_DONE;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "@p bra B2S_DONE;\n"
            "sub.u32 n, n, 1;\n"
            "setp.ne.u32 p, n, 0;\n"
            "@p bra B2S_WAIT;\n" // (falls through with p false: no success yet)
            "B2S_DONE:\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) break;
        spins += 256;
        if (wait_expired(spins, t0, err)) {
            *(volatile int *)err = 1;
            break;
        }
    }
}
// fast path inline (one TRYWAIT: the phase has usually completed), everything else out of line
__device__ __forceinline__ void mbar_wait_tight(uint32_t addr, uint32_t parity, int *err)
{
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (!ok) mbar_wait_retry(addr, parity, err);
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
