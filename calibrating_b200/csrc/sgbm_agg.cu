// sgbm_agg.cu -- semi-global path aggregation S(p,d) = sat16(sum_r L_r(p,d))  (SURVEY.md Appendix A.4), sm_100a.
//
// Replaces the aggregation loops of cv2.StereoSGBM.compute (calibrating/stereo_matching.py:63).
//   L_r(p,d) = C(p,d) + min(L_r(p-r,d), L_r(p-r,d-1)+P1, L_r(p-r,d+1)+P1, minL_r(p-r)+P2) - minL_r(p-r)
// Every direction r is an independent family of scan lines, and the saturating sum over directions is
// order-independent for non-negative L, so each direction is swept by one "scan" kernel:
//   * one warp = one scan line; a lane owns 2*NP consecutive disparities as NP packed int16x2 registers
//   * state kept normalised (L - minL), so a step is  VIMNMX3.S16x2 / VIADDMNMX.S16x2 / VIADD.16x2 (DPX) per register,
//     two SHFLs for the d-1 / d+1 neighbours across lanes and one CREDUX.MIN for minL
//   * the C (and S) chunk of a pixel is 128*NP contiguous bytes; each warp streams its chunks through a private
//     cp.async (LDGSTS) ring in shared memory, STAGES deep, so ~8 KB per warp are in flight and the sweep is HBM-bound
//   * horizontal lines are image rows; vertical and diagonal lines are indexed by their column at the first row and
//     wrap around the image edge with a state reset (an out-of-image predecessor is L = 0, minL = 0), so all lines
//     have the same length and every row of C is read as one contiguous span by neighbouring warps.
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

template <int NP> struct Stages { static constexpr int value = (NP <= 2) ? 16 : (NP == 3 ? 10 : 8); };

struct AggArgs {
    const int16_t *C;
    int16_t *S;
    int H, width1, D;
    int mx, my;
    int P1, P2;
};

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

template <int NP, bool PAD, int MODE>
__global__ void __launch_bounds__(WARPS * 32) agg_scan_kernel(AggArgs a)
{
    constexpr int CH = 128 * NP;                       // bytes of one pixel's d-chunk
    constexpr int NSRC = (MODE == AGG_ACCUM) ? 2 : 1;  // C only, or C and S
    constexpr int STAGE_BYTES = CH * NSRC;
    constexpr int STAGES = Stages<NP>::value;
    constexpr int NSEG = STAGE_BYTES / 16;
    extern __shared__ __align__(16) unsigned char smem[];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int line = blockIdx.x * WARPS + wid;
    const int nlines = a.my == 0 ? a.H : a.width1;
    const int nsteps = a.my == 0 ? a.width1 : a.H;
    if (line >= nlines) return;
    unsigned char *ring = smem + (size_t)wid * STAGES * STAGE_BYTES;
    const int Dp = 64 * NP;

    // positions: (x,y) of the compute cursor and of the prefetch cursor
    int x, y;
    if (a.my == 0) { y = line; x = a.mx > 0 ? 0 : a.width1 - 1; }
    else { x = line; y = a.my > 0 ? 0 : a.H - 1; }
    int px = x, py = y;
    auto advance = [&](int &cx, int &cy) {
        cx += a.mx;
        cy += a.my;
        if (cx < 0) cx = a.width1 - 1;
        else if (cx >= a.width1) cx = 0;
    };
    auto issue = [&](int stage) {
        size_t off = ((size_t)py * a.width1 + px) * Dp;
        unsigned char *dst = ring + stage * STAGE_BYTES;
#pragma unroll
        for (int seg = lane; seg < NSEG; seg += 32) {
            const int16_t *src = (NSRC == 1 || seg < CH / 16) ? a.C + off + seg * 8 : a.S + off + (seg - CH / 16) * 8;
            cp_async16(dst + seg * 16, src);
        }
        advance(px, py);
    };

    uint32_t padmask[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        int d0 = (lane * NP + i) * 2;
        padmask[i] = PAD ? ((d0 >= a.D ? 0x00007FFFu : 0u) | (d0 + 1 >= a.D ? 0x7FFF0000u : 0u)) : 0u;
    }
    const uint32_t P1v = (uint32_t)a.P1 * 0x10001u, P2mP1v = (uint32_t)(a.P2 - a.P1) * 0x10001u;
    const uint32_t BIG = 0x7FFF7FFFu;

#pragma unroll 1
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nsteps) issue(s);
        cp_async_commit();
    }

    uint32_t Ln[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) Ln[i] = padmask[i];

#pragma unroll 1
    for (int k = 0; k < nsteps; k++) {
        __syncwarp();
        if (k + STAGES - 1 < nsteps) issue((k + STAGES - 1) % STAGES);
        cp_async_commit();
        cp_async_wait<STAGES - 1>();
        __syncwarp();

        const uint32_t *st = (const uint32_t *)(ring + (k % STAGES) * STAGE_BYTES);
        uint32_t c[NP], sv[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) c[i] = st[lane * NP + i];
        if (MODE == AGG_ACCUM) {
#pragma unroll
            for (int i = 0; i < NP; i++) sv[i] = st[CH / 4 + lane * NP + i];
        }
        // out-of-image predecessor: L = 0, minL = 0
        bool reset = (k == 0) || (a.my != 0 && ((a.mx > 0 && x == 0) || (a.mx < 0 && x == a.width1 - 1)));
        if (reset) {
#pragma unroll
            for (int i = 0; i < NP; i++) Ln[i] = padmask[i];
        }
        // neighbours d-1 / d+1 (packed): lane boundary values come from the adjacent lanes
        uint32_t up = __shfl_up_sync(0xffffffffu, Ln[NP - 1], 1);
        uint32_t dn = __shfl_down_sync(0xffffffffu, Ln[0], 1);
        if (lane == 0) up = BIG;
        if (lane == 31) dn = BIG;
        uint32_t L[NP];
        uint32_t m = BIG;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            uint32_t lft = __byte_perm(i == 0 ? up : Ln[i - 1], Ln[i], 0x5432);            // L(d-1)
            uint32_t rgt = __byte_perm(Ln[i], i == NP - 1 ? dn : Ln[i + 1], 0x5432);        // L(d+1)
            uint32_t t = __vimin3_s16x2(lft, rgt, P2mP1v);                                  // min(L(d-1), L(d+1), P2-P1)
            t = __viaddmin_s16x2(t, P1v, Ln[i]);                                            // min(. + P1, L(d))  (<= P2)
            L[i] = __vadd2(c[i], t) | padmask[i];                                           // + C, int16 wrap like the cast
            m = __vmins2(m, L[i]);
        }
        int mn = min((int)(short)(m & 0xffffu), ((int)m) >> 16);
        mn = __reduce_min_sync(0xffffffffu, mn);
        const uint32_t negmin = ((uint32_t)(-mn) & 0xffffu) * 0x10001u;
        uint32_t out[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) {
            Ln[i] = __vadd2(L[i], negmin) | padmask[i];
            out[i] = (MODE == AGG_ACCUM) ? __viaddmin_u16x2(sv[i], L[i], BIG) : L[i]; // saturating accumulate (L >= 0)
        }
        store_regs<NP>(a.S + ((size_t)y * a.width1 + x) * Dp + lane * 2 * NP, out);
        advance(x, y);
    }
}

template <int NP, bool PAD, int MODE> cudaError_t launch_scan(b2s_ctx *c, const AggArgs &a)
{
    constexpr int STAGE_BYTES = 128 * NP * (MODE == AGG_ACCUM ? 2 : 1);
    size_t smem = (size_t)WARPS * Stages<NP>::value * STAGE_BYTES;
    static bool configured = false; // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(agg_scan_kernel<NP, PAD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    int nlines = a.my == 0 ? a.H : a.width1;
    agg_scan_kernel<NP, PAD, MODE><<<(nlines + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(a);
    c->launches++;
    return cudaGetLastError();
}

template <int NP, bool PAD> cudaError_t launch_dir(b2s_ctx *c, const AggArgs &a, int mode)
{
    return mode == AGG_INIT ? launch_scan<NP, PAD, AGG_INIT>(c, a) : launch_scan<NP, PAD, AGG_ACCUM>(c, a);
}

cudaError_t launch_dir_np(b2s_ctx *c, const AggArgs &a, int mode)
{
    const bool pad = c->g.D != c->g.Dp;
    switch (c->g.NP) {
    case 1: return pad ? launch_dir<1, true>(c, a, mode) : launch_dir<1, false>(c, a, mode);
    case 2: return pad ? launch_dir<2, true>(c, a, mode) : launch_dir<2, false>(c, a, mode);
    case 3: return pad ? launch_dir<3, true>(c, a, mode) : launch_dir<3, false>(c, a, mode);
    case 4: return pad ? launch_dir<4, true>(c, a, mode) : launch_dir<4, false>(c, a, mode);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace

cudaError_t agg_configure() { return cudaSuccess; }

// Directions as (mx,my) of the MOVE along the path (predecessor = p - move).  cv2 pass 1: (+1,0) (+1,+1) (0,+1) (-1,+1);
// MODE_SGBM adds (-1,0) during the WTA sweep; MODE_HH pass 2 adds (-1,0) (-1,-1) (0,-1) (+1,-1).
cudaError_t launch_aggregate(b2s_ctx *c, int *n_launches)
{
    static const int dirs8[8][2] = {{1, 0}, {-1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, -1}, {0, -1}, {1, -1}};
    const SgbmGeom &g = c->g;
    int nd = g.mode == 1 ? 8 : 5;
    AggArgs a;
    a.C = c->C.as<int16_t>();
    a.S = c->S.as<int16_t>();
    a.H = g.H; a.width1 = g.width1; a.D = g.D; a.P1 = g.P1; a.P2 = g.P2;
    for (int i = 0; i < nd; i++) {
        a.mx = dirs8[i][0];
        a.my = dirs8[i][1];
        cudaError_t e = launch_dir_np(c, a, i == 0 ? AGG_INIT : AGG_ACCUM);
        if (e != cudaSuccess) return e;
    }
    if (n_launches) *n_launches = nd;
    return cudaSuccess;
}
