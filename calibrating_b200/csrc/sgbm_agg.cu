// sgbm_agg.cu -- semi-global path aggregation S(p,d) = sat16(sum_r L_r(p,d))  (SURVEY.md Appendix A.4), sm_100a.
//
// Replaces the aggregation loops of cv2.StereoSGBM.compute (calibrating/stereo_matching.py:63).
//   L_r(p,d) = C(p,d) + min(L_r(p-r,d), L_r(p-r,d-1)+P1, L_r(p-r,d+1)+P1, minL_r(p-r)+P2) - minL_r(p-r)
// The saturating sum over directions is order-independent for non-negative L, so the eight (MODE_HH) or five (MODE_SGBM)
// directions are grouped by what they need rather than by OpenCV's two sweeps:
//   agg_hscan_vsum_kernel       the horizontal path (+1,0): 1080 independent image rows, S = L; forms C on the way from the cost
//                               stage's row sums (the vertical half of the box filter; block sizes <= 5), or
//   agg_hscan_kernel<INIT>      the same path on a finished C
//   agg_vsweep_kernel           every path that crosses rows, in ONE launch: the three top-down paths add to S, the three
//                               bottom-up paths (MODE_HH) go to S2; lock-step strips of columns, see the kernel's comment
//   agg_hscan_kernel<ACCUM[2]>  the horizontal path (-1,0) last: S = sat(S [+ S2] + L), with the winner-take-all fused (default:
//                               S is then never written) or followed by wta_kernel (sgbm_post.cu)
//   agg_scan_kernel             generic one-direction-per-launch scan (any direction; diagonal lines wrap around the image
//                               edge with a state reset): the two vertical paths of MODE_HH4, the legacy schedule for strips
//                               wider than 32 columns, and the cross-check of the tests (B2S_AGG_LEGACY=1)
// Common to all: one warp = one line (or column), a lane owns 2*NP consecutive disparities as NP packed int16x2 registers;
// state kept normalised (L - minL), so a step (sgm_step) is VIMNMX3.S16x2 / VIADDMNMX.S16x2 / VIADD.16x2 (DPX) per register,
// two SHFLs for the d-1 / d+1 neighbours across lanes and one CREDUX.MIN for minL; the C (and S) chunk of a pixel is 128*NP
// contiguous bytes.  The horizontal scans fetch several pixels per bulk copy (UBLKCP + mbarrier), the kernels that move across
// rows stream single pixels through a private cp.async (LDGSTS) ring per warp.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <set>
#include <utility>
#include <vector>
#include <type_traits>

#include "b2s_internal.h"

#include "sgm_common.cuh"

namespace {

struct AggArgs {
    const int16_t *C;
    int16_t *S;
    const int16_t *S2; // AGG_ACCUM2: second partial sum added on the fly (the bottom-up sweep's)
    int H, width1, D;
    int mx, my;
    int P1, P2;
    // fused winner-take-all of the last scan (agg_hscan_kernel, WTA != 0)
    int16_t *raw;
    unsigned *disp2key;
    int W, minX1, minD, uniq;
    const int16_t *hs; // agg_hscan_vsum_kernel: row sums of the per-pixel cost (H, width1, Dp), C = vertical box sum of them
    int SH2;
    int *err;        // device error flag (sgbm_agg.cu: wait_expired), set if a bulk copy never completes
    unsigned uniq_M; // ceil(2^32 / (100 - uniq)): the fused WTA's division by the invariant (100 - uniq) (0 < 100 - uniq <= 100)
};

template <int NP, bool PAD, int MODE>
__global__ void __launch_bounds__(WARPS * 32) agg_scan_kernel(AggArgs a)
{
    constexpr int CH = 128 * NP;                       // bytes of one pixel's d-chunk
    constexpr int NSRC = MODE == AGG_ACCUM2 ? 3 : (MODE == AGG_ACCUM ? 2 : 1); // C only; C and S; C, S and S2
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
            const int which = NSRC == 1 ? 0 : seg / (CH / 16), r = seg - which * (CH / 16);
            const int16_t *src = (which == 0 ? a.C : (which == 1 ? a.S : a.S2)) + off + r * 8;
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
        if (MODE != AGG_INIT) {
#pragma unroll
            for (int i = 0; i < NP; i++) sv[i] = st[CH / 4 + lane * NP + i];
        }
        if (MODE == AGG_ACCUM2) {
#pragma unroll
            for (int i = 0; i < NP; i++) sv[i] = __viaddmin_u16x2(sv[i], st[2 * (CH / 4) + lane * NP + i], BIG);
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
            out[i] = (MODE != AGG_INIT) ? __viaddmin_u16x2(sv[i], L[i], BIG) : L[i]; // saturating accumulate (L >= 0)
        }
        store_regs<NP>(a.S + ((size_t)y * a.width1 + x) * Dp + lane * 2 * NP, out);
        advance(x, y);
    }
}

// Horizontal lines only (my == 0), the two scans of the production schedule: same arithmetic as agg_scan_kernel, but the
// per-step overhead is cut to the bone.  A scan line is ONE warp's in-order instruction stream of 1792 dependent steps and an
// SM holds only 8 such warps, so the kernel's duration is (instructions per step) x (~4 cycles): the C / S / S2 chunks of
// HsChunk<NP>::px consecutive pixels are contiguous in memory and arrive by ONE bulk copy per source (cp.async.bulk -> UBLKCP, issued
// by lane 0, completion counted on an mbarrier per ring slot), which leaves a step with its LDS, the arithmetic and a store.
// WTA: 0 = store S; 1 = winner-take-all fused (A.5), S not stored; 2 = both (B2S_OPT_KEEP_VOLUMES: S stays fetchable)
template <int NP> struct HsChunk { static constexpr int px = NP == 1 ? 16 : (NP == 2 ? 8 : 4); }; // pixels per bulk copy (<= 2 KB per source)
constexpr int HS_SLOTS = 4;                                                                       // ring slots (chunks) per warp
// BLK: the volumes are in the block layout (SgbmGeom::layout 1: word = disparities (16b+j, 16b+8+j)) instead of pairs (2q, 2q+1)
template <int NP, bool PAD, int MODE, int WTA, bool BLK = false>
__global__ void __launch_bounds__(WARPS * 32) agg_hscan_kernel(AggArgs a)
{
    constexpr int CH = 128 * NP;                       // bytes of one pixel's d-chunk
    constexpr int NSRC = MODE == AGG_ACCUM2 ? 3 : (MODE == AGG_ACCUM ? 2 : 1); // C only; C and S; C, S and S2
    constexpr int CPX = HsChunk<NP>::px;
    constexpr int SRC_BYTES = CPX * CH;                // one source's part of a slot
    constexpr int SLOT_BYTES = NSRC * SRC_BYTES;
    constexpr int RING_BYTES = HS_SLOTS * SLOT_BYTES;  // per warp
    extern __shared__ __align__(16) unsigned char smem[];
    // shared memory: rings [WARPS][HS_SLOTS][NSRC][CPX][CH], WTA exchange buffers [WARPS][2][CH], mbarriers [WARPS][HS_SLOTS]
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t mb0 = sm0 + WARPS * RING_BYTES + (WTA != 0 ? WARPS * 2 * CH : 0);
    if (threadIdx.x < WARPS * HS_SLOTS) mbar_init(mb0 + threadIdx.x * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int y = blockIdx.x * WARPS + wid;
    if (y >= a.H) return;
    const int Dp = 64 * NP, nsteps = a.width1;
    const int nchunks = (nsteps + CPX - 1) / CPX;
    const uint32_t ring = sm0 + wid * RING_BYTES, mbw = mb0 + wid * (HS_SLOTS * 8);
    const long long stepE = (long long)a.mx * Dp; // int16 elements to the next pixel of the line
    const long long row0 = (long long)y * a.width1 * Dp;
    const long long o0 = row0 + (long long)(a.mx > 0 ? 0 : a.width1 - 1) * Dp;
    // chunk j = steps [j*CPX, j*CPX + npx): pixels x0 .. x0 + npx - 1 in memory order (x0 = first step's x for +x, last step's for -x)
    auto issue = [&](int j) {
        if (lane == 0) {
            const int k0 = j * CPX, npx = min(CPX, nsteps - k0);
            const int x0 = a.mx > 0 ? k0 : a.width1 - k0 - npx;
            const uint32_t slot = ring + (j % HS_SLOTS) * SLOT_BYTES, mb = mbw + (j % HS_SLOTS) * 8, bytes = (uint32_t)npx * CH;
            const long long o = row0 + (long long)x0 * Dp;
            mbar_arrive_expect_tx(mb, bytes * NSRC);
            bulk_g2s(slot, a.C + o, bytes, mb);
            if (NSRC > 1) bulk_g2s(slot + SRC_BYTES, a.S + o, bytes, mb);
            if (NSRC > 2) bulk_g2s(slot + 2 * SRC_BYTES, a.S2 + o, bytes, mb);
        }
    };
    uint32_t padmask[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const int d0 = b2s_word_d0(BLK ? 1 : 0, lane * NP + i), d1 = d0 + (BLK ? 8 : 1);
        padmask[i] = PAD ? ((d0 >= a.D ? 0x00007FFFu : 0u) | (d1 >= a.D ? 0x7FFF0000u : 0u)) : 0u;
    }
    const uint32_t P1v = (uint32_t)a.P1 * 0x10001u, P2mP1v = (uint32_t)(a.P2 - a.P1) * 0x10001u;
    const uint32_t BIG = 0x7FFF7FFFu;
#pragma unroll 1
    for (int j = 0; j < HS_SLOTS - 1 && j < nchunks; j++) issue(j);
    uint32_t T[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) T[i] = padmask[i];
    int16_t *sp = a.S + o0 + lane * 2 * NP;
    const int xfirst = a.mx > 0 ? 0 : a.width1 - 1;
    // ---- fused winner-take-all (WTA != 0): A.5 of SURVEY.md, same rule as wta_kernel in sgbm_post.cu ----
    // per-warp exchange buffer (two pixels) behind the rings: the sub-pixel fit needs S(best-1), S(best+1) of other lanes
    const uint32_t xbase = sm0 + WARPS * RING_BYTES + wid * (2 * CH);
    const uint32_t xbuf = xbase + lane * NP * 4;
    uint32_t dd[NP], prev[NP], kmin = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const uint32_t d0 = (uint32_t)b2s_word_d0(BLK ? 1 : 0, lane * NP + i);
        dd[i] = d0 | ((d0 + (BLK ? 8 : 1)) << 16);
        prev[i] = 0;
    }
    const int thr = 100 - a.uniq;
    int my_best = -2, my_minS = 0, my_sm = 0, my_sp = 0; // lane i keeps the winner of step (chunk of 32) + i; -2 = rejected
    auto wta_flush = [&](int kp) { // kp = last step of the chunk: one pixel per lane (right-view candidate, sub-pixel fit, store)
        const int kl = (kp & ~31) + lane;
        if (kl <= kp && my_best != -2) {
            const int x = xfirst + a.mx * kl;
            int d = my_best;
            const int x2 = x + a.minX1 - d - a.minD;
            if (my_minS < 32767 && x2 >= 0 && x2 < a.W + 2)
                atomicMin(&a.disp2key[(size_t)y * (a.W + 2) + x2], ((unsigned)my_minS << 16) | (unsigned)(0xFFFF - x));
            if (0 < d && d < a.D - 1) {
                const int den2 = max(my_sm + my_sp - 2 * my_minS, 1);
                d = d * 16 + ((my_sm - my_sp) * 16 + den2) / (den2 * 2);
            } else
                d *= 16;
            a.raw[(size_t)y * a.W + x + a.minX1] = (int16_t)(d + a.minD * 16);
        }
        my_best = -2;
    };
    // second half of the WTA of step kp (prev = its S, kmin = its smallest key); branch-free so that it stays in the basic block
    // of the step and ptxas can interleave it with the next step's dependent chain.  kp = -1 (first iteration) is harmless:
    // lane 31 records a dummy that step 31 overwrites before the first flush.
    auto wta_finish = [&](int kp) {
        const int minS = (int)(kmin >> 16);
        const int best = minS >= 32767 ? -1 : (int)(kmin & 0xffffu); // cv2: strict '<' against MAX_COST never fires
        // uniqueness: reject if some d outside best-1..best+1 has S(d) * (100 - uniq) < minS * 100, i.e. S(d) < ceil(minS * 100 / thr)
        const unsigned n = (unsigned)(minS * 100 + thr - 1);
        const unsigned T = min(thr == 1 ? n : __umulhi(n, a.uniq_M), 32768u);
        const uint32_t Tkey = T << 16;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const int e = b2s_word_d0(BLK ? 1 : 0, lane * NP + i) - best; // d - best of the low half; the high half is (BLK ? 8 : 1) further
            // (the low 16 bits do not matter against a multiple of 65536; padded halves, d >= D, are not candidates)
            bad |= ((prev[i] << 16) < Tkey) && ((unsigned)(e + 1) > 2u) && !(PAD && (padmask[i] & 0xFFFFu));
            bad |= (prev[i] < Tkey) && ((unsigned)(e + (BLK ? 8 : 1) + 1) > 2u) && !(PAD && (padmask[i] >> 16));
        }
        // S(best-1), S(best+1) for the sub-pixel fit: every lane reads the same two halves (broadcast), the recording lane keeps them
        const int bc = min(max(best, 1), a.D - 2);
        const uint32_t nb = xbase + (kp & 1) * CH;
        unsigned short lo, hi;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(lo) : "r"(nb + (uint32_t)b2s_dindex(BLK ? 1 : 0, bc - 1) * 2) : "memory");
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hi) : "r"(nb + (uint32_t)b2s_dindex(BLK ? 1 : 0, bc + 1) * 2) : "memory");
        const bool rej = __any_sync(0xffffffffu, bad);
        const bool rec = lane == (kp & 31);
        my_best = rec ? (rej ? -2 : best) : my_best;
        my_minS = rec ? minS : my_minS;
        my_sm = rec ? (int)(short)lo : my_sm;
        my_sp = rec ? (int)(short)hi : my_sp;
    };
    int k = 0; // step index along the line
#pragma unroll 1
    for (int j = 0; j < nchunks; j++) {
        __syncwarp(); // every lane is done with the slot of chunk j-1: it is refilled now (generic reads before an async-proxy write)
        if (j + HS_SLOTS - 1 < nchunks) issue(j + HS_SLOTS - 1);
        mbar_wait(mbw + (j % HS_SLOTS) * 8, (uint32_t)(j / HS_SLOTS) & 1u, a.err);
        const int npx = min(CPX, nsteps - j * CPX);
        // shared address of this lane's part of the current pixel, walking the chunk in scan order
        uint32_t cur = ring + (j % HS_SLOTS) * SLOT_BYTES + lane * NP * 4 + (a.mx > 0 ? 0 : (npx - 1) * CH);
        const int curstep = a.mx > 0 ? CH : -CH;
#pragma unroll 1
        for (int s = 0; s < npx; s++, k++) {
            uint32_t c[NP], sv[NP], L[NP], out[NP];
            lds_s<NP>(cur, c);
            if (MODE != AGG_INIT) lds_s<NP>(cur + SRC_BYTES, sv);
            if (MODE == AGG_ACCUM2) {
                uint32_t s2[NP];
                lds_s<NP>(cur + 2 * SRC_BYTES, s2);
#pragma unroll
                for (int i = 0; i < NP; i++) sv[i] = __viaddmin_u16x2(sv[i], s2[i], BIG);
            }
            if constexpr (BLK) sgm_step_b32<NP, PAD>(T, c, L, padmask, P1v, P2mP1v, lane);
            else sgm_step<NP, PAD>(T, c, L, padmask, P1v, P2mP1v, lane);
#pragma unroll
            for (int i = 0; i < NP; i++) out[i] = (MODE != AGG_INIT) ? __viaddmin_u16x2(sv[i], L[i], BIG) : L[i];
            if (WTA != 1) stcg_regs<NP>(sp, out);
            if (WTA != 0) {
                // winner-take-all of the pixel just finished, software-pipelined by one step so that its warp reduction, vote and
                // shared-memory round trip overlap the next step's dependent chain instead of extending this one
                __syncwarp(); // the exchange buffer written in the previous step is read by other lanes below
                wta_finish(k - 1);
                sts_s<NP>(xbuf + (k & 1) * CH, out);
                uint32_t key = 0xFFFFFFFFu;
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    // keys (S << 16) | d of the two halves; a padded lane (d >= D) holds S = 0x7FFF and a larger d, so it never wins
                    key = min(key, min(__byte_perm(out[i], dd[i], 0x1054), __byte_perm(out[i], dd[i], 0x3276)));
                    prev[i] = out[i];
                }
                kmin = __reduce_min_sync(0xffffffffu, key);
                if ((k & 31) == 0 && k > 0) wta_flush(k - 1); // steps k-32 .. k-1 are complete
            }
            sp += stepE;
            cur += curstep;
        }
    }
    if (WTA != 0 && nsteps > 0) {
        __syncwarp();
        wta_finish(nsteps - 1);
        wta_flush(nsteps - 1);
    }
}

// The first horizontal scan with the vertical half of the cost's box filter folded in (block sizes up to 5): the cost stage
// leaves the horizontal window sums hs of every row (sgbm_cost.cu), and this kernel forms C(y) = sum_k hs(clamp(y+k)) for
// its row on the fly (NV = 2*SH2+1 bulk copies per chunk, VIADD.16x2 like vsum_kernel), stores it for the later passes and
// scans it.  That removes vsum_kernel's launch and the C read of this scan: 2 B/voxel less DRAM traffic per pair; the NV-fold
// re-read of hs rows by neighbouring warps is served by L2.
template <int NP> struct HvChunk { static constexpr int px = NP == 1 ? 12 : (NP == 2 ? 6 : (NP == 3 ? 4 : 3)); }; // 1.5 KB per hs row
constexpr int HV_SLOTS = 3;
template <int NP, bool PAD, int NV, bool BLK = false>
__global__ void __launch_bounds__(WARPS * 32) agg_hscan_vsum_kernel(AggArgs a)
{
    constexpr int CH = 128 * NP;
    constexpr int CPX = HvChunk<NP>::px;
    constexpr int SRC_BYTES = CPX * CH;
    constexpr int SLOT_BYTES = NV * SRC_BYTES;
    constexpr int RING_BYTES = HV_SLOTS * SLOT_BYTES;
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t mb0 = sm0 + WARPS * RING_BYTES;
    if (threadIdx.x < WARPS * HV_SLOTS) mbar_init(mb0 + threadIdx.x * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int y = blockIdx.x * WARPS + wid;
    if (y >= a.H) return;
    const int Dp = 64 * NP, nsteps = a.width1;
    const int nchunks = (nsteps + CPX - 1) / CPX;
    const uint32_t ring = sm0 + wid * RING_BYTES, mbw = mb0 + wid * (HV_SLOTS * 8);
    const long long rowE = (long long)a.width1 * Dp;
    const int16_t *hrow[NV]; // the NV rows of the window, replicated at the image border (A.3)
#pragma unroll
    for (int v = 0; v < NV; v++) hrow[v] = a.hs + (long long)min(max(y + v - NV / 2, 0), a.H - 1) * rowE;
    auto issue = [&](int j) {
        if (lane == 0) {
            const int k0 = j * CPX, npx = min(CPX, nsteps - k0);
            const uint32_t slot = ring + (j % HV_SLOTS) * SLOT_BYTES, mb = mbw + (j % HV_SLOTS) * 8, bytes = (uint32_t)npx * CH;
            mbar_arrive_expect_tx(mb, bytes * NV);
#pragma unroll
            for (int v = 0; v < NV; v++) bulk_g2s(slot + v * SRC_BYTES, hrow[v] + (long long)k0 * Dp, bytes, mb);
        }
    };
    uint32_t padmask[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        const int d0 = b2s_word_d0(BLK ? 1 : 0, lane * NP + i), d1 = d0 + (BLK ? 8 : 1);
        padmask[i] = PAD ? ((d0 >= a.D ? 0x00007FFFu : 0u) | (d1 >= a.D ? 0x7FFF0000u : 0u)) : 0u;
    }
    const uint32_t P1v = (uint32_t)a.P1 * 0x10001u, P2mP1v = (uint32_t)(a.P2 - a.P1) * 0x10001u;
#pragma unroll 1
    for (int j = 0; j < HV_SLOTS - 1 && j < nchunks; j++) issue(j);
    uint32_t T[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) T[i] = padmask[i];
    const long long o0 = (long long)y * rowE + lane * 2 * NP;
    int16_t *sp = a.S + o0, *cp = const_cast<int16_t *>(a.C) + o0;
    uint32_t ovf = 0;
#pragma unroll 1
    for (int j = 0; j < nchunks; j++) {
        __syncwarp(); // every lane is done with the slot of chunk j-1: it is refilled now
        if (j + HV_SLOTS - 1 < nchunks) issue(j + HV_SLOTS - 1);
        mbar_wait(mbw + (j % HV_SLOTS) * 8, (uint32_t)(j / HV_SLOTS) & 1u, a.err);
        const int npx = min(CPX, nsteps - j * CPX);
        uint32_t cur = ring + (j % HV_SLOTS) * SLOT_BYTES + lane * NP * 4;
#pragma unroll 1
        for (int s = 0; s < npx; s++) {
            uint32_t c[NP], L[NP];
            lds_s<NP>(cur, c);
#pragma unroll
            for (int v = 1; v < NV; v++) {
                uint32_t h[NP];
                lds_s<NP>(cur + v * SRC_BYTES, h);
#pragma unroll
                for (int i = 0; i < NP; i++) c[i] = __vadd2(c[i], h[i]);
            }
            stcg_regs<NP>(cp, c);
#pragma unroll
            for (int i = 0; i < NP; i++) ovf |= c[i]; // a wrapped block sum (C < 0) has bit 15 of its half set
            if constexpr (BLK) sgm_step_b32<NP, PAD>(T, c, L, padmask, P1v, P2mP1v, lane);
            else sgm_step<NP, PAD>(T, c, L, padmask, P1v, P2mP1v, lane);
            stcg_regs<NP>(sp, L);
            sp += Dp;
            cp += Dp;
            cur += CH;
        }
    }
    if (ovf & 0x80008000u) a.err[1] = 1;
}

// the generic one-direction scan (any direction; the only aggregation kernel instantiated for more than 256 disparities)
template <int NP, bool PAD, int MODE> cudaError_t launch_generic(b2s_ctx *c, const AggArgs &a)
{
    constexpr int STAGE_BYTES = 128 * NP * (MODE == AGG_ACCUM2 ? 3 : (MODE == AGG_ACCUM ? 2 : 1));
    const int nlines = a.my == 0 ? a.H : a.width1;
    const size_t smem = (size_t)WARPS * Stages<NP>::value * STAGE_BYTES;
    static std::once_flag once[64];
    cudaError_t e = cudaSuccess;
    std::call_once(once[c->device & 63], [&] { e = cudaFuncSetAttribute(agg_scan_kernel<NP, PAD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    if (e != cudaSuccess) return e;
    agg_scan_kernel<NP, PAD, MODE><<<(nlines + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(a);
    c->launches++;
    return cudaGetLastError();
}
template <int NP, bool PAD> cudaError_t launch_dir_wide(b2s_ctx *c, const AggArgs &a, int mode, int wta)
{
    if (wta != 0 || mode == AGG_ACCUM2) return cudaErrorInvalidValue;
    return mode == AGG_INIT ? launch_generic<NP, PAD, AGG_INIT>(c, a) : launch_generic<NP, PAD, AGG_ACCUM>(c, a);
}

template <int NP, bool PAD, int MODE, int WTA> cudaError_t launch_scan(b2s_ctx *c, const AggArgs &a)
{
    constexpr int STAGE_BYTES = 128 * NP * (MODE == AGG_ACCUM2 ? 3 : (MODE == AGG_ACCUM ? 2 : 1));
    const int nlines = a.my == 0 ? a.H : a.width1;
    if (a.my == 0 && !c->agg_legacy) {
        // horizontal scans: bulk-copy rings + the fused WTA's exchange buffers + one mbarrier per ring slot
        const size_t smem_h = (size_t)WARPS * HS_SLOTS * (STAGE_BYTES * HsChunk<NP>::px) + (WTA != 0 ? (size_t)WARPS * 2 * 128 * NP : 0) + WARPS * HS_SLOTS * 8;
        static std::once_flag once_h[64][2]; // per instantiation and device (the attribute belongs to the device's context)
        cudaError_t e = cudaSuccess;
        if constexpr (NP == 2) {
            if (c->g.layout == 1) { // (the block layout reaches the scans only with NP = 2: the six-path sweep of agg_vsweep6_kernel)
                std::call_once(once_h[c->device & 63][1], [&] { e = cudaFuncSetAttribute(agg_hscan_kernel<NP, PAD, MODE, WTA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h); });
                if (e != cudaSuccess) return e;
                agg_hscan_kernel<NP, PAD, MODE, WTA, true><<<(nlines + WARPS - 1) / WARPS, WARPS * 32, smem_h, c->stream>>>(a);
                c->launches++;
                return cudaGetLastError();
            }
        }
        if (c->g.layout != 0) return cudaErrorInvalidValue;
        std::call_once(once_h[c->device & 63][0], [&] { e = cudaFuncSetAttribute(agg_hscan_kernel<NP, PAD, MODE, WTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h); });
        if (e != cudaSuccess) return e;
        agg_hscan_kernel<NP, PAD, MODE, WTA><<<(nlines + WARPS - 1) / WARPS, WARPS * 32, smem_h, c->stream>>>(a);
    } else if constexpr (WTA != 0) {
        return cudaErrorInvalidValue; // the generic scan has no fused winner-take-all (launch_aggregate never asks for one)
    } else {
        return launch_generic<NP, PAD, MODE>(c, a);
    }
    c->launches++;
    return cudaGetLastError();
}

template <int NP, bool PAD> cudaError_t launch_dir(b2s_ctx *c, const AggArgs &a, int mode, int wta)
{
    if (wta == 1) return mode == AGG_ACCUM2 ? launch_scan<NP, PAD, AGG_ACCUM2, 1>(c, a) : launch_scan<NP, PAD, AGG_ACCUM, 1>(c, a);
    if (wta == 2) return mode == AGG_ACCUM2 ? launch_scan<NP, PAD, AGG_ACCUM2, 2>(c, a) : launch_scan<NP, PAD, AGG_ACCUM, 2>(c, a);
    if (mode == AGG_ACCUM2) return launch_scan<NP, PAD, AGG_ACCUM2, 0>(c, a);
    return mode == AGG_INIT ? launch_scan<NP, PAD, AGG_INIT, 0>(c, a) : launch_scan<NP, PAD, AGG_ACCUM, 0>(c, a);
}

cudaError_t launch_dir_np(b2s_ctx *c, const AggArgs &a, int mode, int wta = 0)
{
    const bool pad = c->g.D != c->g.Dp;
    switch (c->g.NP) {
    case 1: return pad ? launch_dir<1, true>(c, a, mode, wta) : launch_dir<1, false>(c, a, mode, wta);
    case 2: return pad ? launch_dir<2, true>(c, a, mode, wta) : launch_dir<2, false>(c, a, mode, wta);
    case 3: return pad ? launch_dir<3, true>(c, a, mode, wta) : launch_dir<3, false>(c, a, mode, wta);
    case 4: return pad ? launch_dir<4, true>(c, a, mode, wta) : launch_dir<4, false>(c, a, mode, wta);
    // 257 .. 512 disparities: the generic per-direction scan only (launch_aggregate selects the legacy schedule)
    case 5: return pad ? launch_dir_wide<5, true>(c, a, mode, wta) : launch_dir_wide<5, false>(c, a, mode, wta);
    case 6: return pad ? launch_dir_wide<6, true>(c, a, mode, wta) : launch_dir_wide<6, false>(c, a, mode, wta);
    case 7: return pad ? launch_dir_wide<7, true>(c, a, mode, wta) : launch_dir_wide<7, false>(c, a, mode, wta);
    case 8: return pad ? launch_dir_wide<8, true>(c, a, mode, wta) : launch_dir_wide<8, false>(c, a, mode, wta);
    default: return cudaErrorInvalidValue;
    }
}

template <int NP, bool PAD, int NV> cudaError_t launch_hscan_vsum_t(b2s_ctx *c, const AggArgs &a)
{
    const size_t smem = (size_t)WARPS * HV_SLOTS * NV * HvChunk<NP>::px * 128 * NP + WARPS * HV_SLOTS * 8;
    static std::once_flag once[64][2]; // per instantiation and device (the attribute belongs to the device's context)
    cudaError_t e = cudaSuccess;
    if constexpr (NP == 2) {
        if (c->g.layout == 1) {
            std::call_once(once[c->device & 63][1], [&] { e = cudaFuncSetAttribute(agg_hscan_vsum_kernel<NP, PAD, NV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
            if (e != cudaSuccess) return e;
            agg_hscan_vsum_kernel<NP, PAD, NV, true><<<(a.H + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(a);
            c->launches++;
            return cudaGetLastError();
        }
    }
    if (c->g.layout != 0) return cudaErrorInvalidValue;
    std::call_once(once[c->device & 63][0], [&] { e = cudaFuncSetAttribute(agg_hscan_vsum_kernel<NP, PAD, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    if (e != cudaSuccess) return e;
    agg_hscan_vsum_kernel<NP, PAD, NV><<<(a.H + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(a);
    c->launches++;
    return cudaGetLastError();
}
template <int NP, bool PAD> cudaError_t launch_hscan_vsum_p(b2s_ctx *c, const AggArgs &a)
{
    switch (a.SH2) {
    case 0: return launch_hscan_vsum_t<NP, PAD, 1>(c, a);
    case 1: return launch_hscan_vsum_t<NP, PAD, 3>(c, a);
    case 2: return launch_hscan_vsum_t<NP, PAD, 5>(c, a);
    default: return cudaErrorInvalidValue;
    }
}
// the +x scan reading the cost stage's row sums (agg_fuses_vsum)
cudaError_t launch_hscan_vsum(b2s_ctx *c, const AggArgs &a)
{
    const bool pad = c->g.D != c->g.Dp;
    switch (c->g.NP) {
    case 1: return pad ? launch_hscan_vsum_p<1, true>(c, a) : launch_hscan_vsum_p<1, false>(c, a);
    case 2: return pad ? launch_hscan_vsum_p<2, true>(c, a) : launch_hscan_vsum_p<2, false>(c, a);
    case 3: return pad ? launch_hscan_vsum_p<3, true>(c, a) : launch_hscan_vsum_p<3, false>(c, a);
    case 4: return pad ? launch_hscan_vsum_p<4, true>(c, a) : launch_hscan_vsum_p<4, false>(c, a);
    default: return cudaErrorInvalidValue;
    }
}

// ================================================================================================================
// Fused vertical sweep: the three directions that cross image rows, for the top-down sweep (moves (+1,+1) (0,+1) (-1,+1))
// and -- MODE_HH -- the bottom-up sweep (moves (+1,-1) (0,-1) (-1,-1)) in ONE launch that reads C once per sweep
// (DESIGN.md section 4.1-4.2): the top-down sweep adds its three paths to S, the bottom-up sweep writes the sum of its
// three to S2, and the last horizontal scan folds S2 in.
//
//   * the image is cut into G <= #SM vertical strips of n columns; CTA = strip, warp = (column, sweep), lane = 2*NP disparities
//   * the vertical state stays in registers; the two diagonal states move one column per row: through a double-buffered
//     shared-memory slot and an mbarrier pair per warp between warps, and through a 4-deep ring in global memory between
//     neighbouring CTAs.  The ring carries no flags and needs no fences: states are normalised (L - minL, 15 bits), so
//     bit 15 of every int16 is a phase bit ((t>>2)&1) written with the data; the consumer lane polls its own words
//     until the phase matches.  A boundary is crossed by one state in each direction every row, so neighbouring CTAs
//     can never be more than one row apart (lock-step) and a 4-deep ring cannot be overrun.  All G CTAs must be
//     co-resident (G <= #SM, 1 CTA/SM).
//   * C (and S) rows stream through a private cp.async ring per warp, ~50 KB in flight per SM.
struct VsArgs {
    const int16_t *C;
    int16_t *S;   // top-down sweep: S += its three paths
    int16_t *S2;  // bottom-up sweep: S2 = sum of its three paths
    int H, width1, D, P1, P2;
    int n;        // columns (= warps) per CTA
    int up;       // JW == 1 only: 0 = top-down sweep, 1 = bottom-up sweep
    uint32_t *ho; // hand-over rings [J][G-1][2][HO_SLOTS][32*NP] u32
    int *err;     // set to 1 if a hand-over wait timed out (never in a correct run)
    long long *trace; // development aid (B2S_VS2_TRACE): clock64 at fixed points of 16 rows of three CTAs, see scripts/vs2_trace.py
    int dbg;      // what-if timing switches (B2S_VS2_FAKE; results are wrong when set): 1 no hand-over polling, 2 no neighbour waits, 4 late prefetch
};
template <int NP> __device__ __forceinline__ void lds_regs(const uint32_t *src, uint32_t (&v)[NP])
{
    if constexpr (NP == 2) { uint2 t = *(const uint2 *)src; v[0] = t.x; v[1] = t.y; }
    else if constexpr (NP == 4) { uint4 t = *(const uint4 *)src; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) v[i] = src[i];
    }
}
template <int NP> __device__ __forceinline__ void sts_regs(uint32_t *dst, const uint32_t (&v)[NP])
{
    if constexpr (NP == 2) *(uint2 *)dst = make_uint2(v[0], v[1]);
    else if constexpr (NP == 4) *(uint4 *)dst = make_uint4(v[0], v[1], v[2], v[3]);
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) dst[i] = v[i];
    }
}

// consumer side of the hand-over ring: ho_load issues one (non-blocking) load of this lane's words; ho_read takes the loaded
// words, and while any int16 of the warp does not carry the expected phase bit, loads again
template <int NP> __device__ __forceinline__ void ho_load(const uint32_t *p, uint32_t (&v)[NP])
{
    if constexpr (NP == 2) asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "l"(p) : "memory");
    else if constexpr (NP == 4)
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p) : "memory");
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[i]) : "l"(p + i) : "memory");
    }
}
template <int NP> __device__ __forceinline__ void ho_read(const uint32_t *p, uint32_t phase, uint32_t (&T)[NP], int *err, bool preloaded = false)
{
    uint32_t v[NP];
    int spins = 0;
    unsigned long long t0 = 0;
    if (preloaded) {
#pragma unroll
        for (int i = 0; i < NP; i++) v[i] = T[i];
    }
    while (true) {
        if (!preloaded) ho_load<NP>(p, v);
        preloaded = false;
        uint32_t bad = 0;
#pragma unroll
        for (int i = 0; i < NP; i++) bad |= (v[i] ^ phase) & 0x80008000u;
        if (__all_sync(0xffffffffu, bad == 0)) break;
        if (wait_expired(++spins, t0, err)) {
            *(volatile int *)err = 1;
            break;
        }
    }
#pragma unroll
    for (int i = 0; i < NP; i++) T[i] = v[i] & 0x7FFF7FFFu;
}
template <int NP> __device__ __forceinline__ void ho_write(uint32_t *p, uint32_t phase, const uint32_t (&T)[NP])
{
    uint32_t v[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) v[i] = (T[i] & 0x7FFF7FFFu) | phase;
    if constexpr (NP == 2) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v[0]), "r"(v[1]) : "memory");
    else if constexpr (NP == 4)
        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
    else {
#pragma unroll
        for (int i = 0; i < NP; i++) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p + i), "r"(v[i]) : "memory");
    }
}

// JW = sweeps handled by one CTA (2: warps [0,n) run the top-down sweep, warps [n,2n) the bottom-up sweep; 1: a.up selects),
// R = depth of the per-warp cp.async ring.  The top-down sweep adds its three paths to S; the bottom-up sweep (MODE_HH)
// writes the sum of its three paths to S2 without reading anything but C, so the two sweeps share no data at all.
// Warps synchronise only with their two neighbour columns (one mbarrier per warp, phase = row), so the warps of an SM
// drift apart by up to a row per column and keep the issue slots busy while others wait.
template <int NP, bool PAD, int JW, int R>
__global__ void __launch_bounds__(1024, 1) agg_vsweep_kernel(VsArgs a)
{
    constexpr int DW = 32 * NP;           // 32-bit words of one pixel's d-chunk
    constexpr int CHB = DW * 4;           // ... in bytes
    constexpr int HO_DIR = HO_SLOTS * DW; // words of one direction's hand-over ring
    constexpr int NSEG = 2 * DW / 4;      // 16-byte segments of one ring stage: [C | S]
    constexpr int NLD = (NSEG + 31) / 32; // cp.async instructions per lane and step
    // mbarriers [JW*n][2] (8 B each), slots [JW][2 parity][2 dir][n+2][DW], rings [JW*n][R][2*DW]
    extern __shared__ __align__(16) uint32_t vs_smem[];

    const int lane = threadIdx.x & 31;
    const int wi = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); // warp-uniform by construction
    const int n = a.n, G = gridDim.x, H = a.H, Dp = 64 * NP;
    const int jl = JW == 2 ? (wi >= n ? 1 : 0) : 0; // sweep slot inside the CTA
    const int w = wi - jl * n;                      // column inside the strip
    const bool up = JW == 2 ? jl == 1 : a.up != 0;  // bottom-up sweep?
    const bool acc = !up;                           // top-down: S += paths; bottom-up: S2 = paths
    const int x = blockIdx.x * n + w;
    const uint32_t BIG = 0x7FFF7FFFu;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(vs_smem);
    const uint32_t MBB = JW * n * 16; // two mbarriers per warp (even rows, odd rows)
    uint32_t *slots = vs_smem + MBB / 4;

    uint32_t padmask[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        int d0 = (lane * NP + i) * 2;
        padmask[i] = PAD ? ((d0 >= a.D ? 0x00007FFFu : 0u) | (d0 + 1 >= a.D ? 0x7FFF0000u : 0u)) : 0u;
    }
    // every diagonal slot starts as the out-of-image state; the two guard slots (index 0 and n+1) and the slots of idle
    // columns keep it for ever, which is what the first / last image column must read every row
    const int NSLOT = JW * 2 * 2 * (n + 2);
    for (int q = wi; q < NSLOT; q += blockDim.x >> 5) {
#pragma unroll
        for (int i = 0; i < NP; i++) slots[q * DW + lane * NP + i] = padmask[i];
    }
    if (threadIdx.x < JW * n * 2) mbar_init(sbase + threadIdx.x * 8, 1);
    __syncthreads();
    if (x >= a.width1) return; // idle warps of the last strip
    const bool first_col = x == 0, last_col = x == a.width1 - 1;
    const bool left_edge = w == 0, right_edge = w == n - 1;
    const bool out_right = right_edge && !last_col, out_left = left_edge && !first_col;
    const uint32_t P1v = (uint32_t)a.P1 * 0x10001u, P2mP1v = (uint32_t)(a.P2 - a.P1) * 0x10001u;

    // Processing order of the two diagonals.  dir 0 arrives from column x-1 (moves +1 in x), dir 1 from column x+1.  A warp
    // on a CTA boundary computes the diagonal it hands to the neighbour CTA FIRST and polls for the one it receives
    // SECOND, so that the hand-over latency overlaps the rest of the step on both sides.
    const int fd = (out_left && !out_right) ? 1 : 0;
    const uint32_t PSB = 2 * (n + 2) * CHB; // bytes between the two parities of the slots
    bool in_glob[2], out_glob[2];
    uint32_t in_s[2], out_s[2], in_mb[2]; // shared addresses: slot read / written (parity 0), mbarrier of the producing warp (0: none)
    const uint32_t *in_g[2];
    uint32_t *out_g[2];
    const int js = up ? 1 : 0; // hand-over rings are indexed by the sweep's direction
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int dir = k == 0 ? fd : fd ^ 1;
        in_glob[k] = dir == 0 ? (left_edge && !first_col) : (right_edge && !last_col);
        out_glob[k] = dir == 0 ? out_right : out_left;
        const uint32_t sl = sbase + MBB + ((jl * 2 * 2 + dir) * (n + 2)) * CHB + lane * NP * 4;
        in_s[k] = sl + (dir == 0 ? w : w + 2) * CHB;
        out_s[k] = sl + (w + 1) * CHB;
        const int wn = dir == 0 ? w - 1 : w + 1; // producing warp (column) inside the strip
        const bool has = wn >= 0 && wn < n && (int)blockIdx.x * n + wn < a.width1;
        in_mb[k] = has ? sbase + (jl * n + wn) * 16 : 0u;
        // ring of boundary b (between CTA b and b+1), direction dir: ho + ((js*(G-1) + b)*2 + dir) * HO_DIR
        const int b_in = dir == 0 ? (int)blockIdx.x - 1 : (int)blockIdx.x, b_out = dir == 0 ? (int)blockIdx.x : (int)blockIdx.x - 1;
        const size_t HJ = (size_t)(G > 1 ? G - 1 : 1) * 2 * HO_DIR;
        in_g[k] = a.ho + js * HJ + ((size_t)max(b_in, 0) * 2 + dir) * HO_DIR + lane * NP;
        out_g[k] = a.ho + js * HJ + ((size_t)max(b_out, 0) * 2 + dir) * HO_DIR + lane * NP;
    }
    const bool polls = in_glob[1];
    const uint32_t my_mb = sbase + (jl * n + w) * 16;

    // C (and S) rows stream through a private cp.async ring of R stages per warp (a register prefetch does not work: the
    // consumer waits on a scoreboard shared with the younger prefetches, which collapses the prefetch distance)
    const uint32_t ring = sbase + MBB + NSLOT * CHB + wi * (R * 2 * CHB);
    const long long rs = (long long)a.width1 * Dp * (up ? -1 : 1); // int16 elements to the next row of this sweep
    const long long o0 = (long long)x * Dp + (up ? (long long)(H - 1) * a.width1 * Dp : 0);
    const int16_t *src[NLD];
    uint32_t dsto[NLD];
    bool ldok[NLD];
#pragma unroll
    for (int q = 0; q < NLD; q++) {
        const int seg = lane + 32 * q;
        const int isS = seg / (DW / 4), r = seg % (DW / 4);
        src[q] = (isS ? a.S : a.C) + o0 + r * 8;
        dsto[q] = ring + seg * 16;
        ldok[q] = seg < NSEG && (acc || !isS);
    }
    auto issue = [&](int stage) { // load the row `src` points at into `stage`, advance to the next row
#pragma unroll
        for (int q = 0; q < NLD; q++) {
            if (ldok[q]) cp_async16_s(dsto[q] + stage * (2 * CHB), src[q]);
            src[q] += rs;
        }
    };
#pragma unroll 1
    for (int p = 0; p < R - 1; p++) {
        if (p < H) issue(p);
        cp_async_commit();
    }
    int16_t *sp = (acc ? a.S : a.S2) + o0 + lane * 2 * NP; // this lane's output words in the row of step t
    const uint32_t cur0 = ring + lane * NP * 4;

    uint32_t Td[NP], Tpre[NP], c[NP], s[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) Td[i] = Tpre[i] = s[i] = padmask[i];
    cp_async_wait<R - 2>();
    __syncwarp();
    lds_s<NP>(cur0, c);
    if (acc) lds_s<NP>(cur0 + CHB, s);

    int stage = 1 % R, pstage = R - 1; // stage of step t+1, stage that step t+R-1 is loaded into
    uint32_t pin = PSB, pout = 0;      // parity offsets of the slots read (step t-1) and written (step t)
    // The row loop exists twice: warps on a CTA boundary (EDGE) carry the hand-over code, the others do not pay for it.
    auto rows = [&](auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
#pragma unroll 1
        for (int t = 0; t < H; t++) {
            // ---- critical section: from the neighbours' row t-1 states to this column's row t states --------------
            // A warp signals even rows on its first mbarrier and odd rows on the second (a neighbour may run one row
            // ahead; with a single barrier two completed phases would look like none).  At t = 0 the wait is for the
            // phase before the first one, which counts as complete, and the slots hold the out-of-image state.
            const uint32_t mb_off = ((t - 1) & 1) * 8, par_in = ((t - 1) >> 1) & 1;
            uint32_t T0[NP], T1[NP], L0[NP], L1[NP], L2[NP];
            if (EDGE && in_glob[0]) { // (only a one-column strip receives both diagonals from other CTAs)
                if (t == 0) {
#pragma unroll
                    for (int i = 0; i < NP; i++) T0[i] = padmask[i];
                } else ho_read<NP>(in_g[0] + ((t - 1) & (HO_SLOTS - 1)) * DW, (((t - 1) >> 2) & 1) ? 0x80008000u : 0u, T0, a.err);
            } else {
                if (in_mb[0]) mbar_wait(in_mb[0] + mb_off, par_in, a.err);
                lds_s<NP>(in_s[0] + pin, T0);
            }
            if (EDGE && polls) {
                sgm_step<NP, PAD>(T0, c, L0, padmask, P1v, P2mP1v, lane);
                if (out_glob[0]) ho_write<NP>(out_g[0] + (t & (HO_SLOTS - 1)) * DW, ((t >> 2) & 1) ? 0x80008000u : 0u, T0);
                else sts_s<NP>(out_s[0] + pout, T0);
                if (t == 0) {
#pragma unroll
                    for (int i = 0; i < NP; i++) T1[i] = padmask[i];
                } else {
#pragma unroll
                    for (int i = 0; i < NP; i++) T1[i] = Tpre[i]; // loaded at the end of the previous step
                    ho_read<NP>(in_g[1] + ((t - 1) & (HO_SLOTS - 1)) * DW, (((t - 1) >> 2) & 1) ? 0x80008000u : 0u, T1, a.err, true);
                }
                sgm_step<NP, PAD>(T1, c, L1, padmask, P1v, P2mP1v, lane);
            } else {
                if (in_mb[1]) mbar_wait(in_mb[1] + mb_off, par_in, a.err);
                lds_s<NP>(in_s[1] + pin, T1);
                sgm_step<NP, PAD>(T0, c, L0, padmask, P1v, P2mP1v, lane);
                sgm_step<NP, PAD>(T1, c, L1, padmask, P1v, P2mP1v, lane);
                if (EDGE && out_glob[0]) ho_write<NP>(out_g[0] + (t & (HO_SLOTS - 1)) * DW, ((t >> 2) & 1) ? 0x80008000u : 0u, T0);
                else sts_s<NP>(out_s[0] + pout, T0);
            }
            if (EDGE && out_glob[1]) ho_write<NP>(out_g[1] + (t & (HO_SLOTS - 1)) * DW, ((t >> 2) & 1) ? 0x80008000u : 0u, T1);
            else sts_s<NP>(out_s[1] + pout, T1);
            __syncwarp();
            if (lane == 0) mbar_arrive(my_mb + (t & 1) * 8); // this column's row-t states are in their slots
            // the neighbour CTA wrote the state this warp needs in the next step early in ITS step t: fetch it now
            if (EDGE && polls) ho_load<NP>(in_g[1] + (t & (HO_SLOTS - 1)) * DW, Tpre);
            // ---- off the critical path: vertical path, sums, store, next row's C and S -------------------------------
            asm volatile("" : "+r"(Td[0]) : : "memory"); // keeps the compiler from hoisting the vertical path above the hand-over
            sgm_step<NP, PAD>(Td, c, L2, padmask, P1v, P2mP1v, lane);
            // saturating sums (L >= 0, so the order of the directions does not matter)
#pragma unroll
            for (int i = 0; i < NP; i++) {
                uint32_t v = __viaddmin_u16x2(L0[i], L1[i], BIG);
                v = __viaddmin_u16x2(v, L2[i], BIG);
                s[i] = acc ? __viaddmin_u16x2(s[i], v, BIG) : v;
            }
            stcg_regs<NP>(sp, s);
            sp += rs;
            __syncwarp();
            if (t + R - 1 < H) issue(pstage);
            cp_async_commit();
            cp_async_wait<R - 2>(); // row t+1 has landed
            __syncwarp();
            const uint32_t cur = cur0 + stage * (2 * CHB);
            lds_s<NP>(cur, c);
            if (acc) lds_s<NP>(cur + CHB, s);
            pstage = pstage + 1 == R ? 0 : pstage + 1;
            stage = stage + 1 == R ? 0 : stage + 1;
            pin = pout;
            pout ^= PSB;
        }
    };
    if (in_glob[0] || in_glob[1] || out_glob[0] || out_glob[1]) rows(std::true_type{});
    else rows(std::false_type{});
}

// ---------------------------------------------------------------------------------------------------------------------
// agg_vsweep2_kernel: the sweep of agg_vsweep_kernel<NP, PAD, 2, R> (both sweeps in one CTA) after its round-2 profile
// (profiles/r02_agg_raw.csv, source page): the strip runs at the pace of its slowest warp, and the slowest warps were the two per
// sweep on the CTA boundaries (211 instructions per row against 150, busy all 1500 cycles of a row while the other 24 warps
// spent 58 % of their time in the neighbour waits).  Three changes, same arithmetic, same hand-over rings:
//   * a boundary column is split between two warps: the column warp keeps the two diagonal paths and the hand-over to the
//     neighbour CTA and passes sat(L0 + L1) of every row through an 8-deep shared-memory ring to a helper warp, which runs the
//     vertical path of that column, forms the sums and owns the S / S2 traffic.  The eight boundary / helper warps are warps
//     0..7 of the CTA, i.e. two per scheduler.
//   * a warp owns ONE pair of mbarriers (even / odd rows) on which BOTH neighbours arrive, instead of waiting on each
//     neighbour's barrier in turn: one TRYWAIT round trip per row instead of two on the critical path.
//   * the row loop is unrolled by two so that barrier parities and slot parities are immediates, ring stages advance by mask
//     arithmetic, and the retry path of the waits is the 2.75-instruction loop of mbar_wait_tight.
// Needs 2n + 4 <= 32 warps and n >= 3 columns per strip; everything else stays with agg_vsweep_kernel.
template <int NP, bool PAD, int R, bool DBG = false, bool TRACE = false>
__global__ void __launch_bounds__(1024, 1) agg_vsweep2_kernel(VsArgs a)
{
    static_assert(NP == 1 || NP == 2 || NP == 4, "ring stages advance by mask arithmetic");
    constexpr int XR = 8;                 // depth of the column -> helper ring
    constexpr int DW = 32 * NP;           // 32-bit words of one pixel's d-chunk
    constexpr int CHB = DW * 4;           // ... in bytes
    constexpr int HO_DIR = HO_SLOTS * DW; // words of one direction's hand-over ring
    constexpr int NSEG = 2 * DW / 4;      // 16-byte segments of one ring stage: [C | S]
    constexpr int NLD = (NSEG + 31) / 32; // cp.async instructions per lane and step
    constexpr uint32_t SB = 2 * CHB, RMASK = R * SB - 1; // bytes of one ring stage, mask of the whole ring
    // [column mbarriers 2n x 2][helper-ring mbarriers 4 x XR][helper progress 4 x u32][diagonal slots][helper rings][C/S rings]
    extern __shared__ __align__(16) uint32_t vs_smem[];

    const int lane = threadIdx.x & 31;
    const int wi = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); // warp-uniform by construction
    const int n = a.n, G = gridDim.x, H = a.H, Dp = 64 * NP;
    const int dbg = DBG ? a.dbg : 0; // what-if timing switches, compiled out of the production kernel
    // TRACE: lane 0 of every warp of CTAs 60..62 stores clock64 at up to 8 points of rows 512..527 (trace[cta][warp][row][point])
    const bool tr_cta = TRACE && blockIdx.x >= 60 && blockIdx.x < 63 && lane == 0;
    long long *const tr_base = TRACE ? a.trace + ((long long)((int)blockIdx.x - 60) * 32 + wi) * 16 * 8 : nullptr;
    auto mark = [&](int t, int point) {
        if (TRACE) {
            if (tr_cta && (unsigned)(t - 512) < 16u) tr_base[(t - 512) * 8 + point] = clock64();
        }
    };
    int jl, w;
    bool helper = false;
    if (wi < 8) { // the strip's boundary columns and their helpers
        jl = (wi >> 1) & 1;
        w = (wi & 1) ? n - 1 : 0;
        helper = wi >= 4;
    } else {
        jl = (wi - 8) / (n - 2);
        w = 1 + (wi - 8) % (n - 2);
    }
    const bool up = jl == 1, acc = !up;
    const int x = blockIdx.x * n + w;
    const uint32_t BIG = 0x7FFF7FFFu;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(vs_smem);
    const uint32_t OFF_XF = 2 * n * 16, OFF_PROG = OFF_XF + 4 * XR * 8, OFF_SLOT = OFF_PROG + 16;
    const int NSLOT = 2 * 2 * 2 * (n + 2);
    const uint32_t OFF_X = OFF_SLOT + NSLOT * CHB, OFF_RING = OFF_X + 4 * XR * CHB;
    uint32_t *slots = vs_smem + OFF_SLOT / 4;

    uint32_t padmask[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) {
        int d0 = (lane * NP + i) * 2;
        padmask[i] = PAD ? ((d0 >= a.D ? 0x00007FFFu : 0u) | (d0 + 1 >= a.D ? 0x7FFF0000u : 0u)) : 0u;
    }
    for (int q = wi; q < NSLOT; q += blockDim.x >> 5) {
#pragma unroll
        for (int i = 0; i < NP; i++) slots[q * DW + lane * NP + i] = padmask[i];
    }
    if ((int)threadIdx.x < 4 * n) { // column barriers: one arrival per neighbour that lives in this CTA
        const int col = (threadIdx.x >> 1) % n, xx = blockIdx.x * n + col;
        const int cnt = (col > 0 ? 1 : 0) + (col < n - 1 && xx + 1 < a.width1 ? 1 : 0);
        mbar_init(sbase + threadIdx.x * 8, cnt > 0 ? cnt : 1);
    }
    if (threadIdx.x < 4 * XR) mbar_init(sbase + OFF_XF + threadIdx.x * 8, 1);
    if (threadIdx.x < 4) vs_smem[OFF_PROG / 4 + threadIdx.x] = 0;
    __syncthreads();
    if (x >= a.width1) return; // idle columns of the last strip
    const bool first_col = x == 0, last_col = x == a.width1 - 1;
    const bool glob_l = w == 0 && !first_col, glob_r = w == n - 1 && !last_col; // hand-over to / from the neighbour CTA on that side
    const bool split = glob_l || glob_r;                                         // this column is shared with a helper warp
    if (helper && !split) return;
    const bool has_l = w > 0, has_r = w < n - 1 && x + 1 < a.width1; // neighbours inside the CTA
    const uint32_t P1v = (uint32_t)a.P1 * 0x10001u, P2mP1v = (uint32_t)(a.P2 - a.P1) * 0x10001u;

    const uint32_t my_mb = sbase + (jl * n + w) * 16;
    constexpr uint32_t PSB = CHB; // the two parities of a slot are neighbours: [jl][dir][column + 1][parity][DW]
    uint32_t in_s[2], out_s[2];             // slot read / written (parity 0), dir 0 arrives from column x-1, dir 1 from x+1
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        const uint32_t sl = sbase + OFF_SLOT + ((jl * 2 + dir) * (n + 2)) * 2 * CHB + lane * NP * 4;
        in_s[dir] = sl + (dir == 0 ? w : w + 2) * 2 * CHB;
        out_s[dir] = sl + (w + 1) * 2 * CHB;
    }
    const int xi = wi & 3; // column <-> helper pair
    const uint32_t xslot = sbase + OFF_X + xi * XR * CHB + lane * NP * 4, xfull = sbase + OFF_XF + xi * XR * 8, xprog = sbase + OFF_PROG + xi * 4;

    // C (and S) rows stream through a private cp.async ring of R stages per warp
    const uint32_t ring = sbase + OFF_RING + wi * (R * SB);
    const long long rs = (long long)a.width1 * Dp * (up ? -1 : 1); // int16 elements to the next row of this sweep
    const long long o0 = (long long)x * Dp + (up ? (long long)(H - 1) * a.width1 * Dp : 0);
    const bool wantS = acc && !(split && !helper); // the column warp of a split column never touches S
    const int16_t *src[NLD];
    uint32_t dsto[NLD];
    bool ldok[NLD];
#pragma unroll
    for (int q = 0; q < NLD; q++) {
        const int seg = lane + 32 * q;
        const int isS = seg / (DW / 4), r = seg % (DW / 4);
        src[q] = (isS ? a.S : a.C) + o0 + r * 8;
        dsto[q] = ring + seg * 16;
        ldok[q] = seg < NSEG && (wantS || !isS);
    }
    uint32_t o_iss = 0; // stage (byte offset) the next row is loaded into
    // rows past the end of the sweep are "loaded" with a source size of 0 (nothing is read, the stage is zero-filled): no branch
    auto issue = [&](bool live) {
        const uint32_t sz = live ? 16u : 0u;
#pragma unroll
        for (int q = 0; q < NLD; q++) {
            if (ldok[q]) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dsto[q] + o_iss), "l"(src[q]), "r"(sz) : "memory");
            src[q] += rs;
        }
        o_iss = (o_iss + SB) & RMASK;
    };
#pragma unroll 1
    for (int p = 0; p < R - 1; p++) {
        issue(p < H);
        cp_async_commit();
    }
    int16_t *sp = (acc ? a.S : a.S2) + o0 + lane * 2 * NP; // this lane's output words in the row of step t
    const uint32_t cur0 = ring + lane * NP * 4;
    uint32_t o_cur = 0; // stage of row t

    uint32_t c[NP], s[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) s[i] = padmask[i];
    cp_async_wait<R - 2>();
    __syncwarp();
    lds_s<NP>(cur0, c);
    if (wantS) lds_s<NP>(cur0 + CHB, s);

    // end of a row: start the load of row t+R-1, wait for row t+1, fetch it
    auto next_row = [&](int t) {
        __syncwarp();
        issue(t + R - 1 < H);
        cp_async_commit();
        cp_async_wait<R - 2>(); // row t+1 has landed
        __syncwarp();
        o_cur = (o_cur + SB) & RMASK;
        lds_s<NP>(cur0 + o_cur, c);
        if (wantS) lds_s<NP>(cur0 + o_cur + CHB, s);
    };

    if (helper) {
        // ---- helper of a boundary column: vertical path, sums, S / S2 -------------------------------------------------
        uint32_t Td[NP], L2[NP], v[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) Td[i] = padmask[i];
#pragma unroll 1
        for (int t = 0; t < H; t++) {
            mark(t, 0);
            sgm_step<NP, PAD>(Td, c, L2, padmask, P1v, P2mP1v, lane);
            mark(t, 1);
            if (!(dbg & 2)) mbar_wait_tight(xfull + (t & (XR - 1)) * 8, (uint32_t)(t / XR) & 1u, a.err);
            mark(t, 2);
            lds_s<NP>(xslot + (t & (XR - 1)) * CHB, v);
#pragma unroll
            for (int i = 0; i < NP; i++) {
                const uint32_t u = __viaddmin_u16x2(v[i], L2[i], BIG);
                s[i] = acc ? __viaddmin_u16x2(s[i], u, BIG) : u;
            }
            stcg_regs<NP>(sp, s);
            sp += rs;
            __syncwarp(); // every lane has its words of ring slot t: the column warp may write row t+XR into it
            if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(xprog), "r"(t + 1) : "memory");
            mark(t, 3);
            next_row(t);
            mark(t, 4);
        }
        return;
    }

    uint32_t T0[NP], T1[NP], L0[NP], L1[NP];
    if (split) {
        // ---- boundary column: the two diagonal paths and the hand-over to the neighbour CTA ----------------------------
        // The diagonal that leaves the CTA (dir_o) is computed FIRST and written to the neighbour's ring, the one that enters
        // (dir_o ^ 1) is polled for SECOND (its load was issued at the end of the previous row), so that the hand-over latency
        // overlaps the rest of the row on both sides.
        const int dir_o = glob_r ? 0 : 1, dir_i = dir_o ^ 1;
        const int js = up ? 1 : 0;
        const size_t HJ = (size_t)(G > 1 ? G - 1 : 1) * 2 * HO_DIR;
        const int b = glob_r ? (int)blockIdx.x : (int)blockIdx.x - 1; // boundary between CTA b and b+1
        const uint32_t *in_g = a.ho + js * HJ + ((size_t)b * 2 + dir_i) * HO_DIR + lane * NP;
        uint32_t *out_g = a.ho + js * HJ + ((size_t)b * 2 + dir_o) * HO_DIR + lane * NP;
        const uint32_t in_sm = in_s[dir_o], out_sm = out_s[dir_i];
        const uint32_t nb_mb = glob_r ? my_mb - 16 : my_mb + 16; // the one neighbour inside the CTA
        const bool has_nb = glob_r ? has_l : has_r;              // (n >= 3: always)
        uint32_t Tpre[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) Tpre[i] = padmask[i];
        auto row = [&](int t, auto par_tag) {
            constexpr int PAR = decltype(par_tag)::value; // t & 1
            uint32_t prog;
            mark(t, 0);
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(prog) : "r"(xprog) : "memory");
            if ((dbg & 4) && t > 0) ho_load<NP>(in_g + ((t - 1) & (HO_SLOTS - 1)) * DW, Tpre);
            if (t > 0 && has_nb && !(dbg & 2)) mbar_wait_tight(my_mb + (PAR ^ 1) * 8, (uint32_t)((t - 1) >> 1) & 1u, a.err);
            mark(t, 1);
            lds_s<NP>(in_sm + (PAR ^ 1) * PSB, T0);
            sgm_step<NP, PAD>(T0, c, L0, padmask, P1v, P2mP1v, lane);
            ho_write<NP>(out_g + (t & (HO_SLOTS - 1)) * DW, ((t >> 2) & 1) ? 0x80008000u : 0u, T0);
            mark(t, 2);
            if (t == 0) {
#pragma unroll
                for (int i = 0; i < NP; i++) T1[i] = padmask[i];
            } else {
#pragma unroll
                for (int i = 0; i < NP; i++) T1[i] = Tpre[i]; // loaded at the end of the previous row
                if (!(dbg & 1)) ho_read<NP>(in_g + ((t - 1) & (HO_SLOTS - 1)) * DW, (((t - 1) >> 2) & 1) ? 0x80008000u : 0u, T1, a.err, true);
            }
            mark(t, 3);
            sgm_step<NP, PAD>(T1, c, L1, padmask, P1v, P2mP1v, lane);
            sts_s<NP>(out_sm + PAR * PSB, T1);
            __syncwarp();
            if (lane == 0 && has_nb) mbar_arrive(nb_mb + PAR * 8); // this column's row-t state is in its slot
            mark(t, 4);
            if (!(dbg & 4)) ho_load<NP>(in_g + (t & (HO_SLOTS - 1)) * DW, Tpre);   // the neighbour CTA wrote it early in ITS row t
            // sat(L0 + L1) to the helper (ring slot t mod XR is free once the helper has finished row t - XR)
            if (t >= XR && (int)prog < t - XR + 1) {
                int spins = 0;
                unsigned long long t0 = 0;
                do {
                    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(prog) : "r"(xprog) : "memory");
                    if (wait_expired(++spins, t0, a.err)) {
                        *(volatile int *)a.err = 1;
                        break;
                    }
                } while ((int)prog < t - XR + 1);
            }
            uint32_t v[NP];
#pragma unroll
            for (int i = 0; i < NP; i++) v[i] = __viaddmin_u16x2(L0[i], L1[i], BIG);
            sts_s<NP>(xslot + (t & (XR - 1)) * CHB, v);
            __syncwarp();
            if (lane == 0) mbar_arrive(xfull + (t & (XR - 1)) * 8);
            mark(t, 5);
            next_row(t);
            mark(t, 6);
        };
#pragma unroll 1
        for (int t = 0; t < H; t += 2) {
            row(t, std::integral_constant<int, 0>{});
            if (t + 1 < H) row(t + 1, std::integral_constant<int, 1>{});
        }
        return;
    }

    // ---- interior column (or a column on the image border): all three paths -----------------------------------------------
    uint32_t Td[NP], L2[NP];
#pragma unroll
    for (int i = 0; i < NP; i++) Td[i] = padmask[i];
    const bool waits = has_l || has_r;
    const bool arrives = (lane == 0 && has_l) || (lane == 1 && has_r);
    const uint32_t arr_mb = lane == 0 ? my_mb - 16 : my_mb + 16;
    auto row = [&](int t, auto par_tag) {
        constexpr int PAR = decltype(par_tag)::value; // t & 1
        mark(t, 0);
        // both neighbours' row t-1 states are in their slots (they arrive on THIS warp's barrier; even rows on the first
        // barrier, odd rows on the second: a neighbour may run one row ahead)
        if (t > 0 && waits && !(dbg & 2)) mbar_wait_tight(my_mb + (PAR ^ 1) * 8, (uint32_t)((t - 1) >> 1) & 1u, a.err);
        mark(t, 1);
        lds_s<NP>(in_s[0] + (PAR ^ 1) * PSB, T0);
        lds_s<NP>(in_s[1] + (PAR ^ 1) * PSB, T1);
        sgm_step<NP, PAD>(T0, c, L0, padmask, P1v, P2mP1v, lane);
        sgm_step<NP, PAD>(T1, c, L1, padmask, P1v, P2mP1v, lane);
        sts_s<NP>(out_s[0] + PAR * PSB, T0);
        sts_s<NP>(out_s[1] + PAR * PSB, T1);
        __syncwarp();
        if (arrives) mbar_arrive(arr_mb + PAR * 8); // lane 0 on the left neighbour's barrier, lane 1 on the right neighbour's
        mark(t, 2);
        sgm_step<NP, PAD>(Td, c, L2, padmask, P1v, P2mP1v, lane);
#pragma unroll
        for (int i = 0; i < NP; i++) {
            uint32_t v = __viaddmin_u16x2(L0[i], L1[i], BIG);
            v = __viaddmin_u16x2(v, L2[i], BIG);
            s[i] = acc ? __viaddmin_u16x2(s[i], v, BIG) : v;
        }
        stcg_regs<NP>(sp, s);
        sp += rs;
        mark(t, 3);
        next_row(t);
        mark(t, 4);
    };
#pragma unroll 1
    for (int t = 0; t < H; t += 2) {
        row(t, std::integral_constant<int, 0>{});
        if (t + 1 < H) row(t + 1, std::integral_constant<int, 1>{});
    }
}

bool sweep_plain_launch()
{
    const char *e = getenv("B2S_SWEEP_COOPERATIVE");
    if (e) return atoi(e) == 0;
    return getenv("CUDA_MPS_PIPE_DIRECTORY") == nullptr; // shared through MPS: other clients' kernels hold SMs we cannot see
}

// launch of a sweep grid (all G strips co-resident): attribute once per kernel and device, occupancy check, plain or cooperative
template <typename K> cudaError_t launch_sweep_grid(b2s_ctx *c, K kernel, const VsArgs &a, int G, int threads, size_t smem)
{
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done; // (kernel, device): the attribute belongs to the device's context
    cudaError_t e = cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(mu);
        const auto key = std::make_pair((const void *)kernel, c->device);
        if (!done.count(key)) {
            if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)) != cudaSuccess) return e;
            done.insert(key);
        }
    }
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1 || G > occ * c->num_sms) return cudaErrorCooperativeLaunchTooLarge; // all strips must be co-resident
    // The strips spin on their neighbours, so every CTA must be resident.  Two ways to get that:
    //  * plain launch (default): grid <= #SM x occupancy (checked above), sweeps of one process chained per device with an event so
    //    that two launches never interleave their CTAs, and a 2 s timeout in every spin loop that turns a lost hand-over into an
    //    error instead of a hang.  Right for a GPU this process owns.
    //  * cooperative launch (B2S_SWEEP_COOPERATIVE=1, or automatically when an MPS pipe directory is configured): the driver
    //    starts the grid only when all of it fits at once, whatever else shares the GPU.  Measured cost on a dedicated B200: none
    //    for device-resident batches (515 pairs/s either way), 5 % end to end (460 against 484 pairs/s): a cooperative grid does
    //    not overlap the other streams' copies and small kernels as freely.
    if (sweep_plain_launch()) {
        kernel<<<G, threads, smem, c->stream>>>(a);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(G);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if ((e = cudaLaunchKernelEx(&cfg, kernel, a)) != cudaSuccess) return e;
    }
    c->launches++;
    return cudaGetLastError();
}
template <int NP, bool PAD, int JW, int R> cudaError_t launch_vsweep_t(b2s_ctx *c, const VsArgs &a, int G, size_t smem)
{
    return launch_sweep_grid(c, agg_vsweep_kernel<NP, PAD, JW, R>, a, G, a.n * JW * 32, smem);
}
template <int NP, bool PAD, int JW> cudaError_t launch_vsweep_r(b2s_ctx *c, const VsArgs &a, int G)
{
    const size_t fixed = (size_t)JW * a.n * 16 + (size_t)JW * 2 * 2 * (a.n + 2) * 128 * NP, stage = (size_t)JW * a.n * 2 * 128 * NP;
    if (fixed + 8 * stage <= 200 * 1024) return launch_vsweep_t<NP, PAD, JW, 8>(c, a, G, fixed + 8 * stage);
    if (fixed + 4 * stage <= 216 * 1024) return launch_vsweep_t<NP, PAD, JW, 4>(c, a, G, fixed + 4 * stage);
    return cudaErrorInvalidConfiguration;
}
// J sweeps (1 = top-down only, 2 = both): in one CTA when 2n warps fit, else one launch per sweep
// both sweeps with the boundary columns split between a column warp and a helper warp (agg_vsweep2_kernel); false: not applicable
template <int NP, bool PAD> bool launch_vsweep2(b2s_ctx *c, const VsArgs &a, int G, cudaError_t *e)
{
    if constexpr (NP == 1 || NP == 2 || NP == 4) {
        const bool off = getenv("B2S_VSWEEP2") && atoi(getenv("B2S_VSWEEP2")) == 0; // (tests: the same cases through agg_vsweep_kernel)
        const int n = a.n, warps = 2 * n + 4;
        if (off || n < 3 || warps > 32) return false;
        const size_t fixed = (size_t)2 * n * 16 + 4 * 8 * 8 + 16 + (size_t)2 * 2 * 2 * (n + 2) * 128 * NP + (size_t)4 * 8 * 128 * NP;
        const size_t stage = (size_t)warps * 2 * 128 * NP;
        if constexpr (NP == 2 && !PAD) { // what-if timings of the benchmark configuration (scripts/vs_variants.sh)
            if (a.trace && fixed + 8 * stage <= 216 * 1024) {
                *e = launch_sweep_grid(c, agg_vsweep2_kernel<NP, PAD, 8, false, true>, a, G, warps * 32, fixed + 8 * stage);
                return true;
            }
            if (a.dbg && fixed + 8 * stage <= 216 * 1024) {
                *e = launch_sweep_grid(c, agg_vsweep2_kernel<NP, PAD, 8, true>, a, G, warps * 32, fixed + 8 * stage);
                return true;
            }
        }
        if (fixed + 8 * stage <= 216 * 1024) *e = launch_sweep_grid(c, agg_vsweep2_kernel<NP, PAD, 8>, a, G, warps * 32, fixed + 8 * stage);
        else if (fixed + 4 * stage <= 216 * 1024) *e = launch_sweep_grid(c, agg_vsweep2_kernel<NP, PAD, 4>, a, G, warps * 32, fixed + 4 * stage);
        else return false;
        return true;
    } else return false;
}
template <int NP, bool PAD> cudaError_t launch_vsweep_j(b2s_ctx *c, VsArgs &a, int G, int J)
{
    cudaError_t e2 = cudaSuccess;
    if (J == 2 && launch_vsweep2<NP, PAD>(c, a, G, &e2)) return e2;
    if (J == 2 && a.n * 2 <= 32) return launch_vsweep_r<NP, PAD, 2>(c, a, G);
    for (int j = 0; j < J; j++) {
        a.up = j;
        cudaError_t e = launch_vsweep_r<NP, PAD, 1>(c, a, G);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// columns per CTA for the fused vertical sweep, 0 = not applicable (strip wider than 32 columns: legacy per-direction path)
int vsweep_cols(const b2s_ctx *c)
{
    const SgbmGeom &g = c->g;
    if (getenv("B2S_AGG_LEGACY") || g.NP > 4) return 0; // (more than 256 disparities: the generic scans only)
    if (g.layout == 1) return vsweep6_cols(c, g);            // block layout: the six-path sweep's strip width
    int n = (g.width1 + c->num_sms - 1) / c->num_sms;
    if (n < 8) n = g.width1 < 8 ? g.width1 : 8;
    if (const char *e = getenv("B2S_VSWEEP_COLS")) { // test hook: force narrow strips so that small images span several CTAs
        int v = atoi(e);
        if (v >= 1 && v <= 32 && (g.width1 + v - 1) / v <= c->num_sms) n = v;
    }
    return n <= 32 ? n : 0;
}

// The strips of a sweep spin on each other, so all CTAs of a launch must become resident.  Two such launches from
// different streams must not interleave their CTAs (each could hold SMs the other one needs): sweeps of one process
// are chained per device with an event, which costs nothing because they cannot share the SMs anyway.
struct SweepChain {
    std::mutex mu;
    cudaEvent_t ev[64][4] = {};
    unsigned long long count[64] = {};
};
SweepChain g_chain;
// how many sweeps of one process may be in flight on a device at once (1: strictly one after the other)
int sweep_concurrency()
{
    const char *e = getenv("B2S_SWEEP_CONCURRENCY");
    int v = e ? atoi(e) : 1;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
}

cudaError_t launch_vsweep(b2s_ctx *c, int n, int J)
{
    const SgbmGeom &g = c->g;
    const int G = (g.width1 + n - 1) / n;
    VsArgs a;
    a.C = c->C.as<int16_t>();
    a.S = c->S.as<int16_t>();
    a.S2 = c->S2.as<int16_t>();
    a.H = g.H; a.width1 = g.width1; a.D = g.D; a.P1 = g.P1; a.P2 = g.P2; a.n = n; a.up = 0;
    size_t ho_bytes = (size_t)2 * (G > 1 ? G - 1 : 1) * 2 * HO_SLOTS * 32 * g.NP * sizeof(uint32_t);
    cudaError_t e = c->agg_ho.ensure(ho_bytes);
    if (e != cudaSuccess) return e;
    // every int16 of the rings starts with phase 1 (steps 0..3 write phase 0)
    if ((e = cudaMemsetAsync(c->agg_ho.p, 0xFF, ho_bytes, c->stream)) != cudaSuccess) return e;
    a.ho = c->agg_ho.as<uint32_t>();
    a.err = c->agg_err; // (launch_aggregate)
    a.dbg = getenv("B2S_VS2_FAKE") ? atoi(getenv("B2S_VS2_FAKE")) : 0;
    a.trace = nullptr;
    const char *trace_path = getenv("B2S_VS2_TRACE"); // development aid: dump the row timeline of three CTAs (scripts/vs2_trace.py)
    const size_t trace_bytes = (size_t)3 * 32 * 16 * 8 * sizeof(long long);
    if (trace_path) {
        if ((e = cudaMalloc(&a.trace, trace_bytes)) != cudaSuccess) return e;
        cudaMemsetAsync(a.trace, 0, trace_bytes, c->stream);
    }
    const bool pad = g.D != g.Dp;
    const bool chained = sweep_plain_launch(); // (cooperative launches need no ordering between handles)
    std::unique_lock<std::mutex> lock(g_chain.mu, std::defer_lock);
    cudaEvent_t *evp = nullptr;
    if (chained) {
        lock.lock();
        const int dev = c->device & 63, nconc = sweep_concurrency();
        evp = &g_chain.ev[dev][g_chain.count[dev]++ % nconc]; // recorded by the sweep `nconc` launches ago
        if (!*evp) {
            if ((e = cudaEventCreateWithFlags(evp, cudaEventDisableTiming)) != cudaSuccess) return e;
        } else if ((e = cudaStreamWaitEvent(c->stream, *evp, 0)) != cudaSuccess) return e;
    }
    if (g.layout == 1) e = launch_vsweep6(c, n, !chained); // (block layout: only reached when vsweep6_cols(c, g) == n > 0)
    else switch (g.NP) {
    case 1: e = pad ? launch_vsweep_j<1, true>(c, a, G, J) : launch_vsweep_j<1, false>(c, a, G, J); break;
    case 2: e = pad ? launch_vsweep_j<2, true>(c, a, G, J) : launch_vsweep_j<2, false>(c, a, G, J); break;
    case 3: e = pad ? launch_vsweep_j<3, true>(c, a, G, J) : launch_vsweep_j<3, false>(c, a, G, J); break;
    case 4: e = pad ? launch_vsweep_j<4, true>(c, a, G, J) : launch_vsweep_j<4, false>(c, a, G, J); break;
    default: e = cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    if (a.trace) { // (development aid: synchronous on purpose)
        std::vector<long long> host(trace_bytes / sizeof(long long));
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;
        if ((e = cudaMemcpy(host.data(), a.trace, trace_bytes, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
        cudaFree(a.trace);
        if (FILE *f = fopen(trace_path, "wb")) {
            fwrite(host.data(), 1, trace_bytes, f);
            fclose(f);
        }
    }
    return chained ? cudaEventRecord(*evp, c->stream) : cudaSuccess;
}

} // namespace

// True when the cost stage may stop at the row sums (in S2) and leave the vertical box sum to the first horizontal scan:
// BT cost, window height <= 5, the production schedule (not MODE_HH4, whose bottom rows get a constant cost, nor the legacy path).
// Schedules of the eight (five) paths: "sweep" (default: two horizontal scans around the lock-step
// vertical sweep, 20 B/voxel, the fastest for a single pair), "wave" (B2S_OPT_AGG_SCHEDULE = 1 or B2S_AGG_SCHEDULE=wave;
// sgbm_wave.cu: two wavefront sweeps of four paths each at the canonical 8 B/voxel; MODE_HH only), "legacy" (one scan per
// direction, the cross-check of the tests).  The environment variable overrides the option.
bool agg_wave_selected(const b2s_ctx *c, int mode)
{
    if (getenv("B2S_AGG_LEGACY")) return false;
    const char *e = getenv("B2S_AGG_SCHEDULE");
    bool wave = c->agg_schedule == 1;
    if (e) wave = !strcmp(e, "wave");
    return wave && mode == 1; // MODE_SGBM keeps the sweep schedule (its fifth path runs inside the winner-take-all scan)
}

bool agg_fuses_vsum(const b2s_ctx *c)
{
    if (getenv("B2S_NO_VSUM_FUSION")) return false;
    if (c->g.layout != 0 && vsweep6_cols(c, c->g) == 0) return false; // (block layout: only with the six-path sweep schedule)
    if (c->g.layout != 0 && agg_wave_selected(c, c->g.mode)) return false;
    return c->prm.cost == 0 && c->g.SH2 <= 2 && c->g.mode != 3 && vsweep_cols(c) > 0;
}

// Device error flags of a handle, read by agg_poll_error after a stream sync: word 0 = a wait of the aggregation kernels timed
// out (lost hand-over or bulk copy), word 1 = a block sum of the cost volume wrapped past 32767 (C < 0; possible from block 11 with
// three channels on: 121 * 279 = 33759), which the packed unsigned arithmetic of the aggregation does not follow.
cudaError_t agg_error_flags(b2s_ctx *c)
{
    if (c->agg_err) return cudaSuccess;
    cudaError_t e = c->agg_errbuf.ensure(256);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemset(c->agg_errbuf.p, 0, 256)) != cudaSuccess) return e;
    c->agg_err = c->agg_errbuf.as<int>();
    return cudaSuccess;
}

int agg_poll_error(b2s_ctx *c)
{
    if (!c->agg_err) return 0;
    int v[2] = {0, 0};
    if (cudaMemcpy(v, c->agg_err, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    if (v[0] || v[1]) cudaMemset(c->agg_err, 0, sizeof v); // reported once
    return (v[0] ? 1 : 0) | (v[1] ? 2 : 0);
}

// Directions as (mx,my) of the MOVE along the path (predecessor = p - move).  cv2 pass 1: (+1,0) (+1,+1) (0,+1) (-1,+1);
// MODE_SGBM adds (-1,0) during the WTA sweep; MODE_HH pass 2 adds (-1,0) (-1,-1) (0,-1) (+1,-1).  The saturating sum
// over directions is order-independent, so the schedule is: horizontal (+1,0) initialises S; the fused vertical sweep
// adds the three top-down directions to S and (MODE_HH) writes the sum of the three bottom-up directions to S2;
// horizontal (-1,0) comes last and folds S2 in: S = sat(S + S2 + L).
cudaError_t launch_aggregate(b2s_ctx *c, int *n_launches, cudaEvent_t *marks)
{
    // marks (nullable): marks[0] is recorded before the first launch and marks[k] after the k-th (at most B2S_AGG_MAX_PARTS)
    int nm = 0;
    auto mark = [&]() { if (marks && nm <= B2S_AGG_MAX_PARTS) cudaEventRecord(marks[nm++], c->stream); };
    mark();
    static const int dirs8[8][2] = {{1, 0}, {-1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, -1}, {0, -1}, {1, -1}};
    const SgbmGeom &g = c->g;
    AggArgs a;
    a.C = c->C.as<int16_t>();
    a.S = c->S.as<int16_t>();
    a.S2 = c->S2.as<int16_t>();
    a.H = g.H; a.width1 = g.width1; a.D = g.D; a.P1 = g.P1; a.P2 = g.P2;
    a.raw = c->raw.as<int16_t>();
    a.disp2key = c->disp2key.as<unsigned>();
    a.W = g.W; a.minX1 = g.minX1; a.minD = g.minD; a.uniq = g.uniq;
    // device error flag of this aggregation (a wait that timed out: lost strip hand-over or bulk copy), read by agg_poll_error
    // (sticky: cleared when the buffer is created and by agg_poll_error after it reported, never per launch -- several pairs may be
    // queued on the stream between two polls)
    cudaError_t e;
    if ((e = wait_timeout_init()) != cudaSuccess) return e;
    if ((e = agg_error_flags(c)) != cudaSuccess) return e;
    a.err = c->agg_err;
    const int thr = 100 - g.uniq;
    a.uniq_M = thr > 1 ? (unsigned)(((1ull << 32) + thr - 1) / thr) : 0u;
    const bool can_fuse = c->fuse_wta && thr >= 1 && thr <= 100; // (uniquenessRatio >= 100 goes through wta_kernel)
    c->wta_fused = false;
    c->wta_adds_s2 = false;
    if (g.layout == 1 && (agg_wave_selected(c, g.mode) || vsweep6_cols(c, g) == 0)) { // the wavefront schedule, MODE_HH: both sweeps in one launch; wta_kernel forms sat(S + S2)
        c->agg_legacy = false;
        if ((e = launch_wave(c, 2)) != cudaSuccess) return e;
        c->wta_adds_s2 = true;
        mark();
        if (n_launches) *n_launches = 1;
        return cudaSuccess;
    }
    const char *sched = getenv("B2S_AGG_SCHEDULE");
    const int n = (g.mode == 3 || (sched && !strcmp(sched, "legacy"))) ? 0 : vsweep_cols(c);
    c->agg_legacy = n == 0; // the legacy path keeps the generic scan kernel for every direction (it is the cross-check)
    if (n > 0) {
        a.mx = 1; a.my = 0;
        if (c->hs_pending) { // the cost stage left row sums in S2: this scan also forms C (consumed once: S2 is the up-sweep's output)
            a.hs = c->S2.as<int16_t>();
            a.SH2 = g.SH2;
            e = launch_hscan_vsum(c, a);
            c->hs_pending = false;
        } else
            e = launch_dir_np(c, a, AGG_INIT);
        if (e != cudaSuccess) return e;
        mark();
        if ((e = launch_vsweep(c, n, g.mode == 1 ? 2 : 1)) != cudaSuccess) return e;
        mark();
        a.mx = -1;
        // the last scan also does the winner-take-all (its lanes hold the final S of the pixel) and stores S only if the
        // volumes are kept; B2S_OPT_FUSE_WTA = 0 or uniquenessRatio >= 100 leave it to wta_kernel (sgbm_post.cu)
        const int wta = can_fuse ? (c->keep_volumes ? 2 : 1) : 0;
        if ((e = launch_dir_np(c, a, g.mode == 1 ? AGG_ACCUM2 : AGG_ACCUM, wta)) != cudaSuccess) return e;
        c->wta_fused = wta != 0;
        mark();
        if (n_launches) *n_launches = 3;
        return cudaSuccess;
    }
    if (g.mode == 3) {
        // MODE_HH4 (cv2 pass 1: (+1,0) (0,+1); pass 2: (-1,0) (0,-1)): the vertical paths do not couple columns, so each is one
        // launch of the generic scan with a warp per column; the horizontal ones use the fast row scan
        static const int dirs4[4][2] = {{1, 0}, {0, 1}, {0, -1}, {-1, 0}};
        c->agg_legacy = g.NP > 4;
        for (int i = 0; i < 4; i++) {
            a.mx = dirs4[i][0];
            a.my = dirs4[i][1];
            const int wta = (i == 3 && can_fuse && g.NP <= 4) ? (c->keep_volumes ? 2 : 1) : 0;
            if ((e = launch_dir_np(c, a, i == 0 ? AGG_INIT : AGG_ACCUM, wta)) != cudaSuccess) return e;
            if (wta) c->wta_fused = true;
            mark();
        }
        if (n_launches) *n_launches = 4;
        return cudaSuccess;
    }
    int nd = g.mode == 1 ? 8 : 5; // legacy: one scan kernel per direction
    for (int i = 0; i < nd; i++) {
        a.mx = dirs8[i][0];
        a.my = dirs8[i][1];
        if ((e = launch_dir_np(c, a, i == 0 ? AGG_INIT : AGG_ACCUM)) != cudaSuccess) return e;
        mark();
    }
    if (n_launches) *n_launches = nd;
    return cudaSuccess;
}
