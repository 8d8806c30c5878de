// sgbm_cost.cu -- cost volume C(y, x1, d) of cv2.StereoSGBM, sm_100a.
//
// Replaces the calcPixelCostBT + box-sum half of cv2.StereoSGBM.compute, called from
// calibrating/stereo_matching.py:63.  All sums are int16 with two's-complement wrap (packed VIADD.16x2), which
// makes running sums exact whatever the order.
//
// Layout in HBM: planes  (H, 2cn, W) uchar4 {value, lo, hi, 0}   lo/hi = Birchfield-Tomasi half-sample interval
//                hs, C   (H, width1, Dp) int16, d fastest, Dp = 64*NP (padded d-lanes are zero in C)
#include <mutex>
#include <set>
#include <utility>

#include "b2s_internal.h"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// value of plane p (0..cn-1 gradient, cn..2cn-1 raw) at (y, x)
template <int CN>
__device__ __forceinline__ int plane_value(const uint8_t *__restrict__ img, int H, int W, int y, int x, int p, int ftzero)
{
    if (x <= 0 || x >= W - 1) return ftzero;
    if (p >= CN) return img[((size_t)y * W + x) * CN + (p - CN)];
    int ym = max(y - 1, 0), yp = min(y + 1, H - 1);
    const uint8_t *r0 = img + (size_t)y * W * CN + p, *rm = img + (size_t)ym * W * CN + p, *rp = img + (size_t)yp * W * CN + p;
    int s = 2 * ((int)r0[(x + 1) * CN] - (int)r0[(x - 1) * CN]) + ((int)rm[(x + 1) * CN] - (int)rm[(x - 1) * CN]) +
            ((int)rp[(x + 1) * CN] - (int)rp[(x - 1) * CN]);
    return clampi(s, -ftzero, ftzero) + ftzero;
}

// A warp produces 30 consecutive pixels of a row: every lane evaluates the plane values of ONE pixel (lanes 0 and 31 are the
// halo), the half-sample neighbours come from the adjacent lanes.
constexpr int PLANES_WARPS = 8;
template <int CN>
__global__ void __launch_bounds__(PLANES_WARPS * 32) planes_kernel(const uint8_t *__restrict__ img, uchar4 *__restrict__ out, int H, int W, int ftzero)
{
    const int lane = threadIdx.x & 31, y = blockIdx.y;
    const int x = (blockIdx.x * PLANES_WARPS + (threadIdx.x >> 5)) * 30 + lane - 1;
    const int xc = clampi(x, 0, W - 1);
    const bool write = lane >= 1 && lane <= 30 && x < W;
#pragma unroll
    for (int p = 0; p < 2 * CN; p++) {
        const int u = plane_value<CN>(img, H, W, y, xc, p, ftzero);
        const int l = __shfl_up_sync(0xffffffffu, u, 1), r = __shfl_down_sync(0xffffffffu, u, 1);
        if (!write) continue;
        const int ul = x > 0 ? (u + l) / 2 : u;
        const int ur = x < W - 1 ? (u + r) / 2 : u;
        out[((size_t)y * 2 * CN + p) * W + x] = make_uchar4((unsigned char)u, (unsigned char)min(min(ul, ur), u),
                                                            (unsigned char)max(max(ul, ur), u), 0);
    }
}

// One CTA = one image row y and TX cost columns.
// Phase 0: stage the BT operands of the tile in shared memory, already in the form phase 1 consumes:
//   left  pixel  (column j, plane p): four packed int16x2 constants  (u+B, B-u, B-u1, u0+B), B = 256, both halves equal
//   right pixels (plane p): v, v0, v1 as packed PAIRS of neighbouring pixels, indexed so that the pair of a lane's two
//   disparities (d, d+1) is one aligned 32-bit word; two copies (even / odd alignment) cover both parities of the column.
// Phase 1: a lane owns two adjacent disparities of one column.  With the bias B every difference of the
//   Birchfield-Tomasi cost is positive, so packed int16x2 differences are plain 32-bit adds/subtracts (no borrow between
//   the halves), the max(0, ., .) are VIMNMX3.S16x2 against the packed bias, and per plane the cost of two voxels is
//   5 integer adds + 2 VIMNMX3 + 1 VIMNMX (+ shift/mask for the raw planes, whose cost is >> 2):
//       c0 + B = max3((u+B) - v1, v0 + (B-u), B)      c1 + B = max3(v + (B-u1), (u0+B) - v, B)      cost + B = min(.,.)
//   and (cost + B) >> 2 = B/4 + (cost >> 2) exactly, so the biases leave as one constant at the end.
// Phase 2: horizontal box sum (window clamped in cost-volume coordinates) -> hs.
constexpr int TX = 64;
constexpr int COST_THREADS = 256;
constexpr int BT_BIAS = 256;
constexpr int NRP_MAX = (TX + 2 * 5 + 256) / 2 + 3; // packed right-pixel pairs per copy; compile-time strides (template NRP): 104 for D <= 128, NRP_MAX for D <= 256
constexpr int NRP_WIDE = (TX + 2 * 5 + 512) / 2 + 3; // ... and for D <= 512

// LAYOUT 1 (block layout of the wavefront kernel: a word = disparities (d, d+8)): the right-pixel table holds ONE entry per
// reversed index m = (pixel m, pixel m+8) instead of the two alignment copies of neighbouring pairs; 2*NRP entries per plane.
template <int CN, int NRP, int LAYOUT, bool GROUPED = false>
__global__ void __launch_bounds__(COST_THREADS, 4) pixcost_hsum_kernel(const uchar4 *__restrict__ PL, const uchar4 *__restrict__ PR,
                                                                    int16_t *__restrict__ hs, SgbmGeom g)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NPL = 2 * CN;
    const int y = blockIdx.y;
    const int x1_0 = blockIdx.x * TX;
    const int TXH = TX + 2 * g.SW2;
    const int xlo = x1_0 - g.SW2;             // cost column of tile-local index 0
    const int Dp = g.Dp, DW = Dp / 2;         // 32-bit words per column of pix
    uint4 *sL = (uint4 *)smem;                // [TXH][NPL]
    uint4 *sR = sL + NPL * TXH;               // [NPL][2 copies][NRP] {v, v0, v1, -} packed pairs: one LDS.128 per plane and lane
    uint32_t *pixw = (uint32_t *)(sR + NPL * 2 * NRP); // [TXH][DW] packed int16x2
    // right pixel of reversed index m (m grows with d): image column xr(m) = xrmax - m
    const int xrmax = xlo + g.minX1 - g.minD + (TXH - 1);

    // (flat indices are split with compile-time divisors -- TX + 10 >= TXH and NRP >= nrp -- so there is no run-time division)
    constexpr int TXHP = TX + 10;
    for (int i = threadIdx.x; i < NPL * TXHP; i += COST_THREADS) {
        const int p = i / TXHP, j = i % TXHP;
        if (j >= TXH) continue;
        int x = clampi(xlo + j + g.minX1, 0, g.W - 1);
        uchar4 L = PL[((size_t)y * NPL + p) * g.W + x];
        uint32_t u = L.x, u0 = L.y, u1 = L.z;
        sL[j * NPL + p] = make_uint4((u + BT_BIAS) * 0x10001u, (BT_BIAS - u) * 0x10001u, (BT_BIAS - u1) * 0x10001u, (u0 + BT_BIAS) * 0x10001u);
    }
    const int nrp = (TXH + g.D - 1) / 2 + 2; // pairs actually read by phase 1 (<= NRP, checked by the launcher)
    for (int i = threadIdx.x; i < NPL * 2 * NRP; i += COST_THREADS) {
        const int pc = i / NRP, wd = i % NRP, p = pc >> 1, cp = pc & 1;
        const uchar4 *prow = PR + ((size_t)y * NPL + p) * g.W;
        int m0, m1;
        if (LAYOUT == 0) {
            if (wd >= nrp) continue;
            m0 = 2 * wd + cp; // copy 0: pairs (2w, 2w+1); copy 1: pairs (2w+1, 2w+2)
            m1 = m0 + 1;
        } else {
            m0 = cp * NRP + wd; // one table of 2*NRP entries: pair (m, m+8)
            if (m0 >= TXH + g.Dp) continue;
            m1 = m0 + 8;
        }
        uchar4 a = prow[clampi(xrmax - m0, 0, g.W - 1)];
        uchar4 b = prow[clampi(xrmax - m1, 0, g.W - 1)];
        // GROUPED: a lane reads entries 2*lane + const, so the entries of a copy are stored de-interleaved (even ones, then odd ones):
        // the lanes of a quarter-warp then hit eight different 16-byte bank groups
        const int pos = (GROUPED && LAYOUT == 0) ? pc * NRP + (wd >> 1) + (wd & 1) * ((NRP + 1) / 2) : i;
        sR[pos] = make_uint4(a.x | ((uint32_t)b.x << 16), a.y | ((uint32_t)b.y << 16), a.z | ((uint32_t)b.z << 16), 0u);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t Bp = BT_BIAS * 0x10001u, unbias = (uint32_t)(CN * (BT_BIAS + BT_BIAS / 4)) * 0x10001u;
    if constexpr (GROUPED && LAYOUT == 0) {
        // phase 1, grouped (round 2): the kernel is bound by the shared-memory data pipe (89 % busy, profiles/r02_pixcost_a_raw.csv), and
        // most of its wavefronts are the LDS.128 of the right-pixel pairs, one per plane and voxel pair.  The pair that column j needs for
        // word q is the pair that column j+2 needs for word q+1 (same right pixels, same alignment copy), so a warp takes GC = 4 columns of
        // one parity and a lane two adjacent words of each: 5 table loads per plane serve 8 voxel pairs instead of 8.  The loop over the
        // planes is outermost (the left constants of four columns for one plane at a time), eight accumulators stay in registers.
        constexpr int GC = 4;
        const int ngi = ((TXH + 1) / 2 + GC - 1) / GC; // groups per parity
        for (int gidx = wid; gidx < 2 * ngi; gidx += COST_THREADS / 32) {
            const int par = gidx & 1, j0 = par + 2 * GC * (gidx >> 1);
            if (j0 >= TXH) continue;
            const int mj0 = TXH - 1 - j0; // reversed right index of d = 0 for column j0; column j0 + 2c: mj0 - 2c
            const uint4 *rb = sR + (mj0 & 1) * NRP; // the alignment copy of this parity; entry index = (mj0 >> 1) + word - column offset
            for (int w2 = lane; w2 < DW / 2; w2 += 32) { // words 2*w2, 2*w2+1 = disparities 4*w2 .. 4*w2+3
                const int q0 = 2 * w2;
                uint32_t acc[GC][2];
#pragma unroll
                for (int c = 0; c < GC; c++) acc[c][0] = acc[c][1] = 0;
#pragma unroll
                for (int p = 0; p < NPL; p++) {
                    uint4 E[GC + 1]; // entry e - (GC-1): voxel (column c, word k) uses entry k - c
#pragma unroll
                    for (int e = 0; e < GC + 1; e++) {
                        const int idx = max((mj0 >> 1) + q0 + e - (GC - 1), 0); // (negative only for columns past the tile)
                        E[e] = rb[p * 2 * NRP + (idx >> 1) + (idx & 1) * ((NRP + 1) / 2)];
                    }
#pragma unroll
                    for (int c = 0; c < GC; c++) {
                        const uint4 L4 = sL[min(j0 + 2 * c, TXH - 1) * NPL + p];
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const uint4 R4 = E[k - c + (GC - 1)];
                            const uint32_t V = R4.x, V0 = R4.y, V1 = R4.z;
                            uint32_t c0 = __vimax3_s16x2(L4.x - V1, V0 + L4.y, Bp);
                            uint32_t c1 = __vimax3_s16x2(V + L4.z, L4.w - V, Bp);
                            uint32_t cc = __vmins2(c0, c1);
                            if (p >= CN) cc = (cc >> 2) & 0x007F007Fu;
                            acc[c][k] += cc;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < GC; c++) {
                    const int j = j0 + 2 * c, x1 = xlo + j;
                    if (j >= TXH) continue;
                    const bool inside = x1 >= 0 && x1 < g.width1;
                    uint32_t o[2];
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const int d0 = 2 * (q0 + k);
                        uint32_t v = acc[c][k] - unbias;
                        if (d0 + 1 >= g.D) v &= 0x0000FFFFu;
                        o[k] = (inside && d0 < g.D) ? v : 0u; // padded d-lanes and columns outside the cost volume are zero
                    }
                    *(uint2 *)(pixw + j * DW + q0) = make_uint2(o[0], o[1]);
                }
            }
        }
    } else
    // phase 1: warp -> column j, lane -> disparity pairs d0 = 2*(lane + 32*i)
    for (int j = wid; j < TXH; j += COST_THREADS / 32) {
        const int x1 = xlo + j;
        uint32_t *out = pixw + j * DW;
        if (x1 < 0 || x1 >= g.width1) {
            for (int q = lane; q < DW; q += 32) out[q] = 0;
            continue;
        }
        uint4 Lc[NPL];
#pragma unroll
        for (int p = 0; p < NPL; p++) Lc[p] = sL[j * NPL + p];
        const int mj = TXH - 1 - j; // reversed right index of d = 0
        const uint4 *rbase = LAYOUT == 0 ? sR + (mj & 1) * NRP + (mj >> 1) : sR + mj;
        for (int q = lane; q < DW; q += 32) { // word q = disparities (d0, d1)
            const int d0 = b2s_word_d0(LAYOUT, q), d1 = d0 + (LAYOUT == 0 ? 1 : 8);
            if (d0 >= g.D) { // padded d-lanes are zero in C
                out[q] = 0;
                continue;
            }
            uint32_t acc = 0;
#pragma unroll
            for (int p = 0; p < NPL; p++) {
                const uint4 R4 = rbase[p * 2 * NRP + (LAYOUT == 0 ? q : d0)];
                const uint32_t V = R4.x, V0 = R4.y, V1 = R4.z;
                uint32_t c0 = __vimax3_s16x2(Lc[p].x - V1, V0 + Lc[p].y, Bp);
                uint32_t c1 = __vimax3_s16x2(V + Lc[p].z, Lc[p].w - V, Bp);
                uint32_t c = __vmins2(c0, c1);
                if (p >= CN) c = (c >> 2) & 0x007F007Fu;
                acc += c;
            }
            acc -= unbias;
            if (d1 >= g.D) acc &= 0x0000FFFFu;
            out[q] = acc;
        }
    }
    __syncthreads();

    // phase 2: thread -> (segment of 16 columns, packed d pair); sliding window with wrap arithmetic
    const uint32_t *pw = pixw;
    constexpr int SEG = 16;
    if (xlo >= 0 && x1_0 + TX - 1 + g.SW2 <= g.width1 - 1) {
        // every window of the tile lies inside the cost volume (all tiles but the first and the last of a row): no clamping,
        // the added / dropped columns are two pointers that advance with the output
        const int WIN = 2 * g.SW2 + 1;
        for (int i = threadIdx.x; i < (TX / SEG) * DW; i += COST_THREADS) {
            const int seg = i / DW, w = i % DW;
            const uint32_t *psub = pw + (seg * SEG) * DW + w; // column xs - SW2 of the tile
            uint32_t s = 0;
            for (int k = 0; k < WIN; k++) s = __vadd2(s, psub[k * DW]);
            const uint32_t *padd = psub + WIN * DW;
            uint32_t *out = (uint32_t *)(hs + ((size_t)y * g.width1 + x1_0 + seg * SEG) * Dp) + w;
#pragma unroll
            for (int t = 0; t < SEG; t++) {
                *out = s;
                out += DW;
                if (t + 1 < SEG) s = __vsub2(__vadd2(s, *padd), *psub);
                padd += DW;
                psub += DW;
            }
        }
        return;
    }
    for (int i = threadIdx.x; i < (TX / SEG) * DW; i += COST_THREADS) {
        int seg = i / DW, w = i % DW;
        int xs = x1_0 + seg * SEG;
        if (xs >= g.width1) continue;
        uint32_t s = 0;
        for (int k = -g.SW2; k <= g.SW2; k++) s = __vadd2(s, pw[(clampi(xs + k, 0, g.width1 - 1) - xlo) * DW + w]);
        uint32_t *out = (uint32_t *)(hs + ((size_t)y * g.width1 + xs) * Dp) + w;
        int xe = min(xs + SEG, g.width1);
        for (int x1 = xs; x1 < xe; x1++) {
            *out = s;
            out += DW;
            if (x1 + 1 >= xe) break;
            uint32_t add = pw[(clampi(x1 + 1 + g.SW2, 0, g.width1 - 1) - xlo) * DW + w];
            uint32_t sub = pw[(clampi(x1 - g.SW2, 0, g.width1 - 1) - xlo) * DW + w];
            s = __vsub2(__vadd2(s, add), sub);
        }
    }
}

// ---- census cost (b2s_sgbm_params.cost = 1; this engine's own definition, see oracle/sgbm_ref.c: census_desc) --------------
// descriptor = 62 bits, one per neighbour of the 9 (wide) x 7 (tall) window except the centre, row-major, bit = neighbour <
// centre, coordinates clamped to the image.
__global__ void census_kernel(const uint8_t *__restrict__ img, unsigned long long *__restrict__ desc, int H, int W)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int c = img[(size_t)y * W + x];
    unsigned long long bits = 0;
#pragma unroll
    for (int dy = -3; dy <= 3; dy++) {
        const uint8_t *row = img + (size_t)clampi(y + dy, 0, H - 1) * W;
#pragma unroll
        for (int dx = -4; dx <= 4; dx++) {
            if (dy == 0 && dx == 0) continue;
            bits = (bits << 1) | (unsigned long long)(row[clampi(x + dx, 0, W - 1)] < c);
        }
    }
    desc[(size_t)y * W + x] = bits;
}
// C(y, x1, d) = popcount(descL(x) ^ descR(x - d)); warp = cost column, lane = disparity pairs, written as packed int16x2
__global__ void __launch_bounds__(256) census_cost_kernel(const unsigned long long *__restrict__ dl, const unsigned long long *__restrict__ dr,
                                                          uint32_t *__restrict__ C, SgbmGeom g)
{
    const int lane = threadIdx.x & 31, y = blockIdx.y;
    const int x1 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (x1 >= g.width1) return;
    const int x = x1 + g.minX1, DW = g.Dp / 2;
    const unsigned long long l = dl[(size_t)y * g.W + x];
    const unsigned long long *rrow = dr + (size_t)y * g.W;
    uint32_t *out = C + ((size_t)y * g.width1 + x1) * DW;
    for (int q = lane; q < DW; q += 32) {
        const int d0 = b2s_word_d0(g.layout, q), d1 = d0 + (g.layout == 0 ? 1 : 8);
        uint32_t v = 0;
        if (d0 < g.D) v = __popcll(l ^ rrow[x - (d0 + g.minD)]);
        if (d1 < g.D) v |= (uint32_t)__popcll(l ^ rrow[x - (d1 + g.minD)]) << 16;
        out[q] = v; // padded d-lanes are zero in C
    }
}

// C[y] = sum_{k=-SH2..SH2} hs[clamp(y+k, 0, H-1)]; thread = 8 consecutive int16 (uint4), band of rows per blockIdx.y
constexpr int VBAND = 32;
// ovf (word 1 of the handle's error flags) is set when a block sum wrapped past 32767 (C < 0: bit 15 of a packed half): the packed
// unsigned arithmetic of the aggregation kernels (saturating sums, phase bit of the hand-over rings) needs C >= 0.
__global__ void vsum_kernel(const uint4 *__restrict__ hs, uint4 *__restrict__ C, int H, size_t row_vec, int SH2, int *ovf)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_vec) return;
    int y0 = blockIdx.y * VBAND, y1 = min(y0 + VBAND, H);
    uint4 s = make_uint4(0, 0, 0, 0);
    for (int k = -SH2; k <= SH2; k++) {
        uint4 v = hs[(size_t)clampi(y0 + k, 0, H - 1) * row_vec + i];
        s.x = __vadd2(s.x, v.x); s.y = __vadd2(s.y, v.y); s.z = __vadd2(s.z, v.z); s.w = __vadd2(s.w, v.w);
    }
    uint32_t bad = 0;
    for (int y = y0; y < y1; y++) {
        C[(size_t)y * row_vec + i] = s;
        bad |= s.x | s.y | s.z | s.w;
        uint4 a = hs[(size_t)clampi(y + 1 + SH2, 0, H - 1) * row_vec + i];
        uint4 b = hs[(size_t)clampi(y - SH2, 0, H - 1) * row_vec + i];
        s.x = __vsub2(__vadd2(s.x, a.x), b.x); s.y = __vsub2(__vadd2(s.y, a.y), b.y);
        s.z = __vsub2(__vadd2(s.z, a.z), b.z); s.w = __vsub2(__vadd2(s.w, a.w), b.w);
    }
    if (bad & 0x80008000u) *ovf = 1;
}

} // namespace

cudaError_t launch_cost_volume(b2s_ctx *c, const uint8_t *d_left, const uint8_t *d_right)
{
    const SgbmGeom &g = c->g;
    c->hs_pending = false;
    if (c->prm.cost == 1) { // census: descriptors live in the (8 bytes per gray pixel) plane buffers
        dim3 b(128), gd((g.W + 127) / 128, g.H);
        unsigned long long *DL = c->planesL.as<unsigned long long>(), *DR = c->planesR.as<unsigned long long>();
        census_kernel<<<gd, b, 0, c->stream>>>(d_left, DL, g.H, g.W);
        census_kernel<<<gd, b, 0, c->stream>>>(d_right, DR, g.H, g.W);
        dim3 gc((g.width1 + 7) / 8, g.H);
        census_cost_kernel<<<gc, 256, 0, c->stream>>>(DL, DR, c->C.as<uint32_t>(), g);
        c->launches += 3;
        return cudaGetLastError();
    }
    const int NPL = 2 * g.cn;
    dim3 pb(PLANES_WARPS * 32), pg((g.W + PLANES_WARPS * 30 - 1) / (PLANES_WARPS * 30), g.H);
    uchar4 *PL = c->planesL.as<uchar4>(), *PR = c->planesR.as<uchar4>();
    if (g.cn == 3) {
        planes_kernel<3><<<pg, pb, 0, c->stream>>>(d_left, PL, g.H, g.W, g.ftzero);
        planes_kernel<3><<<pg, pb, 0, c->stream>>>(d_right, PR, g.H, g.W, g.ftzero);
    } else {
        planes_kernel<1><<<pg, pb, 0, c->stream>>>(d_left, PL, g.H, g.W, g.ftzero);
        planes_kernel<1><<<pg, pb, 0, c->stream>>>(d_right, PR, g.H, g.W, g.ftzero);
    }
    const int TXH = TX + 2 * g.SW2, need = g.layout == 0 ? (TXH + g.D - 1) / 2 + 2 : (TXH + g.Dp + 1) / 2;
    if (need > NRP_WIDE) return cudaErrorInvalidValue; // (b2s_api.cu rejects such block sizes with a message)
    const int nrp = need <= 104 ? 104 : (need <= NRP_MAX ? NRP_MAX : NRP_WIDE);
    size_t smem = (size_t)NPL * TXH * sizeof(uint4) + (size_t)NPL * 2 * nrp * sizeof(uint4) + (size_t)TXH * g.Dp * sizeof(int16_t);
    dim3 cg((g.width1 + TX - 1) / TX, g.H);
    // row sums: into S (free until aggregation starts), or into S2 when the first horizontal scan takes over the vertical half of
    // the box filter (sgbm_agg.cu: agg_hscan_vsum_kernel; S is that scan's output, S2 is not written before the vertical sweep)
    const bool fuse_vsum = agg_fuses_vsum(c);
    int16_t *hs = fuse_vsum ? c->S2.as<int16_t>() : c->S.as<int16_t>();
    c->hs_pending = fuse_vsum;
    auto launch = [&](auto kern) -> cudaError_t {
        // the attribute is set once per kernel and device (all instantiations share this lambda's type, hence the set keyed by the
        // kernel's address), to a size that covers every geometry the launcher accepts
        static std::mutex mu;
        static std::set<std::pair<const void *, int>> done;
        {
            std::lock_guard<std::mutex> lock(mu);
            if (!done.count({(const void *)kern, c->device})) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
                if (e != cudaSuccess) return e;
                done.insert({(const void *)kern, c->device});
            }
        }
        if (smem > 160 * 1024) return cudaErrorInvalidValue;
        kern<<<cg, COST_THREADS, smem, c->stream>>>(PL, PR, hs, g);
        return cudaSuccess;
    };
    cudaError_t e;
    if (g.layout == 0) {
        if (nrp == NRP_WIDE) e = g.cn == 3 ? launch(pixcost_hsum_kernel<3, NRP_WIDE, 0>) : launch(pixcost_hsum_kernel<1, NRP_WIDE, 0>);
        else if (g.NP >= 2 && !(getenv("B2S_COST_GROUPED") && atoi(getenv("B2S_COST_GROUPED")) == 0)) { // two words per lane need Dp >= 128
            if (g.cn == 3) e = nrp == 104 ? launch(pixcost_hsum_kernel<3, 104, 0, true>) : launch(pixcost_hsum_kernel<3, NRP_MAX, 0, true>);
            else e = nrp == 104 ? launch(pixcost_hsum_kernel<1, 104, 0, true>) : launch(pixcost_hsum_kernel<1, NRP_MAX, 0, true>);
        } else if (g.cn == 3) e = nrp == 104 ? launch(pixcost_hsum_kernel<3, 104, 0>) : launch(pixcost_hsum_kernel<3, NRP_MAX, 0>);
        else e = nrp == 104 ? launch(pixcost_hsum_kernel<1, 104, 0>) : launch(pixcost_hsum_kernel<1, NRP_MAX, 0>);
    } else {
        if (g.cn == 3) e = nrp == 104 ? launch(pixcost_hsum_kernel<3, 104, 1>) : launch(pixcost_hsum_kernel<3, NRP_MAX, 1>);
        else e = nrp == 104 ? launch(pixcost_hsum_kernel<1, 104, 1>) : launch(pixcost_hsum_kernel<1, NRP_MAX, 1>);
    }
    if (e != cudaSuccess) return e;
    c->launches += 3;
    if (fuse_vsum) return cudaGetLastError();
    size_t row_vec = (size_t)g.width1 * g.Dp / 8;
    dim3 vg((unsigned)((row_vec + 255) / 256), (g.H + VBAND - 1) / VBAND);
    cudaError_t ee = agg_error_flags(c);
    if (ee != cudaSuccess) return ee;
    vsum_kernel<<<vg, 256, 0, c->stream>>>((const uint4 *)hs, c->C.as<uint4>(), g.H, row_vec, g.SH2, c->agg_err + 1);
    c->launches++;
    if (g.mode == 3 && g.SH2 > 0 && g.H > 1) {
        // MODE_HH4 of cv2 leaves the cost of the rows whose window reaches below the image (y > 0, y + SH2 >= H) constant:
        // its column-parallel loop skips their update (found by differential testing; oracle/sgbm_ref.c, vertical half of A.3)
        const int y0 = g.H - g.SH2 > 1 ? g.H - g.SH2 : 1;
        cudaError_t em = cudaMemsetAsync(c->C.as<int16_t>() + (size_t)y0 * g.width1 * g.Dp, 0, (size_t)(g.H - y0) * g.width1 * g.Dp * sizeof(int16_t), c->stream);
        if (em != cudaSuccess) return em;
    }
    return cudaGetLastError();
}
