// sgbm_cost.cu -- cost volume C(y, x1, d) of cv2.StereoSGBM, sm_100a.
//
// Replaces the calcPixelCostBT + box-sum half of cv2.StereoSGBM.compute, called from
// calibrating/stereo_matching.py:63.  All sums are int16 with two's-complement wrap (packed VIADD.16x2), which
// makes running sums exact whatever the order.
//
// Layout in HBM: planes  (H, 2cn, W) uchar4 {value, lo, hi, 0}   lo/hi = Birchfield-Tomasi half-sample interval
//                hs, C   (H, width1, Dp) int16, d fastest, Dp = 64*NP (padded d-lanes are zero in C)
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

template <int CN>
__global__ void planes_kernel(const uint8_t *__restrict__ img, uchar4 *__restrict__ out, int H, int W, int ftzero)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W) return;
#pragma unroll
    for (int p = 0; p < 2 * CN; p++) {
        int u = plane_value<CN>(img, H, W, y, x, p, ftzero);
        int ul = x > 0 ? (u + plane_value<CN>(img, H, W, y, x - 1, p, ftzero)) / 2 : u;
        int ur = x < W - 1 ? (u + plane_value<CN>(img, H, W, y, x + 1, p, ftzero)) / 2 : u;
        out[((size_t)y * 2 * CN + p) * W + x] = make_uchar4((unsigned char)u, (unsigned char)min(min(ul, ur), u),
                                                            (unsigned char)max(max(ul, ur), u), 0);
    }
}

// One CTA = one image row y and TX cost columns.  Phase 1: BT pixel cost for TX+2*SW2 columns into shared memory.
// Phase 2: horizontal box sum (window clamped in cost-volume coordinates) -> hs.
constexpr int TX = 64;
constexpr int COST_THREADS = 256;

template <int CN>
__global__ void __launch_bounds__(COST_THREADS) pixcost_hsum_kernel(const uchar4 *__restrict__ PL, const uchar4 *__restrict__ PR,
                                                                    int16_t *__restrict__ hs, SgbmGeom g)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int NPL = 2 * CN;
    const int y = blockIdx.y;
    const int x1_0 = blockIdx.x * TX;
    const int TXH = TX + 2 * g.SW2;
    const int xlo = x1_0 - g.SW2;              // cost column of tile-local index 0
    const int NR = TXH + g.D - 1;              // right-image columns needed
    uchar4 *sL = (uchar4 *)smem;               // [NPL][TXH]
    uchar4 *sR = sL + NPL * TXH;               // [NPL][NR]   index j <-> image column xr0 + j
    int16_t *pix = (int16_t *)(sR + NPL * NR); // [TXH][Dp]
    const int xr0 = xlo + g.minX1 - g.minD - (g.D - 1);

    for (int i = threadIdx.x; i < NPL * TXH; i += COST_THREADS) {
        int p = i / TXH, j = i % TXH;
        int x = clampi(xlo + j + g.minX1, 0, g.W - 1);
        sL[i] = PL[((size_t)y * NPL + p) * g.W + x];
    }
    for (int i = threadIdx.x; i < NPL * NR; i += COST_THREADS) {
        int p = i / NR, j = i % NR;
        int x = clampi(xr0 + j, 0, g.W - 1);
        sR[i] = PR[((size_t)y * NPL + p) * g.W + x];
    }
    __syncthreads();

    // phase 1: thread -> (column j, disparity d); consecutive threads = consecutive d
    const int Dp = g.Dp;
    for (int i = threadIdx.x; i < TXH * Dp; i += COST_THREADS) {
        int j = i / Dp, d = i % Dp;
        int x1 = xlo + j;
        int cost = 0;
        if (d < g.D && x1 >= 0 && x1 < g.width1) {
            int jr = j + (g.D - 1) - d; // image column (x1+minX1) - (d+minD), relative to xr0
#pragma unroll
            for (int p = 0; p < NPL; p++) {
                uchar4 L = sL[p * TXH + j];
                uchar4 R = sR[p * NR + jr];
                int u = L.x, u0 = L.y, u1 = L.z, v = R.x, v0 = R.y, v1 = R.z;
                int c0 = max(0, max(u - v1, v0 - u));
                int c1 = max(0, max(v - u1, u0 - v));
                cost += min(c0, c1) >> (p < CN ? 0 : 2);
            }
        }
        pix[i] = (int16_t)cost;
    }
    __syncthreads();

    // phase 2: thread -> (segment of 16 columns, packed d pair); sliding window with wrap arithmetic
    const int DW = Dp / 2; // 32-bit words per column
    const uint32_t *pw = (const uint32_t *)pix;
    const int SEG = 16;
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

// C[y] = sum_{k=-SH2..SH2} hs[clamp(y+k, 0, H-1)]; thread = 8 consecutive int16 (uint4), band of rows per blockIdx.y
constexpr int VBAND = 32;
__global__ void vsum_kernel(const uint4 *__restrict__ hs, uint4 *__restrict__ C, int H, size_t row_vec, int SH2)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_vec) return;
    int y0 = blockIdx.y * VBAND, y1 = min(y0 + VBAND, H);
    uint4 s = make_uint4(0, 0, 0, 0);
    for (int k = -SH2; k <= SH2; k++) {
        uint4 v = hs[(size_t)clampi(y0 + k, 0, H - 1) * row_vec + i];
        s.x = __vadd2(s.x, v.x); s.y = __vadd2(s.y, v.y); s.z = __vadd2(s.z, v.z); s.w = __vadd2(s.w, v.w);
    }
    for (int y = y0; y < y1; y++) {
        C[(size_t)y * row_vec + i] = s;
        uint4 a = hs[(size_t)clampi(y + 1 + SH2, 0, H - 1) * row_vec + i];
        uint4 b = hs[(size_t)clampi(y - SH2, 0, H - 1) * row_vec + i];
        s.x = __vsub2(__vadd2(s.x, a.x), b.x); s.y = __vsub2(__vadd2(s.y, a.y), b.y);
        s.z = __vsub2(__vadd2(s.z, a.z), b.z); s.w = __vsub2(__vadd2(s.w, a.w), b.w);
    }
}

} // namespace

cudaError_t launch_cost_volume(b2s_ctx *c, const uint8_t *d_left, const uint8_t *d_right)
{
    const SgbmGeom &g = c->g;
    const int NPL = 2 * g.cn;
    dim3 pb(128), pg((g.W + 127) / 128, g.H);
    uchar4 *PL = c->planesL.as<uchar4>(), *PR = c->planesR.as<uchar4>();
    if (g.cn == 3) {
        planes_kernel<3><<<pg, pb, 0, c->stream>>>(d_left, PL, g.H, g.W, g.ftzero);
        planes_kernel<3><<<pg, pb, 0, c->stream>>>(d_right, PR, g.H, g.W, g.ftzero);
    } else {
        planes_kernel<1><<<pg, pb, 0, c->stream>>>(d_left, PL, g.H, g.W, g.ftzero);
        planes_kernel<1><<<pg, pb, 0, c->stream>>>(d_right, PR, g.H, g.W, g.ftzero);
    }
    const int TXH = TX + 2 * g.SW2, NR = TXH + g.D - 1;
    size_t smem = (size_t)NPL * (TXH + NR) * sizeof(uchar4) + (size_t)TXH * g.Dp * sizeof(int16_t);
    dim3 cg((g.width1 + TX - 1) / TX, g.H);
    int16_t *hs = c->S.as<int16_t>(); // S is free until aggregation starts
    cudaError_t e;
    if (g.cn == 3) {
        e = cudaFuncSetAttribute(pixcost_hsum_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pixcost_hsum_kernel<3><<<cg, COST_THREADS, smem, c->stream>>>(PL, PR, hs, g);
    } else {
        e = cudaFuncSetAttribute(pixcost_hsum_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        pixcost_hsum_kernel<1><<<cg, COST_THREADS, smem, c->stream>>>(PL, PR, hs, g);
    }
    size_t row_vec = (size_t)g.width1 * g.Dp / 8;
    dim3 vg((unsigned)((row_vec + 255) / 256), (g.H + VBAND - 1) / VBAND);
    vsum_kernel<<<vg, 256, 0, c->stream>>>((const uint4 *)hs, c->C.as<uint4>(), g.H, row_vec, g.SH2);
    c->launches += 4;
    return cudaGetLastError();
}
