// remap.cu -- the remap family of Stereo.get_depth, sm_100a.
//   Stereo.rectify          calibrating/stereo_camera.py:216-242   cv2.remap(u8, f32 maps, INTER_LANCZOS4) + right-image shift
//   Stereo.undistort_img    calibrating/stereo_camera.py:430-431   cv2.undistort == CV_16SC2 fixed maps + bilinear
//   disparity_to_depth      calibrating/stereo_camera.py:408-413   (+ the += min_disparity / valid-mask lines :510-512)
//   unrectify_depth         calibrating/utils.py:173-200           z' = z*(m.[x,y,1]) then INTER_NEAREST remap
// cv2 semantics restated: coordinates quantised to 1/32 px with round-half-even, 15-bit
// fixed-point weight tables summing to exactly 32768, rounding (sum + 2^14) >> 15, constant-0 border.
#include <math.h>
#include <string.h>

#include <stdlib.h>

#include <mutex>

#include "b2s_internal.h"

// ---- host: cv::initInterTab2D(INTER_LANCZOS4, fixpt) restated ------------------------------------------------------
static void lanczos4_coeffs(float x, float *co)
{
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    if (x < 1.1920928955078125e-07f) {
        for (int i = 0; i < 8; i++) co[i] = 0;
        co[3] = 1;
        return;
    }
    float sum = 0;
    double y0 = -(x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; i++) {
        double y = -(x + 3 - i) * 3.1415926535897932384626433832795 * 0.25;
        co[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        sum += co[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; i++) co[i] *= sum;
}

void build_lanczos4_table(int16_t *tab)
{
    float t1[32][8];
    for (int i = 0; i < 32; i++) lanczos4_coeffs((float)i * (1.f / 32), t1[i]);
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            int16_t *it = tab + (size_t)(i * 32 + j) * 64;
            int isum = 0;
            for (int a = 0; a < 8; a++)
                for (int b = 0; b < 8; b++) {
                    float v = t1[i][a] * t1[j][b];
                    long r = lrintf(v * 32768.f);
                    r = r > 32767 ? 32767 : (r < -32768 ? -32768 : r);
                    it[a * 8 + b] = (int16_t)r;
                    isum += (int)r;
                }
            if (isum != 32768) {
                int diff = isum - 32768;
                int Mk = 4 * 8 + 4, mk = 4 * 8 + 4;
                for (int a = 4; a < 6; a++)
                    for (int b = 4; b < 6; b++) {
                        if (it[a * 8 + b] < it[mk]) mk = a * 8 + b;
                        else if (it[a * 8 + b] > it[Mk]) Mk = a * 8 + b;
                    }
                if (diff < 0) it[Mk] = (int16_t)(it[Mk] - diff);
                else it[mk] = (int16_t)(it[mk] - diff);
            }
        }
}

namespace {

__device__ __forceinline__ int sat_short(int v) { return min(max(v, -32768), 32767); }
__device__ __forceinline__ unsigned char fix_cast(int sum) { return (unsigned char)min(max((sum + (1 << 14)) >> 15, 0), 255); }

// two taps per instruction: d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1 (IDP.2A.LO.S16.U8) / bytes 2, 3 (HI)
__device__ __forceinline__ int dp2a_lo(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// thread = one destination pixel (all channels).  xshift: destination column x samples the map at x - xshift
// (Stereo.rectify's translation of rectify_img2, stereo_camera.py:230-240); vacated columns are 0.
template <int CN, int INTERP>
__global__ void remap_u8_kernel(const uint8_t *__restrict__ src, int sH, int sW, const float *__restrict__ mapx,
                                const float *__restrict__ mapy, int dH, int dW, int xshift, const int16_t *__restrict__ tab,
                                uint8_t *__restrict__ dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dW) return;
    uint8_t *o = dst + ((size_t)y * dW + x) * CN;
    int xm = x - xshift;
    if (xm < 0 || xm >= dW) {
#pragma unroll
        for (int c = 0; c < CN; c++) o[c] = 0;
        return;
    }
    size_t mi = (size_t)y * dW + xm;
    int sx = __float2int_rn(mapx[mi] * 32.f), sy = __float2int_rn(mapy[mi] * 32.f);
    int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
    int fx = sx & 31, fy = sy & 31;
    int acc[CN];
#pragma unroll
    for (int c = 0; c < CN; c++) acc[c] = 0;
    if (INTERP == 0) {
        const int x0 = ix - 3, y0 = iy - 3;
        if (x0 >= sW || x0 + 8 <= 0 || y0 >= sH || y0 + 8 <= 0) {
#pragma unroll
            for (int c = 0; c < CN; c++) o[c] = 0;
            return;
        }
        const uint4 *w4 = (const uint4 *)(tab + (size_t)(fy * 32 + fx) * 64);
        const bool inside = x0 >= 0 && x0 + 8 <= sW && y0 >= 0 && y0 + 8 <= sH;
        if (inside) {
            // whole 8x8 window inside the image (all but a border of 4 pixels): a row of the window is 8*CN contiguous bytes at an
            // arbitrary byte address (sW * CN is not a multiple of 4 in general, so the alignment changes from row to row).
            // Read it as aligned 32-bit words, realign with funnel shifts, and take two taps per IDP.2A: the packed int16 weight
            // pairs as they lie in the table times two u8 samples gathered by one PRMT.
            constexpr int NW = 2 * CN; // 32-bit words of one window row
            const size_t a0 = ((size_t)y0 * sW + x0) * CN, rstep = (size_t)sW * CN;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const uint4 wq = w4[r];
                const int wpair[4] = {(int)wq.x, (int)wq.y, (int)wq.z, (int)wq.w}; // taps (0,1) (2,3) (4,5) (6,7)
                const size_t a = a0 + r * rstep;
                const uint32_t *wp = (const uint32_t *)(src + (a & ~(size_t)3));
                const unsigned s = (unsigned)(a & 3) * 8;
                uint32_t W[NW + 1], B[NW];
#pragma unroll
                for (int i = 0; i < NW; i++) W[i] = __ldg(wp + i);
                W[NW] = s ? __ldg(wp + NW) : 0u; // (an aligned row ends with its last word: nothing is read past the window)
#pragma unroll
                for (int i = 0; i < NW; i++) B[i] = __funnelshift_r(W[i], W[i + 1], s);
                if (CN == 1) {
                    acc[0] = dp2a_lo(wpair[0], B[0], acc[0]);
                    acc[0] = dp2a_hi(wpair[1], B[0], acc[0]);
                    acc[0] = dp2a_lo(wpair[2], B[1], acc[0]);
                    acc[0] = dp2a_hi(wpair[3], B[1], acc[0]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int c = 0; c < CN; c++) {
                            const int ja = 2 * q * CN + c, jb = ja + CN; // bytes of taps 2q and 2q+1 of channel c in the window row
                            const uint32_t pr = __byte_perm(B[ja >> 2], B[jb >> 2], (ja & 3) | (((jb & 3) + 4) << 4));
                            acc[c] = dp2a_lo(wpair[q], pr, acc[c]);
                        }
                }
            }
        } else {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint4 wq = w4[r];
            int w[8] = {(short)(wq.x & 0xffff), ((int)wq.x) >> 16, (short)(wq.y & 0xffff), ((int)wq.y) >> 16,
                        (short)(wq.z & 0xffff), ((int)wq.z) >> 16, (short)(wq.w & 0xffff), ((int)wq.w) >> 16};
            int yy = y0 + r;
            if (yy < 0 || yy >= sH) continue;
            const uint8_t *row = src + ((size_t)yy * sW + x0) * CN;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (x0 + k < 0 || x0 + k >= sW) continue;
#pragma unroll
                for (int c = 0; c < CN; c++) acc[c] += (int)row[k * CN + c] * w[k];
            }
        }
        }
    } else {
        if (ix >= sW || ix + 2 <= 0 || iy >= sH || iy + 2 <= 0) {
#pragma unroll
            for (int c = 0; c < CN; c++) o[c] = 0;
            return;
        }
        // (1-fx)(1-fy)*32768 etc. are exact integers for fx,fy multiples of 1/32
        int w[4] = {(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
#pragma unroll
        for (int r = 0; r < 2; r++) {
            int yy = iy + r;
            if (yy < 0 || yy >= sH) continue;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                int xx = ix + k;
                if (xx < 0 || xx >= sW) continue;
                const uint8_t *p = src + ((size_t)yy * sW + xx) * CN;
#pragma unroll
                for (int c = 0; c < CN; c++) acc[c] += (int)p[c] * w[r * 2 + k];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CN; c++) o[c] = fix_cast(acc[c]);
}

// ---- LANCZOS4, production version: persistent CTAs, the fixed-point weight table in shared memory ------------------------
// remap_u8_kernel<CN, 0> above is bound by its weight fetch: every pixel reads the 128-byte row (fy*32+fx) of the 128 KB table,
// 8 x LDG.128 per thread with 32 different cache lines per warp instruction.  Here every CTA first pulls the whole table into
// shared memory with bulk copies (cp.async.bulk -> UBLKCP, one mbarrier), rows padded to 144 bytes so that the 8 lanes of a
// quarter-warp reading the same 16-byte column of 8 different rows hit 8 different bank groups, then walks destination tiles of
// 32 x RL_WARPS pixels (warp = 32 consecutive pixels of a row: their 8-row source windows overlap, and so do those of the
// neighbouring rows held by the other warps of the CTA, which keeps the source reads in L1).
// ANALYTIC (rig set by parameters, b2s_set_rig_params): the map pixel is evaluated on the spot in float64 -- cv2's operation
// order, rounded to float32 exactly like the stored planes -- so that no map plane is read at all (the
// fused undistort + rectify remap).
__device__ __forceinline__ void map_eval(const b2s_map_params &p, int j, int i, double &u, double &v); // (below, with gen_maps_kernel)
constexpr int RL_WARPS = 16;
constexpr int RL_ROWB = 144;                 // padded bytes of one table row
constexpr int RL_TABB = 1024 * RL_ROWB;      // 147,456 bytes
template <int CN, bool ANALYTIC>
__global__ void __launch_bounds__(RL_WARPS * 32, 1) remap_lz4_kernel(const uint8_t *__restrict__ src, int sH, int sW, const float *__restrict__ mapx,
                                                                    const float *__restrict__ mapy, b2s_map_params prm, int dH, int dW, int xshift,
                                                                    const unsigned char *__restrict__ tabp, uint8_t *__restrict__ dst)
{
    extern __shared__ __align__(128) unsigned char rl_smem[];
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(rl_smem), mb = sm0 + RL_TABB;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(mb), "r"((uint32_t)RL_TABB) : "memory");
#pragma unroll 1
        for (int k = 0; k < RL_TABB; k += 16384) {
            const uint32_t bytes = min(16384, RL_TABB - k);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm0 + k), "l"(tabp + k), "r"(bytes), "r"(mb)
                         : "memory");
        }
    }
    __syncthreads();
    {
        uint32_t ok;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb) : "memory");
        } while (!ok);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tx_n = (dW + 31) / 32, ty_n = (dH + RL_WARPS - 1) / RL_WARPS;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < tx_n * ty_n; tile += gridDim.x) {
        const int x = (tile % tx_n) * 32 + lane, y = (tile / tx_n) * RL_WARPS + wid;
        if (x >= dW || y >= dH) continue;
        uint8_t *o = dst + ((size_t)y * dW + x) * CN;
        const int xm = x - xshift;
        int acc[CN];
#pragma unroll
        for (int c = 0; c < CN; c++) acc[c] = 0;
        bool zero = xm < 0 || xm >= dW;
        int x0 = 0, y0 = 0, fx = 0, fy = 0;
        if (!zero) {
            float mu, mv;
            if (ANALYTIC) {
                double u, v;
                map_eval(prm, xm, y, u, v);
                mu = (float)u;
                mv = (float)v;
            } else {
                const size_t mi = (size_t)y * dW + xm;
                mu = mapx[mi];
                mv = mapy[mi];
            }
            const int sx = __float2int_rn(mu * 32.f), sy = __float2int_rn(mv * 32.f);
            x0 = sat_short(sx >> 5) - 3;
            y0 = sat_short(sy >> 5) - 3;
            fx = sx & 31;
            fy = sy & 31;
            zero = x0 >= sW || x0 + 8 <= 0 || y0 >= sH || y0 + 8 <= 0;
        }
        if (zero) {
#pragma unroll
            for (int c = 0; c < CN; c++) o[c] = 0;
            continue;
        }
        const uint32_t wrow = sm0 + (uint32_t)(fy * 32 + fx) * RL_ROWB;
        if (x0 >= 0 && x0 + 8 <= sW && y0 >= 0 && y0 + 8 <= sH) {
            constexpr int NW = 2 * CN; // 32-bit words of one window row
            const size_t a0 = ((size_t)y0 * sW + x0) * CN, rstep = (size_t)sW * CN;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                int wpair[4]; // taps (0,1) (2,3) (4,5) (6,7) as packed int16 pairs
                asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(wpair[0]), "=r"(wpair[1]), "=r"(wpair[2]), "=r"(wpair[3]) : "r"(wrow + r * 16));
                const size_t a = a0 + r * rstep;
                const uint32_t *wp = (const uint32_t *)(src + (a & ~(size_t)3));
                const unsigned s = (unsigned)(a & 3) * 8;
                uint32_t Wd[NW + 1], B[NW];
#pragma unroll
                for (int i = 0; i < NW; i++) Wd[i] = __ldg(wp + i);
                Wd[NW] = s ? __ldg(wp + NW) : 0u; // (an aligned row ends with its last word: nothing is read past the window)
#pragma unroll
                for (int i = 0; i < NW; i++) B[i] = __funnelshift_r(Wd[i], Wd[i + 1], s);
                if (CN == 1) {
                    acc[0] = dp2a_lo(wpair[0], B[0], acc[0]);
                    acc[0] = dp2a_hi(wpair[1], B[0], acc[0]);
                    acc[0] = dp2a_lo(wpair[2], B[1], acc[0]);
                    acc[0] = dp2a_hi(wpair[3], B[1], acc[0]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int c = 0; c < CN; c++) {
                            const int ja = 2 * q * CN + c, jb = ja + CN; // bytes of taps 2q and 2q+1 of channel c in the window row
                            const uint32_t pr = __byte_perm(B[ja >> 2], B[jb >> 2], (ja & 3) | (((jb & 3) + 4) << 4));
                            acc[c] = dp2a_lo(wpair[q], pr, acc[c]);
                        }
                }
            }
        } else { // window crosses the image border: taps outside contribute 0 (BORDER_CONSTANT)
#pragma unroll 1
            for (int r = 0; r < 8; r++) {
                const int yy = y0 + r;
                if (yy < 0 || yy >= sH) continue;
                const uint8_t *row = src + ((size_t)yy * sW + x0) * CN;
#pragma unroll 1
                for (int k = 0; k < 8; k++) {
                    if (x0 + k < 0 || x0 + k >= sW) continue;
                    short wv;
                    asm volatile("ld.shared.s16 %0, [%1];" : "=h"(wv) : "r"(wrow + (r * 8 + k) * 2));
#pragma unroll
                    for (int c = 0; c < CN; c++) acc[c] += (int)row[k * CN + c] * (int)wv;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CN; c++) o[c] = fix_cast(acc[c]);
    }
}

// cv2.undistort: pre-quantised CV_16SC2 maps (xy = integer part, fxy = fy*32+fx), bilinear, border 0
template <int CN>
__global__ void undistort_u8_kernel(const uint8_t *__restrict__ src, int H, int W, const short2 *__restrict__ xy,
                                    const uint16_t *__restrict__ fxy, uint8_t *__restrict__ dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t i = (size_t)y * W + x;
    short2 p = xy[i];
    int f = fxy[i] & 1023, fx = f & 31, fy = f >> 5;
    int ix = p.x, iy = p.y;
    uint8_t *o = dst + i * CN;
    int acc[CN];
#pragma unroll
    for (int c = 0; c < CN; c++) acc[c] = 0;
    if (!(ix >= W || ix + 2 <= 0 || iy >= H || iy + 2 <= 0)) {
        int w[4] = {(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
#pragma unroll
        for (int r = 0; r < 2; r++) {
            int yy = iy + r;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                int xx = ix + k;
                if (xx < 0 || xx >= W) continue;
                const uint8_t *q = src + ((size_t)yy * W + xx) * CN;
#pragma unroll
                for (int c = 0; c < CN; c++) acc[c] += (int)q[c] * w[r * 2 + k];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CN; c++) o[c] = fix_cast(acc[c]);
}

// stereo_camera.py:510-513: disparity += min_disparity (every pixel); *= valid mask; depth = fx*B/disparity in f64
// (np.float64 scalar / f32 array under NumPy 2), > max_depth -> 0, < 0 -> 0 (inf > max_depth -> 0).
__global__ void disp_to_depth_kernel(const float *__restrict__ disp_in, const uint8_t *__restrict__ mask, float *__restrict__ disp_out,
                                     double *__restrict__ depth, size_t n, float add, double fxb, double max_depth)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d = disp_in[i] + add;
    if (mask) d = mask[i] ? d : d * 0.0f;
    if (disp_out) disp_out[i] = d;
    double z = fxb / (double)d;
    if (z > max_depth) z = 0.0;
    if (z < 0.0) z = 0.0;
    if (z != z) z = 0.0 * z; // keep NaN as NaN like NumPy (0/0 cannot occur: fxb > 0)
    depth[i] = z;
}

// utils.rotate_depth_by_remap: new_z(x,y) = z(x,y)*(m0*x + m1*y + m2); out(u,v) = new_z[rint(mapy), rint(mapx)], outside -> 0
__global__ void unrectify_kernel(const double *__restrict__ depth, int H, int W, const float *__restrict__ mapx,
                                 const float *__restrict__ mapy, double *__restrict__ out, int H1, int W1, double m0, double m1, double m2)
{
    int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= W1) return;
    size_t i = (size_t)v * W1 + u;
    int sx = sat_short(__float2int_rn(mapx[i])), sy = sat_short(__float2int_rn(mapy[i]));
    double r = 0.0;
    if (sx >= 0 && sx < W && sy >= 0 && sy < H) {
        double z = depth[(size_t)sy * W + sx];
        r = (m0 * ((double)sx * z) + m1 * ((double)sy * z)) + m2 * z;
    }
    out[i] = r;
}

} // namespace

namespace {
// cv2.initUndistortRectifyMap restated, float64 without fused multiply-add so that every
// operation rounds exactly where OpenCV's scalar code does; thread = one map pixel.
// one map pixel (j, i): the float64 source coordinates (u, v) of cv2.initUndistortRectifyMap
__device__ __forceinline__ void map_eval(const b2s_map_params &p, int j, int i, double &u, double &v)
{
    const double dj = (double)j, di = (double)i;
    const double X = __dadd_rn(__dmul_rn(dj, p.iR[0]), __dadd_rn(__dmul_rn(di, p.iR[1]), p.iR[2]));
    const double Y = __dadd_rn(__dmul_rn(dj, p.iR[3]), __dadd_rn(__dmul_rn(di, p.iR[4]), p.iR[5]));
    const double Wh = __dadd_rn(__dmul_rn(dj, p.iR[6]), __dadd_rn(__dmul_rn(di, p.iR[7]), p.iR[8]));
    const double w = __ddiv_rn(1.0, Wh), x = __dmul_rn(X, w), y = __dmul_rn(Y, w);
    const double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y), r2 = __dadd_rn(x2, y2), _2xy = __dmul_rn(__dmul_rn(2.0, x), y);
    const double k1 = p.k[0], k2 = p.k[1], p1 = p.k[2], p2 = p.k[3], k3 = p.k[4], k4 = p.k[5], k5 = p.k[6], k6 = p.k[7];
    const double s1 = p.k[8], s2 = p.k[9], s3 = p.k[10], s4 = p.k[11];
    auto poly = [&](double a, double b, double c) { // 1 + ((a*r2 + b)*r2 + c)*r2
        return __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(a, r2), b), r2), c), r2));
    };
    const double kr = __ddiv_rn(poly(k3, k2, k1), poly(k6, k5, k4));
    double xd = __dadd_rn(__dmul_rn(x, kr), __dmul_rn(p1, _2xy));
    xd = __dadd_rn(xd, __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(2.0, x2))));
    xd = __dadd_rn(xd, __dmul_rn(s1, r2));
    xd = __dadd_rn(xd, __dmul_rn(__dmul_rn(s2, r2), r2));
    double yd = __dadd_rn(__dmul_rn(y, kr), __dmul_rn(p1, __dadd_rn(r2, __dmul_rn(2.0, y2))));
    yd = __dadd_rn(yd, __dmul_rn(p2, _2xy));
    yd = __dadd_rn(yd, __dmul_rn(s3, r2));
    yd = __dadd_rn(yd, __dmul_rn(__dmul_rn(s4, r2), r2));
    u = __dadd_rn(__dmul_rn(p.fx, xd), p.cx);
    v = __dadd_rn(__dmul_rn(p.fy, yd), p.cy);
}

__global__ void gen_maps_kernel(b2s_map_params p, float *__restrict__ mapx, float *__restrict__ mapy, uint8_t *__restrict__ mask,
                                int mW, int mH, short2 *__restrict__ xy16, unsigned short *__restrict__ fxy16)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= p.W) return;
    double u, v;
    map_eval(p, j, i, u, v);
    const size_t o = (size_t)i * p.W + j;
    if (mapx) {
        const float fu = (float)u, fv = (float)v;
        mapx[o] = fu;
        mapy[o] = fv;
        // rectify_valid_mask1 (stereo_camera.py:167-176), on the float32 maps like the reference
        if (mask) mask[o] = (-0.5f < fu) && (fu < (float)mW - 0.5f) && (-0.5f < fv) && (fv < (float)mH - 0.5f);
    }
    if (xy16) { // CV_16SC2 + interpolation-table index, as cv2 builds them for cv2.undistort (from the float64 values)
        const int iu = __double2int_rn(__dmul_rn(u, 32.0)), iv = __double2int_rn(__dmul_rn(v, 32.0));
        xy16[o] = make_short2((short)sat_short(iu >> 5), (short)sat_short(iv >> 5));
        fxy16[o] = (unsigned short)((iv & 31) * 32 + (iu & 31));
    }
}
} // namespace

namespace {
// cv2.undistortPoints(points f32, K, None) then cv2.projectPoints(., 0, 0, K, D) restated: float64 in OpenCV's
// operation order without FMA, rounded to float32 where OpenCV stores float32 (the normalised point and the image point),
// then the reference's astype(int32) truncation.  Winner per target = smallest source index (atomicMin).
__global__ void distort_scatter_kernel(int W, int H, double fx, double fy, double cx, double cy, const double *__restrict__ kk,
                                       unsigned *__restrict__ key)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= W) return;
    const double ifx = __ddiv_rn(1.0, fx), ify = __ddiv_rn(1.0, fy);
    const double x = (double)(float)__dmul_rn(__dadd_rn((double)u, -cx), ifx), y = (double)(float)__dmul_rn(__dadd_rn((double)v, -cy), ify);
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), r4 = __dmul_rn(r2, r2), r6 = __dmul_rn(r4, r2);
    const double a1 = __dmul_rn(__dmul_rn(2.0, x), y), a2 = __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x)),
                 a3 = __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y));
    const double cdist = __dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(kk[0], r2)), __dmul_rn(kk[1], r4)), __dmul_rn(kk[4], r6));
    const double icdist2 = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(kk[5], r2)), __dmul_rn(kk[6], r4)), __dmul_rn(kk[7], r6)));
    double xd = __dadd_rn(__dmul_rn(__dmul_rn(x, cdist), icdist2), __dmul_rn(kk[2], a1));
    xd = __dadd_rn(__dadd_rn(__dadd_rn(xd, __dmul_rn(kk[3], a2)), __dmul_rn(kk[8], r2)), __dmul_rn(kk[9], r4));
    double yd = __dadd_rn(__dmul_rn(__dmul_rn(y, cdist), icdist2), __dmul_rn(kk[2], a3));
    yd = __dadd_rn(__dadd_rn(__dadd_rn(yd, __dmul_rn(kk[3], a1)), __dmul_rn(kk[10], r2)), __dmul_rn(kk[11], r4));
    const float mx = (float)__dadd_rn(__dmul_rn(xd, fx), cx), my = (float)__dadd_rn(__dmul_rn(yd, fy), cy);
    const int tx = (int)mx, ty = (int)my; // astype(np.int32): truncation toward zero
    if (mx <= -1.f || my <= -1.f || tx < 0 || ty < 0 || tx >= W || ty >= H) return;
    atomicMin(&key[(size_t)ty * W + tx], (unsigned)(v * W + u));
}
__global__ void distort_gather_kernel(const double *__restrict__ depth, const unsigned *__restrict__ key, double *__restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned k = key[i];
    out[i] = k == 0xFFFFFFFFu ? 0.0 : depth[k];
}
} // namespace

cudaError_t launch_distort_depth(b2s_ctx *c, const double *d_depth, double *d_out)
{
    const int W = c->rW1, H = c->rH1;
    const size_t n = (size_t)W * H;
    cudaError_t e = c->dkey.ensure(n * 4 + 12 * 8 + 8);
    if (e != cudaSuccess) return e;
    double *kk = (double *)((char *)c->dkey.p + n * 4 + (8 - (n * 4) % 8) % 8);
    if ((e = cudaMemsetAsync(c->dkey.p, 0xFF, n * 4, c->stream)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(kk, c->cam1_k, 12 * 8, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return e;
    dim3 b(128), g((W + 127) / 128, H);
    distort_scatter_kernel<<<g, b, 0, c->stream>>>(W, H, c->cam1_f[0], c->cam1_f[1], c->cam1_f[2], c->cam1_f[3], kk, c->dkey.as<unsigned>());
    distort_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_depth, c->dkey.as<unsigned>(), d_out, n);
    c->launches += 2;
    return cudaGetLastError();
}

namespace {
struct ProjArgs {
    double Kinv[9], T[12], K1[9];
    double rate, sx, sy; // sx = W2 / W2u, sy = H2 / H2u: cv2.resize's source-per-destination scale
    int W2, H2, W2u, H2u, W1, H1;
};
// order-preserving map double -> uint64 (smaller double = smaller key), and back
__device__ __forceinline__ unsigned long long dkey(double z)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(z);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k)
{
    return __longlong_as_double((long long)((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
}
// thread = one sample of the up-sampled depth image
__global__ void project_scatter_kernel(const double *__restrict__ depth2, ProjArgs p, unsigned long long *__restrict__ key)
{
    const int uu = blockIdx.x * blockDim.x + threadIdx.x, vv = blockIdx.y;
    if (uu >= p.W2u) return;
    // cv2.resize INTER_NEAREST: source index = min(floor(dst * scale), size - 1)
    const int su = min((int)floor(uu * p.sx), p.W2 - 1), sv = min((int)floor(vv * p.sy), p.H2 - 1);
    const double z = depth2[(size_t)sv * p.W2 + su];
    if (z == 0.0) return;
    const double u = (double)uu / p.rate, v = (double)vv / p.rate;
    const double a0 = u * z, a1 = v * z;
    const double X = p.Kinv[0] * a0 + p.Kinv[1] * a1 + p.Kinv[2] * z, Y = p.Kinv[3] * a0 + p.Kinv[4] * a1 + p.Kinv[5] * z,
                 Z = p.Kinv[6] * a0 + p.Kinv[7] * a1 + p.Kinv[8] * z;
    const double x1 = p.T[0] * X + p.T[1] * Y + p.T[2] * Z + p.T[3], y1 = p.T[4] * X + p.T[5] * Y + p.T[6] * Z + p.T[7],
                 z1 = p.T[8] * X + p.T[9] * Y + p.T[10] * Z + p.T[11];
    const double px = p.K1[0] * x1 + p.K1[1] * y1 + p.K1[2] * z1, py = p.K1[3] * x1 + p.K1[4] * y1 + p.K1[5] * z1,
                 pz = p.K1[6] * x1 + p.K1[7] * y1 + p.K1[8] * z1;
    const double fu = rint(px / pz), fv = rint(py / pz); // np.round: half to even
    if (!(fu >= 0.0 && fu < (double)p.W1 && fv >= 0.0 && fv < (double)p.H1)) return;
    atomicMin(&key[(size_t)(int)fv * p.W1 + (int)fu], dkey(pz));
}
__global__ void project_gather_kernel(const unsigned long long *__restrict__ key, double *__restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = key[i];
    out[i] = k == 0xFFFFFFFFFFFFFFFFull ? 0.0 : dunkey(k);
}
} // namespace

cudaError_t launch_project_depth(b2s_ctx *c, const double *d_depth2, int W2, int H2, double rate, const double *Kinv, const double *T,
                                 const double *K1, int W1, int H1, unsigned long long *d_key, double *d_out)
{
    ProjArgs p;
    memcpy(p.Kinv, Kinv, sizeof p.Kinv);
    memcpy(p.T, T, sizeof p.T);
    memcpy(p.K1, K1, sizeof p.K1);
    p.rate = rate;
    p.W2 = W2; p.H2 = H2; p.W1 = W1; p.H1 = H1;
    p.W2u = rate == 1.0 ? W2 : (int)nearbyint(W2 * rate);
    p.H2u = rate == 1.0 ? H2 : (int)nearbyint(H2 * rate);
    if (p.W2u <= 0 || p.H2u <= 0) return cudaErrorInvalidValue;
    p.sx = 1.0 / ((double)p.W2u / (double)W2); // cv::resize: ifx = 1 / inv_scale_x, inv_scale_x = dsize.width / ssize.width
    p.sy = 1.0 / ((double)p.H2u / (double)H2);
    const size_t n = (size_t)W1 * H1;
    cudaError_t e = cudaMemsetAsync(d_key, 0xFF, n * 8, c->stream);
    if (e != cudaSuccess) return e;
    dim3 b(128), g((p.W2u + 127) / 128, p.H2u);
    project_scatter_kernel<<<g, b, 0, c->stream>>>(d_depth2, p, d_key);
    project_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_key, d_out, n);
    c->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_gen_maps(b2s_ctx *c, const b2s_map_params &p, float *mapx, float *mapy, uint8_t *mask, int mW, int mH, int16_t *xy16,
                            uint16_t *fxy16)
{
    dim3 b(128), g((p.W + 127) / 128, p.H);
    gen_maps_kernel<<<g, b, 0, c->stream>>>(p, mapx, mapy, mask, mW, mH, (short2 *)xy16, fxy16);
    c->launches++;
    return cudaGetLastError();
}

// prm != NULL: evaluate the map from its parameters inside the kernel (LANCZOS4 only; mapx / mapy are not read)
cudaError_t launch_remap_u8(b2s_ctx *c, const uint8_t *src, int sH, int sW, int cn, const float *mapx, const float *mapy, int dH,
                            int dW, int xshift, int interp, uint8_t *dst, const b2s_map_params *prm)
{
    if (interp == 0 && c->lanczos_tabp.p && (cn == 1 || cn == 3) && !getenv("B2S_REMAP_SIMPLE")) {
        const size_t smem = RL_TABB + 16;
        static std::once_flag once[64][4];
        cudaError_t e = cudaSuccess;
        auto go = [&](auto kern, int slot) -> cudaError_t {
            std::call_once(once[c->device & 63][slot], [&] { e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
            if (e != cudaSuccess) return e;
            const int tiles = ((dW + 31) / 32) * ((dH + RL_WARPS - 1) / RL_WARPS);
            const int grid = tiles < c->num_sms ? tiles : c->num_sms;
            static const b2s_map_params none = {};
            kern<<<grid, RL_WARPS * 32, smem, c->stream>>>(src, sH, sW, mapx, mapy, prm ? *prm : none, dH, dW, xshift, c->lanczos_tabp.as<unsigned char>(), dst);
            c->launches++;
            return cudaGetLastError();
        };
        if (cn == 3) return prm ? go(remap_lz4_kernel<3, true>, 0) : go(remap_lz4_kernel<3, false>, 1);
        return prm ? go(remap_lz4_kernel<1, true>, 2) : go(remap_lz4_kernel<1, false>, 3);
    }
    dim3 b(128), g((dW + 127) / 128, dH);
    const int16_t *tab = c->lanczos_tab.as<int16_t>();
    if (cn == 3 && interp == 0) remap_u8_kernel<3, 0><<<g, b, 0, c->stream>>>(src, sH, sW, mapx, mapy, dH, dW, xshift, tab, dst);
    else if (cn == 3) remap_u8_kernel<3, 1><<<g, b, 0, c->stream>>>(src, sH, sW, mapx, mapy, dH, dW, xshift, tab, dst);
    else if (interp == 0) remap_u8_kernel<1, 0><<<g, b, 0, c->stream>>>(src, sH, sW, mapx, mapy, dH, dW, xshift, tab, dst);
    else remap_u8_kernel<1, 1><<<g, b, 0, c->stream>>>(src, sH, sW, mapx, mapy, dH, dW, xshift, tab, dst);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_undistort_u8(b2s_ctx *c, const uint8_t *src, int H, int W, int cn, const int16_t *xy, const uint16_t *fxy, uint8_t *dst)
{
    dim3 b(128), g((W + 127) / 128, H);
    if (cn == 3) undistort_u8_kernel<3><<<g, b, 0, c->stream>>>(src, H, W, (const short2 *)xy, fxy, dst);
    else undistort_u8_kernel<1><<<g, b, 0, c->stream>>>(src, H, W, (const short2 *)xy, fxy, dst);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_depth_bare(b2s_ctx *c, const float *d_disp, double *d_depth)
{
    size_t n = (size_t)c->rH * c->rW;
    disp_to_depth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_disp, nullptr, nullptr, d_depth, n, 0.f, c->r_fxb, c->r_max_depth);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_unrectify(b2s_ctx *c, const double *d_depth, double *d_out)
{
    dim3 b(128), g((c->rW1 + 127) / 128, c->rH1);
    unrectify_kernel<<<g, b, 0, c->stream>>>(d_depth, c->rH, c->rW, c->umapx.as<float>(), c->umapy.as<float>(), d_out, c->rH1, c->rW1,
                                             c->r_m[0], c->r_m[1], c->r_m[2]);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_depth(b2s_ctx *c, const float *d_disp_in, int add_min_disp, int want_unrectify)
{
    size_t n = (size_t)c->rH * c->rW;
    disp_to_depth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_disp_in, c->vmask.as<uint8_t>(), c->dispfinal.as<float>(),
                                                                            c->rdepth.as<double>(), n, add_min_disp ? (float)c->r_min_disp : 0.f,
                                                                            c->r_fxb, c->r_max_depth);
    c->launches++;
    if (want_unrectify) {
        dim3 b(128), g((c->rW1 + 127) / 128, c->rH1);
        unrectify_kernel<<<g, b, 0, c->stream>>>(c->rdepth.as<double>(), c->rH, c->rW, c->umapx.as<float>(), c->umapy.as<float>(),
                                                 c->udepth.as<double>(), c->rH1, c->rW1, c->r_m[0], c->r_m[1], c->r_m[2]);
        c->launches++;
    }
    return cudaGetLastError();
}
