// resize.cu -- the down-scale / up-scale around the matcher (reference default max_size = 1000), sm_100a.
//
// Replaces the two boxx.resize calls of SemiGlobalBlockMatching.__call__ (calibrating/stereo_matching.py:61-62, 65-69): the
// rectified pair is reduced so that its longest side is max_size before cv2.StereoSGBM.compute, and the float disparity is
// brought back to the full size and multiplied by w / sw.  boxx is an un-vendored dependency whose interpolation is unpinned
// (SURVEY.md section 8(c)); the restatement pinned here is cv2.resize(INTER_LINEAR), see oracle/resize.py:
//   uint8:   bit-exact with cv2 (11-bit fixed-point coefficients, the (S >> 4) * b >> 16 vertical pass of OpenCV's resize.cpp)
//   float32: bilinear with float64 coefficients (what cv2's IPP path does to within 2 ulp; tolerance 1e-6 in the tests)
// HBM-bound gathers: one thread per output pixel, 4 taps per channel.
#include "b2s_internal.h"

namespace {

// source index and weight of OpenCV's INTER_LINEAR for destination index d: f = (d + 0.5) * scale - 0.5, clamped at both ends
__device__ __forceinline__ void lin_coef_f32(int d, double scale, int n, int &s, float &f)
{
    f = (float)(((double)d + 0.5) * scale - 0.5);
    s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n - 1) { s = n - 1; f = 0.f; }
}

template <int CN>
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t *__restrict__ src, int sH, int sW, uint8_t *__restrict__ dst, int dH, int dW,
                                                        double scale_x, double scale_y)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dW) return;
    int sx, sy;
    float fx, fy;
    lin_coef_f32(dx, scale_x, sW, sx, fx);
    lin_coef_f32(dy, scale_y, sH, sy, fy);
    // saturate_cast<short>(f * INTER_RESIZE_COEF_SCALE): round half to even
    const int ax1 = __float2int_rn(fx * 2048.f), ax0 = __float2int_rn((1.f - fx) * 2048.f);
    const int ay1 = __float2int_rn(fy * 2048.f), ay0 = __float2int_rn((1.f - fy) * 2048.f);
    const int sx1 = min(sx + 1, sW - 1), sy1 = min(sy + 1, sH - 1);
    const uint8_t *r0 = src + (size_t)sy * sW * CN, *r1 = src + (size_t)sy1 * sW * CN;
    uint8_t *o = dst + ((size_t)dy * dW + dx) * CN;
#pragma unroll
    for (int c = 0; c < CN; c++) {
        const int S0 = r0[sx * CN + c] * ax0 + r0[sx1 * CN + c] * ax1;
        const int S1 = r1[sx * CN + c] * ax0 + r1[sx1 * CN + c] * ax1;
        o[c] = (uint8_t)((((ay0 * (S0 >> 4)) >> 16) + ((ay1 * (S1 >> 4)) >> 16) + 2) >> 2);
    }
}

__device__ __forceinline__ void lin_coef_f64(int d, double scale, int n, int &s, double &f)
{
    f = ((double)d + 0.5) * scale - 0.5;
    const double fl = floor(f);
    s = (int)fl;
    f -= fl;
    if (s < 0) { s = 0; f = 0.0; }
    if (s >= n - 1) { s = n - 1; f = 0.0; }
}

// dst = bilinear(src) * mul / div, the two float32 roundings of `resize(sdisparity / 16.0, (h, w)) * w / sw`
__global__ void __launch_bounds__(256) resize_f32_kernel(const float *__restrict__ src, int sH, int sW, float *__restrict__ dst, int dH, int dW,
                                                         double scale_x, double scale_y, float mul, float div)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dW) return;
    int sx, sy;
    double fx, fy;
    lin_coef_f64(dx, scale_x, sW, sx, fx);
    lin_coef_f64(dy, scale_y, sH, sy, fy);
    const int sx1 = min(sx + 1, sW - 1), sy1 = min(sy + 1, sH - 1);
    const float *r0 = src + (size_t)sy * sW, *r1 = src + (size_t)sy1 * sW;
    // (no contraction into fused multiply-adds: the oracle is plain float64 numpy)
    const double S0 = __dadd_rn(__dmul_rn((double)r0[sx], 1.0 - fx), __dmul_rn((double)r0[sx1], fx));
    const double S1 = __dadd_rn(__dmul_rn((double)r1[sx], 1.0 - fx), __dmul_rn((double)r1[sx1], fx));
    const float v = (float)__dadd_rn(__dmul_rn(S0, 1.0 - fy), __dmul_rn(S1, fy));
    dst[(size_t)dy * dW + dx] = __fdiv_rn(__fmul_rn(v, mul), div);
}

// cv2.resize(float32, INTER_NEAREST) times a factor: source index = min(floor(dst * scale), size - 1)
__global__ void __launch_bounds__(256) resize_nearest_f32_kernel(const float *__restrict__ src, int sH, int sW, float *__restrict__ dst, int dH, int dW,
                                                                 double scale_x, double scale_y, float mul)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= dW) return;
    const int sx = min((int)floor(dx * scale_x), sW - 1), sy = min((int)floor(dy * scale_y), sH - 1);
    dst[(size_t)dy * dW + dx] = __fmul_rn(src[(size_t)sy * sW + sx], mul);
}

} // namespace

cudaError_t launch_resize_nearest_f32(b2s_ctx *c, const float *src, int sH, int sW, float *dst, int dH, int dW, float mul)
{
    const dim3 b(256), g((dW + 255) / 256, dH);
    resize_nearest_f32_kernel<<<g, b, 0, c->stream>>>(src, sH, sW, dst, dH, dW, 1.0 / ((double)dW / sW), 1.0 / ((double)dH / sH), mul);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_resize_u8(b2s_ctx *c, const uint8_t *src, int sH, int sW, int cn, uint8_t *dst, int dH, int dW)
{
    const dim3 b(256), g((dW + 255) / 256, dH);
    const double sx = (double)sW / dW, sy = (double)sH / dH;
    if (cn == 3) resize_u8_kernel<3><<<g, b, 0, c->stream>>>(src, sH, sW, dst, dH, dW, sx, sy);
    else if (cn == 1) resize_u8_kernel<1><<<g, b, 0, c->stream>>>(src, sH, sW, dst, dH, dW, sx, sy);
    else return cudaErrorInvalidValue;
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_resize_f32(b2s_ctx *c, const float *src, int sH, int sW, float *dst, int dH, int dW, float mul, float div)
{
    const dim3 b(256), g((dW + 255) / 256, dH);
    resize_f32_kernel<<<g, b, 0, c->stream>>>(src, sH, sW, dst, dH, dW, (double)sW / dW, (double)sH / dH, mul, div);
    c->launches++;
    return cudaGetLastError();
}
