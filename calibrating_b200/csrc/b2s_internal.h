// b2s_internal.h -- shared declarations of libb2s.so (not part of the public ABI; see include/b2s.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "b2s.h"

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return (T *)p; }
};

// Normalised matcher geometry (== StereoSGBM parameter defaulting in OpenCV).
struct SgbmGeom {
    int H, W, cn;
    int minD, maxD, D, Dp, NP; // Dp = 64*NP >= D: d-lanes padded to a whole warp of packed int16 pairs
    int minX1, width1;
    int SW2, SH2, ftzero, uniq, d12, P1, P2, invalid;
    int speckle_window, speckle_range, mode;
    // order of the disparities inside a pixel's d-chunk of C / S (32-bit words of two int16):
    //   0: word q = disparities (2q, 2q+1)                                  (scan / sweep kernels of sgbm_agg.cu)
    //   1: word q = disparities (16b+j, 16b+8+j), b = q/8, j = q%8  ("block layout", wavefront kernel of sgbm_wave.cu)
    int layout;
};
// int16 index of disparity d inside a pixel's d-chunk
__host__ __device__ inline int b2s_dindex(int layout, int d) { return layout == 0 ? d : ((d >> 4) << 4) + ((d & 7) << 1) + ((d >> 3) & 1); }
// first disparity of word q (the second one is +1 in layout 0, +8 in layout 1)
__host__ __device__ inline int b2s_word_d0(int layout, int q) { return layout == 0 ? 2 * q : ((q >> 3) << 4) + (q & 7); }

enum { AGG_INIT = 0, AGG_ACCUM = 1, AGG_ACCUM2 = 2 };
constexpr int B2S_AGG_MAX_PARTS = 8;

struct b2s_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;

    bool have_prm = false;
    b2s_sgbm_params prm{};
    SgbmGeom g{};
    bool have_volume = false;
    const uint8_t *last_dl = nullptr, *last_dr = nullptr; // device images of the last matcher run (bench: the cost stage is re-run untimed)
    bool keep_volumes = false; // b2s_set_option(B2S_OPT_KEEP_VOLUMES): a fused last pass also stores S
    bool fuse_wta = true;      // b2s_set_option(B2S_OPT_FUSE_WTA): on by default
    int max_size = 0;          // b2s_set_option(B2S_OPT_MAX_SIZE): longest image side the matcher works on (0 = no limit), stereo_matching.py:26,61
    int agg_schedule = 0;      // b2s_set_option(B2S_OPT_AGG_SCHEDULE): 0 = scans + lock-step sweep (sgbm_agg.cu), 1 = wavefront sweeps (sgbm_wave.cu)
    bool agg_legacy = false;   // launch_aggregate: per-direction scan kernels (strips too wide, or B2S_AGG_LEGACY)
    bool hs_pending = false;   // launch_cost_volume stopped at the row sums (in S2): the first horizontal scan forms C (agg_fuses_vsum)
    bool wta_fused = false;    // set by launch_aggregate when the last scan already did the winner-take-all
    bool wta_adds_s2 = false;  // set by launch_aggregate (wavefront schedule, MODE_HH): the aggregated volume is sat(S + S2), formed by wta_kernel

    // matcher buffers
    DevBuf left, right;       // (H,W,cn) u8
    DevBuf planesL, planesR;  // (H, 2cn, W) uchar4 = (value, lo, hi, 0)
    DevBuf C, S;              // (H, width1, Dp) int16;  S doubles as the horizontal-sum scratch
    DevBuf S2;                // MODE_HH: partial sum of the bottom-up sweep, same shape
    DevBuf raw, disp16;       // (H,W) int16
    DevBuf disp2key;          // (H,W+2) u32: (minS<<16)|(0xFFFF-x1) of the winning left pixel, 0xFFFFFFFF = none
    DevBuf labels, sizes;     // (H,W) int32 each (speckle filter)
    DevBuf med;               // (H,W) int16 (median output before speckle)
    DevBuf dispf;             // (H,W) f32
    DevBuf sleft, sright, sdispf; // the down-scaled pair and its disparity (B2S_OPT_MAX_SIZE)
    DevBuf agg_ho;            // hand-over rings of the fused vertical sweep + its error flag (sgbm_agg.cu)
    DevBuf agg_errbuf;        // the aggregation kernels' device error flag
    int *agg_err = nullptr;   // device address of that flag (valid after an aggregation was enqueued)

    // rig
    bool have_rig = false;
    int rW = 0, rH = 0, rW1 = 0, rH1 = 0, rW2 = 0, rH2 = 0, r_min_disp = 0, r_interp = 0;
    double r_m[3] = {0, 0, 0}, r_fxb = 0, r_max_depth = 0;
    DevBuf map1x, map1y, map2x, map2y, vmask, umapx, umapy, und_xy, und_fxy;
    DevBuf img1, img2, rect1, rect2, und1; // raw inputs and remapped outputs
    DevBuf dispfinal, rdepth, udepth;      // (H,W) f32, (H,W) f64, (H1,W1) f64
    DevBuf lanczos_tab;                    // (1024, 8, 8) int16
    DevBuf lanczos_tabp;                   // the same with rows padded to 144 bytes (remap_lz4_kernel's shared-memory copy)
    bool have_map_params = false;          // rig set by b2s_set_rig_params: rectification evaluates the maps analytically
    b2s_map_params mp_rect1{}, mp_rect2{};
    DevBuf cl_pts, cl_aux;                 // depth_to_point_cloud / point_cloud_to_depth: points, row counts + offsets + K
    DevBuf pkey, pin, pout;                // project_depth: winner key per target pixel, staged input / output
    DevBuf dkey, ddepth;                   // distort_depth: winner index per target pixel (+ the 12 coefficients); (H1,W1) f64 result
    bool have_cam1 = false;
    double cam1_f[4] = {0, 0, 0, 0}, cam1_k[12] = {0};
    DevBuf stage_f32;                      // upload scratch for b2s_depth_from_disparity

    cudaEvent_t ev[8] = {};
    cudaEvent_t uev[4] = {};
    cudaEvent_t aev[B2S_AGG_MAX_PARTS + 1] = {};
    b2s_timing timing{};
};

// ---- kernel launchers (each enqueues on ctx->stream and bumps ctx->launches) --------------------------------------
// sgbm_cost.cu
cudaError_t launch_cost_volume(b2s_ctx *c, const uint8_t *d_left, const uint8_t *d_right);
// sgbm_agg.cu
cudaError_t launch_aggregate(b2s_ctx *c, int *n_launches, cudaEvent_t *marks = nullptr);
bool agg_fuses_vsum(const b2s_ctx *c);
bool agg_wave_selected(const b2s_ctx *c, int mode); // the wavefront schedule (sgbm_wave.cu) applies to this matcher mode (decides SgbmGeom::layout)
// sgbm_wave.cu
cudaError_t launch_wave(b2s_ctx *c, int ndirs);
// sgbm_sweep6.cu: the six row-crossing paths of MODE_HH, path-parallel (block layout, 65..128 disparities)
int vsweep6_cols(const b2s_ctx *c, const SgbmGeom &g); // columns per CTA, 0 = not applicable
cudaError_t launch_vsweep6(b2s_ctx *c, int n, bool cooperative);
cudaError_t agg_error_flags(b2s_ctx *c); // creates the handle's device error flags (c->agg_err) on first use
int agg_poll_error(b2s_ctx *c); // after a stream sync: bit 0 = a wait of the aggregation kernels timed out, bit 1 = cost volume outside the int16 exactness domain
// sgbm_post.cu
cudaError_t launch_wta_prepare(b2s_ctx *c); // before the aggregation: clears the WTA outputs
cudaError_t launch_wta(b2s_ctx *c);         // after it: stand-alone WTA kernel unless the aggregation fused it
cudaError_t launch_post(b2s_ctx *c, int16_t *d_out_disp16, float *d_out_disp);
// remap.cu
void build_lanczos4_table(int16_t *tab /* 1024*64 */);
cudaError_t launch_remap_u8(b2s_ctx *c, const uint8_t *src, int sH, int sW, int cn, const float *mapx, const float *mapy,
                            int dH, int dW, int xshift, int interp, uint8_t *dst, const b2s_map_params *prm = nullptr);
cudaError_t launch_undistort_u8(b2s_ctx *c, const uint8_t *src, int H, int W, int cn, const int16_t *xy, const uint16_t *fxy,
                                uint8_t *dst);
cudaError_t launch_depth(b2s_ctx *c, const float *d_disp_in, int add_min_disp, int want_unrectify);
cudaError_t launch_distort_depth(b2s_ctx *c, const double *d_depth, double *d_out);
cudaError_t launch_project_depth(b2s_ctx *c, const double *d_depth2, int W2, int H2, double rate, const double *Kinv, const double *T,
                                 const double *K1, int W1, int H1, unsigned long long *d_key, double *d_out);
cudaError_t launch_gen_maps(b2s_ctx *c, const b2s_map_params &p, float *mapx, float *mapy, uint8_t *mask, int mW, int mH, int16_t *xy16,
                            uint16_t *fxy16);
// resize.cu
cudaError_t launch_resize_u8(b2s_ctx *c, const uint8_t *src, int sH, int sW, int cn, uint8_t *dst, int dH, int dW);
cudaError_t launch_rbf_fill(b2s_ctx *c, const double *d_uvw, int n, const uint8_t *d_mask, int H, int W, double *d_out);
cudaError_t launch_resize_nearest_f32(b2s_ctx *c, const float *src, int sH, int sW, float *dst, int dH, int dW, float mul);
cudaError_t launch_resize_f32(b2s_ctx *c, const float *src, int sH, int sW, float *dst, int dH, int dW, float mul, float div);
// cloud.cu
cudaError_t launch_depth_to_cloud(b2s_ctx *c, const double *d_depth, int H, int W, double rate, const double *Kinv, int cols, double *d_out,
                                  unsigned long long capacity, unsigned *d_rowcount, unsigned long long *d_rowoff, int *Hu_out);
cudaError_t launch_cloud_to_depth(b2s_ctx *c, const double *d_pts, unsigned long long n, const double *d_K, int W, int H, unsigned long long *d_key,
                                  double *d_out, double bg);
cudaError_t launch_plane_fill(b2s_ctx *c, const uint8_t *d_mask, int H, int W, double a, double b, double cc, float *d_out);
cudaError_t launch_nearest_fill(b2s_ctx *c, const double *d_uvz, int n, const uint8_t *d_mask, int H, int W, double distance, float *d_out);
cudaError_t launch_depth_bare(b2s_ctx *c, const float *d_disp, double *d_depth);
cudaError_t launch_unrectify(b2s_ctx *c, const double *d_depth, double *d_out);
