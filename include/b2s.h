/*
 * b2s.h -- C-ABI of the B200-native `Stereo.get_depth` engine (libb2s.so).
 *
 * This is the drop-in boundary for the hot path of DIYer22/calibrating:
 *   Stereo.get_depth            calibrating/stereo_camera.py:492-533
 *   Stereo.rectify              calibrating/stereo_camera.py:216-242   (cv2.remap INTER_LANCZOS4 x2 + shift)
 *   SemiGlobalBlockMatching     calibrating/stereo_matching.py:22-70   (cv2.StereoSGBM.compute + post-processing)
 *   Stereo.disparity_to_depth   calibrating/stereo_camera.py:408-413
 *   Stereo.unrectify_depth      calibrating/stereo_camera.py:415-428 -> utils.rotate_depth_by_remap, utils.py:173-200
 *   Stereo.undistort_img        calibrating/stereo_camera.py:430-431   (cv2.undistort)
 * The reference has no FFI of its own (it is pure Python over cv2); the binding a maintainer adds is the
 * ctypes stub shown in INTEGRATION.md (calibrating_b200/_ffi.py is that stub).
 *
 * Conventions: every entry point returns 0 on success or a negative B2S_E* code; b2s_last_error() gives the
 * message.  No C++ exception crosses the boundary.  A handle owns one CUDA stream and all device buffers; it
 * is not thread-safe, distinct handles may be driven from distinct threads.  "host" pointers are ordinary
 * (pageable or pinned) caller-owned memory, C-contiguous, alive until the call (or the matching b2s_sync for
 * *_async calls) returns.  Every image / map / result pointer may equally be DEVICE memory of the handle's GPU
 * (copies use cudaMemcpyDefault): that is how the multi-GPU host layer hands over NCCL-broadcast rig constants
 * (b2s_set_rig) and collects depth maps for the all-gather (b2s_get_depth_async + b2s_sync) without a host round
 * trip.  Images are (H,W,cn) uint8, cn in {1,3}.  There is no CPU fallback: without a
 * CUDA device b2s_create fails with B2S_ECUDA.
 */
#ifndef B2S_H
#define B2S_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK 0
#define B2S_EINVAL (-1)   /* bad argument / unsupported parameter combination */
#define B2S_ESIZE (-2)    /* cv2's precondition: "input images are too small for your window size and max disparity" */
#define B2S_ECUDA (-3)    /* CUDA runtime error (message has the CUDA error string) */
#define B2S_ESTATE (-4)   /* call order: rig / parameters not set */

typedef struct b2s_ctx *b2s_handle;

/* Same fields and meaning as cv2.StereoSGBM_create (calibrating/stereo_matching.py:48-58);
 * mode: 0 = MODE_SGBM (5 paths, the reference default), 1 = MODE_HH (8 paths), 3 = MODE_HH4 (4 paths: the two horizontal
 * and the two vertical ones).  cv2's MODE_SGBM_3WAY (2) is rejected: its result depends on cv2's thread count. */
typedef struct {
    int min_disparity, num_disparities, block_size;
    int P1, P2, disp12_max_diff, pre_filter_cap, uniqueness_ratio;
    int speckle_window_size, speckle_range, mode;
    /* Extension (not in cv2, parity unpinned: BASELINE config 4): 0 = cv2's Birchfield-Tomasi cost + block sum,
     * 1 = 9x7 census transform / Hamming distance, gray images only, block_size unused.  Everything after the cost volume
     * (aggregation, winner-take-all, L/R check, median, speckle) is the same code. */
    int cost;
} b2s_sgbm_params;

/* Per-rig constants produced once on the host by Stereo._get_undistort_rectify_map
 * (calibrating/stereo_camera.py:125-177) and Stereo.set_stereo_matching (:466-489). */
typedef struct {
    int W, H;                   /* rectified size  (Stereo.xy) */
    int W1, H1;                 /* cam1 raw size   (cam1.xy)   */
    int W2, H2;                 /* cam2 raw size   (cam2.xy)   */
    const float *map1x, *map1y; /* (H,W) f32: undistort_rectify_map1 */
    const float *map2x, *map2y; /* (H,W) f32: undistort_rectify_map2 */
    const uint8_t *valid_mask1; /* (H,W) u8 0/1: rectify_valid_mask1 */
    const float *unrect_mapx, *unrect_mapy; /* (H1,W1) f32: maps of rotate_depth_by_remap (utils.py:184-191) */
    const int16_t *undist_xy;   /* (H1,W1,2) i16 and                                                    */
    const uint16_t *undist_fxy; /* (H1,W1) u16: CV_16SC2 maps of cv2.undistort(img1, cam1.K, cam1.D)     */
    double unrect_m[3];         /* third row of R1^T * K^-1: z' = z*(m0*x + m1*y + m2) (utils.py:192-197) */
    double fx_baseline;         /* K[0,0] * |t|   (disparity_to_depth, stereo_camera.py:408-410)         */
    double max_depth;           /* Stereo.get_max_depth()                                                */
    int min_disparity;          /* Stereo.min_disparity (translation of rectify_img2), 0 when disabled   */
    int interp;                 /* 0 = INTER_LANCZOS4 (reference), 1 = INTER_LINEAR (fast mode)          */
} b2s_rig;

/* One cv2.initUndistortRectifyMap(cameraMatrix, distCoeffs, R, newCameraMatrix, (W,H)) call described by its inputs, for
 * device-side map generation (b2s_set_rig_params): pixel (j,i) of the map -> [X Y Wh] = iR*[j i 1], x = X/Wh, y = Y/Wh,
 * OpenCV distortion model k = (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4), u = fx*xd + cx, v = fy*yd + cy, evaluated in
 * float64 in cv2's operation order (no fused multiply-add) and rounded to float32 once. */
typedef struct {
    int W, H;              /* map size */
    double fx, fy, cx, cy; /* cameraMatrix: intrinsics of the image the map samples */
    double k[12];          /* distortion coefficients, zero-padded (the tilt terms tauX/tauY are not supported) */
    double iR[9];          /* inv(newCameraMatrix * R), row-major */
} b2s_map_params;

/* The rig of b2s_rig described by parameters instead of arrays: the engine generates the four rectification map planes,
 * the valid mask (stereo_camera.py:167-176), the unrectify maps (utils.py:184-191) and the CV_16SC2 undistort maps on
 * the device.  SURVEY.md section 8(f) rank 1. */
typedef struct {
    int W, H, W1, H1, W2, H2;
    b2s_map_params rect1, rect2; /* (cam1.K, cam1.D, R1, K), (cam2.K, cam2.D, R2, K) at (W,H)            */
    b2s_map_params unrect;       /* (K, none, R1^T, cam1.K) at (W1,H1)                                  */
    b2s_map_params undist;       /* (cam1.K, cam1.D, I, cam1.K) at (W1,H1)                              */
    double unrect_m[3], fx_baseline, max_depth;
    int min_disparity, interp;
} b2s_rig_params;

/* Which arrays b2s_get_depth copies back; NULL pointers are skipped. */
typedef struct {
    uint8_t *rectify_img1, *rectify_img2; /* (H,W,cn) u8 */
    float *disparity;                     /* (H,W) f32, after += min_disparity and the valid mask */
    double *rectify_depth;                /* (H,W) f64 */
    double *unrectify_depth;              /* (H1,W1) f64 */
    uint8_t *undistort_img1;              /* (H1,W1,cn) u8 */
    int16_t *disp16;                      /* (H,W) i16: raw StereoSGBM output (debug / parity) */
    double *distort_depth;                /* (H1,W1) f64: Stereo.distort_depth(unrectify_depth) (return_distort_depth);
                                             needs b2s_set_cam1_model and want_unrectify */
} b2s_depth_out;

typedef struct {
    float rectify_ms, cost_ms, aggregate_ms, wta_ms, post_ms, depth_ms, total_ms;
    int aggregate_launches, total_launches;
} b2s_timing;

/* ---- lifecycle ------------------------------------------------------------------------------------------- */
int b2s_device_count(void);
int b2s_create(int device, b2s_handle *out);
int b2s_destroy(b2s_handle h);
const char *b2s_last_error(b2s_handle h); /* h may be NULL: error of the last failed b2s_create */
int b2s_sync(b2s_handle h);               /* wait for the handle's stream */

/* pinned host memory for the async API (cudaHostAlloc / cudaFreeHost) */
int b2s_host_alloc(size_t bytes, void **out);
int b2s_host_free(void *p);

/* ---- matcher: replaces cv2.StereoSGBM_create + .compute (stereo_matching.py:48-63) ----------------------- */
int b2s_set_sgbm_params(b2s_handle h, const b2s_sgbm_params *p);
/* left/right host (H,W,cn) u8.  out_disp16 (nullable): (H,W) i16 = 16*disparity exactly as cv2 returns it.
 * out_disp (nullable): (H,W) f32 with the reference's post-processing (stereo_matching.py:63-64 and /16):
 * clip(0), values < 16*minDisparity zeroed, divided by 16.  Synchronous. */
int b2s_compute_disparity(b2s_handle h, const uint8_t *left, const uint8_t *right, int H, int W, int cn,
                          int16_t *out_disp16, float *out_disp);
/* Same, enqueued on the handle's stream; host buffers should be pinned; complete after b2s_sync(). */
int b2s_compute_disparity_async(b2s_handle h, const uint8_t *left, const uint8_t *right, int H, int W, int cn,
                                int16_t *out_disp16, float *out_disp);
/* Device-resident variant: all pointers are device memory on the handle's device; enqueues and returns. */
int b2s_compute_disparity_dev(b2s_handle h, const uint8_t *d_left, const uint8_t *d_right, int H, int W, int cn,
                              int16_t *d_out_disp16, float *d_out_disp);

/* ---- full chain: replaces Stereo.rectify / get_depth / unrectify_depth / undistort_img ------------------- */
int b2s_set_rig(b2s_handle h, const b2s_rig *rig);
/* Same rig, maps generated on the device from the calibration parameters (bit-identical to cv2's maps wherever
 * float64 arithmetic is: tests/test_gpu_chain.py); replaces the cv2.initUndistortRectifyMap calls of
 * stereo_camera.py:159-165, utils.py:184-191 and the ~100 MB upload / broadcast of their results. */
int b2s_set_rig_params(b2s_handle h, const b2s_rig_params *rig);
/* Stereo.rectify (stereo_camera.py:216-242): img1 (H1,W1,cn), img2 (H2,W2,cn) host -> out1/out2 (H,W,cn). */
int b2s_rectify(b2s_handle h, const uint8_t *img1, const uint8_t *img2, int cn, uint8_t *out1, uint8_t *out2);
/* Stereo.get_depth with the built-in matcher (stereo_camera.py:492-533). want_unrectify: also run
 * unrectify_depth + undistort_img (return_unrectify_depth, default True in the reference). */
int b2s_get_depth(b2s_handle h, const uint8_t *img1, const uint8_t *img2, int cn, int want_unrectify,
                  const b2s_depth_out *out);
int b2s_get_depth_async(b2s_handle h, const uint8_t *img1, const uint8_t *img2, int cn, int want_unrectify,
                        const b2s_depth_out *out);
/* Tail of get_depth for a foreign MetaStereoMatching plugin: disparity (H,W) f32 host, as returned by the
 * plugin (stereo_camera.py:506-533 after the plugin call). img1 (nullable) feeds undistort_img. */
int b2s_depth_from_disparity(b2s_handle h, const float *disparity, const uint8_t *img1, int cn,
                             int want_unrectify, const b2s_depth_out *out);

/* Stand-alone stages, same meaning as the reference methods of the same name (host buffers, synchronous):
 * Stereo.disparity_to_depth (stereo_camera.py:408-413): (H,W) f32 -> (H,W) f64, no mask, no offset. */
int b2s_disparity_to_depth(b2s_handle h, const float *disparity, double *depth);
/* Stereo.unrectify_depth (stereo_camera.py:415-428): (H,W) f64 -> (H1,W1) f64. */
int b2s_unrectify_depth(b2s_handle h, const double *rectify_depth, double *out);
/* cam1's own intrinsics and distortion model (fx fy cx cy; k = k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4, zero-padded), used by
 * b2s_distort_depth.  b2s_set_rig_params sets it implicitly (its `undist` map describes the same camera). */
int b2s_set_cam1_model(b2s_handle h, double fx, double fy, double cx, double cy, const double k[12]);
/* Stereo.distort_depth (stereo_camera.py:433-464): forward splat of the undistorted (H1,W1) f64 depth image into the raw
 * distorted cam1 image: every pixel is projected through the distortion model exactly as cv2.undistortPoints +
 * cv2.projectPoints do (float64, float32 where they round to float32), truncated to int; of the pixels landing on the
 * same target the smallest source index wins (np.unique(..., return_index=True)); targets nobody hits are 0.  Source
 * pixels that project outside the image are dropped (the reference would raise IndexError or wrap around). */
int b2s_distort_depth(b2s_handle h, const double *unrectify_depth, double *out);
/* Cam.project_cam2_depth (calibrating/camera.py:298-309 -> utils.depth_to_point_cloud, apply_T_to_point_cloud,
 * point_cloud_to_depth, utils.py:152-161, 213-317): the depth image of a second camera seen from this one.  depth2 (H2,W2)
 * f64 (0 = no measurement) is up-sampled nearest-neighbour by `rate` like cv2.resize(INTER_NEAREST) to
 * (round(W2*rate), round(H2*rate)), every non-zero sample is un-projected with K2inv (row-major 3x3), moved by the rigid
 * transform T (row-major 4x4, cam2 -> cam1), projected with K1, rounded half-to-even to the (H1,W1) grid; the smallest z
 * landing on a pixel wins (the reference writes in order of descending z), pixels nobody hits are 0.  float64 throughout;
 * the reference's matrix products go through BLAS, so z agrees to ~1e-15 relative, not bit for bit.  SURVEY 8(f) rank 3.
 * Independent of the rig: needs only a handle. */
int b2s_project_depth(b2s_handle h, const double *depth2, int W2, int H2, double rate, const double K2inv[9], const double T[16],
                      const double K1[9], int W1, int H1, double *out);
/* Stereo.undistort_img (stereo_camera.py:430-431): (H1,W1,cn) u8 -> (H1,W1,cn) u8. */
int b2s_undistort_img(b2s_handle h, const uint8_t *img1, int cn, uint8_t *out);

/* ---- introspection ----------------------------------------------------------------------------------------- */
#define B2S_FETCH_C 0    /* cost volume (H,width1,Dp) i16 */
#define B2S_FETCH_S 1    /* aggregated volume (H,width1,Dp) i16 */
#define B2S_FETCH_RAW 2  /* (H,W) i16 disparity before median/speckle */
#define B2S_FETCH_RIG 16 /* + k: k-th rig array in b2s_rig order: map1x map1y map2x map2y valid_mask1 unrect_mapx unrect_mapy undist_xy undist_fxy */
/* Options.
 * B2S_OPT_FUSE_WTA (default 1): the winner-take-all step runs inside the last aggregation pass; the aggregated volume S is
 *   then never written unless B2S_OPT_KEEP_VOLUMES (default 0) is also set (B2S_FETCH_S fails otherwise).  With 0 the pass
 *   stores S and a separate kernel picks the winners.  Results are identical either way.
 * B2S_OPT_MAX_SIZE (default 0 = off): `max_size` of SemiGlobalBlockMatching (calibrating/stereo_matching.py:26,60-70; the reference's
 *   default is 1000).  When the longest image side exceeds it, b2s_compute_disparity* / b2s_get_depth* reduce the pair by
 *   min(max_size / max(h, w), 1) on the device (cv2.resize INTER_LINEAR semantics, bit-exact for uint8), match at that size and
 *   return the float disparity at full size times w / sw; the int16 disparity is not defined then and must not be requested.
 * B2S_OPT_AGG_SCHEDULE (default 0): how the eight paths of MODE_HH are scheduled.  0 = two horizontal scans around a
 *   lock-step vertical sweep (20 bytes of DRAM traffic per cost voxel; the fastest for one pair); 1 = two wavefront sweeps of four
 *   paths each in OpenCV's own order (the canonical 8 bytes per voxel; sgbm_wave.cu).  Results are identical; other modes
 *   ignore it.  The environment variable B2S_AGG_SCHEDULE=sweep|wave overrides the option. */
#define B2S_OPT_KEEP_VOLUMES 1
#define B2S_OPT_FUSE_WTA 2
#define B2S_OPT_AGG_SCHEDULE 3
#define B2S_OPT_MAX_SIZE 4
/* calibrating/utils.py:213-250 depth_to_point_cloud: the non-zero pixels of `depth` ((H,W) f64 host; the caller applies the
 * reference's uint16 -> float32(depth / 1000) rule first), optionally nearest-neighbour up-sampled by `rate` (cv2.resize
 * INTER_NEAREST to (round(W*rate), round(H*rate))), un-projected with Kinv = inv(K) in NumPy's row-major order:
 * out[k] = Kinv @ (u z, v z, z) [+ (u, v) when with_uv], u = column / rate.  `capacity` = rows of `out`; *n_out = points found
 * (B2S_ESIZE when it exceeds the capacity).  z within 1e-12 of the reference (its 3x3 product goes through BLAS). */
int b2s_depth_to_point_cloud(b2s_handle h, const double *depth, int H, int W, double rate, const double Kinv[9], int with_uv, double *out,
                             unsigned long long capacity, unsigned long long *n_out);
/* calibrating/utils.py:254-317 point_cloud_to_depth (= point_cloud_to_arr2d without values): points (n,3) f64 projected with K,
 * rounded half-to-even to the pixel grid of (W,H); of the points landing on a pixel the smallest z survives (the reference
 * writes in the order of descending z), pixels without a point get bg_value. */
int b2s_point_cloud_to_depth(b2s_handle h, const double *points, unsigned long long n, const double K[9], int W, int H, double bg_value, double *out);
/* dst (dH,dW) f32 = cv2.resize(src (sH,sW) f32, INTER_NEAREST) * mul: the up-scale of FeatureMatchingAsStereoMatching
 * (calibrating/stereo_matching.py:133-140: `disparity * hw[1] / resize_shape[1]`, then a nearest-neighbour resize). */
int b2s_resize_nearest_f32(b2s_handle h, const float *src, int sH, int sW, float *dst, int dH, int dW, float mul);
/* calibrating/utils.py:347-415 interpolate_uvzs, the dense half (what MatchingByBoard and
 * FeatureMatchingAsStereoMatching densify their sparse disparities with).  out: (H,W) f32.  mask (nullable, (H,W) u8): the pixels to
 * fill (the convex hull the caller rasterised), others are 0.
 *   inter_type 0 ("lstsq"):   out = abc[0]*x + abc[1]*y + abc[2] (the caller fitted the plane; float64, stored float32)
 *   inter_type 1 ("nearest"): out = z of the nearest of the n points uvz (n,3) f64 if it is closer than `distance`, else 0 */
int b2s_interpolate_sparse(b2s_handle h, int inter_type, const double *uvz, int n, const double abc[3], const uint8_t *mask, int H, int W,
                           double distance, float *out);
/* inter_type "rbf" of interpolate_uvzs (utils.py:373-387: scipy.interpolate.Rbf, thin plate): out (H,W) f64 =
 * sum_i w_i * r_i^2 * log(r_i), r_i = distance of the pixel to sample i; uvw (n,3) f64 = (u, v, node weight w): the caller solved the
 * n x n system (phi(r_ij) - smooth * I) w = z on the host.  mask as above. */
int b2s_interpolate_rbf(b2s_handle h, const double *uvw, int n, const uint8_t *mask, int H, int W, double *out);
int b2s_set_option(b2s_handle h, int option, int value);
/* sha256 (first 16 hex digits) over the CUDA sources this library was built from (calibrating_b200/build.py); ties a
 * profile under profiles/ to the binary it was measured on. */
const char *b2s_build_hash(void);
int b2s_volume_dims(b2s_handle h, int *H, int *width1, int *D, int *Dp);
int b2s_debug_fetch(b2s_handle h, int which, void *dst, size_t bytes);
int b2s_timings(b2s_handle h, b2s_timing *t);      /* CUDA-event stage times of the last synchronous call */
int b2s_launch_count(b2s_handle h, long long *n);  /* kernels launched by this handle since creation */
/* The aggregation GROUP (path aggregation + winner-take-all: everything between the cost stage and the raw winner map) timed
 * alone, in its production form, on the images of the last call (which must still be valid device memory: the handle's own
 * copies after b2s_compute_disparity, the caller's after b2s_compute_disparity_dev): every repetition re-runs the cost stage
 * first, outside the timed interval, so that the first aggregation launch is the one a real pair runs (the +x scan that also
 * forms C from the cost stage's row sums).  Returns the mean ms per repetition over `iters` CUDA-event intervals. */
int b2s_bench_aggregate(b2s_handle h, int iters, float *ms_per_iter);
/* Same, with one CUDA-event interval per kernel launch of the group in launch order (wavefront schedule, MODE_HH: the
 * two-sweep launch, then the winner-take-all that adds the two sums; MODE_SGBM: the top-down sweep, then the (-1,0) scan with
 * the fused winner-take-all; B2S_AGG_SCHEDULE=sweep: +x scan, fused vertical sweep, -x scan).  ms_parts[k] = mean ms of
 * launch k over `iters` repetitions, *n_parts = number of launches (<= max_parts). */
int b2s_bench_aggregate_parts(b2s_handle h, int iters, float *ms_parts, int max_parts, int *n_parts);
/* Enqueue `iters` repetitions of the group on the handle's stream without waiting (several handles in flight; bracket with
 * b2s_event_record / b2s_event_elapsed).  Runs on the finished cost volume: the first launch is the plain +x scan. */
int b2s_enqueue_aggregate(b2s_handle h, int iters);
/* User CUDA events on the handle's stream (slot 0..3), for device-side timing of caller-defined regions. */
int b2s_event_record(b2s_handle h, int slot);
/* ms between event `slot_a` of handle a and event `slot_b` of handle b (same device); waits for event b. */
int b2s_event_elapsed(b2s_handle a, int slot_a, b2s_handle b, int slot_b, float *ms);
/* Re-read the stage timings of the last *_async call (call after b2s_sync). */
int b2s_collect_timings(b2s_handle h, int chain);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H */
