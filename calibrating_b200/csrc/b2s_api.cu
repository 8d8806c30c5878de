// b2s_api.cu -- the C-ABI of libb2s.so (include/b2s.h): handle lifecycle, parameter normalisation, buffer
// management and the per-pair kernel schedule.  No torch types, no C++ exceptions across the boundary.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "b2s_internal.h"
#if __has_include("build_hash.h")
#include "build_hash.h" // written by calibrating_b200/build.py: sha256 over the CUDA sources
#endif
#ifndef B2S_BUILD_HASH
#define B2S_BUILD_HASH "unknown"
#endif

static thread_local std::string g_create_err;

static int fail(b2s_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_err = buf;
    return code;
}
// device error flags of the matcher (sgbm_agg.cu: agg_poll_error)
static int agg_fail(b2s_ctx *c, int flags)
{
    if (flags & 1) return fail(c, B2S_ECUDA, "aggregation hand-over timed out (a wait inside the aggregation kernels expired)");
    return fail(c, B2S_EINVAL, "a block sum of the cost volume wrapped past 32767 (C < 0): outside the int16 domain this matcher is exact "
                               "on; use a smaller blockSize");
}
#define CK(c, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t _e = (call);                                                                             \
        if (_e != cudaSuccess) return fail(c, B2S_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" {

int b2s_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *b2s_build_hash(void) { return B2S_BUILD_HASH; }

const char *b2s_last_error(b2s_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int b2s_create(int device, b2s_handle *out)
{
    if (!out) return fail(nullptr, B2S_EINVAL, "b2s_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(nullptr, B2S_ECUDA, "b2s_create: no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(nullptr, B2S_EINVAL, "b2s_create: device %d out of range [0,%d)", device, n);
    b2s_ctx *c = new (std::nothrow) b2s_ctx();
    if (!c) return fail(nullptr, B2S_EINVAL, "b2s_create: out of host memory");
    c->device = device;
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (c->num_sms <= 0) c->num_sms = 1;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return fail(nullptr, B2S_ECUDA, "b2s_create: %s", cudaGetErrorString(e));
    }
    for (auto &ev : c->ev) cudaEventCreate(&ev);
    for (auto &ev : c->uev) cudaEventCreate(&ev);
    for (auto &ev : c->aev) cudaEventCreate(&ev);
    // Lanczos4 fixed-point table (1024 x 8 x 8 int16 = 128 KB), shared by both rectify remaps
    std::vector<int16_t> tab(1024 * 64);
    build_lanczos4_table(tab.data());
    if ((e = c->lanczos_tab.ensure(tab.size() * 2)) != cudaSuccess ||
        (e = cudaMemcpy(c->lanczos_tab.p, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) {
        b2s_destroy(c);
        return fail(nullptr, B2S_ECUDA, "b2s_create: %s", cudaGetErrorString(e));
    }
    {   // rows padded to 144 bytes for the shared-memory copy of remap_lz4_kernel (remap.cu)
        std::vector<unsigned char> padded(1024 * 144, 0);
        for (int r = 0; r < 1024; r++) memcpy(&padded[(size_t)r * 144], &tab[(size_t)r * 64], 128);
        if ((e = c->lanczos_tabp.ensure(padded.size())) != cudaSuccess ||
            (e = cudaMemcpy(c->lanczos_tabp.p, padded.data(), padded.size(), cudaMemcpyHostToDevice)) != cudaSuccess) {
            b2s_destroy(c);
            return fail(nullptr, B2S_ECUDA, "b2s_create: %s", cudaGetErrorString(e));
        }
    }
    *out = c;
    return B2S_OK;
}

int b2s_destroy(b2s_handle c)
{
    if (!c) return B2S_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->left, &c->right, &c->planesL, &c->planesR, &c->C, &c->S, &c->S2, &c->raw, &c->disp16, &c->disp2key, &c->labels,
                      &c->sizes, &c->med, &c->dispf, &c->sleft, &c->sright, &c->sdispf, &c->agg_ho, &c->agg_errbuf, &c->map1x, &c->map1y, &c->map2x, &c->map2y, &c->vmask, &c->umapx, &c->umapy,
                      &c->und_xy, &c->und_fxy, &c->img1, &c->img2, &c->rect1, &c->rect2, &c->und1, &c->dispfinal, &c->rdepth,
                      &c->udepth, &c->lanczos_tab, &c->lanczos_tabp, &c->stage_f32, &c->dkey, &c->ddepth, &c->pkey, &c->pin, &c->pout, &c->cl_pts, &c->cl_aux};
    for (DevBuf *b : bufs) b->release();
    for (auto &ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->uev)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->aev)
        if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return B2S_OK;
}

int b2s_sync(b2s_handle c)
{
    if (!c) return B2S_EINVAL;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    if (int ae = agg_poll_error(c)) return agg_fail(c, ae);
    return B2S_OK;
}

int b2s_host_alloc(size_t bytes, void **out)
{
    if (!out) return B2S_EINVAL;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(nullptr, B2S_ECUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return B2S_OK;
}
int b2s_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? B2S_OK : B2S_ECUDA; }

int b2s_set_sgbm_params(b2s_handle c, const b2s_sgbm_params *p)
{
    if (!c || !p) return B2S_EINVAL;
    if (p->num_disparities <= 0) return fail(c, B2S_EINVAL, "numDisparities must be > 0 (got %d)", p->num_disparities);
    if (p->num_disparities > 512) return fail(c, B2S_EINVAL, "numDisparities > 512 is not supported (got %d)", p->num_disparities);
    if (p->mode != 0 && p->mode != 1 && p->mode != 3)
        return fail(c, B2S_EINVAL, "mode must be 0 (MODE_SGBM), 1 (MODE_HH) or 3 (MODE_HH4), got %d (MODE_SGBM_3WAY = 2 depends on cv2's thread count and is not offered)", p->mode);
    if (p->cost != 0 && p->cost != 1) return fail(c, B2S_EINVAL, "cost must be 0 (Birchfield-Tomasi, cv2) or 1 (census), got %d", p->cost);
    int P1 = p->P1 > 0 ? p->P1 : 2;
    int P2 = p->P2 > 0 ? p->P2 : 5;
    if (P2 < P1 + 1) P2 = P1 + 1;
    if (P2 > 32767 - 1) return fail(c, B2S_EINVAL, "P2 = %d does not fit the int16 cost type", P2);
    c->prm = *p;
    c->have_prm = true;
    c->have_volume = false;
    return B2S_OK;
}

} // extern "C"

// SURVEY.md Appendix A.1
static int make_geom(b2s_ctx *c, int H, int W, int cn)
{
    if (!c->have_prm) return fail(c, B2S_ESTATE, "b2s_set_sgbm_params has not been called");
    if (cn != 1 && cn != 3) return fail(c, B2S_EINVAL, "images must have 1 or 3 channels (got %d)", cn);
    if (c->prm.cost == 1 && cn != 1) return fail(c, B2S_EINVAL, "the census cost takes gray images (got %d channels)", cn);
    if (H <= 0 || W <= 0 || W > 65535) return fail(c, B2S_EINVAL, "bad image size %dx%d", W, H);
    const b2s_sgbm_params &p = c->prm;
    SgbmGeom g;
    g.H = H; g.W = W; g.cn = cn;
    g.minD = p.min_disparity;
    g.maxD = g.minD + p.num_disparities;
    g.D = p.num_disparities;
    g.mode = p.mode;
    g.minX1 = g.maxD > 0 ? g.maxD : 0;                        // A.1: minX1 = max(maxD, 0), maxX1 = W + min(minD, 0)
    g.width1 = (W + (g.minD < 0 ? g.minD : 0)) - g.minX1;
    // the block layout: the wavefront schedule, or the default schedule when its six-path sweep applies (sgbm_sweep6.cu)
    g.layout = (agg_wave_selected(c, p.mode) && g.D <= 256) ? 1 : 0;
    if (g.layout == 0 && p.cost == 0 && vsweep6_cols(c, g) > 0) g.layout = 1;
    g.NP = g.layout == 1 ? 2 * ((g.D + 127) / 128) : (g.D + 63) / 64; // the block layout takes whole blocks of 128 disparities
    g.Dp = 64 * g.NP;
    int bs = p.block_size > 0 ? p.block_size : 5;
    g.SW2 = g.SH2 = bs / 2;
    g.ftzero = (p.pre_filter_cap > 15 ? p.pre_filter_cap : 15) | 1;
    g.uniq = p.uniqueness_ratio >= 0 ? p.uniqueness_ratio : 10;
    g.d12 = p.disp12_max_diff > 0 ? p.disp12_max_diff : 1;
    g.P1 = p.P1 > 0 ? p.P1 : 2;
    g.P2 = p.P2 > 0 ? p.P2 : 5;
    if (g.P2 < g.P1 + 1) g.P2 = g.P1 + 1;
    g.minX1 = g.maxD > 0 ? g.maxD : 0;                        // A.1: minX1 = max(maxD, 0), maxX1 = W + min(minD, 0)
    g.width1 = (W + (g.minD < 0 ? g.minD : 0)) - g.minX1;
    g.invalid = (g.minD - 1) * 16;
    g.speckle_window = p.speckle_window_size;
    g.speckle_range = p.speckle_range;
    g.mode = p.mode;
    if (W - g.maxD <= g.SW2 || g.width1 <= 0) return fail(c, B2S_ESIZE, "input images are too small for your window size and max disparity");
    if (g.ftzero > 127) return fail(c, B2S_EINVAL, "preFilterCap %d too large", p.pre_filter_cap);
    if ((g.layout == 0 ? (64 + 2 * g.SW2 + g.D - 1) / 2 + 2 : (64 + 2 * g.SW2 + g.Dp + 1) / 2) > (g.D > 256 ? 296 : 168)) // shared-memory tile of the cost kernel (sgbm_cost.cu: TX, NRP)
        return fail(c, B2S_EINVAL, "blockSize %d is too large for numDisparities %d (supported: blockSize + numDisparities <= 269)", bs, g.D);
    c->g = g;
    size_t npx = (size_t)H * W, vol = (size_t)H * g.width1 * g.Dp * sizeof(int16_t);
    CK(c, c->planesL.ensure(npx * 2 * cn * 4));
    CK(c, c->planesR.ensure(npx * 2 * cn * 4));
    CK(c, c->C.ensure(vol));
    CK(c, c->S.ensure(vol));
    CK(c, c->S2.ensure(vol)); // MODE_HH: the bottom-up sweep's partial sum; before that, the cost stage's row sums (agg_fuses_vsum)
    CK(c, c->raw.ensure(npx * 2));
    CK(c, c->disp16.ensure(npx * 2));
    CK(c, c->med.ensure(npx * 2));
    CK(c, c->disp2key.ensure((size_t)H * (W + 2) * 4));
    CK(c, c->labels.ensure(npx * 4));
    CK(c, c->sizes.ensure(npx * 4));
    CK(c, c->dispf.ensure(npx * 4));
    return B2S_OK;
}

// events: 0 start, 1 after rectify, 2 after cost, 3 after aggregate, 4 after wta, 5 after post, 6 after depth
static int matcher_dev(b2s_ctx *c, const uint8_t *dl, const uint8_t *dr, int16_t *d_out16, float *d_outf, bool timed)
{
    if (timed) cudaEventRecord(c->ev[1], c->stream);
    c->last_dl = dl;
    c->last_dr = dr;
    CK(c, launch_cost_volume(c, dl, dr));
    if (timed) cudaEventRecord(c->ev[2], c->stream);
    int nl = 0;
    CK(c, launch_wta_prepare(c));
    CK(c, launch_aggregate(c, &nl));
    c->timing.aggregate_launches = nl;
    if (timed) cudaEventRecord(c->ev[3], c->stream);
    CK(c, launch_wta(c));
    if (timed) cudaEventRecord(c->ev[4], c->stream);
    CK(c, launch_post(c, d_out16, d_outf));
    if (timed) cudaEventRecord(c->ev[5], c->stream);
    c->have_volume = true;
    return B2S_OK;
}

// SemiGlobalBlockMatching.__call__ (stereo_matching.py:60-70) on device data: with B2S_OPT_MAX_SIZE set and an image whose
// longest side exceeds it, the pair is reduced by min(max_size / max(h, w), 1) (sides rounded half-to-even like Python's
// round), matched at that size, and the float disparity comes back at full size times w / sw.  Sets the matcher geometry.
// d_out16: device destination of the int16 disparity, or B2S_OWN16 = the handle's own buffer c->disp16 (resolved after make_geom,
// which may reallocate it), or NULL = not wanted.
#define B2S_OWN16 ((int16_t *)(uintptr_t)1)
static int matcher_scaled_dev(b2s_ctx *c, const uint8_t *dl, const uint8_t *dr, int H, int W, int cn, int16_t *d_out16, float *d_outf, bool timed)
{
    int nh = H, nw = W;
    const int longest = H > W ? H : W;
    if (c->max_size > 0 && longest > c->max_size) {
        const double ratio = (double)c->max_size / (double)longest;
        nh = (int)nearbyint((double)H * ratio);
        nw = (int)nearbyint((double)W * ratio);
    }
    if (nh == H && nw == W) {
        int rc = make_geom(c, H, W, cn);
        if (rc) return rc;
        return matcher_dev(c, dl, dr, (d_out16 && d_out16 != B2S_OWN16) ? d_out16 : c->disp16.as<int16_t>(), d_outf, timed);
    }
    if (d_out16) return fail(c, B2S_EINVAL, "the int16 disparity is not defined when the matcher works on a down-scaled pair (max_size %d < %d): ask for the float disparity", c->max_size, longest);
    if (!d_outf) return fail(c, B2S_EINVAL, "no output requested");
    if (nh <= 0 || nw <= 0) return fail(c, B2S_ESIZE, "max_size %d leaves no image", c->max_size);
    int rc = make_geom(c, nh, nw, cn);
    if (rc) return rc;
    const size_t ns = (size_t)nh * nw;
    CK(c, c->sleft.ensure(ns * cn));
    CK(c, c->sright.ensure(ns * cn));
    CK(c, c->sdispf.ensure(ns * 4));
    CK(c, launch_resize_u8(c, dl, H, W, cn, c->sleft.as<uint8_t>(), nh, nw));
    CK(c, launch_resize_u8(c, dr, H, W, cn, c->sright.as<uint8_t>(), nh, nw));
    if ((rc = matcher_dev(c, c->sleft.as<uint8_t>(), c->sright.as<uint8_t>(), c->disp16.as<int16_t>(), c->sdispf.as<float>(), timed))) return rc;
    CK(c, launch_resize_f32(c, c->sdispf.as<float>(), nh, nw, d_outf, H, W, (float)W, (float)nw));
    if (timed) cudaEventRecord(c->ev[5], c->stream); // (the up-scale belongs to the post stage)
    return B2S_OK;
}

static void collect_timing(b2s_ctx *c, bool chain)
{
    auto ms = [&](int a, int b) { float t = 0; cudaEventElapsedTime(&t, c->ev[a], c->ev[b]); return t; };
    c->timing.rectify_ms = chain ? ms(0, 1) : 0.f;
    c->timing.cost_ms = ms(1, 2);
    c->timing.aggregate_ms = ms(2, 3);
    c->timing.wta_ms = ms(3, 4);
    c->timing.post_ms = ms(4, 5);
    c->timing.depth_ms = chain ? ms(5, 6) : 0.f;
    c->timing.total_ms = ms(0, chain ? 6 : 5);
}

static int compute_disparity_host(b2s_ctx *c, const uint8_t *left, const uint8_t *right, int H, int W, int cn, int16_t *out16, float *outf,
                                  bool sync)
{
    if (!c || !left || !right) return B2S_EINVAL;
    CK(c, cudaSetDevice(c->device));
    if (cn != 1 && cn != 3) return fail(c, B2S_EINVAL, "images must have 1 or 3 channels (got %d)", cn);
    if (H <= 0 || W <= 0 || W > 65535) return fail(c, B2S_EINVAL, "bad image size %dx%d", W, H);
    size_t nb = (size_t)H * W * cn, npx = (size_t)H * W;
    CK(c, c->left.ensure(nb));
    CK(c, c->right.ensure(nb));
    if (outf) CK(c, c->dispf.ensure(npx * 4));
    long long l0 = c->launches;
    cudaEventRecord(c->ev[0], c->stream);
    CK(c, cudaMemcpyAsync(c->left.p, left, nb, cudaMemcpyDefault, c->stream));
    CK(c, cudaMemcpyAsync(c->right.p, right, nb, cudaMemcpyDefault, c->stream));
    int rc = matcher_scaled_dev(c, c->left.as<uint8_t>(), c->right.as<uint8_t>(), H, W, cn, out16 ? B2S_OWN16 : nullptr,
                                outf ? c->dispf.as<float>() : nullptr, true);
    if (rc) return rc;
    if (out16) CK(c, cudaMemcpyAsync(out16, c->disp16.p, npx * 2, cudaMemcpyDefault, c->stream));
    if (outf) CK(c, cudaMemcpyAsync(outf, c->dispf.p, npx * 4, cudaMemcpyDefault, c->stream));
    c->timing.total_launches = (int)(c->launches - l0);
    if (sync) {
        CK(c, cudaStreamSynchronize(c->stream));
        collect_timing(c, false);
        if (int ae = agg_poll_error(c)) return agg_fail(c, ae);
    }
    return B2S_OK;
}

extern "C" {

int b2s_compute_disparity(b2s_handle c, const uint8_t *left, const uint8_t *right, int H, int W, int cn, int16_t *out16, float *outf)
{
    return compute_disparity_host(c, left, right, H, W, cn, out16, outf, true);
}
int b2s_compute_disparity_async(b2s_handle c, const uint8_t *left, const uint8_t *right, int H, int W, int cn, int16_t *out16, float *outf)
{
    return compute_disparity_host(c, left, right, H, W, cn, out16, outf, false);
}
int b2s_compute_disparity_dev(b2s_handle c, const uint8_t *dl, const uint8_t *dr, int H, int W, int cn, int16_t *d_out16, float *d_outf)
{
    if (!c || !dl || !dr) return B2S_EINVAL;
    CK(c, cudaSetDevice(c->device));
    long long l0 = c->launches;
    int rc = matcher_scaled_dev(c, dl, dr, H, W, cn, d_out16, d_outf, false);
    c->timing.total_launches = (int)(c->launches - l0);
    return rc;
}

int b2s_set_rig(b2s_handle c, const b2s_rig *r)
{
    if (!c || !r) return B2S_EINVAL;
    if (!r->map1x || !r->map1y || !r->map2x || !r->map2y || !r->valid_mask1) return fail(c, B2S_EINVAL, "b2s_set_rig: rectify maps / mask missing");
    if (r->W <= 0 || r->H <= 0 || r->W1 <= 0 || r->H1 <= 0 || r->W2 <= 0 || r->H2 <= 0) return fail(c, B2S_EINVAL, "b2s_set_rig: bad sizes");
    if (r->interp != 0 && r->interp != 1) return fail(c, B2S_EINVAL, "b2s_set_rig: interp must be 0 (LANCZOS4) or 1 (LINEAR)");
    CK(c, cudaSetDevice(c->device));
    size_t n = (size_t)r->W * r->H, n1 = (size_t)r->W1 * r->H1;
    struct Up { DevBuf *b; const void *src; size_t bytes; } ups[] = {
        {&c->map1x, r->map1x, n * 4}, {&c->map1y, r->map1y, n * 4}, {&c->map2x, r->map2x, n * 4}, {&c->map2y, r->map2y, n * 4},
        {&c->vmask, r->valid_mask1, n}, {&c->umapx, r->unrect_mapx, n1 * 4}, {&c->umapy, r->unrect_mapy, n1 * 4},
        {&c->und_xy, r->undist_xy, n1 * 4}, {&c->und_fxy, r->undist_fxy, n1 * 2}};
    for (auto &u : ups) {
        if (!u.src) continue;
        CK(c, u.b->ensure(u.bytes));
        CK(c, cudaMemcpyAsync(u.b->p, u.src, u.bytes, cudaMemcpyDefault, c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    c->rW = r->W; c->rH = r->H; c->rW1 = r->W1; c->rH1 = r->H1; c->rW2 = r->W2; c->rH2 = r->H2;
    c->r_min_disp = r->min_disparity; c->r_interp = r->interp;
    c->r_m[0] = r->unrect_m[0]; c->r_m[1] = r->unrect_m[1]; c->r_m[2] = r->unrect_m[2];
    c->r_fxb = r->fx_baseline; c->r_max_depth = r->max_depth;
    c->have_rig = true;
    c->have_map_params = false;
    CK(c, c->dispfinal.ensure(n * 4));
    CK(c, c->rdepth.ensure(n * 8));
    CK(c, c->udepth.ensure(n1 * 8));
    return B2S_OK;
}

int b2s_set_rig_params(b2s_handle c, const b2s_rig_params *r)
{
    if (!c || !r) return B2S_EINVAL;
    if (r->W <= 0 || r->H <= 0 || r->W1 <= 0 || r->H1 <= 0 || r->W2 <= 0 || r->H2 <= 0) return fail(c, B2S_EINVAL, "b2s_set_rig_params: bad sizes");
    if (r->interp != 0 && r->interp != 1) return fail(c, B2S_EINVAL, "b2s_set_rig_params: interp must be 0 (LANCZOS4) or 1 (LINEAR)");
    if (r->rect1.W != r->W || r->rect1.H != r->H || r->rect2.W != r->W || r->rect2.H != r->H || r->unrect.W != r->W1 || r->unrect.H != r->H1 ||
        r->undist.W != r->W1 || r->undist.H != r->H1)
        return fail(c, B2S_EINVAL, "b2s_set_rig_params: map sizes do not match the rig sizes");
    CK(c, cudaSetDevice(c->device));
    size_t n = (size_t)r->W * r->H, n1 = (size_t)r->W1 * r->H1;
    CK(c, c->map1x.ensure(n * 4)); CK(c, c->map1y.ensure(n * 4)); CK(c, c->map2x.ensure(n * 4)); CK(c, c->map2y.ensure(n * 4));
    CK(c, c->vmask.ensure(n)); CK(c, c->umapx.ensure(n1 * 4)); CK(c, c->umapy.ensure(n1 * 4));
    CK(c, c->und_xy.ensure(n1 * 4)); CK(c, c->und_fxy.ensure(n1 * 2));
    CK(c, launch_gen_maps(c, r->rect1, c->map1x.as<float>(), c->map1y.as<float>(), c->vmask.as<uint8_t>(), r->W1, r->H1, nullptr, nullptr));
    CK(c, launch_gen_maps(c, r->rect2, c->map2x.as<float>(), c->map2y.as<float>(), nullptr, 0, 0, nullptr, nullptr));
    CK(c, launch_gen_maps(c, r->unrect, c->umapx.as<float>(), c->umapy.as<float>(), nullptr, 0, 0, nullptr, nullptr));
    CK(c, launch_gen_maps(c, r->undist, nullptr, nullptr, nullptr, 0, 0, c->und_xy.as<int16_t>(), c->und_fxy.as<uint16_t>()));
    CK(c, cudaStreamSynchronize(c->stream));
    c->rW = r->W; c->rH = r->H; c->rW1 = r->W1; c->rH1 = r->H1; c->rW2 = r->W2; c->rH2 = r->H2;
    c->r_min_disp = r->min_disparity; c->r_interp = r->interp;
    c->r_m[0] = r->unrect_m[0]; c->r_m[1] = r->unrect_m[1]; c->r_m[2] = r->unrect_m[2];
    c->r_fxb = r->fx_baseline; c->r_max_depth = r->max_depth;
    c->have_rig = true;
    c->have_map_params = true;
    c->mp_rect1 = r->rect1;
    c->mp_rect2 = r->rect2;
    c->cam1_f[0] = r->undist.fx; c->cam1_f[1] = r->undist.fy; c->cam1_f[2] = r->undist.cx; c->cam1_f[3] = r->undist.cy;
    memcpy(c->cam1_k, r->undist.k, sizeof c->cam1_k);
    c->have_cam1 = true;
    CK(c, c->dispfinal.ensure(n * 4));
    CK(c, c->rdepth.ensure(n * 8));
    CK(c, c->udepth.ensure(n1 * 8));
    return B2S_OK;
}

int b2s_project_depth(b2s_handle c, const double *depth2, int W2, int H2, double rate, const double K2inv[9], const double T[16],
                      const double K1[9], int W1, int H1, double *out)
{
    if (!c || !depth2 || !K2inv || !T || !K1 || !out) return B2S_EINVAL;
    if (W2 <= 0 || H2 <= 0 || W1 <= 0 || H1 <= 0 || !(rate > 0) || rate > 64) return fail(c, B2S_EINVAL, "b2s_project_depth: bad sizes or rate");
    CK(c, cudaSetDevice(c->device));
    const size_t n2 = (size_t)W2 * H2, n1 = (size_t)W1 * H1;
    CK(c, c->pin.ensure(n2 * 8));
    CK(c, c->pout.ensure(n1 * 8));
    CK(c, c->pkey.ensure(n1 * 8));
    CK(c, cudaMemcpyAsync(c->pin.p, depth2, n2 * 8, cudaMemcpyDefault, c->stream));
    CK(c, launch_project_depth(c, c->pin.as<double>(), W2, H2, rate, K2inv, T, K1, W1, H1, c->pkey.as<unsigned long long>(), c->pout.as<double>()));
    CK(c, cudaMemcpyAsync(out, c->pout.p, n1 * 8, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_depth_to_point_cloud(b2s_handle c, const double *depth, int H, int W, double rate, const double Kinv[9], int with_uv, double *out,
                             unsigned long long capacity, unsigned long long *n_out)
{
    if (!c || !depth || !Kinv || !out || !n_out) return B2S_EINVAL;
    if (W <= 0 || H <= 0 || !(rate > 0) || rate > 64) return fail(c, B2S_EINVAL, "b2s_depth_to_point_cloud: bad size or rate");
    CK(c, cudaSetDevice(c->device));
    const int cols = with_uv ? 5 : 3;
    const int Hu = rate == 1.0 ? H : (int)nearbyint(H * rate), Wu = rate == 1.0 ? W : (int)nearbyint(W * rate);
    if (Hu <= 0 || Wu <= 0) return fail(c, B2S_EINVAL, "b2s_depth_to_point_cloud: rate %g leaves no image", rate);
    const unsigned long long maxn = (unsigned long long)Hu * Wu, cap = capacity < maxn ? capacity : maxn;
    CK(c, c->pin.ensure((size_t)H * W * 8));
    CK(c, c->cl_pts.ensure((size_t)(cap ? cap : 1) * cols * 8));
    CK(c, c->cl_aux.ensure((size_t)Hu * 4 + 256 + ((size_t)Hu + 1) * 8));
    unsigned *rowcount = c->cl_aux.as<unsigned>();
    unsigned long long *rowoff = (unsigned long long *)((char *)c->cl_aux.p + (((size_t)Hu * 4 + 255) / 256) * 256);
    CK(c, cudaMemcpyAsync(c->pin.p, depth, (size_t)H * W * 8, cudaMemcpyDefault, c->stream));
    int hu = 0;
    CK(c, launch_depth_to_cloud(c, c->pin.as<double>(), H, W, rate, Kinv, cols, c->cl_pts.as<double>(), cap, rowcount, rowoff, &hu));
    unsigned long long n = 0;
    CK(c, cudaMemcpyAsync(&n, rowoff + hu, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *n_out = n;
    if (n > capacity) return fail(c, B2S_ESIZE, "b2s_depth_to_point_cloud: %llu points, capacity %llu", n, capacity);
    if (n) CK(c, cudaMemcpy(out, c->cl_pts.p, (size_t)n * cols * 8, cudaMemcpyDefault));
    return B2S_OK;
}

int b2s_point_cloud_to_depth(b2s_handle c, const double *points, unsigned long long n, const double K[9], int W, int H, double bg_value, double *out)
{
    if (!c || (!points && n) || !K || !out) return B2S_EINVAL;
    if (W <= 0 || H <= 0) return fail(c, B2S_EINVAL, "b2s_point_cloud_to_depth: bad size");
    CK(c, cudaSetDevice(c->device));
    const size_t npx = (size_t)W * H;
    CK(c, c->cl_pts.ensure((size_t)(n ? n : 1) * 24));
    CK(c, c->cl_aux.ensure(256));
    CK(c, c->pkey.ensure(npx * 8));
    CK(c, c->pout.ensure(npx * 8));
    if (n) CK(c, cudaMemcpyAsync(c->cl_pts.p, points, (size_t)n * 24, cudaMemcpyDefault, c->stream));
    CK(c, cudaMemcpyAsync(c->cl_aux.p, K, 72, cudaMemcpyHostToDevice, c->stream));
    CK(c, launch_cloud_to_depth(c, c->cl_pts.as<double>(), n, c->cl_aux.as<double>(), W, H, c->pkey.as<unsigned long long>(), c->pout.as<double>(), bg_value));
    CK(c, cudaMemcpyAsync(out, c->pout.p, npx * 8, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_resize_nearest_f32(b2s_handle c, const float *src, int sH, int sW, float *dst, int dH, int dW, float mul)
{
    if (!c || !src || !dst) return B2S_EINVAL;
    if (sH <= 0 || sW <= 0 || dH <= 0 || dW <= 0) return fail(c, B2S_EINVAL, "b2s_resize_nearest_f32: bad sizes");
    CK(c, cudaSetDevice(c->device));
    CK(c, c->pin.ensure((size_t)sH * sW * 4));
    CK(c, c->pout.ensure((size_t)dH * dW * 4));
    CK(c, cudaMemcpyAsync(c->pin.p, src, (size_t)sH * sW * 4, cudaMemcpyDefault, c->stream));
    CK(c, launch_resize_nearest_f32(c, c->pin.as<float>(), sH, sW, c->pout.as<float>(), dH, dW, mul));
    CK(c, cudaMemcpyAsync(dst, c->pout.p, (size_t)dH * dW * 4, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_interpolate_sparse(b2s_handle c, int inter_type, const double *uvz, int n, const double abc[3], const uint8_t *mask, int H, int W,
                           double distance, float *out)
{
    if (!c || !out) return B2S_EINVAL;
    if (W <= 0 || H <= 0) return fail(c, B2S_EINVAL, "b2s_interpolate_sparse: bad size");
    if (inter_type != 0 && inter_type != 1) return fail(c, B2S_EINVAL, "b2s_interpolate_sparse: inter_type 0 (lstsq plane) or 1 (nearest), got %d", inter_type);
    if (inter_type == 0 && !abc) return fail(c, B2S_EINVAL, "b2s_interpolate_sparse: the plane coefficients are missing");
    if (inter_type == 1 && (n < 0 || (n > 0 && !uvz))) return fail(c, B2S_EINVAL, "b2s_interpolate_sparse: the points are missing");
    CK(c, cudaSetDevice(c->device));
    const size_t npx = (size_t)W * H;
    CK(c, c->pout.ensure(npx * 8));
    uint8_t *dmask = nullptr;
    if (mask) {
        CK(c, c->pin.ensure(npx));
        CK(c, cudaMemcpyAsync(c->pin.p, mask, npx, cudaMemcpyDefault, c->stream));
        dmask = c->pin.as<uint8_t>();
    }
    if (inter_type == 0) {
        CK(c, launch_plane_fill(c, dmask, H, W, abc[0], abc[1], abc[2], c->pout.as<float>()));
    } else {
        CK(c, c->cl_pts.ensure((size_t)(n ? n : 1) * 24));
        if (n) CK(c, cudaMemcpyAsync(c->cl_pts.p, uvz, (size_t)n * 24, cudaMemcpyDefault, c->stream));
        CK(c, launch_nearest_fill(c, c->cl_pts.as<double>(), n, dmask, H, W, distance, c->pout.as<float>()));
    }
    CK(c, cudaMemcpyAsync(out, c->pout.p, npx * 4, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_interpolate_rbf(b2s_handle c, const double *uvw, int n, const uint8_t *mask, int H, int W, double *out)
{
    if (!c || !out) return B2S_EINVAL;
    if (W <= 0 || H <= 0) return fail(c, B2S_EINVAL, "b2s_interpolate_rbf: bad size");
    if (n < 0 || (n > 0 && !uvw)) return fail(c, B2S_EINVAL, "b2s_interpolate_rbf: the samples are missing");
    CK(c, cudaSetDevice(c->device));
    const size_t npx = (size_t)W * H;
    CK(c, c->pout.ensure(npx * 8));
    uint8_t *dmask = nullptr;
    if (mask) {
        CK(c, c->pin.ensure(npx));
        CK(c, cudaMemcpyAsync(c->pin.p, mask, npx, cudaMemcpyDefault, c->stream));
        dmask = c->pin.as<uint8_t>();
    }
    CK(c, c->cl_pts.ensure((size_t)(n ? n : 1) * 24));
    if (n) CK(c, cudaMemcpyAsync(c->cl_pts.p, uvw, (size_t)n * 24, cudaMemcpyDefault, c->stream));
    CK(c, launch_rbf_fill(c, c->cl_pts.as<double>(), n, dmask, H, W, c->pout.as<double>()));
    CK(c, cudaMemcpyAsync(out, c->pout.p, npx * 8, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_set_cam1_model(b2s_handle c, double fx, double fy, double cx, double cy, const double k[12])
{
    if (!c || !k) return B2S_EINVAL;
    if (fx == 0 || fy == 0) return fail(c, B2S_EINVAL, "b2s_set_cam1_model: zero focal length");
    c->cam1_f[0] = fx; c->cam1_f[1] = fy; c->cam1_f[2] = cx; c->cam1_f[3] = cy;
    memcpy(c->cam1_k, k, sizeof c->cam1_k);
    c->have_cam1 = true;
    return B2S_OK;
}

int b2s_distort_depth(b2s_handle c, const double *depth, double *out)
{
    if (!c || !depth || !out) return B2S_EINVAL;
    if (!c->have_rig || !c->have_cam1) return fail(c, B2S_ESTATE, "rig and cam1 model (b2s_set_cam1_model) have not been set");
    CK(c, cudaSetDevice(c->device));
    size_t n1 = (size_t)c->rW1 * c->rH1;
    CK(c, c->udepth.ensure(n1 * 8));
    CK(c, c->ddepth.ensure(n1 * 8));
    CK(c, cudaMemcpyAsync(c->udepth.p, depth, n1 * 8, cudaMemcpyDefault, c->stream));
    CK(c, launch_distort_depth(c, c->udepth.as<double>(), c->ddepth.as<double>()));
    CK(c, cudaMemcpyAsync(out, c->ddepth.p, n1 * 8, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

} // extern "C"

static int rectify_dev(b2s_ctx *c, const uint8_t *img1, const uint8_t *img2, int cn)
{
    size_t n1 = (size_t)c->rW1 * c->rH1 * cn, n2 = (size_t)c->rW2 * c->rH2 * cn, n = (size_t)c->rW * c->rH * cn;
    CK(c, c->img1.ensure(n1));
    CK(c, c->img2.ensure(n2));
    CK(c, c->rect1.ensure(n));
    CK(c, c->rect2.ensure(n));
    CK(c, cudaMemcpyAsync(c->img1.p, img1, n1, cudaMemcpyDefault, c->stream));
    CK(c, cudaMemcpyAsync(c->img2.p, img2, n2, cudaMemcpyDefault, c->stream));
    const bool analytic = c->have_map_params && c->r_interp == 0; // rig given by parameters: the remap evaluates the maps itself
    CK(c, launch_remap_u8(c, c->img1.as<uint8_t>(), c->rH1, c->rW1, cn, c->map1x.as<float>(), c->map1y.as<float>(), c->rH, c->rW, 0,
                          c->r_interp, c->rect1.as<uint8_t>(), analytic ? &c->mp_rect1 : nullptr));
    CK(c, launch_remap_u8(c, c->img2.as<uint8_t>(), c->rH2, c->rW2, cn, c->map2x.as<float>(), c->map2y.as<float>(), c->rH, c->rW,
                          c->r_min_disp, c->r_interp, c->rect2.as<uint8_t>(), analytic ? &c->mp_rect2 : nullptr));
    return B2S_OK;
}

static int depth_tail(b2s_ctx *c, const float *d_disp, int add_min, const uint8_t *d_img1, int cn, int want_unrectify, const b2s_depth_out *o)
{
    if (want_unrectify && (!c->umapx.p || !c->umapy.p)) return fail(c, B2S_ESTATE, "rig has no unrectify maps");
    CK(c, launch_depth(c, d_disp, add_min, want_unrectify));
    size_t n = (size_t)c->rW * c->rH, n1 = (size_t)c->rW1 * c->rH1;
    bool und = want_unrectify && d_img1 && o->undistort_img1;
    if (und) {
        if (!c->und_xy.p || !c->und_fxy.p) return fail(c, B2S_ESTATE, "rig has no undistort maps");
        CK(c, c->und1.ensure(n1 * cn));
        CK(c, launch_undistort_u8(c, d_img1, c->rH1, c->rW1, cn, c->und_xy.as<int16_t>(), c->und_fxy.as<uint16_t>(), c->und1.as<uint8_t>()));
    }
    const bool dist = want_unrectify && o->distort_depth;
    if (dist) {
        if (!c->have_cam1) return fail(c, B2S_ESTATE, "distort_depth needs b2s_set_cam1_model");
        CK(c, c->ddepth.ensure(n1 * 8));
        CK(c, launch_distort_depth(c, c->udepth.as<double>(), c->ddepth.as<double>()));
    }
    cudaEventRecord(c->ev[6], c->stream);
    if (dist) CK(c, cudaMemcpyAsync(o->distort_depth, c->ddepth.p, n1 * 8, cudaMemcpyDefault, c->stream));
    if (o->disparity) CK(c, cudaMemcpyAsync(o->disparity, c->dispfinal.p, n * 4, cudaMemcpyDefault, c->stream));
    if (o->rectify_depth) CK(c, cudaMemcpyAsync(o->rectify_depth, c->rdepth.p, n * 8, cudaMemcpyDefault, c->stream));
    if (want_unrectify && o->unrectify_depth) CK(c, cudaMemcpyAsync(o->unrectify_depth, c->udepth.p, n1 * 8, cudaMemcpyDefault, c->stream));
    if (und) CK(c, cudaMemcpyAsync(o->undistort_img1, c->und1.p, n1 * cn, cudaMemcpyDefault, c->stream));
    return B2S_OK;
}

static int get_depth_impl(b2s_ctx *c, const uint8_t *img1, const uint8_t *img2, int cn, int want_unrectify, const b2s_depth_out *o, bool sync)
{
    if (!c || !img1 || !img2 || !o) return B2S_EINVAL;
    if (!c->have_rig) return fail(c, B2S_ESTATE, "b2s_set_rig has not been called");
    CK(c, cudaSetDevice(c->device));
    int rc;
    if (cn != 1 && cn != 3) return fail(c, B2S_EINVAL, "images must have 1 or 3 channels (got %d)", cn);
    const bool scaled = c->max_size > 0 && (c->rH > c->rW ? c->rH : c->rW) > c->max_size;
    if (scaled && o->disp16) return fail(c, B2S_EINVAL, "the int16 disparity is not defined when the matcher works on a down-scaled pair (max_size)");
    CK(c, c->dispf.ensure((size_t)c->rW * c->rH * 4));
    long long l0 = c->launches;
    cudaEventRecord(c->ev[0], c->stream);
    if ((rc = rectify_dev(c, img1, img2, cn))) return rc;
    if ((rc = matcher_scaled_dev(c, c->rect1.as<uint8_t>(), c->rect2.as<uint8_t>(), c->rH, c->rW, cn, scaled ? nullptr : B2S_OWN16,
                                 c->dispf.as<float>(), true)))
        return rc;
    if ((rc = depth_tail(c, c->dispf.as<float>(), 1, c->img1.as<uint8_t>(), cn, want_unrectify, o))) return rc;
    size_t n = (size_t)c->rW * c->rH;
    if (o->rectify_img1) CK(c, cudaMemcpyAsync(o->rectify_img1, c->rect1.p, n * cn, cudaMemcpyDefault, c->stream));
    if (o->rectify_img2) CK(c, cudaMemcpyAsync(o->rectify_img2, c->rect2.p, n * cn, cudaMemcpyDefault, c->stream));
    if (o->disp16) CK(c, cudaMemcpyAsync(o->disp16, c->disp16.p, n * 2, cudaMemcpyDefault, c->stream));
    c->timing.total_launches = (int)(c->launches - l0);
    if (sync) {
        CK(c, cudaStreamSynchronize(c->stream));
        collect_timing(c, true);
        if (int ae = agg_poll_error(c)) return agg_fail(c, ae);
    }
    return B2S_OK;
}

extern "C" {

int b2s_rectify(b2s_handle c, const uint8_t *img1, const uint8_t *img2, int cn, uint8_t *out1, uint8_t *out2)
{
    if (!c || !img1 || !img2) return B2S_EINVAL;
    if (!c->have_rig) return fail(c, B2S_ESTATE, "b2s_set_rig has not been called");
    if (cn != 1 && cn != 3) return fail(c, B2S_EINVAL, "images must have 1 or 3 channels (got %d)", cn);
    CK(c, cudaSetDevice(c->device));
    int rc = rectify_dev(c, img1, img2, cn);
    if (rc) return rc;
    size_t n = (size_t)c->rW * c->rH * cn;
    if (out1) CK(c, cudaMemcpyAsync(out1, c->rect1.p, n, cudaMemcpyDefault, c->stream));
    if (out2) CK(c, cudaMemcpyAsync(out2, c->rect2.p, n, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_get_depth(b2s_handle c, const uint8_t *img1, const uint8_t *img2, int cn, int want_unrectify, const b2s_depth_out *o)
{
    return get_depth_impl(c, img1, img2, cn, want_unrectify, o, true);
}
int b2s_get_depth_async(b2s_handle c, const uint8_t *img1, const uint8_t *img2, int cn, int want_unrectify, const b2s_depth_out *o)
{
    return get_depth_impl(c, img1, img2, cn, want_unrectify, o, false);
}

int b2s_depth_from_disparity(b2s_handle c, const float *disparity, const uint8_t *img1, int cn, int want_unrectify, const b2s_depth_out *o)
{
    if (!c || !disparity || !o) return B2S_EINVAL;
    if (!c->have_rig) return fail(c, B2S_ESTATE, "b2s_set_rig has not been called");
    if (img1 && cn != 1 && cn != 3) return fail(c, B2S_EINVAL, "images must have 1 or 3 channels (got %d)", cn);
    CK(c, cudaSetDevice(c->device));
    size_t n = (size_t)c->rW * c->rH, n1 = (size_t)c->rW1 * c->rH1;
    CK(c, c->stage_f32.ensure(n * 4));
    CK(c, cudaMemcpyAsync(c->stage_f32.p, disparity, n * 4, cudaMemcpyDefault, c->stream));
    if (img1) {
        CK(c, c->img1.ensure(n1 * cn));
        CK(c, cudaMemcpyAsync(c->img1.p, img1, n1 * cn, cudaMemcpyDefault, c->stream));
    }
    int rc = depth_tail(c, c->stage_f32.as<float>(), 1, img1 ? c->img1.as<uint8_t>() : nullptr, cn, want_unrectify, o);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_disparity_to_depth(b2s_handle c, const float *disparity, double *depth)
{
    if (!c || !disparity || !depth) return B2S_EINVAL;
    if (!c->have_rig) return fail(c, B2S_ESTATE, "b2s_set_rig has not been called");
    CK(c, cudaSetDevice(c->device));
    size_t n = (size_t)c->rW * c->rH;
    CK(c, c->stage_f32.ensure(n * 4));
    CK(c, cudaMemcpyAsync(c->stage_f32.p, disparity, n * 4, cudaMemcpyDefault, c->stream));
    CK(c, launch_depth_bare(c, c->stage_f32.as<float>(), c->rdepth.as<double>()));
    CK(c, cudaMemcpyAsync(depth, c->rdepth.p, n * 8, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_unrectify_depth(b2s_handle c, const double *rectify_depth, double *out)
{
    if (!c || !rectify_depth || !out) return B2S_EINVAL;
    if (!c->have_rig || !c->umapx.p || !c->umapy.p) return fail(c, B2S_ESTATE, "rig (with unrectify maps) has not been set");
    CK(c, cudaSetDevice(c->device));
    size_t n = (size_t)c->rW * c->rH, n1 = (size_t)c->rW1 * c->rH1;
    CK(c, cudaMemcpyAsync(c->rdepth.p, rectify_depth, n * 8, cudaMemcpyDefault, c->stream));
    CK(c, launch_unrectify(c, c->rdepth.as<double>(), c->udepth.as<double>()));
    CK(c, cudaMemcpyAsync(out, c->udepth.p, n1 * 8, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_undistort_img(b2s_handle c, const uint8_t *img1, int cn, uint8_t *out)
{
    if (!c || !img1 || !out) return B2S_EINVAL;
    if (cn != 1 && cn != 3) return fail(c, B2S_EINVAL, "images must have 1 or 3 channels (got %d)", cn);
    if (!c->have_rig || !c->und_xy.p || !c->und_fxy.p) return fail(c, B2S_ESTATE, "rig (with undistort maps) has not been set");
    CK(c, cudaSetDevice(c->device));
    size_t n1 = (size_t)c->rW1 * c->rH1 * cn;
    CK(c, c->img1.ensure(n1));
    CK(c, c->und1.ensure(n1));
    CK(c, cudaMemcpyAsync(c->img1.p, img1, n1, cudaMemcpyDefault, c->stream));
    CK(c, launch_undistort_u8(c, c->img1.as<uint8_t>(), c->rH1, c->rW1, cn, c->und_xy.as<int16_t>(), c->und_fxy.as<uint16_t>(), c->und1.as<uint8_t>()));
    CK(c, cudaMemcpyAsync(out, c->und1.p, n1, cudaMemcpyDefault, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int b2s_volume_dims(b2s_handle c, int *H, int *width1, int *D, int *Dp)
{
    if (!c || !c->have_volume) return c ? fail(c, B2S_ESTATE, "no cost volume resident") : B2S_EINVAL;
    if (H) *H = c->g.H;
    if (width1) *width1 = c->g.width1;
    if (D) *D = c->g.D;
    if (Dp) *Dp = c->g.Dp;
    return B2S_OK;
}

int b2s_debug_fetch(b2s_handle c, int which, void *dst, size_t bytes)
{
    if (!c || !dst) return B2S_EINVAL;
    if (which < B2S_FETCH_RIG && !c->have_volume) return fail(c, B2S_ESTATE, "no cost volume resident");
    CK(c, cudaSetDevice(c->device));
    const SgbmGeom &g = c->g;
    size_t vol = (size_t)g.H * g.width1 * g.Dp * 2, need;
    const void *src;
    switch (which) {
    case B2S_FETCH_C: src = c->C.p; need = vol; break;
    case B2S_FETCH_S:
        if ((c->wta_fused || (c->wta_adds_s2 && c->fuse_wta)) && !c->keep_volumes)
            return fail(c, B2S_ESTATE, "the aggregated volume was not stored (winner-take-all is fused into the last scan); "
                                       "call b2s_set_option(h, B2S_OPT_KEEP_VOLUMES, 1) before computing");
        src = c->S.p; need = vol; break;
    case B2S_FETCH_RAW: src = c->raw.p; need = (size_t)g.H * g.W * 2; break;
    default:
        if (which >= B2S_FETCH_RIG && which < B2S_FETCH_RIG + 9 && c->have_rig) {
            const size_t n = (size_t)c->rW * c->rH, n1 = (size_t)c->rW1 * c->rH1;
            const DevBuf *bufs[9] = {&c->map1x, &c->map1y, &c->map2x, &c->map2y, &c->vmask, &c->umapx, &c->umapy, &c->und_xy, &c->und_fxy};
            const size_t sizes[9] = {n * 4, n * 4, n * 4, n * 4, n, n1 * 4, n1 * 4, n1 * 4, n1 * 2};
            src = bufs[which - B2S_FETCH_RIG]->p; need = sizes[which - B2S_FETCH_RIG];
            if (!src) return fail(c, B2S_ESTATE, "b2s_debug_fetch: that rig array was not set");
            break;
        }
        return fail(c, B2S_EINVAL, "b2s_debug_fetch: unknown selector %d", which);
    }
    if (bytes != need) return fail(c, B2S_EINVAL, "b2s_debug_fetch: need %zu bytes, got %zu", need, bytes);
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost));
    if ((which == B2S_FETCH_C || which == B2S_FETCH_S) && g.layout != 0) {
        // the ABI promises disparities in natural order: undo the device layout (SgbmGeom::layout) pixel by pixel
        int16_t *px = (int16_t *)dst;
        std::vector<int16_t> tmp(g.Dp);
        for (size_t i = 0, n = (size_t)g.H * g.width1; i < n; i++, px += g.Dp) {
            for (int d = 0; d < g.Dp; d++) tmp[d] = px[b2s_dindex(g.layout, d)];
            memcpy(px, tmp.data(), (size_t)g.Dp * 2);
        }
    }
    return B2S_OK;
}

int b2s_set_option(b2s_handle c, int option, int value)
{
    if (!c) return B2S_EINVAL;
    switch (option) {
    case B2S_OPT_KEEP_VOLUMES: c->keep_volumes = value != 0; return B2S_OK;
    case B2S_OPT_FUSE_WTA: c->fuse_wta = value != 0; return B2S_OK;
    case B2S_OPT_MAX_SIZE:
        if (value < 0) return fail(c, B2S_EINVAL, "B2S_OPT_MAX_SIZE: >= 0 (0 = no limit), got %d", value);
        c->max_size = value;
        return B2S_OK;
    case B2S_OPT_AGG_SCHEDULE:
        if (value != 0 && value != 1) return fail(c, B2S_EINVAL, "B2S_OPT_AGG_SCHEDULE: 0 (scans + sweep) or 1 (wavefront), got %d", value);
        c->agg_schedule = value;
        return B2S_OK;
    default: return fail(c, B2S_EINVAL, "b2s_set_option: unknown option %d", option);
    }
}

int b2s_timings(b2s_handle c, b2s_timing *t)
{
    if (!c || !t) return B2S_EINVAL;
    *t = c->timing;
    return B2S_OK;
}

int b2s_launch_count(b2s_handle c, long long *n)
{
    if (!c || !n) return B2S_EINVAL;
    *n = c->launches;
    return B2S_OK;
}

// One production repetition of the aggregation group on the handle's stream: the cost stage is re-run first (untimed by the
// callers: it leaves the row sums the first aggregation launch consumes, exactly as in a real pair), then the group
// = launch_aggregate + the stand-alone winner-take-all when the last scan did not fuse it.  marks: see launch_aggregate.
static int agg_repetition(b2s_ctx *c, int *nl, cudaEvent_t *marks, cudaEvent_t before)
{
    if (c->last_dl && c->last_dr) CK(c, launch_cost_volume(c, c->last_dl, c->last_dr));
    CK(c, launch_wta_prepare(c));
    if (before) CK(c, cudaEventRecord(before, c->stream));
    CK(c, launch_aggregate(c, nl, marks));
    if (!c->wta_fused) {
        CK(c, launch_wta(c));
        if (marks && *nl < B2S_AGG_MAX_PARTS) CK(c, cudaEventRecord(marks[++*nl], c->stream));
    }
    return B2S_OK;
}

int b2s_bench_aggregate(b2s_handle c, int iters, float *ms_per_iter)
{
    if (!c || !ms_per_iter || iters <= 0) return B2S_EINVAL;
    if (!c->have_volume) return fail(c, B2S_ESTATE, "no cost volume resident (run a disparity computation first)");
    CK(c, cudaSetDevice(c->device));
    int nl = 0, rc;
    if ((rc = agg_repetition(c, &nl, nullptr, nullptr))) return rc; // warm-up
    double total = 0;
    for (int i = 0; i < iters; i++) {
        if ((rc = agg_repetition(c, &nl, nullptr, c->ev[0]))) return rc;
        CK(c, cudaEventRecord(c->ev[7], c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        CK(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[7]));
        total += ms;
    }
    *ms_per_iter = (float)(total / iters);
    if (int ae = agg_poll_error(c)) return agg_fail(c, ae);
    return B2S_OK;
}

int b2s_enqueue_aggregate(b2s_handle c, int iters)
{
    if (!c || iters <= 0) return B2S_EINVAL;
    if (!c->have_volume) return fail(c, B2S_ESTATE, "no cost volume resident (run a disparity computation first)");
    CK(c, cudaSetDevice(c->device));
    // (several handles in flight; the cost stage cannot be kept out of this figure, so the group runs on the finished C:
    // its first launch is then the plain +x scan instead of the scan that also forms C)
    int nl = 0;
    for (int i = 0; i < iters; i++) {
        CK(c, launch_aggregate(c, &nl));
        CK(c, launch_wta(c));
    }
    return B2S_OK;
}

int b2s_bench_aggregate_parts(b2s_handle c, int iters, float *ms_parts, int max_parts, int *n_parts)
{
    if (!c || !ms_parts || !n_parts || iters <= 0 || max_parts <= 0) return B2S_EINVAL;
    if (!c->have_volume) return fail(c, B2S_ESTATE, "no cost volume resident (run a disparity computation first)");
    CK(c, cudaSetDevice(c->device));
    int nl = 0, rc;
    if ((rc = agg_repetition(c, &nl, nullptr, nullptr))) return rc; // warm-up
    for (int k = 0; k < max_parts; k++) ms_parts[k] = 0.f;
    for (int i = 0; i < iters; i++) {
        if ((rc = agg_repetition(c, &nl, c->aev, nullptr))) return rc;
        CK(c, cudaStreamSynchronize(c->stream));
        for (int k = 0; k < nl && k < max_parts && k < B2S_AGG_MAX_PARTS; k++) {
            float ms = 0;
            CK(c, cudaEventElapsedTime(&ms, c->aev[k], c->aev[k + 1]));
            ms_parts[k] += ms / iters;
        }
    }
    *n_parts = nl < max_parts ? nl : max_parts;
    if (int ae = agg_poll_error(c)) return agg_fail(c, ae);
    return B2S_OK;
}

int b2s_event_record(b2s_handle c, int slot)
{
    if (!c || slot < 0 || slot >= 4) return B2S_EINVAL;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaEventRecord(c->uev[slot], c->stream));
    return B2S_OK;
}

int b2s_event_elapsed(b2s_handle a, int slot_a, b2s_handle b, int slot_b, float *ms)
{
    if (!a || !b || !ms || slot_a < 0 || slot_a >= 4 || slot_b < 0 || slot_b >= 4) return B2S_EINVAL;
    CK(b, cudaSetDevice(b->device));
    CK(b, cudaEventSynchronize(b->uev[slot_b]));
    CK(b, cudaEventElapsedTime(ms, a->uev[slot_a], b->uev[slot_b]));
    return B2S_OK;
}

int b2s_collect_timings(b2s_handle c, int chain)
{
    if (!c) return B2S_EINVAL;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    collect_timing(c, chain != 0);
    return B2S_OK;
}

} // extern "C"
