// sgbm_post.cu -- winner-take-all, uniqueness, sub-pixel, right-view map, L/R check, 3x3 median, speckle filter and the
// reference's float post-processing, sm_100a.
//
// Replaces the tail of cv2.StereoSGBM.compute (calibrating/stereo_matching.py:63) and the NumPy post-processing at
// calibrating/stereo_matching.py:63-64 (+ the /16 of :66).
#include "b2s_internal.h"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- A.5: one warp per pixel ----------------------------------------------------------------------------------
// cv2 visits x1 from width1-1 down to 0 and keeps, per right-image column x2, the candidate with the smallest minS
// (strict '>' => among equal costs the largest x1 wins).  That order-dependent rule is an atomicMin on the key
// (minS << 16) | (0xFFFF - x1).
constexpr int WTA_CHUNK = 32;
// ADD2: the aggregated volume is sat16(S + S2) (the two sweeps of the wavefront schedule, sgbm_wave.cu), formed here on
// the fly; keep != 0 also writes it back to S (B2S_OPT_KEEP_VOLUMES: S stays fetchable).
template <int NP, bool ADD2, int LAYOUT>
__global__ void __launch_bounds__(256) wta_kernel(int16_t *__restrict__ S, const int16_t *__restrict__ S2, int16_t *__restrict__ raw,
                                                  unsigned *__restrict__ disp2key, SgbmGeom g, int keep)
{
    // a warp walks WTA_CHUNK consecutive pixels of one row (no index divisions); blockIdx.y = row
    const int lane = threadIdx.x & 31, y = blockIdx.y;
    const int xb = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * WTA_CHUNK, xe = min(xb + WTA_CHUNK, g.width1);
    if (xb >= xe) return;
    const int Dp = 64 * NP;
    int16_t *Srow = S + (size_t)y * g.width1 * Dp;
    const int16_t *S2row = S2 + (size_t)y * g.width1 * Dp;
    auto load_one = [&](const int16_t *Sp, uint32_t (&v)[NP]) {
        if constexpr (NP == 1) v[0] = __ldcs((const uint32_t *)(Sp + lane * 2));
        else if constexpr (NP == 2) { uint2 t = __ldcs((const uint2 *)(Sp + lane * 4)); v[0] = t.x; v[1] = t.y; }
        else if constexpr (NP == 4) { uint4 t = __ldcs((const uint4 *)(Sp + lane * 8)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else {
#pragma unroll
            for (int i = 0; i < NP; i++) v[i] = __ldcs((const uint32_t *)(Sp + lane * 2 * NP) + i);
        }
    };
    auto load_px = [&](int x, uint32_t (&v)[NP]) {
        load_one(Srow + (size_t)x * Dp, v);
        if constexpr (ADD2) {
            uint32_t v2[NP];
            load_one(S2row + (size_t)x * Dp, v2);
#pragma unroll
            for (int i = 0; i < NP; i++) v[i] = __viaddmin_u16x2(v[i], v2[i], 0x7FFF7FFFu);
            if (keep) {
#pragma unroll
                for (int i = 0; i < NP; i++) *((uint32_t *)(Srow + (size_t)x * Dp + lane * 2 * NP) + i) = v[i];
            }
        }
    };
    // three pixels in flight per warp: the body is one dependent chain per pixel, and with a single 256-byte load
    // outstanding per warp the kernel would be bound by DRAM latency, not bandwidth
    uint32_t buf[3][NP];
#pragma unroll
    for (int u = 0; u < 2; u++)
        if (xb + u < xe) load_px(xb + u, buf[u]);
    const int thr_mul = 100 - g.uniq;
    int my_best = -2, my_minS = 0; // lane i keeps the winner of pixel xb+i (-2: rejected by the uniqueness test)
    for (int base = xb; base < xe; base += 3) {
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int x = base + u;
            if (x >= xe) break;
            if (x + 2 < xe) load_px(x + 2, buf[(u + 2) % 3]);
            int sv[2 * NP];
            unsigned key = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < NP; i++) {
                sv[2 * i] = (int)(short)(buf[u][i] & 0xffffu);
                sv[2 * i + 1] = ((int)buf[u][i]) >> 16;
            }
#pragma unroll
            for (int j = 0; j < 2 * NP; j++) {
                const int d = b2s_word_d0(LAYOUT, lane * NP + (j >> 1)) + (j & 1) * (LAYOUT == 0 ? 1 : 8);
                if (d < g.D) key = min(key, ((unsigned)(sv[j] & 0xffff) << 16) | (unsigned)d);
            }
            key = __reduce_min_sync(0xffffffffu, key);
            const int minS = (int)(key >> 16);
            int best = (int)(key & 0xffffu);
            if (minS >= 32767) best = -1; // cv2: strict '<' against MAX_COST never fires
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 2 * NP; j++) {
                const int d = b2s_word_d0(LAYOUT, lane * NP + (j >> 1)) + (j & 1) * (LAYOUT == 0 ? 1 : 8);
                if (d < g.D && sv[j] * thr_mul < minS * 100 && abs(best - d) > 1) bad = true;
            }
            const bool rej = __any_sync(0xffffffffu, bad);
            if (lane == x - xb) {
                my_best = rej ? -2 : best;
                my_minS = minS;
            }
        }
    }
    if (ADD2 && keep) {
        __threadfence_block();
        __syncwarp();
    }
    // the per-pixel tail (right-view candidate, sub-pixel interpolation, store) runs once for the warp's 32 pixels, one per lane
    const int x = xb + lane;
    if (x < xe && my_best != -2) {
        int d = my_best;
        const int minS = my_minS;
        const int16_t *Sp = Srow + (size_t)x * Dp;
        int x2 = x + g.minX1 - d - g.minD;
        if (minS < 32767 && x2 >= 0 && x2 < g.W + 2)
            atomicMin(&disp2key[(size_t)y * (g.W + 2) + x2], ((unsigned)minS << 16) | (unsigned)(0xFFFF - x));
        if (0 < d && d < g.D - 1) {
            const int im = b2s_dindex(LAYOUT, d - 1), ip = b2s_dindex(LAYOUT, d + 1), i0 = b2s_dindex(LAYOUT, d);
            int sm = Sp[im], sp = Sp[ip], s0 = Sp[i0];
            if constexpr (ADD2) {
                if (!keep) { // (with keep the sums are already in S; the __syncwarp above orders the write-back before these reads)
                    const int16_t *Sq = S2row + (size_t)x * Dp;
                    sm = min(sm + Sq[im], 32767); sp = min(sp + Sq[ip], 32767); s0 = min(s0 + Sq[i0], 32767);
                }
            }
            int den2 = max(sm + sp - 2 * s0, 1);
            d = d * 16 + ((sm - sp) * 16 + den2) / (den2 * 2);
        } else
            d *= 16;
        raw[(size_t)y * g.W + x + g.minX1] = (int16_t)(d + g.minD * 16);
    }
}

__global__ void fill_i16_kernel(int16_t *p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- A.6 + A.7a: L/R check applied on load, then 3x3 median with replicated border --------------------------
__device__ __forceinline__ int disp2_at(const unsigned *__restrict__ keys, const SgbmGeom &g, int y, int xq)
{
    unsigned k = keys[(size_t)y * (g.W + 2) + xq];
    if (k == 0xFFFFFFFFu) return g.invalid; // never written: keeps the SCALED invalid value (A.6 quirk)
    int x1 = 0xFFFF - (int)(k & 0xffffu);
    return x1 + g.minX1 - xq; // best + minD
}

__device__ __forceinline__ int lr_checked(const int16_t *__restrict__ raw, const unsigned *__restrict__ keys, const SgbmGeom &g, int y, int x)
{
    int d1 = raw[(size_t)y * g.W + x];
    if (d1 == g.invalid || x < g.minX1) return d1;
    int _d = d1 >> 4, d_ = (d1 + 15) >> 4;
    int _x = x - _d, x_ = x - d_;
    if (0 <= _x && _x < g.W && 0 <= x_ && x_ < g.W) {
        int a = disp2_at(keys, g, y, _x), b = disp2_at(keys, g, y, x_);
        if (a >= g.minD && abs(a - _d) > g.d12 && b >= g.minD && abs(b - d_) > g.d12) return g.invalid;
    }
    return d1;
}

#define CSWAP(a, b) { int _t = min(a, b); b = max(a, b); a = _t; }
// CTA = 32 x 8 output pixels; the L/R-checked values of the 34 x 10 halo tile are computed ONCE into shared memory (each is 2
// right-view look-ups), then every thread takes the median of its 3 x 3 neighbourhood from there
constexpr int MED_TX = 32, MED_TY = 8;
__global__ void __launch_bounds__(MED_TX * MED_TY) lr_median_kernel(const int16_t *__restrict__ raw, const unsigned *__restrict__ keys,
                                                                    int16_t *__restrict__ out, SgbmGeom g)
{
    __shared__ int16_t tile[MED_TY + 2][MED_TX + 2];
    const int x0 = blockIdx.x * MED_TX, y0 = blockIdx.y * MED_TY;
    for (int i = threadIdx.y * MED_TX + threadIdx.x; i < (MED_TY + 2) * (MED_TX + 2); i += MED_TX * MED_TY) {
        const int ty = i / (MED_TX + 2), tx = i % (MED_TX + 2);
        tile[ty][tx] = (int16_t)lr_checked(raw, keys, g, clampi(y0 + ty - 1, 0, g.H - 1), clampi(x0 + tx - 1, 0, g.W - 1));
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    int v[9];
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) v[dy * 3 + dx] = tile[threadIdx.y + dy][threadIdx.x + dx];
    // median-of-9 exchange network
    CSWAP(v[1], v[2]); CSWAP(v[4], v[5]); CSWAP(v[7], v[8]); CSWAP(v[0], v[1]); CSWAP(v[3], v[4]); CSWAP(v[6], v[7]);
    CSWAP(v[1], v[2]); CSWAP(v[4], v[5]); CSWAP(v[7], v[8]); CSWAP(v[0], v[3]); CSWAP(v[5], v[8]); CSWAP(v[4], v[7]);
    CSWAP(v[3], v[6]); CSWAP(v[1], v[4]); CSWAP(v[2], v[5]); CSWAP(v[4], v[7]); CSWAP(v[4], v[2]); CSWAP(v[6], v[4]);
    CSWAP(v[4], v[2]);
    out[(size_t)y * g.W + x] = (int16_t)v[4];
}

// ---- A.7b: cv2.filterSpeckles as connected-component labelling ------------------------------------------------------
// Components are 4-connected sets of valid pixels whose neighbouring values differ by at most maxDiff; a component of at
// most maxSize pixels is replaced by newVal.  Labelling works on horizontal RUNS (maximal linked pixel sequences of a row):
//   rows  : label of every pixel = index of its run's first pixel (one block-wide max-scan per row)
//   merge : union-find over run starts for vertically linked pixels (skipped where the left pixel pair made the same union)
//   count : the last pixel of every run adds the run length to its root's size (few atomics, compresses the path)
//   apply : pixel -> run start -> root -> size test
__device__ __forceinline__ int uf_find(int *lab, int a)
{
    const volatile int *vl = lab; // other threads re-parent roots concurrently (atomicMin in uf_union)
    int p = vl[a];
    while (p != a) { a = p; p = vl[a]; }
    return a;
}
__device__ __forceinline__ void uf_union(int *lab, int a, int b)
{
    while (true) {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&lab[a], b); // a > b: hang the larger root under the smaller
        if (old == a) return;
        a = old;
    }
}
__device__ __forceinline__ bool linked(int a, int b, int newVal, int maxDiff) { return a != newVal && b != newVal && abs(a - b) <= maxDiff; }

constexpr int CCL_T = 256;
__global__ void __launch_bounds__(CCL_T) ccl_rows_kernel(const int16_t *__restrict__ img, int *__restrict__ lab, int *__restrict__ sizes, int W,
                                                         int newVal, int maxDiff)
{
    __shared__ int wmax[CCL_T / 32];
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int16_t *row = img + (size_t)y * W;
    const int ppt = (W + CCL_T - 1) / CCL_T, x0 = tid * ppt, x1 = min(x0 + ppt, W);
    // last run start inside this thread's segment
    int last = -1, prev = x0 > 0 && x0 < W ? row[x0 - 1] : newVal;
    for (int x = x0; x < x1; x++) {
        int v = row[x];
        if (!linked(v, prev, newVal, maxDiff)) last = x;
        prev = v;
    }
    // exclusive max-scan of `last` over the threads of the block = run start carried into the segment
    int inc = last;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    if (lane == 31) wmax[wid] = inc;
    __syncthreads();
    int carry = -1;
    for (int k = 0; k < wid; k++) carry = max(carry, wmax[k]);
    int exc = __shfl_up_sync(0xffffffffu, inc, 1);
    carry = lane == 0 ? carry : max(carry, exc);
    // labels
    int start = carry;
    prev = x0 > 0 && x0 < W ? row[x0 - 1] : newVal;
    const int base = y * W;
    for (int x = x0; x < x1; x++) {
        int v = row[x];
        if (!linked(v, prev, newVal, maxDiff)) start = x;
        lab[base + x] = v == newVal ? -1 : base + start;
        sizes[base + x] = 0;
        prev = v;
    }
}
__global__ void ccl_merge_kernel(const int16_t *__restrict__ img, int *lab, int H, int W, int newVal, int maxDiff)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y + 1;
    if (x >= W || y >= H) return;
    int i = y * W + x;
    int v = img[i], u = img[i - W];
    if (!linked(v, u, newVal, maxDiff)) return;
    if (x > 0) {
        int vl = img[i - 1], ul = img[i - W - 1];
        if (linked(v, vl, newVal, maxDiff) && linked(u, ul, newVal, maxDiff) && linked(vl, ul, newVal, maxDiff)) return; // implied
    }
    uf_union(lab, lab[i], lab[i - W]); // (a run start's label is its parent, which is in the same set)
}
__global__ void ccl_count_kernel(const int16_t *__restrict__ img, int *lab, int *sizes, int H, int W, int newVal, int maxDiff)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int i = y * W + x;
    int v = img[i];
    if (v == newVal) return;
    if (x + 1 < W && linked(img[i + 1], v, newVal, maxDiff)) return; // not the last pixel of its run
    const bool is_start = x == 0 || !linked(v, img[i - 1], newVal, maxDiff);
    const int rs = is_start ? i : lab[i];
    const int root = uf_find(lab, rs);
    atomicAdd(&sizes[root], i - rs + 1);
    if (root != rs) atomicMin(&lab[rs], root); // path compression for ccl_apply (only ever moves towards the root)
}
__global__ void ccl_apply_kernel(const int16_t *__restrict__ img, int *lab, const int *__restrict__ sizes, int16_t *__restrict__ out,
                                 int H, int W, int newVal, int maxDiff, int maxSize)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int i = y * W + x;
    int v = img[i];
    if (v != newVal) {
        const bool is_start = x == 0 || !linked(v, img[i - 1], newVal, maxDiff);
        const int rs = is_start ? i : lab[i];
        if (sizes[uf_find(lab, rs)] <= maxSize) v = newVal;
    }
    out[i] = (int16_t)v;
}

// ---- A.8: reference post-processing (stereo_matching.py:63-64, /16) ------------------------------------------
__global__ void disp_to_float_kernel(const int16_t *__restrict__ d16, float *__restrict__ out, size_t n, int minD16)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float f = fmaxf((float)d16[i], 0.f);
    if (f < (float)minD16) f = 0.f;
    out[i] = f / 16.0f;
}

} // namespace

cudaError_t launch_wta_prepare(b2s_ctx *c)
{
    const SgbmGeom &g = c->g;
    size_t n = (size_t)g.H * g.W;
    fill_i16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->raw.as<int16_t>(), n, (int16_t)g.invalid);
    c->launches++;
    return cudaMemsetAsync(c->disp2key.p, 0xFF, (size_t)g.H * (g.W + 2) * sizeof(unsigned), c->stream);
}

cudaError_t launch_wta(b2s_ctx *c)
{
    if (c->wta_fused) return cudaSuccess; // the last aggregation scan already selected the winners (sgbm_agg.cu)
    const SgbmGeom &g = c->g;
    dim3 blocks((g.width1 + 8 * WTA_CHUNK - 1) / (8 * WTA_CHUNK), g.H);
    int16_t *S = c->S.as<int16_t>();
    const int16_t *S2 = c->S2.as<int16_t>();
    int16_t *raw = c->raw.as<int16_t>();
    unsigned *keys = c->disp2key.as<unsigned>();
    const int keep = (c->keep_volumes || !c->fuse_wta) ? 1 : 0; // (B2S_OPT_FUSE_WTA = 0 promises a stored S, like the unfused scans)
#define B2S_WTA(NPV)                                                                                                                      \
    if (g.layout == 1 && c->wta_adds_s2) wta_kernel<NPV, true, 1><<<blocks, 256, 0, c->stream>>>(S, S2, raw, keys, g, keep);              \
    else if (g.layout == 1) wta_kernel<NPV, false, 1><<<blocks, 256, 0, c->stream>>>(S, S2, raw, keys, g, 0);                             \
    else if (c->wta_adds_s2) wta_kernel<NPV, true, 0><<<blocks, 256, 0, c->stream>>>(S, S2, raw, keys, g, keep);                          \
    else wta_kernel<NPV, false, 0><<<blocks, 256, 0, c->stream>>>(S, S2, raw, keys, g, 0);
    switch (g.NP) {
    case 1: B2S_WTA(1) break;
    case 2: B2S_WTA(2) break;
    case 3: B2S_WTA(3) break;
    case 4: B2S_WTA(4) break;
    case 5: B2S_WTA(5) break;
    case 6: B2S_WTA(6) break;
    case 7: B2S_WTA(7) break;
    case 8: B2S_WTA(8) break;
    default: return cudaErrorInvalidValue;
    }
#undef B2S_WTA
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_post(b2s_ctx *c, int16_t *d_out_disp16, float *d_out_disp)
{
    const SgbmGeom &g = c->g;
    size_t n = (size_t)g.H * g.W;
    dim3 b2(128), g2((g.W + 127) / 128, g.H);
    unsigned nb = (unsigned)((n + 255) / 256);
    int16_t *final16 = d_out_disp16 ? d_out_disp16 : c->disp16.as<int16_t>();
    int16_t *med = g.speckle_window > 0 ? c->med.as<int16_t>() : final16;
    lr_median_kernel<<<dim3((g.W + MED_TX - 1) / MED_TX, (g.H + MED_TY - 1) / MED_TY), dim3(MED_TX, MED_TY), 0, c->stream>>>(
        c->raw.as<int16_t>(), c->disp2key.as<unsigned>(), med, g);
    c->launches++;
    if (g.speckle_window > 0) {
        int *lab = c->labels.as<int>(), *sizes = c->sizes.as<int>();
        const int md = 16 * g.speckle_range;
        dim3 gm((g.W + 127) / 128, g.H > 1 ? g.H - 1 : 1);
        ccl_rows_kernel<<<g.H, CCL_T, 0, c->stream>>>(med, lab, sizes, g.W, g.invalid, md);
        if (g.H > 1) ccl_merge_kernel<<<gm, b2, 0, c->stream>>>(med, lab, g.H, g.W, g.invalid, md);
        ccl_count_kernel<<<g2, b2, 0, c->stream>>>(med, lab, sizes, g.H, g.W, g.invalid, md);
        ccl_apply_kernel<<<g2, b2, 0, c->stream>>>(med, lab, sizes, final16, g.H, g.W, g.invalid, md, g.speckle_window);
        c->launches += g.H > 1 ? 4 : 3;
    }
    if (d_out_disp) {
        disp_to_float_kernel<<<nb, 256, 0, c->stream>>>(final16, d_out_disp, n, g.minD * 16);
        c->launches++;
    }
    return cudaGetLastError();
}
