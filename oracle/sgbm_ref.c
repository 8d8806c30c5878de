/*
 * oracle/sgbm_ref.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * Plain-C restatement of what the reference's default matcher executes:
 *   calibrating/stereo_matching.py:48-63  ->  cv2.StereoSGBM_create(...).compute(img1, img2)
 * The arithmetic lives in OpenCV (opencv-contrib-python>=4.7.0.72, requirements.txt:2; installed here:
 * opencv-python-headless 4.13.0.92), modules/calib3d/src/stereosgbm.cpp (computeDisparitySGBM, modes
 * MODE_SGBM=0 / MODE_HH=1), followed by medianBlur(3) and filterSpeckles inside StereoSGBM::compute.
 * The source is not vendored under /root/reference; this file restates the published algorithm as
 * specified in SURVEY.md Appendix A and is pinned by differential tests against the installed cv2
 * (tests/test_oracle.py) and by the committed golden vectors in tests/golden/.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libsgbm_ref.so oracle/sgbm_ref.c   (oracle/build.py)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int16_t cost_t;

typedef struct {
    int min_disparity, num_disparities, block_size;
    int P1, P2, disp12_max_diff, pre_filter_cap, uniqueness_ratio;
    int speckle_window_size, speckle_range, mode; /* 0 = MODE_SGBM (5 paths), 1 = MODE_HH (8 paths), 3 = MODE_HH4 (4 paths) */
    int cost; /* 0 = Birchfield-Tomasi + box sum (cv2); 1 = 9x7 census / Hamming (BASELINE config 4; NOT in cv2: parity unpinned) */
} sgbm_params;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline cost_t sat16(int v) { return (cost_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }

/* A.2: per-row BT planes. planes: [2*cn][W] value, lo (min of half-sample interval), hi (max). */
static void build_planes(const uint8_t *img, int H, int W, int cn, int y, int ftzero,
                         uint8_t *val, uint8_t *lo, uint8_t *hi)
{
    int ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
    const uint8_t *r0 = img + (size_t)y * W * cn, *rm = img + (size_t)ym * W * cn, *rp = img + (size_t)yp * W * cn;
    for (int c = 0; c < cn; c++) {
        uint8_t *g = val + (size_t)c * W, *r = val + (size_t)(c + cn) * W;
        g[0] = g[W - 1] = (uint8_t)ftzero;
        r[0] = r[W - 1] = (uint8_t)ftzero;
        for (int x = 1; x < W - 1; x++) {
            int s = 2 * ((int)r0[(x + 1) * cn + c] - r0[(x - 1) * cn + c])
                  + ((int)rm[(x + 1) * cn + c] - rm[(x - 1) * cn + c])
                  + ((int)rp[(x + 1) * cn + c] - rp[(x - 1) * cn + c]);
            g[x] = (uint8_t)(clampi(s, -ftzero, ftzero) + ftzero);
            r[x] = r0[x * cn + c];
        }
    }
    for (int p = 0; p < 2 * cn; p++) {
        const uint8_t *v = val + (size_t)p * W;
        for (int x = 0; x < W; x++) {
            int u = v[x];
            int ul = x > 0 ? (u + v[x - 1]) / 2 : u;
            int ur = x < W - 1 ? (u + v[x + 1]) / 2 : u;
            lo[(size_t)p * W + x] = (uint8_t)imin(imin(ul, ur), u);
            hi[(size_t)p * W + x] = (uint8_t)imax(imax(ul, ur), u);
        }
    }
}

/* Census cost (cost = 1), this engine's own definition (there is no reference implementation: parity unpinned):
 * descriptor of a gray pixel = 62 bits, one per neighbour of the 9 (wide) x 7 (tall) window except the centre, row-major,
 * bit = neighbour < centre, coordinates clamped to the image; C(x, d) = popcount(descL(x) ^ descR(x - d)), no box sum. */
static uint64_t census_desc(const uint8_t *img, int H, int W, int y, int x)
{
    uint64_t bits = 0;
    int c = img[(size_t)y * W + x];
    for (int dy = -3; dy <= 3; dy++)
        for (int dx = -4; dx <= 4; dx++) {
            if (dy == 0 && dx == 0) continue;
            int v = img[(size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)];
            bits = (bits << 1) | (uint64_t)(v < c);
        }
    return bits;
}

/* A.4 one step of the path recurrence.  Lp/Ln are D-long; minLp is min over Lp. returns min over Ln. */
static inline int path_step(const cost_t *Cp, const cost_t *Lp, int minLp, cost_t *Ln, int D, int P1, int P2)
{
    int minL = 32767;
    for (int d = 0; d < D; d++) {
        int lm = d > 0 ? Lp[d - 1] : 32767;
        int lp = d < D - 1 ? Lp[d + 1] : 32767;
        int L = Cp[d] + imin((int)Lp[d], imin(lm + P1, imin(lp + P1, minLp + P2))) - minLp;
        Ln[d] = (cost_t)L;
        if ((cost_t)L < minL) minL = (cost_t)L;
    }
    return minL;
}

/* 3x3 median, replicated border (cv2.medianBlur(int16, 3)) */
static void median3(const int16_t *src, int16_t *dst, int H, int W)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int16_t v[9]; int n = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    v[n++] = src[(size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)];
            for (int i = 1; i < 9; i++) { int16_t k = v[i]; int j = i - 1; while (j >= 0 && v[j] > k) { v[j + 1] = v[j]; j--; } v[j + 1] = k; }
            dst[(size_t)y * W + x] = v[4];
        }
}

/* cv2.filterSpeckles: 4-connected components under |dv| <= maxDiff, pixels == newVal skipped,
 * components of size <= maxSpeckleSize set to newVal. */
static void filter_speckles(int16_t *img, int H, int W, int newVal, int maxSpeckleSize, int maxDiff)
{
    size_t N = (size_t)H * W;
    int32_t *labels = (int32_t *)calloc(N, sizeof(int32_t));
    int32_t *stack = (int32_t *)malloc(N * sizeof(int32_t));
    uint8_t *small = (uint8_t *)calloc(N + 1, 1);
    int cur = 0;
    for (size_t p0 = 0; p0 < N; p0++) {
        if (img[p0] == newVal) continue;
        if (labels[p0]) { if (small[labels[p0]]) img[p0] = (int16_t)newVal; continue; }
        int sp = 0, count = 0; cur++;
        stack[sp++] = (int32_t)p0; labels[p0] = cur;
        while (sp) {
            int32_t p = stack[--sp]; count++;
            int y = p / W, x = p % W; int v = img[p];
            int nb[4]; int k = 0;
            if (y < H - 1) nb[k++] = p + W;
            if (y > 0) nb[k++] = p - W;
            if (x < W - 1) nb[k++] = p + 1;
            if (x > 0) nb[k++] = p - 1;
            for (int i = 0; i < k; i++) {
                int q = nb[i];
                if (!labels[q] && img[q] != newVal && abs(v - img[q]) <= maxDiff) { labels[q] = cur; stack[sp++] = q; }
            }
        }
        if (count <= maxSpeckleSize) { small[cur] = 1; img[p0] = (int16_t)newVal; }
    }
    free(labels); free(stack); free(small);
}

/* exported helpers so the stages can be checked one at a time */
void oracle_median3(const int16_t *src, int16_t *dst, int H, int W) { median3(src, dst, H, W); }
void oracle_filter_speckles(int16_t *img, int H, int W, int newVal, int maxSpeckleSize, int maxDiff)
{ filter_speckles(img, H, W, newVal, maxSpeckleSize, maxDiff); }

/*
 * Full StereoSGBM::compute.  left/right: (H,W,cn) uint8.  disp: (H,W) int16 = 16*disparity.
 * Optional outputs (NULL to skip): C_out, S_out: (H,width1,D) int16;  raw_out: (H,W) int16 disparity before
 * median/speckle.  Returns 0, or -1 for the cv2 precondition failure (W - maxD > SW2 violated), -2 for
 * unsupported parameters.  Negative min_disparity (minX1 = max(maxD, 0), maxX1 = W + min(minD, 0), A.1) is covered and pinned
 * against cv2 in tests/test_oracle.py.
 */
int oracle_sgbm_compute(const uint8_t *left, const uint8_t *right, int H, int W, int cn,
                        const sgbm_params *prm, int16_t *disp, int16_t *C_out, int16_t *S_out, int16_t *raw_out)
{
    /* A.1 parameter normalisation */
    int minD = prm->min_disparity, maxD = minD + prm->num_disparities;
    int bs = prm->block_size > 0 ? prm->block_size : 5;
    int SW2 = bs / 2, SH2 = bs / 2;
    int ftzero = imax(prm->pre_filter_cap, 15) | 1;
    int uniq = prm->uniqueness_ratio >= 0 ? prm->uniqueness_ratio : 10;
    int d12 = prm->disp12_max_diff > 0 ? prm->disp12_max_diff : 1;
    int P1 = prm->P1 > 0 ? prm->P1 : 2;
    int P2 = imax(prm->P2 > 0 ? prm->P2 : 5, P1 + 1);
    if (prm->mode != 0 && prm->mode != 1 && prm->mode != 3) return -2; /* MODE_SGBM_3WAY depends on cv2's thread count: not restated */
    int minX1 = imax(maxD, 0), maxX1 = W + imin(minD, 0);
    int D = maxD - minD, width1 = maxX1 - minX1;
    const int DISP_SHIFT = 4, DISP_SCALE = 16;
    int INVALID = (minD - 1) * DISP_SCALE;
    if (W - maxD <= SW2 || D <= 0) return -1;

    const int hh4 = prm->mode == 3; /* MODE_HH4: two passes, only the horizontal and the vertical path of each */
    size_t row = (size_t)width1 * D, vol = row * H;
    cost_t *Cv = (cost_t *)malloc(vol * sizeof(cost_t));
    cost_t *Sv = (cost_t *)calloc(vol, sizeof(cost_t));
    if (prm->cost == 1) {
        if (cn != 1) { free(Cv); free(Sv); return -2; }
        uint64_t *cl = (uint64_t *)malloc((size_t)W * 8), *cr = (uint64_t *)malloc((size_t)W * 8);
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < W; x++) { cl[x] = census_desc(left, H, W, y, x); cr[x] = census_desc(right, H, W, y, x); }
            for (int x = minX1; x < maxX1; x++)
                for (int d = minD; d < maxD; d++)
                    Cv[(size_t)y * row + (size_t)(x - minX1) * D + (d - minD)] = (cost_t)__builtin_popcountll(cl[x] ^ cr[x - d]);
        }
        free(cl); free(cr);
    } else if (prm->cost != 0) { free(Cv); free(Sv); return -2; }
    else {
    cost_t *hs = (cost_t *)malloc(vol * sizeof(cost_t));
    int np = 2 * cn;
    uint8_t *lv = (uint8_t *)malloc((size_t)np * W * 3), *rv = (uint8_t *)malloc((size_t)np * W * 3);
    uint8_t *llo = lv + (size_t)np * W, *lhi = llo + (size_t)np * W;
    uint8_t *rlo = rv + (size_t)np * W, *rhi = rlo + (size_t)np * W;
    cost_t *pix = (cost_t *)malloc(row * sizeof(cost_t));

    /* A.2 + horizontal half of A.3 */
    for (int y = 0; y < H; y++) {
        build_planes(left, H, W, cn, y, ftzero, lv, llo, lhi);
        build_planes(right, H, W, cn, y, ftzero, rv, rlo, rhi);
        memset(pix, 0, row * sizeof(cost_t));
        for (int p = 0; p < np; p++) {
            int sh = p < cn ? 0 : 2;
            for (int x = minX1; x < maxX1; x++) {
                int u = lv[(size_t)p * W + x], u0 = llo[(size_t)p * W + x], u1 = lhi[(size_t)p * W + x];
                cost_t *pc = pix + (size_t)(x - minX1) * D;
                for (int d = minD; d < maxD; d++) {
                    int xr = x - d;
                    int v = rv[(size_t)p * W + xr], v0 = rlo[(size_t)p * W + xr], v1 = rhi[(size_t)p * W + xr];
                    int c0 = imax(0, imax(u - v1, v0 - u));
                    int c1 = imax(0, imax(v - u1, u0 - v));
                    pc[d - minD] = (cost_t)(pc[d - minD] + (imin(c0, c1) >> sh));
                }
            }
        }
        cost_t *h = hs + (size_t)y * row;
        for (int x1 = 0; x1 < width1; x1++)
            for (int d = 0; d < D; d++) {
                int s = 0;
                for (int k = -SW2; k <= SW2; k++) s += pix[(size_t)clampi(x1 + k, 0, width1 - 1) * D + d];
                h[(size_t)x1 * D + d] = (cost_t)s;
            }
    }
    /* vertical half of A.3.  MODE_HH4 quirk (cv2 4.13, found by differential testing, tests/test_oracle.py): the rows whose
     * window reaches below the image (y > 0 and y + SH2 >= H) keep a constant cost -- cv2's column-parallel cost loop skips
     * the "k >= height" update -- which is the same as C = 0 for them (L and S of such a row only see the path terms). */
    for (int y = 0; y < H; y++) {
        cost_t *c = Cv + (size_t)y * row;
        if (hh4 && y > 0 && y + SH2 >= H) { memset(c, 0, row * sizeof(cost_t)); continue; }
        for (size_t i = 0; i < row; i++) {
            int s = 0;
            for (int k = -SH2; k <= SH2; k++) s += hs[(size_t)clampi(y + k, 0, H - 1) * row + i];
            c[i] = (cost_t)s;
        }
    }
    free(hs); free(pix); free(lv); free(rv);
    }

    /* A.4 aggregation */
    int16_t *raw = (int16_t *)malloc((size_t)H * W * sizeof(int16_t));
    for (size_t i = 0; i < (size_t)H * W; i++) raw[i] = (int16_t)INVALID;
    int npass = (prm->mode == 1 || prm->mode == 3) ? 2 : 1;
    /* Lr rows: [2 rows][4 dirs][width1+2][D], minLr likewise; x index shifted by +1 so that x=-1 and x=width1 are zero */
    size_t lrow = (size_t)(width1 + 2) * D;
    cost_t *Lr = (cost_t *)calloc(2 * 4 * lrow, sizeof(cost_t));
    cost_t *mLr = (cost_t *)calloc(2 * 4 * (size_t)(width1 + 2), sizeof(cost_t));
    int16_t *disp2 = (int16_t *)malloc((size_t)(W + 2) * sizeof(int16_t));
    cost_t *disp2cost = (cost_t *)malloc((size_t)(W + 2) * sizeof(cost_t));
    cost_t *L4prev = (cost_t *)malloc((size_t)D * sizeof(cost_t)), *L4cur = (cost_t *)malloc((size_t)D * sizeof(cost_t));
    cost_t *Srow = (cost_t *)malloc((size_t)D * sizeof(cost_t));

    for (int pass = 1; pass <= npass; pass++) {
        int y1 = pass == 1 ? 0 : H - 1, y2 = pass == 1 ? H : -1, dy = pass == 1 ? 1 : -1;
        int x1s = pass == 1 ? 0 : width1 - 1, x2s = pass == 1 ? width1 : -1, dx = pass == 1 ? 1 : -1;
        memset(Lr, 0, 2 * 4 * lrow * sizeof(cost_t));
        memset(mLr, 0, 2 * 4 * (size_t)(width1 + 2) * sizeof(cost_t));
        for (int y = y1; y != y2; y += dy) {
            int cur = (y & 1), prv = cur ^ 1;
            cost_t *Lc = Lr + (size_t)cur * 4 * lrow, *Lp = Lr + (size_t)prv * 4 * lrow;
            cost_t *mc = mLr + (size_t)cur * 4 * (width1 + 2), *mp = mLr + (size_t)prv * 4 * (width1 + 2);
            /* clear the border slots of the current row (predecessors outside the image are zero) */
            for (int r = 0; r < 4; r++) {
                memset(Lc + r * lrow, 0, D * sizeof(cost_t)); memset(Lc + r * lrow + (size_t)(width1 + 1) * D, 0, D * sizeof(cost_t));
                mc[r * (width1 + 2)] = 0; mc[r * (width1 + 2) + width1 + 1] = 0;
            }
            const cost_t *Crow = Cv + (size_t)y * row;
            cost_t *Sr = Sv + (size_t)y * row;
            for (int x = x1s; x != x2s; x += dx) {
                int xi = x + 1;
                const cost_t *Cp = Crow + (size_t)x * D;
                cost_t *Sp = Sr + (size_t)x * D;
                /* dir 0: (x-dx, y), dir 1: (x-dx, y-dy), dir 2: (x, y-dy), dir 3: (x+dx, y-dy) */
                cost_t *o0 = Lc + 0 * lrow + (size_t)xi * D, *o1 = Lc + 1 * lrow + (size_t)xi * D;
                cost_t *o2 = Lc + 2 * lrow + (size_t)xi * D, *o3 = Lc + 3 * lrow + (size_t)xi * D;
                mc[0 * (width1 + 2) + xi] = (cost_t)path_step(Cp, Lc + 0 * lrow + (size_t)(xi - dx) * D, mc[0 * (width1 + 2) + xi - dx], o0, D, P1, P2);
                if (hh4) {
                    mc[2 * (width1 + 2) + xi] = (cost_t)path_step(Cp, Lp + 2 * lrow + (size_t)xi * D, mp[2 * (width1 + 2) + xi], o2, D, P1, P2);
                    for (int d = 0; d < D; d++) Sp[d] = sat16((int)Sp[d] + o0[d] + o2[d]);
                    continue;
                }
                mc[1 * (width1 + 2) + xi] = (cost_t)path_step(Cp, Lp + 1 * lrow + (size_t)(xi - dx) * D, mp[1 * (width1 + 2) + xi - dx], o1, D, P1, P2);
                mc[2 * (width1 + 2) + xi] = (cost_t)path_step(Cp, Lp + 2 * lrow + (size_t)xi * D, mp[2 * (width1 + 2) + xi], o2, D, P1, P2);
                mc[3 * (width1 + 2) + xi] = (cost_t)path_step(Cp, Lp + 3 * lrow + (size_t)(xi + dx) * D, mp[3 * (width1 + 2) + xi + dx], o3, D, P1, P2);
                for (int d = 0; d < D; d++)
                    Sp[d] = sat16((int)Sp[d] + o0[d] + o1[d] + o2[d] + o3[d]);
            }
            if (pass == npass) {
                /* A.5 WTA for this row, x1 descending */
                int16_t *d1 = raw + (size_t)y * W;
                for (int x = 0; x < W + 2; x++) { disp2[x] = (int16_t)INVALID; disp2cost[x] = 32767; }
                int minL4 = 0; memset(L4prev, 0, D * sizeof(cost_t));
                for (int x = width1 - 1; x >= 0; x--) {
                    const cost_t *Sp = Sr + (size_t)x * D;
                    if (npass == 1) {
                        /* MODE_SGBM: fifth direction, predecessor (x+1, y) */
                        minL4 = path_step(Crow + (size_t)x * D, L4prev, minL4, L4cur, D, P1, P2);
                        for (int d = 0; d < D; d++) Srow[d] = sat16((int)Sp[d] + L4cur[d]);
                        memcpy(L4prev, L4cur, D * sizeof(cost_t));
                        memcpy(Sr + (size_t)x * D, Srow, D * sizeof(cost_t));
                    }
                    int minS = 32767, best = -1;
                    for (int d = 0; d < D; d++) if (Sp[d] < minS) { minS = Sp[d]; best = d; }
                    int d;
                    for (d = 0; d < D; d++)
                        if (Sp[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
                    if (d < D) continue;
                    d = best;
                    int x2 = x + minX1 - d - minD;
                    if (x2 >= 0 && x2 < W + 2 && disp2cost[x2] > minS) { disp2cost[x2] = (cost_t)minS; disp2[x2] = (int16_t)(d + minD); }
                    if (0 < d && d < D - 1) {
                        int den2 = imax(Sp[d - 1] + Sp[d + 1] - 2 * Sp[d], 1);
                        d = d * DISP_SCALE + ((Sp[d - 1] - Sp[d + 1]) * DISP_SCALE + den2) / (den2 * 2);
                    } else d *= DISP_SCALE;
                    d1[x + minX1] = (int16_t)(d + minD * DISP_SCALE);
                }
                /* A.6 left-right check */
                for (int x = minX1; x < maxX1; x++) {
                    int dd = d1[x];
                    if (dd == INVALID) continue;
                    int _d = dd >> DISP_SHIFT, d_ = (dd + DISP_SCALE - 1) >> DISP_SHIFT;
                    int _x = x - _d, x_ = x - d_;
                    if (0 <= _x && _x < W && disp2[_x] >= minD && abs(disp2[_x] - _d) > d12 &&
                        0 <= x_ && x_ < W && disp2[x_] >= minD && abs(disp2[x_] - d_) > d12)
                        d1[x] = (int16_t)INVALID;
                }
            }
        }
    }
    if (C_out) memcpy(C_out, Cv, vol * sizeof(cost_t));
    if (S_out) memcpy(S_out, Sv, vol * sizeof(cost_t));
    if (raw_out) memcpy(raw_out, raw, (size_t)H * W * sizeof(int16_t));
    /* A.7 post filters */
    median3(raw, disp, H, W);
    if (prm->speckle_window_size > 0)
        filter_speckles(disp, H, W, INVALID, prm->speckle_window_size, DISP_SCALE * prm->speckle_range);
    free(Cv); free(Sv); free(raw); free(Lr); free(mLr); free(disp2); free(disp2cost); free(L4prev); free(L4cur); free(Srow);
    return 0;
}
