"""NumPy restatement of the cv2.remap / cv2.undistort variants on the hot path -- TEST INFRASTRUCTURE ONLY.

Reference call sites: calibrating/stereo_camera.py:217-228 (remap, INTER_LANCZOS4, u8c3, f32 map pair,
default BORDER_CONSTANT 0), :431 (cv2.undistort = bilinear remap), calibrating/utils.py:199 (remap,
INTER_NEAREST on float64).  The arithmetic is OpenCV's imgproc (not vendored): coordinates quantised to
1/32 px (INTER_BITS=5), 15-bit fixed-point weight tables (INTER_REMAP_COEF_BITS=15) for u8 sources.
Pinned by tests/test_oracle_vs_cv2.py against the installed cv2.
"""
import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS


def _lanczos4_coeffs(x):
    """cv::interpolateLanczos4 (float x in [0,1)) -> 8 float32 weights."""
    x = np.float32(x)
    co = np.zeros(8, np.float32)
    if x < np.finfo(np.float32).eps:
        co[3] = 1
        return co
    s45 = 0.70710678118654752440084436210485
    cs = [(1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45)]
    y0 = -(float(x) + 3) * np.pi * 0.25
    s0, c0 = np.sin(y0), np.cos(y0)
    total = np.float32(0)
    for i in range(8):
        y = -(float(x) + 3 - i) * np.pi * 0.25
        co[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        total = np.float32(total + co[i])
    inv = np.float32(np.float32(1) / total)
    return (co * inv).astype(np.float32)


def _tab1d(kind):
    scale = np.float32(1.0 / INTER_TAB_SIZE)
    if kind == "lanczos4":
        return np.stack([_lanczos4_coeffs(np.float32(i) * scale) for i in range(INTER_TAB_SIZE)])
    if kind == "linear":
        x = (np.arange(INTER_TAB_SIZE, dtype=np.float32) * scale).astype(np.float32)
        return np.stack([np.float32(1) - x, x], 1).astype(np.float32)
    raise ValueError(kind)


_TABS = {}


def fixed_point_table(kind):
    """cv::initInterTab2D(method, fixpt=true): int16 table [32*32][k][k], every entry sums to 32768."""
    if kind in _TABS:
        return _TABS[kind]
    t1 = _tab1d(kind)
    k = t1.shape[1]
    tab = np.zeros((INTER_TAB_SIZE * INTER_TAB_SIZE, k, k), np.int16)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = (t1[i][:, None] * t1[j][None, :]).astype(np.float32)
            it = np.clip(np.rint(v * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int32)
            isum = int(it.sum())
            if isum != COEF_SCALE and k > 2:  # bilinear products are exact multiples of 1/1024: never needs it
                diff = isum - COEF_SCALE
                k2 = k // 2
                Mk = mk = (k2, k2)
                for a in range(k2, k2 + 2):
                    for b in range(k2, k2 + 2):
                        if it[a, b] < it[mk]:
                            mk = (a, b)
                        elif it[a, b] > it[Mk]:
                            Mk = (a, b)
                if diff < 0:
                    it[Mk] -= diff
                else:
                    it[mk] -= diff
            tab[i * INTER_TAB_SIZE + j] = it.astype(np.int16)
    _TABS[kind] = tab
    return tab


def quantise_maps(mapx, mapy):
    """float32 maps -> (ix, iy, fxy) exactly as cv::remap's CV_32FC1 pair conversion (cvRound(v*32))."""
    sx = np.rint(mapx.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    sy = np.rint(mapy.astype(np.float32) * np.float32(INTER_TAB_SIZE)).astype(np.int64)
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    fxy = (sy & (INTER_TAB_SIZE - 1)) * INTER_TAB_SIZE + (sx & (INTER_TAB_SIZE - 1))
    return ix, iy, fxy


def _remap_fixed(src, ix, iy, fxy, kind):
    tab = fixed_point_table(kind).astype(np.int64)
    k = tab.shape[1]
    off = k // 2 - 1
    H, W = src.shape[:2]
    s = src.reshape(H, W, -1).astype(np.int64)
    acc = np.zeros(ix.shape + (s.shape[2],), np.int64)
    w = tab[fxy]
    for a in range(k):
        yy = iy - off + a
        oky = (yy >= 0) & (yy < H)
        for b in range(k):
            xx = ix - off + b
            ok = oky & (xx >= 0) & (xx < W)
            v = s[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok[..., None]
            acc += v * w[..., a, b][..., None]
    out = np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)
    return out.reshape(ix.shape + src.shape[2:])


def remap_lanczos4_u8(src, mapx, mapy):
    return _remap_fixed(src, *quantise_maps(mapx, mapy), "lanczos4")


def remap_linear_u8(src, mapx, mapy):
    return _remap_fixed(src, *quantise_maps(mapx, mapy), "linear")


def remap_linear_u8_fixed(src, ix, iy, fxy):
    """Bilinear remap from already-quantised maps (what cv2.undistort builds internally, CV_16SC2)."""
    return _remap_fixed(src, ix.astype(np.int64), iy.astype(np.int64), fxy.astype(np.int64), "linear")


def remap_nearest(src, mapx, mapy):
    """INTER_NEAREST with f32 maps: index = round-half-even, outside -> 0."""
    H, W = src.shape[:2]
    ix = np.rint(mapx.astype(np.float32)).astype(np.int64)
    iy = np.rint(mapy.astype(np.float32)).astype(np.int64)
    ok = (ix >= 0) & (ix < W) & (iy >= 0) & (iy < H)
    out = src[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)]
    return np.where(ok.reshape(ok.shape + (1,) * (src.ndim - 2)), out, 0).astype(src.dtype)
