"""CPU restatement of the resize calls around the matcher -- TEST INFRASTRUCTURE ONLY.

`SemiGlobalBlockMatching.__call__` (calibrating/stereo_matching.py:61-62, 65-69) down-scales the rectified pair and up-scales
the disparity with `boxx.resize`, an un-vendored dependency (boxx>=0.10.6, requirements.txt:1) whose interpolation is
UNPINNED.  The stand-in pinned here is cv2.resize(INTER_LINEAR) of the installed OpenCV 4.13:
`resize_u8` is bit-exact with it (tests/test_oracle.py), `resize_f32` agrees with its IPP float path to 2 ulp."""
import numpy as np


def scaled_size(h, w, max_size):
    """(nh, nw) of `boxx.resize(img, min(max_size / max(h, w), 1))`: int(round(side * ratio)), Python's round-half-even."""
    ratio = min(max_size / max(h, w), 1)
    return int(round(h * ratio)), int(round(w * ratio))


def _coef(src, dst, dtype):
    scale = np.float64(src) / dst
    f = ((np.arange(dst) + 0.5) * scale - 0.5).astype(dtype)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(dtype)
    lo = s < 0
    f[lo], s[lo] = 0, 0
    hi = s >= src - 1
    f[hi], s[hi] = 0, src - 1
    return s, f


def resize_u8(img, nh, nw):
    """cv2.resize(uint8, (nw, nh), interpolation=cv2.INTER_LINEAR): 11-bit fixed-point coefficients (OpenCV resize.cpp:
    HResizeLinear / VResizeLinear<uchar>)."""
    h, w = img.shape[:2]
    sx, fx = _coef(w, nw, np.float32)
    sy, fy = _coef(h, nh, np.float32)
    q = lambda f: np.rint(f * np.float32(2048)).astype(np.int32)  # saturate_cast<short>
    ax1, ax0, ay1, ay0 = q(fx), q(np.float32(1) - fx), q(fy), q(np.float32(1) - fy)
    sx1, sy1 = np.minimum(sx + 1, w - 1), np.minimum(sy + 1, h - 1)
    im = img.astype(np.int32)
    tail = (1,) * (img.ndim - 2)
    S = im[:, sx] * ax0.reshape(1, -1, *tail) + im[:, sx1] * ax1.reshape(1, -1, *tail)
    b0, b1 = ay0.reshape(-1, 1, *tail), ay1.reshape(-1, 1, *tail)
    return ((((b0 * (S[sy] >> 4)) >> 16) + ((b1 * (S[sy1] >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def resize_f32(img, nh, nw):
    """Bilinear with float64 coefficients and accumulation, rounded to float32 once."""
    h, w = img.shape[:2]
    sx, fx = _coef(w, nw, np.float64)
    sy, fy = _coef(h, nh, np.float64)
    sx1, sy1 = np.minimum(sx + 1, w - 1), np.minimum(sy + 1, h - 1)
    im = img.astype(np.float64)
    S = im[:, sx] * (1 - fx)[None] + im[:, sx1] * fx[None]
    return (S[sy] * (1 - fy)[:, None] + S[sy1] * fy[:, None]).astype(np.float32)


def scaled_matcher(compute_float, img1, img2, max_size):
    """stereo_matching.py:60-70 with the resize stand-in: compute_float(simg1, simg2) -> float32 disparity in pixels."""
    h, w = img1.shape[:2]
    nh, nw = scaled_size(h, w, max_size)
    if (nh, nw) == (h, w):
        return compute_float(img1, img2)
    sd = compute_float(resize_u8(img1, nh, nw), resize_u8(img2, nh, nw))
    return resize_f32(sd, h, w) * np.float32(w) / np.float32(nw)
