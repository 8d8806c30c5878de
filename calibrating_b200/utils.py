"""Device versions of the reference's depth <-> point-cloud helpers (calibrating/utils.py:213-317), SURVEY.md section 8(f) rank 3.

Same names, arguments and result conventions as the reference; the arithmetic runs in libb2s.so (csrc/cloud.cu) and there is
no CPU fallback.  `Cam.project_cam2_depth` (stereo_camera.py) is the fused form of depth_to_point_cloud -> rigid transform ->
point_cloud_to_depth."""
import ctypes

import numpy as np

from . import _ffi


def _handle(device):
    from .stereo_camera import _module_handle
    return _module_handle(device)


def _mat9(m):
    return (ctypes.c_double * 9)(*np.float64(m).reshape(9))


def depth_to_point_cloud(depth, K, interpolation_rate=1, return_xyzuv=False, device=0):
    """calibrating/utils.py:213-250.  depth: (h,w) uint16 (millimetres -> float32 metres, like the reference), float32 or float64;
    returns the (n,3) float64 point cloud of the non-zero pixels in row-major order, or (n,5) xyzuv with return_xyzuv."""
    depth = np.asarray(depth)
    assert depth.ndim == 2
    if depth.dtype == np.uint16:
        depth = np.float32(depth / 1000.0)
    d64 = np.ascontiguousarray(depth, np.float64)  # (float32 -> float64 is exact: the products below are float64 in the reference too)
    h, w = d64.shape
    rate = float(interpolation_rate)
    hu, wu = (h, w) if rate == 1 else (int(round(h * rate)), int(round(w * rate)))
    cols = 5 if return_xyzuv else 3
    cap = max(int(np.count_nonzero(d64)) if rate == 1 else hu * wu, 1)
    out = np.empty((cap, cols), np.float64)
    n = ctypes.c_ulonglong(0)
    _handle(device).call("b2s_depth_to_point_cloud", _ffi.ptr(d64), h, w, rate, _mat9(np.linalg.inv(np.float64(K))), int(bool(return_xyzuv)),
                         _ffi.ptr(out), ctypes.c_ulonglong(cap), ctypes.byref(n))
    return out[:n.value]


def point_cloud_to_arr2d(points, K, xy, values=None, bg_value=0, device=0):
    """calibrating/utils.py:258-288 for the depth case (values=None): z-buffered projection of (n,3) points to a (h,w) image."""
    if values is not None:
        raise NotImplementedError("per-point `values` (utils.py:284-288) are not offered on the device; only the depth image (values=None)")
    pts = np.ascontiguousarray(points, np.float64)
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("points must be (n, 3)")
    w, h = int(xy[0]), int(xy[1])
    out = np.empty((h, w), np.float64)
    _handle(device).call("b2s_point_cloud_to_depth", _ffi.ptr(pts) if len(pts) else None, ctypes.c_ulonglong(len(pts)), _mat9(K), w, h, float(bg_value),
                         _ffi.ptr(out))
    return out


def point_cloud_to_depth(points, K, xy, device=0):
    """calibrating/utils.py:254-255."""
    return point_cloud_to_arr2d(points, K, xy, device=device)


__all__ = ["depth_to_point_cloud", "point_cloud_to_depth", "point_cloud_to_arr2d"]
