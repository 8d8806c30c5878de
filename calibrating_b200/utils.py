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




def interpolate_uvzs(uvzs, hw=None, constrained_type=None, inter_type="lstsq", distance=2, device=0):
    """calibrating/utils.py:356-411: densify sparse (u, v, z) samples to an (h, w) float32 image.  The sparse half (a 3-parameter
    least-squares fit on a few hundred points, the convex hull of the samples) stays on the host like the reference's; the dense
    half -- one value per pixel -- runs on the device.  inter_type "lstsq", "nearest" (float32 result) or "rbf" (thin plate, float64 result)."""
    import cv2
    uvzs = np.asarray(uvzs)
    if hw is None:
        hw = int(uvzs[:, 1].max()) + 2, int(uvzs[:, 0].max()) + 2
    hw = (int(hw[0]), int(hw[1]))
    if not uvzs.size:
        return np.zeros(hw, uvzs.dtype)
    mask = None
    if constrained_type is not None and constrained_type:
        mask = np.zeros(hw, np.uint8)
        hull = cv2.convexHull(np.int32(uvzs[:, :2].round()))
        cv2.drawContours(mask, [hull], -1, 1, -1)
    out = np.empty(hw, np.float32)
    h = _handle(device)
    if inter_type == "lstsq":
        A = np.float64(uvzs).copy()
        A[:, 2] = 1
        abc = np.linalg.lstsq(A, np.float64(uvzs[:, 2]), rcond=None)[0]
        h.call("b2s_interpolate_sparse", 0, None, 0, (ctypes.c_double * 3)(*abc), None if mask is None else _ffi.ptr(mask), hw[0], hw[1], 0.0, _ffi.ptr(out))
    elif inter_type == "nearest":
        pts = np.ascontiguousarray(uvzs[:, :3], np.float64)
        h.call("b2s_interpolate_sparse", 1, _ffi.ptr(pts), len(pts), None, None if mask is None else _ffi.ptr(mask), hw[0], hw[1], float(distance),
               _ffi.ptr(out))
    elif inter_type == "rbf":
        # scipy.interpolate.Rbf(u, v, z, function="thin_plate", smooth=0.5) (utils.py:373-387; the reference's `episilon=5` is a
        # misspelt keyword that scipy stores and ignores): the n x n solve on the host, the dense evaluation on the device, float64
        p = np.float64(uvzs[:, :2])
        r = np.sqrt(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1))
        with np.errstate(divide="ignore", invalid="ignore"):
            A = np.where(r > 0, r ** 2 * np.log(r), 0.0) - np.eye(len(p)) * 0.5
        uvw = np.ascontiguousarray(np.concatenate([p, np.linalg.solve(A, np.float64(uvzs[:, 2]))[:, None]], 1))
        z = np.empty(hw, np.float64)
        h.call("b2s_interpolate_rbf", _ffi.ptr(uvw), len(uvw), None if mask is None else _ffi.ptr(mask), hw[0], hw[1], _ffi.ptr(z))
        # the reference appends the result to the (u, v, 0) grid rows and hands all of it to uvzs_to_arr2d, which therefore returns TWO
        # channels (utils.py:383-387, 410): channel 0 is the grid's zero column, channel 1 the interpolated surface.  Kept as it is.
        out = np.stack([np.zeros(hw, np.float64), z], -1)
    else:
        raise NotImplementedError("inter_type %r: 'lstsq', 'nearest' and 'rbf'" % (inter_type,))
    return out


def interpolate_sparse2d(sparse2d, constrained_type=None, inter_type="lstsq", device=0):
    """calibrating/utils.py:347-353: the non-zero finite pixels of `sparse2d` as samples of interpolate_uvzs."""
    sparse2d = np.asarray(sparse2d)
    m = (sparse2d != 0) & np.isfinite(sparse2d)
    ys, xs = np.nonzero(m)  # row-major, the order of arr2d_to_uvzs (utils.py:318-328)
    uvzs = np.array([xs, ys, sparse2d[m]]).T
    return interpolate_uvzs(uvzs, sparse2d.shape[:2], constrained_type, inter_type, device=device)


__all__ = ["depth_to_point_cloud", "point_cloud_to_depth", "point_cloud_to_arr2d", "interpolate_uvzs", "interpolate_sparse2d"]
