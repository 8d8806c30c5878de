"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's sparse -> dense interpolation,
calibrating/utils.py:347-411 (`interpolate_sparse2d`, `interpolate_uvzs`; inter_type "lstsq", "nearest" and "rbf").

Pinned against outputs of the real reference (tests/golden/sparse_small.npz, written by tests/golden/make_golden.py).
"nearest" is a brute-force search instead of scipy's KDTree: equal to it wherever the nearest sample is unique (of equally near
samples the lowest index is taken here; KDTree's pick depends on its tree).
"""
import numpy as np


def hull_mask(uvzs, hw):
    """utils.py:366-369: filled convex hull of the rounded sample positions (cv2.convexHull + cv2.drawContours)."""
    import cv2
    mask = np.zeros(hw, np.uint8)
    cv2.drawContours(mask, [cv2.convexHull(np.int32(uvzs[:, :2].round()))], -1, 1, -1)
    return mask


def interpolate_uvzs(uvzs, hw=None, constrained_type=None, inter_type="lstsq", distance=2):
    uvzs = np.asarray(uvzs)
    if hw is None:  # utils.py:361-362
        hw = int(uvzs[:, 1].max()) + 2, int(uvzs[:, 0].max()) + 2
    hw = (int(hw[0]), int(hw[1]))
    if not uvzs.size:  # utils.py:363-364
        return np.zeros(hw, uvzs.dtype)
    mask = hull_mask(uvzs, hw).astype(bool) if constrained_type else np.ones(hw, bool)
    ys, xs = np.mgrid[:hw[0], :hw[1]]
    if inter_type == "lstsq":  # utils.py:388-394: z = a*u + b*v + c, evaluated in float64, stored as float32
        A = uvzs.copy()
        A[:, 2] = 1
        abc = np.linalg.lstsq(A, uvzs[:, 2], rcond=None)[0]
        val = np.float32(np.float64(xs) * abc[0] + np.float64(ys) * abc[1] + abc[2])
    elif inter_type == "nearest":  # utils.py:395-405
        val = np.zeros(hw, np.float32)
        u, v = np.float64(uvzs[:, 0]), np.float64(uvzs[:, 1])
        for y in range(hw[0]):
            d2 = (u[None, :] - np.float64(xs[y])[:, None]) ** 2 + (v[None, :] - float(y)) ** 2
            k = d2.argmin(1)
            near = np.sqrt(d2[np.arange(hw[1]), k]) < distance
            val[y, near] = np.float32(uvzs[k[near], 2])
    elif inter_type == "rbf":  # utils.py:373-387: scipy.interpolate.Rbf(u, v, z, function="thin_plate", smooth=0.5), restated with numpy:
        # nodes = solve(phi(r_ij) - 0.5 I, z), phi(r) = r^2 log r; value(p) = sum_i nodes_i phi(|p - p_i|).  The reference appends the
        # result to the (u, v, 0) grid rows, so its dense array has two channels: zeros and the surface (float64).
        p = np.float64(uvzs[:, :2])
        r = np.sqrt(((p[:, None, :] - p[None, :, :]) ** 2).sum(-1))
        with np.errstate(divide="ignore", invalid="ignore"):
            A = np.where(r > 0, r ** 2 * np.log(r), 0.0) - np.eye(len(p)) * 0.5
            w = np.linalg.solve(A, np.float64(uvzs[:, 2]))
            d = np.sqrt((xs[..., None] - p[:, 0]) ** 2 + (ys[..., None] - p[:, 1]) ** 2)
            z = np.where(d > 0, d ** 2 * np.log(d), 0.0) @ w
        return np.stack([np.zeros(hw), np.where(mask, z, 0.0)], -1)
    else:
        raise NotImplementedError(inter_type)
    return np.where(mask, val, np.float32(0))


def interpolate_sparse2d(sparse2d, constrained_type=None, inter_type="lstsq"):
    """utils.py:347-353 (+ arr2d_to_uvzs, utils.py:318-328: samples in row-major order)."""
    m = (sparse2d != 0) & np.isfinite(sparse2d)
    ys, xs = np.nonzero(m)
    return interpolate_uvzs(np.array([xs, ys, sparse2d[m]]).T, sparse2d.shape[:2], constrained_type, inter_type)
