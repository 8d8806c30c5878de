"""CPU restatement of `Cam.project_cam2_depth` -- TEST INFRASTRUCTURE ONLY.

Follows calibrating/camera.py:298-309 and calibrating/utils.py:203-210 (_get_appropriate_interpolation_rate), :213-250
(depth_to_point_cloud), :152-161 (apply_T_to_point_cloud), :254-288 (point_cloud_to_arr2d) and :291-317 (uvzs_to_arr2d).
Pinned by tests/test_oracle_chain.py against tests/golden/rig320_project.npz, produced by the REAL reference package.
"""
import cv2
import numpy as np


def interpolation_rate(K1, K2, interpolation=1.5):
    if not interpolation:
        return 1
    rate = K1[0, 0] / K2[0, 0] * interpolation
    return max(rate, 1) if interpolation >= 1 else rate


def project_cam2_depth(K1, xy1, K2, depth2, T, interpolation=1.5):
    """Depth image of cam2 seen from cam1: nearest-neighbour up-sampling by `rate`, un-projection with K2, rigid transform T
    (cam2 -> cam1), projection with K1, rounding to the pixel grid; of the points landing on a pixel the last one written in
    the order of descending z survives, i.e. the smallest z."""
    K1, K2, T = np.float64(K1), np.float64(K2), np.float64(T)
    if depth2.dtype == np.uint16:
        depth2 = np.float32(depth2 / 1000.0)
    rate = interpolation_rate(K1, K2, interpolation)
    y, x = depth2.shape
    if rate == 1:
        mask = depth2 != 0
        vs, us = np.mgrid[:y, :x][:, mask]
        d = depth2[mask]
    else:
        y_, x_ = int(round(y * rate)), int(round(x * rate))
        depth_ = cv2.resize(depth2, (x_, y_), interpolation=cv2.INTER_NEAREST)
        mask = depth_ != 0
        vs, us = np.mgrid[:y_, :x_][:, mask] / rate
        d = depth_[mask]
    pts = (np.linalg.inv(K2) @ (np.array([us, vs, np.ones_like(us)]) * d)).T
    pts4 = np.ones((len(pts), 4))
    pts4[:, :3] = pts
    pts1 = (T @ pts4.T).T[:, :3]
    xyz = pts1 @ K1.T
    xyz[:, :2] /= xyz[:, 2:]
    xyz = xyz[np.argsort(-xyz[:, 2])]
    w, h = xy1
    out = np.ones((h, w), xyz.dtype) * 0
    xs, ys = np.int32(xyz[:, :2].round()).T
    ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
    out[ys[ok], xs[ok]] = xyz[ok, 2]
    return out


def depth_to_point_cloud(depth, K, interpolation_rate=1, return_xyzuv=False):
    """calibrating/utils.py:213-250 restated."""
    K = np.float64(K)
    y, x = depth.shape
    if depth.dtype == np.uint16:
        depth = np.float32(depth / 1000.0)
    if interpolation_rate == 1:
        mask = depth != 0
        vs, us = np.mgrid[:y, :x][:, mask]
        pts = (np.array([us, vs, np.ones_like(us)]) * depth[mask]).T
    else:
        y_, x_ = int(round(y * interpolation_rate)), int(round(x * interpolation_rate))
        depth_ = cv2.resize(depth, (x_, y_), interpolation=cv2.INTER_NEAREST)
        mask = depth_ != 0
        vs, us = np.mgrid[:y_, :x_][:, mask] / interpolation_rate
        pts = (np.array([us, vs, np.ones_like(us)]) * depth_[mask]).T
    cloud = (np.linalg.inv(K) @ pts.T).T
    return np.concatenate([cloud, us[:, None], vs[:, None]], -1) if return_xyzuv else cloud


def point_cloud_to_depth(points, K, xy):
    """calibrating/utils.py:254-317 restated (values=None, bg_value=0)."""
    xyz = np.float64(points) @ np.float64(K).T
    with np.errstate(all="ignore"):
        xyz[:, :2] /= xyz[:, 2:]
    xyz = xyz[np.argsort(-xyz[:, 2])]
    w, h = xy
    out = np.ones((h, w), xyz.dtype) * 0
    with np.errstate(all="ignore"):
        xs, ys = np.int32(xyz[:, :2].round()).T
    ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
    out[ys[ok], xs[ok]] = xyz[ok, 2]
    return out
