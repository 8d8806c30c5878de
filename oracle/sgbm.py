"""ctypes front-end of oracle/sgbm_ref.c -- TEST INFRASTRUCTURE ONLY.

Restates `cv2.StereoSGBM.compute` as called by calibrating/stereo_matching.py:63.
"""
import ctypes

import numpy as np

from . import build as _build

MODE_SGBM, MODE_HH, MODE_HH4 = 0, 1, 3


class SgbmParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "min_disparity", "num_disparities", "block_size", "P1", "P2", "disp12_max_diff",
        "pre_filter_cap", "uniqueness_ratio", "speckle_window_size", "speckle_range", "mode", "cost")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.oracle_sgbm_compute.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def sgbm_compute(left, right, *, min_disparity=0, num_disparities=16, block_size=3, P1=0, P2=0,
                 disp12_max_diff=0, pre_filter_cap=0, uniqueness_ratio=0, speckle_window_size=0,
                 speckle_range=0, mode=MODE_SGBM, cost=0, want_volumes=False, want_raw=False):
    """Same keyword meaning as cv2.StereoSGBM_create (cost=1: 9x7 census / Hamming cost, this engine's own definition,
    not in cv2 -- parity unpinned).  Returns int16 (H,W) disparity*16, or a dict
    with the C / S volumes (H, width1, D) and the pre-median disparity when asked."""
    left = np.ascontiguousarray(left, np.uint8)
    right = np.ascontiguousarray(right, np.uint8)
    assert left.shape == right.shape
    H, W = left.shape[:2]
    cn = 1 if left.ndim == 2 else left.shape[2]
    prm = SgbmParams(min_disparity, num_disparities, block_size, P1, P2, disp12_max_diff,
                     pre_filter_cap, uniqueness_ratio, speckle_window_size, speckle_range, mode, cost)
    disp = np.empty((H, W), np.int16)
    maxd = min_disparity + num_disparities
    width1 = (W + min(min_disparity, 0)) - max(maxd, 0)  # maxX1 - minX1
    C = S = raw = None
    if want_volumes and width1 > 0:
        C = np.empty((H, width1, num_disparities), np.int16)
        S = np.empty_like(C)
    if want_raw or want_volumes:
        raw = np.empty((H, W), np.int16)
    rc = lib().oracle_sgbm_compute(_p(left), _p(right), H, W, cn, ctypes.byref(prm), _p(disp), _p(C), _p(S), _p(raw))
    if rc == -1:
        raise ValueError("input images are too small for your window size and max disparity")
    if rc != 0:
        raise ValueError("unsupported parameters (rc=%d)" % rc)
    if want_volumes or want_raw:
        return dict(disp=disp, C=C, S=S, raw=raw)
    return disp


def median3(a):
    a = np.ascontiguousarray(a, np.int16)
    out = np.empty_like(a)
    lib().oracle_median3(_p(a), _p(out), a.shape[0], a.shape[1])
    return out


def filter_speckles(a, new_val, max_speckle_size, max_diff):
    a = np.array(a, np.int16, order="C", copy=True)
    lib().oracle_filter_speckles(_p(a), a.shape[0], a.shape[1], int(new_val), int(max_speckle_size), int(max_diff))
    return a
