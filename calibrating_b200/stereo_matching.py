"""Matcher plugins: host-side mirror of calibrating/stereo_matching.py:10-70 over the B200 engine.

`MetaStereoMatching` is the reference's plugin protocol verbatim in meaning (cfg dict; `__call__(img1, img2)` with
RGB uint8 (h,w,3) in, float (h,w) disparity or dict(disparity=...) out).  `SemiGlobalBlockMatching` is the drop-in for
the reference class of the same name: same constructor, same defaults (minDisparity 2, numDisparities 218, block 11,
uniqueness 5, speckle 200/2, disp12MaxDiff 0, P1/P2 = 968/3872, MODE_SGBM), same post-processing -- but `compute`
runs on the GPU through the C-ABI (include/b2s.h) instead of cv2.StereoSGBM.  Extra cfg keys expose the cv2 parameters
the reference hard-codes.
"""
import ctypes

import numpy as np

from . import _ffi

MODE_SGBM, MODE_HH, MODE_HH4 = 0, 1, 3  # cv2's MODE_SGBM_3WAY (2) depends on cv2's thread count and is not offered

# calibrating/stereo_matching.py:29-58
REFERENCE_DEFAULTS = dict(min_disparity=2, num_disparities=218, block_size=11, uniqueness_ratio=5, speckle_window_size=200,
                          speckle_range=2, disp12_max_diff=0, P1=8 * 1 * 11 * 11, P2=32 * 1 * 11 * 11, pre_filter_cap=0,
                          mode=MODE_SGBM, cost=0)
COST_BT, COST_CENSUS = 0, 1  # cost: 0 = cv2's Birchfield-Tomasi + block sum; 1 = 9x7 census / Hamming (extension, gray images)


class MetaStereoMatching:
    def __init__(self, cfg=None):
        self.cfg = cfg

    def __call__(self, img1, img2):
        # input: RGB uint8 (h, w, 3); output: float disparity (h, w) in pixels, or dict(disparity=disparity)
        raise NotImplementedError()


class StereoSGBM:
    """The part of the cv2.StereoSGBM object surface the reference touches (create / compute / getMinDisparity),
    executed by libb2s.so.  `compute` returns int16 16*disparity exactly like cv2."""

    def __init__(self, device=0, handle=None, **params):
        unknown = set(params) - set(REFERENCE_DEFAULTS)
        if unknown:
            raise TypeError("unknown StereoSGBM parameters: %s" % sorted(unknown))
        self.params = dict(min_disparity=0, num_disparities=16, block_size=3, P1=0, P2=0, disp12_max_diff=0, pre_filter_cap=0,
                           uniqueness_ratio=0, speckle_window_size=0, speckle_range=0, mode=MODE_SGBM, cost=0)
        self.params.update(params)
        self.handle = handle or _ffi.Handle(device)
        self._push()

    def _push(self):
        p = _ffi.SgbmParams(**{k: int(v) for k, v in self.params.items()})
        self.handle.call("b2s_set_sgbm_params", ctypes.byref(p))

    def getMinDisparity(self):
        return self.params["min_disparity"]

    def getNumDisparities(self):
        return self.params["num_disparities"]

    @staticmethod
    def _check_pair(left, right):
        left = np.ascontiguousarray(left)
        right = np.ascontiguousarray(right)
        if left.dtype != np.uint8 or right.dtype != np.uint8:
            raise ValueError("images must be uint8")
        if left.shape != right.shape or left.ndim not in (2, 3):
            raise ValueError("left/right shapes differ or are not images: %s vs %s" % (left.shape, right.shape))
        cn = 1 if left.ndim == 2 else left.shape[2]
        return left, right, left.shape[0], left.shape[1], cn

    def compute(self, left, right):
        left, right, H, W, cn = self._check_pair(left, right)
        out = np.empty((H, W), np.int16)
        self.handle.call("b2s_compute_disparity", _ffi.ptr(left), _ffi.ptr(right), H, W, cn, _ffi.ptr(out), None)
        return out

    def compute_float(self, left, right):
        """compute + the reference's post-processing (stereo_matching.py:63-64, /16) fused on the device."""
        left, right, H, W, cn = self._check_pair(left, right)
        out = np.empty((H, W), np.float32)
        self.handle.call("b2s_compute_disparity", _ffi.ptr(left), _ffi.ptr(right), H, W, cn, None, _ffi.ptr(out))
        return out


def StereoSGBM_create(minDisparity=0, numDisparities=16, blockSize=3, P1=0, P2=0, disp12MaxDiff=0, preFilterCap=0,
                      uniquenessRatio=0, speckleWindowSize=0, speckleRange=0, mode=MODE_SGBM, device=0, handle=None, cost=0):
    """Keyword-compatible with cv2.StereoSGBM_create (MODE_SGBM=0, MODE_HH=1 and MODE_HH4=3); cost=COST_CENSUS is an extension."""
    return StereoSGBM(device=device, handle=handle, min_disparity=minDisparity, num_disparities=numDisparities, block_size=blockSize,
                      P1=P1, P2=P2, disp12_max_diff=disp12MaxDiff, pre_filter_cap=preFilterCap, uniqueness_ratio=uniquenessRatio,
                      speckle_window_size=speckleWindowSize, speckle_range=speckleRange, mode=mode, cost=cost)


class SemiGlobalBlockMatching(MetaStereoMatching):
    """Drop-in for calibrating.SemiGlobalBlockMatching (calibrating/stereo_matching.py:22-70)."""

    def __init__(self, cfg=None, device=0, handle=None):
        if cfg is None:
            cfg = {}
        self.cfg = cfg
        self.max_size = self.cfg.get("max_size", 1000)
        params = dict(REFERENCE_DEFAULTS)
        params.update({k: cfg[k] for k in REFERENCE_DEFAULTS if k in cfg})
        self.stereo_sgbm = StereoSGBM(device=device, handle=handle, **params)

    def _max_size_option(self):
        """`max_size` as the engine takes it (B2S_OPT_MAX_SIZE; 0 = no limit)."""
        return int(min(max(self.max_size, 1), 2 ** 31 - 1))

    def __call__(self, img1, img2):
        """stereo_matching.py:60-70 in ONE C-ABI call: the down-scale to `max_size` (the reference's default is 1000), the
        matcher, `/16`, `clip(0)`, the `< minD * 16` zeroing, the up-scale and the `* w / sw` all run on the device.  The two
        `boxx.resize` calls of the reference are an un-vendored, unpinned dependency; the engine follows
        cv2.resize(INTER_LINEAR) (uint8 bit-exact; see oracle/resize.py)."""
        h = self.stereo_sgbm.handle
        h.call("b2s_set_option", 4, self._max_size_option())
        try:
            return self.stereo_sgbm.compute_float(img1, img2)
        finally:
            h.call("b2s_set_option", 4, 0)  # (the cv2-style StereoSGBM object sharing this handle never scales)


    def compute_batch(self, pairs, streams=2):
        """Throughput entry point (not in the reference): full-resolution disparities for a list of (img1, img2)."""
        from .batch import DisparityBatchEngine
        eng = getattr(self, "_batch_engine", None)
        if eng is None or len(eng.handles) != streams:
            eng = self._batch_engine = DisparityBatchEngine(self.stereo_sgbm.params, self.stereo_sgbm.handle.device, streams)
        return eng.compute_batch(pairs)


B200StereoMatching = SemiGlobalBlockMatching


class MatchingByBoard(MetaStereoMatching):
    """calibrating/stereo_matching.py:73-110: disparity from a calibration board's corners seen in both rectified images, densified
    over the board's convex hull.  The board (any object with the reference's `find_image_points(d)`, e.g. `calibrating.Chessboard`)
    and its detector stay on the host -- a few hundred points; the dense disparity is evaluated on the device
    (`utils.interpolate_sparse2d` -> b2s_interpolate_sparse)."""

    def __init__(self, board, dense_predict=True, device=0):
        self.board = board
        self.dense_predict = dense_predict
        self.device = device

    def __call__(self, img1, img2):
        from .utils import interpolate_sparse2d
        d1, d2 = dict(img=img1), dict(img=img2)
        self.board.find_image_points(d1)
        self.board.find_image_points(d2)
        p1, p2 = d1["image_points"], d2["image_points"]
        if isinstance(p1, dict):
            keys = sorted(set(p1).intersection(p2))
            p1 = np.concatenate([d1["image_points"][k] for k in keys], 0)
            p2 = np.concatenate([d2["image_points"][k] for k in keys], 0)
        rectify_std = np.std((p2 - p1)[:, 1])
        xyds = np.append(p1, (p1 - p2)[:, :1], axis=-1)
        h, w = img1.shape[:2]
        sparse = np.zeros((h, w), xyds.dtype)  # uvzs_to_arr2d (utils.py:291-317): rounded pixel, later points overwrite earlier ones
        xs, ys = np.int32(xyds[:, :2].round()).T
        ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
        sparse[ys[ok], xs[ok]] = xyds[ok, 2]
        if not self.dense_predict:
            return dict(disparity=sparse, rectify_std=rectify_std)
        with np.errstate(divide="ignore"):
            dense = 1 / interpolate_sparse2d(1 / sparse, "convex_hull", device=self.device)
        return dict(disparity=dense, rectify_std=rectify_std)


class FeatureMatchingAsStereoMatching(MetaStereoMatching):
    """calibrating/stereo_matching.py:113-142: sparse matches of a learned feature matcher (`feature_matching(img1, img2)` ->
    dict(uvs1, uvs2) in normalised coordinates, cfg["shape"]) -> nearest-neighbour densification at 1/8 resolution -> up-scale.
    The matcher is the user's; densification and up-scale run on the device."""

    def __init__(self, feature_matching, device=0):
        self.feature_matching = feature_matching
        self.device = device

    def __call__(self, img1, img2):
        from .utils import interpolate_uvzs
        matched = self.feature_matching(img1, img2)
        hw = img1.shape[:2]
        shape = self.feature_matching.cfg.get("shape", hw)
        small = (shape[0] // 8, shape[1] // 8)
        uvs1, uvs2 = matched["uvs1"] * small[::-1], matched["uvs2"] * small[::-1]
        uvds = np.concatenate((uvs1, (uvs1 - uvs2)[:, :1]), 1)
        disparity = interpolate_uvzs(uvds, small, constrained_type=None, inter_type="nearest", device=self.device)
        if tuple(small) != tuple(hw):
            from .stereo_camera import _module_handle
            # `disparity * hw[1] / resize_shape[1]` (stereo_matching.py:133-137) is evaluated left to right on the host so that
            # the two roundings stay two; the nearest-neighbour up-scale runs on the device.
            src = np.ascontiguousarray(disparity * hw[1] / small[1], np.float32)
            out = np.empty(tuple(hw), np.float32)
            _module_handle(self.device).call("b2s_resize_nearest_f32", _ffi.ptr(src), small[0], small[1], _ffi.ptr(out), hw[0], hw[1], 1.0)
            disparity = out
        return dict(disparity=disparity, matched=matched)

