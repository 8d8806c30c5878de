"""CPU restatement of the reference's `Stereo.get_depth` chain with direct cv2 calls -- TEST INFRASTRUCTURE ONLY.

Follows calibrating/stereo_camera.py:125-177 (maps), :199-214 (own rectification), :216-242 (rectify), :408-413
(disparity_to_depth), :415-431 (unrectify_depth / undistort_img), :466-533 (set_stereo_matching / get_depth),
calibrating/stereo_matching.py:22-70 (SemiGlobalBlockMatching) and calibrating/utils.py:139-149, 173-200.
Pinned by tests/test_oracle_chain.py against tests/golden/rig320_*.npz, which were produced by the REAL reference
package (tests/golden/make_golden.py).  cv2 is the reference's own arithmetic, so this module is also what
`bench.py --impl reference` and the cpu_baseline leg time.
"""
import cv2
import numpy as np

EPS = 1e-8


def _proj(v, n):
    return v - np.dot(v, n) / (np.linalg.norm(n) ** 2) * n


def _rot(v1, v2):
    cross = np.cross(v1, v2)
    rad = np.arccos((v1 * v2).sum() / np.linalg.norm(v1) / np.linalg.norm(v2))
    return cv2.Rodrigues(rad * cross / (np.linalg.norm(cross) + EPS))[0]


def _K(c):
    return np.float64([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]]) if "K" not in c else np.float64(c["K"])


class SgbmPlugin:
    """stereo_matching.py:22-70 with the cv2 parameters exposed; defaults = the reference's hard-coded values."""

    def __init__(self, max_size=1000, **kw):
        p = dict(minDisparity=2, numDisparities=218, blockSize=11, uniquenessRatio=5, speckleWindowSize=200, speckleRange=2,
                 disp12MaxDiff=0, P1=8 * 11 * 11, P2=32 * 11 * 11)
        p.update(kw)
        self.max_size = max_size
        self.sgbm = cv2.StereoSGBM_create(**p)

    def _compute_float(self, a, b):
        d = self.sgbm.compute(a, b).astype(np.float32).clip(0)
        d[d < self.sgbm.getMinDisparity() * 16] = 0
        return d / 16.0

    def __call__(self, img1, img2):
        """stereo_matching.py:60-70; the two boxx.resize calls (un-vendored, unpinned) are stood in for by oracle/resize.py
        (= cv2.resize INTER_LINEAR)."""
        from . import resize
        return resize.scaled_matcher(self._compute_float, img1, img2, self.max_size)


class RefStereo:
    def __init__(self, rig):
        self.cam1, self.cam2 = rig["cam1"], rig["cam2"]
        self.K1, self.K2 = _K(self.cam1), _K(self.cam2)
        self.D1, self.D2 = np.float64(self.cam1["D"]), np.float64(self.cam2["D"])
        self.R = np.float64(rig["R"])
        self.t = np.float64(rig["t"]).reshape(3, 1)
        # stereo_recitfy (stereo_camera.py:199-214)
        z, nx, t = np.array([0, 0, 1.0]), np.array([-1.0, 0, 0]), self.t.squeeze()
        z2, z1 = _proj(z, t), _proj(self.R @ z, t)
        zp = z2 / np.linalg.norm(z2) + z1 / np.linalg.norm(z1)
        Rx = _rot(nx, t)
        Rz = _rot(Rx @ z, zp)
        self.R2 = (Rz @ Rx).T
        self.R1 = self.R2 @ self.R
        # _get_undistort_rectify_map (stereo_camera.py:125-177), xy_target=None, K_target=1
        self.xy = xy = tuple(self.cam1["xy"])
        self.K = self.K1.copy()
        self.K[:2, :2] *= 1
        self.K[:2, 2] += (np.array(xy) - self.cam1["xy"]) / 2

        def centre(cxy, cK, R):
            c = np.array([[0, 0, 1], [cxy[0], 0, 1], list(cxy) + [1], [0, cxy[1], 1]])
            uv = ((c @ np.linalg.inv(cK).T) @ R.T) @ self.K.T
            uv = uv[:, :2] / uv[:, 2:]
            return uv.mean(0) - self.K[:2, 2]

        ctr = (centre(self.cam1["xy"], self.K1, self.R1) + centre(self.cam2["xy"], self.K2, self.R2)) / 2
        self.K[:2, 2] = np.array(xy) / 2 - ctr
        self.map1 = cv2.initUndistortRectifyMap(self.K1, self.D1, self.R1, self.K, xy, cv2.CV_32FC1)
        self.map2 = cv2.initUndistortRectifyMap(self.K2, self.D2, self.R2, self.K, xy, cv2.CV_32FC1)
        w1, h1 = self.cam1["xy"]
        self.valid1 = (-0.5 < self.map1[0]) & (self.map1[0] < w1 - 0.5) & (-0.5 < self.map1[1]) & (self.map1[1] < h1 - 0.5)
        self.baseline = np.sum(self.t ** 2) ** 0.5
        self._umaps = None

    def set_stereo_matching(self, plugin, max_depth=None, translation_rectify_img=None):
        self.plugin = plugin
        self.translate = bool(max_depth) if translation_rectify_img is None else translation_rectify_img
        self.max_depth = max_depth or 1000
        self.min_disparity = int(self.K1[0, 0] * self.baseline / self.max_depth)
        return self

    def rectify(self, img1, img2, interp=cv2.INTER_LANCZOS4):
        r1 = cv2.remap(img1, self.map1[0], self.map1[1], interp)
        r2 = cv2.remap(img2, self.map2[0], self.map2[1], interp)
        if getattr(self, "translate", None) and self.min_disparity > 0:
            md = self.min_disparity
            r2[:, md:] = r2[:, :-md]
            r2[:, :md] = 0
        return r1, r2

    def disparity_to_depth(self, disparity):
        with np.errstate(all="ignore"):
            depth = 1.0 * self.baseline * self.K[0, 0] / disparity
            depth[depth > self.max_depth] = 0
            depth[depth < 0] = 0
        return depth

    def unrectify_depth(self, depth):
        R = self.R1.T
        if self._umaps is None:
            self._umaps = cv2.initUndistortRectifyMap(self.K, None, R, self.K1.copy(), tuple(self.cam1["xy"]), cv2.CV_32FC1)
        y, x = depth.shape
        ys, xs = np.mgrid[:y, :x]
        pts = np.array([xs.ravel(), ys.ravel(), np.ones(x * y, dtype=xs.dtype)]) * depth.ravel()[None]
        newz = (R @ np.linalg.inv(self.K) @ pts).T[:, 2].reshape(y, x)
        return cv2.remap(newz, self._umaps[0], self._umaps[1], cv2.INTER_NEAREST)

    def undistort_img(self, img1):
        return cv2.undistort(img1, self.K1, self.D1)

    def distort_depth(self, depth):
        """stereo_camera.py:433-464: forward splat of the undistorted depth image into the raw (distorted) cam1 image.  Every
        pixel is projected through the distortion model (cv2.undistortPoints with no distortion, then cv2.projectPoints
        with D), truncated to int, and of the pixels that land on the same target the one with the smallest source index
        wins (np.unique(..., return_index=True))."""
        w, h = self.cam1["xy"]
        res = np.zeros((h, w), dtype=depth.dtype)
        u, v = np.meshgrid(np.arange(w, dtype=np.int32), np.arange(h, dtype=np.int32))
        pts = np.stack([u.ravel(), v.ravel()], -1).astype(np.float32)
        zero = np.zeros(3, np.float32)
        und = cv2.undistortPoints(pts, self.K1, None)
        img_pts, _ = cv2.projectPoints(cv2.convertPointsToHomogeneous(und), zero, zero, self.K1, self.D1, und)
        img_pts = img_pts.reshape(-1, 2).astype(np.int32)
        uniq, index = np.unique(img_pts, axis=0, return_index=True)
        res[uniq[:, 1], uniq[:, 0]] = depth.ravel()[index]
        return res

    def get_depth(self, img1, img2, interp=cv2.INTER_LANCZOS4, return_distort_depth=False):
        r1, r2 = self.rectify(img1, img2, interp)
        disparity = self.plugin(r1, r2)
        if self.translate:
            disparity += self.min_disparity
        disparity = self.valid1 * disparity
        depth = self.disparity_to_depth(disparity)
        res = dict(rectify_img1=r1, rectify_img2=r2, disparity=disparity, rectify_depth=depth,
                   unrectify_depth=self.unrectify_depth(depth), undistort_img1=self.undistort_img(img1))
        if return_distort_depth:
            res.update(distort_img1=img1, distort_depth=self.distort_depth(res["unrectify_depth"]))
        return res
