"""`Stereo`: host-side mirror of calibrating/stereo_camera.py for the `get_depth` hot path, executed on a B200.

Same surface as the reference class for this path (`load/dump`, `rectify`, `set_stereo_matching`, `get_depth`,
`disparity_to_depth`, `depth_to_disparity`, `unrectify_depth`, `undistort_img`, attributes `R t R1 R2 K xy
undistort_rectify_map1/2 rectify_valid_mask1 min_disparity max_depth translation_rectify_img`), same result-dict keys,
shapes and dtypes (calibrating/stereo_camera.py:492-533).  What differs is where the arithmetic runs: every per-pair
array operation goes through the C-ABI of libb2s.so (include/b2s.h); the once-per-rig setup (3x3 algebra and
`cv2.initUndistortRectifyMap`, stereo_camera.py:125-177, 199-214) stays on the host exactly as SURVEY.md section 2 scopes
it.  Calibration (`Stereo(cam1, cam2)` from image points, stereo_camera.py:55-123) is out of scope: rigs are built
from the reference's own yaml/dict schema via `Stereo.load`, or from (cam1, cam2, R, t).
"""
import copy as _copy
import ctypes
import weakref

import cv2
import numpy as np
import yaml

from . import _ffi
from .stereo_matching import MetaStereoMatching, SemiGlobalBlockMatching

_EPS = 1e-8  # calibrating/utils.py:12


class Cam:
    """Intrinsics holder with the reference's dict schema {fx,fy,cx,cy,D,xy,name} or {K,D,xy} (camera.py:407-448)."""

    def __init__(self, K=None, D=None, xy=None, name=None):
        self.K = None if K is None else np.float64(K).reshape(3, 3)
        self.D = np.zeros((1, 5)) if D is None else np.float64(D)
        self.xy = None if xy is None else tuple(int(v) for v in xy)
        self.name = name

    @classmethod
    def load(cls, src):
        if isinstance(src, Cam):
            return src.copy()
        if not isinstance(src, dict):
            if "\n" in src:
                src = yaml.safe_load(src)
            else:
                with open(src) as f:
                    src = yaml.safe_load(f)
        d = _copy.deepcopy(src)
        if "K" in d:
            K = np.float64(d["K"])
        else:
            K = np.float64([[d["fx"], 0, d["cx"]], [0, d["fy"], d["cy"]], [0, 0, 1]])
        assert "xy" in d and len(d["xy"]), "Need xy"
        return cls(K, d.get("D"), d["xy"], d.get("name"))

    def dump(self, path="", return_dict=False):
        dic = dict(D=self.D.tolist(), xy=list(self.xy), fx=float(self.K[0, 0]), fy=float(self.K[1, 1]),
                   cx=float(self.K[0, 2]), cy=float(self.K[1, 2]))
        if self.name is not None:
            dic["name"] = self.name
        if return_dict:
            return dic
        s = yaml.safe_dump(dic)
        if path:
            with open(path, "w") as f:
                f.write(s)
        return s

    def copy(self):
        return Cam(self.K.copy(), self.D.copy(), self.xy, self.name)

    def project_cam2_depth(cam1, cam2, depth2, T=None, interpolation=1.5, device=0):
        """calibrating/camera.py:298-309 on the device: the depth image `depth2` of `cam2` seen from this camera, T = pose of
        cam2 in this camera's frame (4x4).  The reference derives a missing T from shared calibration boards
        (`get_T_cam2_in_self`, calibration is out of scope here), so T is required."""
        if T is None:
            raise NotImplementedError("T from calibration boards (camera.py:283-296) is out of scope: pass T")
        depth2 = np.asarray(depth2)
        if depth2.ndim != 2:
            raise ValueError("depth2 must be a 2-D depth image")
        if depth2.dtype == np.uint16:
            depth2 = np.float32(depth2 / 1000.0)  # utils.py:219-220
        rate = 1
        if interpolation:  # utils._get_appropriate_interpolation_rate (utils.py:203-210)
            rate = cam1.K[0, 0] / cam2.K[0, 0] * interpolation
            if interpolation >= 1:
                rate = max(rate, 1)
        depth2 = np.ascontiguousarray(depth2, np.float64)
        w1, h1 = cam1.xy
        out = np.empty((h1, w1), np.float64)
        h = _module_handle(device)
        arr = lambda m: (ctypes.c_double * m.size)(*np.float64(m).ravel())
        h.call("b2s_project_depth", _ffi.ptr(depth2), depth2.shape[1], depth2.shape[0], float(rate), arr(np.linalg.inv(cam2.K)),
               arr(np.float64(T).reshape(4, 4)), arr(cam1.K), int(w1), int(h1), _ffi.ptr(out))
        return out


_HANDLES = {}


def _module_handle(device):
    """One shared engine handle per device for the rig-independent helpers (Cam.project_cam2_depth)."""
    if device not in _HANDLES:
        _HANDLES[device] = _ffi.Handle(device)
    return _HANDLES[device]


def _project_on_plane(v, plane_normal):
    # calibrating/utils.py:139-140
    return v - np.dot(v, plane_normal) / (np.linalg.norm(plane_normal) ** 2) * plane_normal


def _shortest_rotation(v1, v2):
    # calibrating/utils.py:143-149
    axis = np.cross(v1, v2)
    angle = np.arccos((v1 * v2).sum() / np.linalg.norm(v1) / np.linalg.norm(v2))
    return cv2.Rodrigues(angle * axis / (np.linalg.norm(axis) + _EPS))[0]


class Stereo:
    MAX_DEPTH = 1000
    DUMP_ATTRS = ["R", "t", "retval"]

    def __init__(self, cam1=None, cam2=None, xy_target=None, K_target=1, R=None, t=None, device=0, interp="lanczos4", maps="host"):
        """interp: "lanczos4" (what the reference's rectify uses) or "linear" (north_star's fast bilinear mode).
        maps: "host" uploads the cv2.initUndistortRectifyMap results (stereo_camera.py:159-165); "device" sends only the
        calibration parameters and the engine generates the same maps itself (b2s_set_rig_params)."""
        self.xy_target = xy_target
        self.K_target = K_target
        self.device = device
        self.interp = interp
        self.maps = maps
        self._handle = None
        self._rig_dirty = True
        if cam1 is None:
            return
        if R is None or t is None:
            raise NotImplementedError("stereo calibration from image points (cv2.stereoCalibrate, stereo_camera.py:95-123) is "
                                      "out of scope; pass R and t, or use Stereo.load(yaml_or_dict)")
        self.cam1, self.cam2 = cam1, cam2
        self.R = np.float64(R).reshape(3, 3)
        self.t = np.float64(t).reshape(3, 1)
        self._get_undistort_rectify_map()

    # ---- persistence (stereo_camera.py:244-297) --------------------------------------------------------------
    def dump(self, path="", return_dict=False):
        dic = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items() if k in self.DUMP_ATTRS}
        dic["cam1"] = self.cam1.dump(return_dict=True)
        dic["cam2"] = self.cam2.dump(return_dict=True)
        if return_dict:
            return dic
        s = yaml.safe_dump(dic)
        if path:
            with open(path, "w") as f:
                f.write(s)
        return s

    def load(self, src=None, **kw):
        if src is None or not isinstance(self, Stereo):  # called as Stereo.load(x)
            src, self = (self if src is None else src), Stereo(**kw)
        if isinstance(src, Stereo):
            return src.copy()
        if not isinstance(src, (list, dict)):
            if "\n" in src:
                dic = yaml.safe_load(src)
            else:
                with open(src) as f:
                    dic = yaml.safe_load(f)
        else:
            dic = dict(src)
        self.cam1 = Cam.load(dic.pop("cam1"))
        self.cam2 = Cam.load(dic.pop("cam2"))
        if "R" not in dic and "T" in dic:
            T = np.float64(dic.pop("T"))
            dic["r"], dic["t"] = cv2.Rodrigues(T[:3, :3])[0], T[:3, 3:]
        if "R" not in dic and "r" in dic:
            dic["R"] = cv2.Rodrigues(np.float64(dic.pop("r")))[0]
        dic.setdefault("R", np.eye(3))
        dic.pop("_calibrating_version", None)
        for k, v in dic.items():
            setattr(self, k, np.array(v) if k in self.DUMP_ATTRS else v)
        self.R = np.float64(self.R)
        self.t = np.float64(self.t).reshape(3, 1)
        self._get_undistort_rectify_map()
        return self

    def copy(self):
        new = type(self)(device=self.device, interp=self.interp, maps=self.maps)
        new.load(self.dump(return_dict=True))
        return new

    # ---- once-per-rig setup on the host (stereo_camera.py:125-177, 199-214) -----------------------------------
    def stereo_recitfy(self):
        z_axis = np.array([0, 0, 1.0])
        neg_x = np.array([-1.0, 0, 0])
        t = self.t.squeeze()
        z2 = _project_on_plane(z_axis, t)
        z1 = _project_on_plane(self.R @ z_axis, t)
        z_mean = z2 / np.linalg.norm(z2) + z1 / np.linalg.norm(z1)
        R_align_x = _shortest_rotation(neg_x, t)
        R_align_z = _shortest_rotation(R_align_x @ z_axis, z_mean)
        self.R2 = (R_align_z @ R_align_x).T
        self.R1 = self.R2 @ self.R[:3, :3]

    def _get_undistort_rectify_map(self):
        self.stereo_recitfy()
        if self.xy_target is None:
            self.xy_target = self.cam1.xy
        if isinstance(self.xy_target, (int, float)):
            self.xy_target = [int(round(i * self.xy_target)) for i in self.cam1.xy]
        self.xy = xy = tuple(self.xy_target)
        self.K = self.K_target
        if isinstance(self.K_target, (int, float)):
            self.K = self.cam1.K.copy()
            self.K[:2, :2] *= self.K_target
            self.K[:2, 2] += (np.array(xy) - self.cam1.xy) / 2
        if not isinstance(self.K_target, np.ndarray):
            # "better_cx_cy" (stereo_camera.py:137-157): centre the projected corners of both raw images
            def centre_offset(cam_xy, cam_K, R):
                corners = np.array([[0, 0, 1], [cam_xy[0], 0, 1], list(cam_xy) + [1], [0, cam_xy[1], 1]])
                rays = corners @ np.linalg.inv(cam_K).T
                uv = (rays @ R.T) @ self.K.T
                uv = uv[:, :2] / uv[:, 2:]
                return uv.mean(0) - self.K[:2, 2]

            c = (centre_offset(self.cam1.xy, self.cam1.K, self.R1) + centre_offset(self.cam2.xy, self.cam2.K, self.R2)) / 2
            self.K[:2, 2] = np.array(xy) / 2 - c
        self.undistort_rectify_map1 = cv2.initUndistortRectifyMap(self.cam1.K, self.cam1.D, self.R1, self.K, xy, cv2.CV_32FC1)
        self.undistort_rectify_map2 = cv2.initUndistortRectifyMap(self.cam2.K, self.cam2.D, self.R2, self.K, xy, cv2.CV_32FC1)
        mx, my = self.undistort_rectify_map1
        w1, h1 = self.cam1.xy
        self.rectify_valid_mask1 = (-0.5 < mx) & (mx < w1 - 0.5) & (-0.5 < my) & (my < h1 - 0.5)
        self._unrectify_depth_maps = None
        self._rig_dirty = True

    # ---- small properties (stereo_camera.py:386-406) -------------------------------------------------------------
    def get_max_depth(self):
        return getattr(self, "max_depth", self.MAX_DEPTH)

    @property
    def D(self):
        return np.zeros((1, 5))

    @property
    def T(self):
        T = np.eye(4)
        T[:3, :3] = self.R
        T[:3, 3] = self.t.squeeze()
        return T

    @property
    def baseline(self):
        return np.sum(self.t ** 2) ** 0.5

    def depth_to_disparity(self, depth):
        return 1.0 * self.baseline * self.K[0, 0] / depth

    # ---- engine plumbing ---------------------------------------------------------------------------------------------
    @property
    def handle(self):
        if self._handle is None:
            sm = getattr(self, "stereo_matching", None)
            own = getattr(getattr(sm, "stereo_sgbm", None), "handle", None)
            self._handle = own if own is not None and own.device == self.device else _ffi.Handle(self.device)
        return self._handle

    def _unrectify_maps(self):
        # maps of rotate_depth_by_remap (utils.py:184-191), cached like Stereo._unrectify_depth_maps
        if self._unrectify_depth_maps is None:
            self._unrectify_depth_maps = cv2.initUndistortRectifyMap(self.K, None, self.R1.T, self.cam1.K.copy(),
                                                                     tuple(self.cam1.xy), cv2.CV_32FC1)
        return self._unrectify_depth_maps

    @staticmethod
    def _map_params(K, D, R, Knew, size):
        """Inputs of one cv2.initUndistortRectifyMap(K, D, R, Knew, size) call as a b2s_map_params struct."""
        D = np.zeros(0) if D is None else np.float64(D).ravel()
        if D.size > 12 and np.any(D[12:] != 0):
            raise NotImplementedError("tilted sensor model (tauX, tauY) is not supported by the device-side map generation")
        k = np.zeros(12)
        k[:min(D.size, 12)] = D[:12]
        iR = np.linalg.inv(np.float64(Knew) @ (np.eye(3) if R is None else np.float64(R)))
        K = np.float64(K)
        return _ffi.MapParams(int(size[0]), int(size[1]), K[0, 0], K[1, 1], K[0, 2], K[1, 2], (ctypes.c_double * 12)(*k),
                              (ctypes.c_double * 9)(*iR.ravel()))

    def _push_rig_params(self, handle=None):
        w1, h1 = self.cam1.xy
        rp = _ffi.RigParams()
        rp.W, rp.H = self.xy
        rp.W1, rp.H1 = w1, h1
        rp.W2, rp.H2 = self.cam2.xy
        rp.rect1 = self._map_params(self.cam1.K, self.cam1.D, self.R1, self.K, self.xy)
        rp.rect2 = self._map_params(self.cam2.K, self.cam2.D, self.R2, self.K, self.xy)
        rp.unrect = self._map_params(self.K, None, self.R1.T, self.cam1.K, (w1, h1))
        rp.undist = self._map_params(self.cam1.K, self.cam1.D, None, self.cam1.K, (w1, h1))
        M = self.R1.T @ np.linalg.inv(self.K)
        rp.unrect_m = (ctypes.c_double * 3)(*M[2])
        rp.fx_baseline = float(1.0 * self.baseline * self.K[0, 0])
        rp.max_depth = float(self.get_max_depth())
        rp.min_disparity = int(self.min_disparity) if getattr(self, "translation_rectify_img", None) else 0
        rp.interp = {"lanczos4": 0, "linear": 1}[self.interp]
        h = handle or self.handle
        h.call("b2s_set_rig_params", ctypes.byref(rp))
        h._rig_owner = weakref.ref(self)
        if handle is None:
            self._rig_dirty = False

    def _push_cam1_model(self, handle=None):
        D = np.float64(self.cam1.D).ravel()
        k = np.zeros(12)
        k[:min(D.size, 12)] = D[:12]
        K = self.cam1.K
        (handle or self.handle).call("b2s_set_cam1_model", float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]),
                                     (ctypes.c_double * 12)(*k))

    def _push_rig(self, handle=None):
        """Upload the per-rig constants to the engine (once per rig and handle).  A handle holds ONE rig: when several
        Stereo objects share a matcher (and with it the engine handle), as example/test_different_stereo.py does, each of
        them re-uploads its rig whenever another one was the last to do so."""
        if handle is None:
            owner = getattr(self.handle, "_rig_owner", None)
            if not self._rig_dirty and owner is not None and owner() is self:
                return
            self._drop_batch()  # (the handles of get_depth_batch carry the old rig)
        if self.maps == "device":
            return self._push_rig_params(handle)
        h = handle or self.handle
        m1x, m1y = (np.ascontiguousarray(m, np.float32) for m in self.undistort_rectify_map1)
        m2x, m2y = (np.ascontiguousarray(m, np.float32) for m in self.undistort_rectify_map2)
        mask = np.ascontiguousarray(self.rectify_valid_mask1, np.uint8)
        umx, umy = (np.ascontiguousarray(m, np.float32) for m in self._unrectify_maps())
        w1, h1 = self.cam1.xy
        und_xy, und_fxy = cv2.initUndistortRectifyMap(self.cam1.K, self.cam1.D, None, self.cam1.K, (w1, h1), cv2.CV_16SC2)
        und_xy, und_fxy = np.ascontiguousarray(und_xy, np.int16), np.ascontiguousarray(und_fxy, np.uint16)
        M = self.R1.T @ np.linalg.inv(self.K)
        rig = _ffi.Rig()
        rig.W, rig.H = self.xy
        rig.W1, rig.H1 = w1, h1
        rig.W2, rig.H2 = self.cam2.xy
        keep = [m1x, m1y, m2x, m2y, mask, umx, umy, und_xy, und_fxy]
        (rig.map1x, rig.map1y, rig.map2x, rig.map2y, rig.valid_mask1, rig.unrect_mapx, rig.unrect_mapy, rig.undist_xy,
         rig.undist_fxy) = [a.ctypes.data for a in keep]
        rig.unrect_m = (ctypes.c_double * 3)(*M[2])
        rig.fx_baseline = float(1.0 * self.baseline * self.K[0, 0])
        rig.max_depth = float(self.get_max_depth())
        rig.min_disparity = int(self.min_disparity) if getattr(self, "translation_rectify_img", None) else 0
        rig.interp = {"lanczos4": 0, "linear": 1}[self.interp]
        h.call("b2s_set_rig", ctypes.byref(rig))
        h._rig_owner = weakref.ref(self)
        self._push_cam1_model(handle)
        if handle is None:
            self._rig_dirty = False

    def _drop_batch(self):
        """Close the handles of get_depth_batch and give their pinned staging buffers back."""
        for b in getattr(self, "_batch", None) or []:
            try:
                b["handle"].sync()
            except Exception:
                pass
            for a in b["pin"].values():
                _ffi.pinned_free(a)
            b["pin"].clear()
            b["handle"].close()
        self._batch = None

    @staticmethod
    def _get_img(path_or_np):
        if isinstance(path_or_np, str):
            return cv2.imread(path_or_np)[..., ::-1]
        return path_or_np

    @staticmethod
    def _prep(img):
        img = np.ascontiguousarray(Stereo._get_img(img))
        if img.dtype != np.uint8 or img.ndim not in (2, 3):
            raise ValueError("images must be uint8 (h,w) or (h,w,3)")
        return img, (1 if img.ndim == 2 else img.shape[2])

    def _check_raw(self, img1, img2):
        (w1, h1), (w2, h2) = self.cam1.xy, self.cam2.xy
        if img1.shape[:2] != (h1, w1) or img2.shape[:2] != (h2, w2):
            raise ValueError("image sizes %s/%s do not match cam1.xy/cam2.xy %s/%s" % (img1.shape, img2.shape, self.cam1.xy, self.cam2.xy))

    # ---- the hot path -------------------------------------------------------------------------------------------------------
    def rectify(self, img1, img2):
        """stereo_camera.py:216-242 -- two INTER_LANCZOS4 remaps (+ the min_disparity shift of the right image)."""
        img1, cn = self._prep(img1)
        img2, cn2 = self._prep(img2)
        if cn != cn2:
            raise ValueError("img1/img2 channel counts differ")
        self._check_raw(img1, img2)
        self._push_rig()
        w, h = self.xy
        shape = (h, w) if img1.ndim == 2 else (h, w, cn)
        out1, out2 = np.empty(shape, np.uint8), np.empty(shape, np.uint8)
        self.handle.call("b2s_rectify", _ffi.ptr(img1), _ffi.ptr(img2), cn, _ffi.ptr(out1), _ffi.ptr(out2))
        return [out1, out2]

    def disparity_to_depth(self, disparity):
        """stereo_camera.py:408-413 (float64 result, like NumPy 2 promotion of the reference expression)."""
        self._push_rig()
        disparity = np.ascontiguousarray(disparity, np.float32)
        w, h = self.xy
        if disparity.shape != (h, w):
            raise ValueError("disparity must have shape %s" % ((h, w),))
        depth = np.empty((h, w), np.float64)
        self.handle.call("b2s_disparity_to_depth", _ffi.ptr(disparity), _ffi.ptr(depth))
        return depth

    def unrectify_depth(self, depth):
        """stereo_camera.py:415-428 -> utils.rotate_depth_by_remap (utils.py:173-200)."""
        self._push_rig()
        depth = np.ascontiguousarray(depth, np.float64)
        w, h = self.xy
        w1, h1 = self.cam1.xy
        if depth.shape != (h, w):
            raise ValueError("depth must have shape %s" % ((h, w),))
        out = np.empty((h1, w1), np.float64)
        self.handle.call("b2s_unrectify_depth", _ffi.ptr(depth), _ffi.ptr(out))
        return out

    def undistort_img(self, img1):
        """stereo_camera.py:430-431 (cv2.undistort) on the device."""
        self._push_rig()
        img1, cn = self._prep(img1)
        w1, h1 = self.cam1.xy
        if img1.shape[:2] != (h1, w1):
            raise ValueError("img1 must be cam1-sized %s" % ((h1, w1),))
        out = np.empty(img1.shape, np.uint8)
        self.handle.call("b2s_undistort_img", _ffi.ptr(img1), cn, _ffi.ptr(out))
        return out

    def distort_depth(self, depth):
        """stereo_camera.py:433-464 on the device (the reference's version is documented as "OOM warning and very slow")."""
        self._push_rig()
        depth = np.ascontiguousarray(depth, np.float64)
        w1, h1 = self.cam1.xy
        if depth.shape != (h1, w1):
            raise ValueError("depth must be cam1-sized %s" % ((h1, w1),))
        out = np.empty((h1, w1), np.float64)
        self.handle.call("b2s_distort_depth", _ffi.ptr(depth), _ffi.ptr(out))
        return out

    def set_stereo_matching(self, stereo_matching, max_depth=None, translation_rectify_img=None):
        """stereo_camera.py:466-489."""
        self.stereo_matching = stereo_matching
        self.translation_rectify_img = bool(max_depth) if translation_rectify_img is None else translation_rectify_img
        self.max_depth = max_depth or self.MAX_DEPTH
        self.min_disparity = int(self.cam1.K[0, 0] * self.baseline / self.max_depth)
        self._rig_dirty = True
        return self

    def get_depth(self, img1, img2, return_unrectify_depth=True, return_distort_depth=False):
        """stereo_camera.py:492-533.  With the built-in `SemiGlobalBlockMatching` (any `max_size`) the whole chain runs
        in one C-ABI call (one upload, one stream of kernels, one download); with a foreign `MetaStereoMatching` plugin
        the rectify half and the depth half run on the device around the plugin's host call."""
        assert hasattr(self, "stereo_matching"), "Please stereo.set_stereo_matching(stereo_matching)"
        img1, cn = self._prep(img1)
        img2, cn2 = self._prep(img2)
        if cn != cn2:
            raise ValueError("img1/img2 channel counts differ")
        self._check_raw(img1, img2)
        w, h = self.xy
        w1, h1 = self.cam1.xy
        want = bool(return_unrectify_depth or return_distort_depth)
        sm = self.stereo_matching
        fused = isinstance(sm, SemiGlobalBlockMatching) and sm.stereo_sgbm.handle is self.handle
        self._push_rig()
        ishape = (h, w) if img1.ndim == 2 else (h, w, cn)
        result = {}
        out = _ffi.DepthOut()
        disparity = np.empty((h, w), np.float32)
        rectify_depth = np.empty((h, w), np.float64)
        out.disparity, out.rectify_depth = disparity.ctypes.data, rectify_depth.ctypes.data
        if want:
            unrectify_depth = np.empty((h1, w1), np.float64)
            undistort_img1 = np.empty(img1.shape, np.uint8)
            out.unrectify_depth, out.undistort_img1 = unrectify_depth.ctypes.data, undistort_img1.ctypes.data
        if return_distort_depth:
            distort_depth = np.empty((h1, w1), np.float64)
            out.distort_depth = distort_depth.ctypes.data
        if fused:
            rectify_img1, rectify_img2 = np.empty(ishape, np.uint8), np.empty(ishape, np.uint8)
            out.rectify_img1, out.rectify_img2 = rectify_img1.ctypes.data, rectify_img2.ctypes.data
            self.handle.call("b2s_set_option", 4, sm._max_size_option())  # the matcher's max_size (stereo_matching.py:26,61)
            try:
                self.handle.call("b2s_get_depth", _ffi.ptr(img1), _ffi.ptr(img2), cn, int(want), ctypes.byref(out))
            finally:
                self.handle.call("b2s_set_option", 4, 0)
        else:
            rectify_img1, rectify_img2 = self.rectify(img1, img2)
            plug = sm(rectify_img1, rectify_img2)
            if isinstance(plug, dict):
                result.update(plug)
                plug = plug["disparity"]
            plug = np.asarray(plug)
            if plug.ndim != 2 or plug.shape != (h, w):
                # (the reference would fail in `rectify_valid_mask1 * disparity`, stereo_camera.py:512)
                raise ValueError("the stereo matching plugin returned a disparity of shape %s, expected %s" % (plug.shape, (h, w)))
            plug = np.ascontiguousarray(plug, np.float32)  # (the engine computes on float32 disparities, like cv2's)
            self.handle.call("b2s_depth_from_disparity", _ffi.ptr(plug), _ffi.ptr(img1), cn, int(want), ctypes.byref(out))
        result.update(rectify_img1=rectify_img1, rectify_depth=rectify_depth, disparity=disparity, rectify_img2=rectify_img2)
        if want:
            result.update(unrectify_depth=unrectify_depth, undistort_img1=undistort_img1)
        if return_distort_depth:
            result.update(distort_img1=img1, distort_depth=distort_depth)
        return result


    # ---- throughput: several pairs in flight (not in the reference) -------------------------------
    def get_depth_batch(self, pairs, streams=4, keys=("unrectify_depth",), out=None):
        """`get_depth` for a list of (img1, img2) with `streams` engine handles (= CUDA streams) in flight: upload, the ~20
        kernels and the download of different pairs overlap.  Needs the built-in `SemiGlobalBlockMatching`
        (the one-call path of `get_depth`).  keys: which result arrays to return, any of rectify_img1, rectify_img2, disparity,
        rectify_depth, unrectify_depth, undistort_img1, distort_depth.  Returns a list of dicts.
        Host copies are avoided when the caller supplies pinned memory (`_ffi.pinned_empty`): pinned input images are
        uploaded in place, and `out` (a list of dicts of pinned arrays, one per pair, same keys) receives the results directly."""
        assert hasattr(self, "stereo_matching"), "Please stereo.set_stereo_matching(stereo_matching)"
        sm = self.stereo_matching
        w, h = self.xy
        w1, h1 = self.cam1.xy
        if not isinstance(sm, SemiGlobalBlockMatching):
            raise ValueError("get_depth_batch needs the built-in SemiGlobalBlockMatching")
        self._push_rig()
        batch = getattr(self, "_batch", None)
        if batch is not None and getattr(self, "_batch_max_size", None) != sm._max_size_option():
            batch = None
        self._batch_max_size = sm._max_size_option()
        if batch is None or len(batch) != streams:
            self._drop_batch()
            from .stereo_matching import StereoSGBM
            batch = []
            for _ in range(int(streams)):
                hd = _ffi.Handle(self.device)
                StereoSGBM(handle=hd, **sm.stereo_sgbm.params)
                hd.call("b2s_set_option", 4, sm._max_size_option())
                self._push_rig(hd)
                batch.append(dict(handle=hd, pin={}, pending=None))
            self._batch = batch
        want = int(any(k in keys for k in ("unrectify_depth", "undistort_img1", "distort_depth")))
        results = [None] * len(pairs)

        def collect(slot):
            b = batch[slot]
            if b["pending"] is None:
                return
            b["handle"].sync()
            i, bufs = b["pending"]
            results[i] = bufs if out is not None else {k: v.copy() for k, v in bufs.items()}
            if "distort_depth" in keys:
                results[i]["distort_img1"] = pairs[i][0]
            b["pending"] = None

        for i, (img1, img2) in enumerate(pairs):
            slot = i % len(batch)
            collect(slot)
            b = batch[slot]
            img1, cn = self._prep(img1)
            img2, cn2 = self._prep(img2)
            if cn != cn2:
                raise ValueError("img1/img2 channel counts differ")
            self._check_raw(img1, img2)
            ishape = (h, w) if img1.ndim == 2 else (h, w, cn)
            spec = dict(rectify_img1=(ishape, np.uint8), rectify_img2=(ishape, np.uint8), disparity=((h, w), np.float32),
                        rectify_depth=((h, w), np.float64), unrectify_depth=((h1, w1), np.float64), undistort_img1=(img1.shape, np.uint8),
                        distort_depth=((h1, w1), np.float64))
            dout, bufs = _ffi.DepthOut(), {}

            def pinned(name, shape, dtype):
                a = b["pin"].get(name)
                if a is None or a.shape != tuple(shape) or a.dtype != np.dtype(dtype):
                    if a is not None:
                        _ffi.pinned_free(a)
                    a = b["pin"][name] = _ffi.pinned_empty(shape, dtype)
                return a

            for k in keys:
                if k not in spec:
                    raise ValueError("unknown result key %r" % (k,))
                if out is not None:
                    a = out[i][k]
                    if a.shape != tuple(spec[k][0]) or a.dtype != np.dtype(spec[k][1]) or not a.flags.c_contiguous:
                        raise ValueError("out[%d][%r] must be a C-contiguous %s array of shape %s" % (i, k, np.dtype(spec[k][1]), spec[k][0]))
                    bufs[k] = a
                else:
                    bufs[k] = pinned(k, *spec[k])
                setattr(dout, k, bufs[k].ctypes.data)
            in1, in2 = img1, img2
            if img1.ctypes.data not in _ffi._PINNED:
                in1 = pinned("_in1", img1.shape, np.uint8)
                np.copyto(in1, img1)
            if img2.ctypes.data not in _ffi._PINNED:
                in2 = pinned("_in2", img2.shape, np.uint8)
                np.copyto(in2, img2)
            b["handle"].call("b2s_get_depth_async", _ffi.ptr(in1), _ffi.ptr(in2), cn, want, ctypes.byref(dout))
            b["pending"] = (i, bufs)
        for slot in range(len(batch)):
            collect(slot)
        return results


__all__ = ["Stereo", "Cam", "MetaStereoMatching", "SemiGlobalBlockMatching"]
