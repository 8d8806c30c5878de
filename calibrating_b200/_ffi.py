"""ctypes binding of libb2s.so (include/b2s.h).  This is the stub a maintainer of the reference would add to call
the engine.  There is no CPU fallback: a missing library or a missing GPU raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2s.so")

c_int, c_void_p, c_size_t, c_double, c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_double, ctypes.c_float


class B2SError(RuntimeError):
    pass


class SgbmParams(ctypes.Structure):
    _fields_ = [(n, c_int) for n in (
        "min_disparity", "num_disparities", "block_size", "P1", "P2", "disp12_max_diff", "pre_filter_cap",
        "uniqueness_ratio", "speckle_window_size", "speckle_range", "mode", "cost")]


class Rig(ctypes.Structure):
    _fields_ = [("W", c_int), ("H", c_int), ("W1", c_int), ("H1", c_int), ("W2", c_int), ("H2", c_int),
                ("map1x", c_void_p), ("map1y", c_void_p), ("map2x", c_void_p), ("map2y", c_void_p),
                ("valid_mask1", c_void_p), ("unrect_mapx", c_void_p), ("unrect_mapy", c_void_p),
                ("undist_xy", c_void_p), ("undist_fxy", c_void_p),
                ("unrect_m", c_double * 3), ("fx_baseline", c_double), ("max_depth", c_double),
                ("min_disparity", c_int), ("interp", c_int)]


class MapParams(ctypes.Structure):
    _fields_ = [("W", c_int), ("H", c_int), ("fx", c_double), ("fy", c_double), ("cx", c_double), ("cy", c_double),
                ("k", c_double * 12), ("iR", c_double * 9)]


class RigParams(ctypes.Structure):
    _fields_ = [("W", c_int), ("H", c_int), ("W1", c_int), ("H1", c_int), ("W2", c_int), ("H2", c_int),
                ("rect1", MapParams), ("rect2", MapParams), ("unrect", MapParams), ("undist", MapParams),
                ("unrect_m", c_double * 3), ("fx_baseline", c_double), ("max_depth", c_double),
                ("min_disparity", c_int), ("interp", c_int)]


class DepthOut(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("rectify_img1", "rectify_img2", "disparity", "rectify_depth",
                                        "unrectify_depth", "undistort_img1", "disp16", "distort_depth")]


class Timing(ctypes.Structure):
    _fields_ = [(n, c_float) for n in ("rectify_ms", "cost_ms", "aggregate_ms", "wta_ms", "post_ms", "depth_ms", "total_ms")] + \
               [("aggregate_launches", c_int), ("total_launches", c_int)]


# every symbol include/b2s.h declares, with its argument types (tests check the export list against the header)
SIGNATURES = {
    "b2s_device_count": (c_int, []),
    "b2s_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "b2s_destroy": (c_int, [c_void_p]),
    "b2s_last_error": (ctypes.c_char_p, [c_void_p]),
    "b2s_sync": (c_int, [c_void_p]),
    "b2s_host_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p)]),
    "b2s_host_free": (c_int, [c_void_p]),
    "b2s_set_sgbm_params": (c_int, [c_void_p, ctypes.POINTER(SgbmParams)]),
    "b2s_compute_disparity": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b2s_compute_disparity_async": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b2s_compute_disparity_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b2s_set_rig": (c_int, [c_void_p, ctypes.POINTER(Rig)]),
    "b2s_set_rig_params": (c_int, [c_void_p, ctypes.POINTER(RigParams)]),
    "b2s_rectify": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b2s_get_depth": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(DepthOut)]),
    "b2s_get_depth_async": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(DepthOut)]),
    "b2s_depth_from_disparity": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(DepthOut)]),
    "b2s_disparity_to_depth": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b2s_unrectify_depth": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b2s_undistort_img": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "b2s_set_cam1_model": (c_int, [c_void_p, c_double, c_double, c_double, c_double, ctypes.POINTER(c_double)]),
    "b2s_distort_depth": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b2s_project_depth": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, ctypes.POINTER(c_double), ctypes.POINTER(c_double),
                                  ctypes.POINTER(c_double), c_int, c_int, c_void_p]),
    "b2s_set_option": (c_int, [c_void_p, c_int, c_int]),
    "b2s_build_hash": (ctypes.c_char_p, []),
    "b2s_resize_nearest_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float]),
    "b2s_interpolate_rbf": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b2s_interpolate_sparse": (c_int, [c_void_p, c_int, c_void_p, c_int, ctypes.POINTER(c_double), c_void_p, c_int, c_int, c_double, c_void_p]),
    "b2s_depth_to_point_cloud": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, ctypes.POINTER(c_double), c_int, c_void_p, ctypes.c_ulonglong,
                                         ctypes.POINTER(ctypes.c_ulonglong)]),
    "b2s_point_cloud_to_depth": (c_int, [c_void_p, c_void_p, ctypes.c_ulonglong, ctypes.POINTER(c_double), c_int, c_int, c_double, c_void_p]),
    "b2s_volume_dims": (c_int, [c_void_p] + [ctypes.POINTER(c_int)] * 4),
    "b2s_debug_fetch": (c_int, [c_void_p, c_int, c_void_p, c_size_t]),
    "b2s_timings": (c_int, [c_void_p, ctypes.POINTER(Timing)]),
    "b2s_launch_count": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_longlong)]),
    "b2s_bench_aggregate": (c_int, [c_void_p, c_int, ctypes.POINTER(c_float)]),
    "b2s_bench_aggregate_parts": (c_int, [c_void_p, c_int, ctypes.POINTER(c_float), c_int, ctypes.POINTER(c_int)]),
    "b2s_enqueue_aggregate": (c_int, [c_void_p, c_int]),
    "b2s_event_record": (c_int, [c_void_p, c_int]),
    "b2s_event_elapsed": (c_int, [c_void_p, c_int, c_void_p, c_int, ctypes.POINTER(c_float)]),
    "b2s_collect_timings": (c_int, [c_void_p, c_int]),
}

_lib = None


def lib():
    """Load libb2s.so (built in-tree by `python -m calibrating_b200.build`).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2SError("libb2s.so is not built (%s); run `python -m calibrating_b200.build` -- "
                           "there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def check(rc, handle=None):
    if rc == 0:
        return
    msg = lib().b2s_last_error(handle)
    msg = msg.decode() if msg else "error %d" % rc
    if rc in (-1, -2):
        raise ValueError(msg)
    raise B2SError("b2s error %d: %s" % (rc, msg))


class Handle:
    """One engine instance = one CUDA stream + its device buffers on one GPU."""

    def __init__(self, device=0):
        self._h = c_void_p()
        self._lib = lib()
        rc = self._lib.b2s_create(int(device), ctypes.byref(self._h))
        if rc != 0:
            self._h = c_void_p()
            check(rc, None)
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.b2s_destroy(self._h)
            self._h = c_void_p()

    __del__ = close

    def call(self, name, *args):
        check(getattr(self._lib, name)(self._h, *args), self._h)

    def sync(self):
        self.call("b2s_sync")

    def timings(self):
        t = Timing()
        self.call("b2s_timings", ctypes.byref(t))
        return {n: getattr(t, n) for n, _ in Timing._fields_}

    def launch_count(self):
        n = ctypes.c_longlong()
        self.call("b2s_launch_count", ctypes.byref(n))
        return n.value

    def keep_volumes(self, on=True):
        """Make the last aggregation pass store S as well (it fuses the winner-take-all by default), so that fetch_volume(1) works."""
        self.call("b2s_set_option", 1, int(bool(on)))

    def fuse_wta(self, on=True):
        """Winner-take-all inside the last aggregation pass (default) or as a separate kernel (same results; see include/b2s.h)."""
        self.call("b2s_set_option", 2, int(bool(on)))

    def agg_schedule(self, wave=False):
        """B2S_OPT_AGG_SCHEDULE: False = scans + lock-step sweep (default), True = wavefront sweeps (MODE_HH)."""
        self.call("b2s_set_option", 3, int(bool(wave)))

    def volume_dims(self):
        v = [c_int() for _ in range(4)]
        self.call("b2s_volume_dims", *[ctypes.byref(x) for x in v])
        return tuple(x.value for x in v)

    def fetch_volume(self, which):
        """which: 0 = cost volume C, 1 = aggregated volume S; returns (H, width1, D) int16."""
        H, width1, D, Dp = self.volume_dims()
        out = np.empty((H, width1, Dp), np.int16)
        self.call("b2s_debug_fetch", int(which), ptr(out), out.nbytes)
        return out[..., :D]

    def fetch_raw(self, H, W):
        """(H, W) int16 disparity before the median / speckle filters."""
        out = np.empty((H, W), np.int16)
        self.call("b2s_debug_fetch", 2, ptr(out), out.nbytes)
        return out

    RIG_ARRAYS = ["map1x", "map1y", "map2x", "map2y", "valid_mask1", "unrect_mapx", "unrect_mapy", "undist_xy", "undist_fxy"]

    def fetch_rig(self, name, shape, dtype):
        """Device copy of a rig array (as uploaded by b2s_set_rig or generated by b2s_set_rig_params)."""
        out = np.empty(shape, dtype)
        self.call("b2s_debug_fetch", 16 + self.RIG_ARRAYS.index(name), ptr(out), out.nbytes)
        return out

    def event_record(self, slot):
        self.call("b2s_event_record", int(slot))

    def event_elapsed(self, slot_a, other, slot_b):
        """ms from this handle's event slot_a to `other`'s event slot_b (waits for the latter)."""
        ms = c_float()
        check(self._lib.b2s_event_elapsed(self._h, int(slot_a), other._h, int(slot_b), ctypes.byref(ms)), other._h)
        return ms.value

    def bench_aggregate(self, iters=10):
        ms = c_float()
        self.call("b2s_bench_aggregate", int(iters), ctypes.byref(ms))
        return ms.value

    def enqueue_aggregate(self, iters=1):
        self.call("b2s_enqueue_aggregate", int(iters))

    def bench_aggregate_parts(self, iters=10):
        """mean ms of every kernel launch of the aggregation group, in launch order"""
        parts = (c_float * 8)()
        n = c_int()
        self.call("b2s_bench_aggregate_parts", int(iters), parts, 8, ctypes.byref(n))
        return [parts[i] for i in range(n.value)]


def pinned_empty(shape, dtype):
    """numpy array backed by cudaHostAlloc memory (for the *_async entry points)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = c_void_p()
    check(lib().b2s_host_alloc(max(n, 1), ctypes.byref(p)))
    buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p is not None:
        lib().b2s_host_free(p)
