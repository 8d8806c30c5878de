"""Batch front-end: several image pairs in flight on several CUDA streams of one GPU (BASELINE config 3: "one pair
per stream").  The reference has no batching (one `get_depth` per call); pairs are independent,
so a batch is just the same per-pair path issued asynchronously through the *_async entry points of include/b2s.h."""
import ctypes

import numpy as np

from . import _ffi
from .stereo_matching import REFERENCE_DEFAULTS, StereoSGBM


def _optr(o):
    """Address of an output buffer: a numpy array (host, ideally pinned) or anything with data_ptr() (a torch CUDA tensor:
    the result then stays on the device, e.g. as the input of an NCCL all-gather)."""
    return ctypes.c_void_p(o.data_ptr()) if hasattr(o, "data_ptr") else _ffi.ptr(o)


class DisparityBatchEngine:
    """`streams` independent engine handles (stream + buffers) on one device, fed round-robin."""

    def __init__(self, params=None, device=0, streams=2):
        p = dict(REFERENCE_DEFAULTS)
        p.update(params or {})
        self.matchers = [StereoSGBM(device=device, **p) for _ in range(int(streams))]
        self.handles = [m.handle for m in self.matchers]
        self.device = device
        self._pin = {}

    def _pinned(self, key, shape, dtype):
        a = self._pin.get(key)
        if a is None or a.shape != tuple(shape) or a.dtype != np.dtype(dtype):
            if a is not None:
                _ffi.pinned_free(a)
            a = self._pin[key] = _ffi.pinned_empty(shape, dtype)
        return a

    def close(self):
        for a in self._pin.values():
            _ffi.pinned_free(a)
        self._pin = {}
        for h in self.handles:
            h.close()

    def launch_count(self):
        return sum(h.launch_count() for h in self.handles)

    def compute_batch(self, pairs, out=None, as_float=True):
        """pairs: list of (left, right) uint8 host arrays (pinned arrays from `_ffi.pinned_empty` are used in place,
        others are staged; `out`: one destination per pair, numpy arrays or torch CUDA tensors).  Returns a list of (H,W) float32 disparities (reference post-processing applied) or int16
        16*disparity when as_float=False.  H2D copy, kernels and D2H copy of different pairs overlap across streams."""
        n, S = len(pairs), len(self.handles)
        results = [None] * n
        busy = [False] * S  # handle has calls in flight that use its staging buffers
        copy_back = []
        for i in range(n):
            slot = i % S
            h = self.handles[slot]
            left, right, H, W, cn = StereoSGBM._check_pair(*pairs[i])
            staged = left.ctypes.data not in _ffi._PINNED or right.ctypes.data not in _ffi._PINNED or out is None
            if staged and busy[slot]:
                h.sync()  # the previous call of this handle still reads / writes the per-slot staging buffers
                busy[slot] = False
                for j, o in copy_back:
                    if j % S == slot and results[j] is o:
                        results[j] = o.copy()
            if left.ctypes.data not in _ffi._PINNED:
                buf = self._pinned(("l", slot), left.shape, np.uint8)
                np.copyto(buf, left)
                left = buf
            if right.ctypes.data not in _ffi._PINNED:
                buf = self._pinned(("r", slot), right.shape, np.uint8)
                np.copyto(buf, right)
                right = buf
            if out is not None:
                o = out[i]
            else:
                o = self._pinned(("o", slot), (H, W), np.float32 if as_float else np.int16)
                copy_back.append((i, o))
            # calls on one handle are stream-ordered (upload, kernels, download), so pinned caller buffers need no sync in between
            if as_float:
                h.call("b2s_compute_disparity_async", _ffi.ptr(left), _ffi.ptr(right), H, W, cn, None, _optr(o))
            else:
                h.call("b2s_compute_disparity_async", _ffi.ptr(left), _ffi.ptr(right), H, W, cn, _optr(o), None)
            results[i] = o
            busy[slot] = busy[slot] or staged
        for h in self.handles:
            h.sync()
        for j, o in copy_back:
            if results[j] is o:
                results[j] = o.copy()
        return results

    def compute_batch_dev(self, dev_pairs, H, W, cn, dev_out16):
        """Device-resident variant: dev_pairs = [(left_ptr, right_ptr)], dev_out16 = [int16 out ptr]; raw device
        addresses (e.g. torch.Tensor.data_ptr()).  Enqueues everything, then waits for all streams."""
        S = len(self.handles)
        for i, ((lp, rp), op) in enumerate(zip(dev_pairs, dev_out16)):
            self.handles[i % S].call("b2s_compute_disparity_dev", ctypes.c_void_p(lp), ctypes.c_void_p(rp), H, W, cn,
                                     ctypes.c_void_p(op), None)
        for h in self.handles:
            h.sync()
