"""Pair-batch sharding across GPUs: one process per GPU (`torchrun`), `torch.distributed` for the plumbing.

The reference has no multi-device path (one `Stereo.get_depth` per call, calibrating/stereo_camera.py:492-533).  Image
pairs are independent, so the path shards by pair with no data-path collective; exactly two
collectives exist, both outside the per-pair kernels:

  * ONE broadcast of the per-rig constant block from rank 0 -- the four rectification map planes, the two unrectify map
    planes, the valid mask, the undistort maps and the scalars (`Stereo._get_undistort_rectify_map`,
    stereo_camera.py:125-177; ~100 MB at 1080p) -- so that only rank 0 runs the host-side map generation.  With the
    NCCL backend the block is broadcast GPU-to-GPU over NVLink and handed to the engine as device pointers
    (`b2s_set_rig` accepts them), with gloo (CPU tests) it travels as a host tensor.
  * ONE all-gather of the per-pair results (`unrectify_depth`, float64 (H1,W1) per pair) so that every rank holds the
    batch result in global pair order.

Pair i of a global batch belongs to rank i % world_size (`shard_indices`).  The compute engine behind a rank is pluggable
(`engine_factory`) so that the host logic is testable on CPU with the oracle (tests/test_sharded.py, gloo, world_size 2);
the default engine is the CUDA one and fails loudly without a GPU.
"""
import ctypes
import json

import numpy as np

from . import _ffi

_MAGIC = b"B2SRIG01"


def shard_indices(n, rank, world):
    """Global pair indices handled by `rank`: i % world == rank (static round-robin)."""
    return list(range(rank, n, world))


# ---- the rig block: everything a rank needs to run get_depth, in one contiguous byte buffer ------------------------
_ARRAYS = [("map1x", np.float32), ("map1y", np.float32), ("map2x", np.float32), ("map2y", np.float32), ("valid_mask1", np.uint8),
           ("unrect_mapx", np.float32), ("unrect_mapy", np.float32), ("undist_xy", np.int16), ("undist_fxy", np.uint16)]


def rig_arrays(stereo):
    """The per-rig constant arrays and scalars of `stereo` (host side; what `Stereo._push_rig` uploads)."""
    import cv2
    m1x, m1y = stereo.undistort_rectify_map1
    m2x, m2y = stereo.undistort_rectify_map2
    umx, umy = stereo._unrectify_maps()
    w1, h1 = stereo.cam1.xy
    und_xy, und_fxy = cv2.initUndistortRectifyMap(stereo.cam1.K, stereo.cam1.D, None, stereo.cam1.K, (w1, h1), cv2.CV_16SC2)
    M = stereo.R1.T @ np.linalg.inv(stereo.K)
    arrays = dict(map1x=m1x, map1y=m1y, map2x=m2x, map2y=m2y, valid_mask1=stereo.rectify_valid_mask1, unrect_mapx=umx, unrect_mapy=umy,
                  undist_xy=und_xy, undist_fxy=und_fxy)
    arrays = {k: np.ascontiguousarray(arrays[k], dt) for k, dt in _ARRAYS}
    scalars = dict(W=int(stereo.xy[0]), H=int(stereo.xy[1]), W1=int(w1), H1=int(h1), W2=int(stereo.cam2.xy[0]), H2=int(stereo.cam2.xy[1]),
                   unrect_m=[float(v) for v in M[2]], fx_baseline=float(1.0 * stereo.baseline * stereo.K[0, 0]),
                   max_depth=float(stereo.get_max_depth()),
                   min_disparity=int(stereo.min_disparity) if getattr(stereo, "translation_rectify_img", None) else 0,
                   interp={"lanczos4": 0, "linear": 1}[stereo.interp])
    return arrays, scalars


def rig_params(stereo):
    """The inputs of the four cv2.initUndistortRectifyMap calls as plain lists (for `Stereo(maps="device")` rigs: the ranks
    regenerate the maps on their GPU with b2s_set_rig_params, so the broadcast block carries ~1 KB instead of ~100 MB)."""
    w1, h1 = stereo.cam1.xy
    specs = dict(rect1=(stereo.cam1.K, stereo.cam1.D, stereo.R1, stereo.K, stereo.xy), rect2=(stereo.cam2.K, stereo.cam2.D, stereo.R2, stereo.K, stereo.xy),
                 unrect=(stereo.K, None, stereo.R1.T, stereo.cam1.K, (w1, h1)), undist=(stereo.cam1.K, stereo.cam1.D, None, stereo.cam1.K, (w1, h1)))
    out = {}
    for name, (K, D, R, Knew, size) in specs.items():
        m = stereo._map_params(K, D, R, Knew, size)
        out[name] = dict(W=m.W, H=m.H, fx=m.fx, fy=m.fy, cx=m.cx, cy=m.cy, k=list(m.k), iR=list(m.iR))
    return out


def pack_rig_block(stereo, matcher_cfg=None):
    """-> uint8 array: magic | u64 header length | JSON header (scalars, matcher cfg, array offsets) | 256-aligned arrays.
    For a `maps="device"` rig the header carries the map parameters instead and there are no arrays."""
    arrays, scalars = rig_arrays(stereo)
    params = rig_params(stereo) if getattr(stereo, "maps", "host") == "device" else None
    off, table = 0, {}
    for name, dt in ([] if params else _ARRAYS):
        a = arrays[name]
        table[name] = dict(offset=off, shape=list(a.shape), dtype=np.dtype(dt).str)
        off += (a.nbytes + 255) // 256 * 256
    header = json.dumps(dict(scalars=dict(scalars, map_params=params), matcher=matcher_cfg or {}, arrays=table)).encode()
    head_len = (len(_MAGIC) + 8 + len(header) + 255) // 256 * 256
    buf = np.zeros(head_len + off, np.uint8)
    buf[:8] = np.frombuffer(_MAGIC, np.uint8)
    buf[8:16] = np.frombuffer(np.uint64(len(header)).tobytes(), np.uint8)
    buf[16:16 + len(header)] = np.frombuffer(header, np.uint8)
    for name in table:
        a = arrays[name]
        o = head_len + table[name]["offset"]
        buf[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
    return buf


def unpack_rig_block(buf):
    """-> (scalars, matcher_cfg, {name: (byte offset into buf, shape, dtype)}); `buf` is the host copy of the header at least."""
    head = np.asarray(buf[:16]).tobytes()
    if head[:8] != _MAGIC:
        raise ValueError("not a rig block")
    n = int(np.frombuffer(head[8:16], np.uint64)[0])
    meta = json.loads(np.asarray(buf[16:16 + n]).tobytes().decode())
    head_len = (16 + n + 255) // 256 * 256
    table = {k: (head_len + v["offset"], tuple(v["shape"]), np.dtype(v["dtype"])) for k, v in meta["arrays"].items()}
    return meta["scalars"], meta["matcher"], table


def rig_params_struct(scalars):
    """ctypes `b2s_rig_params` from the scalars of a `maps="device"` rig block."""
    rp = _ffi.RigParams()
    for k in ("W", "H", "W1", "H1", "W2", "H2", "min_disparity", "interp"):
        setattr(rp, k, int(scalars[k]))
    rp.unrect_m = (ctypes.c_double * 3)(*scalars["unrect_m"])
    rp.fx_baseline, rp.max_depth = float(scalars["fx_baseline"]), float(scalars["max_depth"])
    for name, m in scalars["map_params"].items():
        setattr(rp, name, _ffi.MapParams(int(m["W"]), int(m["H"]), m["fx"], m["fy"], m["cx"], m["cy"], (ctypes.c_double * 12)(*m["k"]),
                                         (ctypes.c_double * 9)(*m["iR"])))
    return rp


def rig_struct(scalars, table, base_address):
    """ctypes `b2s_rig` whose array pointers are base_address + offset (host or device memory alike)."""
    rig = _ffi.Rig()
    for k in ("W", "H", "W1", "H1", "W2", "H2", "min_disparity", "interp"):
        setattr(rig, k, int(scalars[k]))
    rig.unrect_m = (ctypes.c_double * 3)(*scalars["unrect_m"])
    rig.fx_baseline, rig.max_depth = float(scalars["fx_baseline"]), float(scalars["max_depth"])
    for name, _ in _ARRAYS:
        setattr(rig, name, base_address + table[name][0])
    return rig


def check_pair(img1, img2, scalars):
    """The checks of `Stereo._prep` / `Stereo._check_raw` against the broadcast rig: C-contiguous uint8 (h,w) or (h,w,3) images of
    cam1's / cam2's size with equal channel counts.  Returns (img1, img2, cn).  A wrong image would be an out-of-bounds host
    read inside the engine, so every rank checks its own pairs before the FFI call."""
    img1, img2 = np.ascontiguousarray(img1), np.ascontiguousarray(img2)
    for im in (img1, img2):
        if im.dtype != np.uint8 or im.ndim not in (2, 3) or (im.ndim == 3 and im.shape[2] != 3):
            raise ValueError("images must be uint8 (h,w) or (h,w,3), got %s %s" % (im.dtype, im.shape))
    if img1.ndim != img2.ndim:
        raise ValueError("img1/img2 channel counts differ")
    if img1.shape[:2] != (scalars["H1"], scalars["W1"]) or img2.shape[:2] != (scalars["H2"], scalars["W2"]):
        raise ValueError("image sizes %s/%s do not match the rig's cam1/cam2 sizes %s/%s" % (
            img1.shape, img2.shape, (scalars["H1"], scalars["W1"]), (scalars["H2"], scalars["W2"])))
    return img1, img2, (1 if img1.ndim == 2 else 3)


class CudaEngine:
    """Default per-rank engine: libb2s.so on `device`, `streams` engine handles (= CUDA streams) in flight like
    `Stereo.get_depth_batch`.  The rig block stays where the broadcast left it (a torch CUDA tensor); results are written
    straight into the caller's device tensor.  The matcher honours the cfg's `max_size` (the reference's default 1000
    included): the down-scale runs on the device (B2S_OPT_MAX_SIZE), so a sharded batch equals single-GPU `get_depth`."""

    def __init__(self, device, streams=3):
        self.handles = [_ffi.Handle(device) for _ in range(int(streams))]
        self.handle = self.handles[0]
        self.device = device
        self._block = None
        self._pin = [dict() for _ in self.handles]

    def set_rig_block(self, block, scalars, matcher_cfg, table):
        from .stereo_matching import SemiGlobalBlockMatching
        self._block = block  # keep the storage alive
        self.matchers = []
        for h in self.handles:
            if scalars.get("map_params"):  # maps="device": the rank generates the maps itself from ~1 KB of parameters
                h.call("b2s_set_rig_params", ctypes.byref(rig_params_struct(scalars)))
            else:
                base = block.data_ptr() if hasattr(block, "data_ptr") else block.ctypes.data
                h.call("b2s_set_rig", ctypes.byref(rig_struct(scalars, table, base)))
            sm = SemiGlobalBlockMatching(dict(matcher_cfg), handle=h)
            h.call("b2s_set_option", 4, sm._max_size_option())
            self.matchers.append(sm)
        self.matcher = self.matchers[0]
        self.scalars = scalars

    def _pinned(self, slot, key, shape):
        a = self._pin[slot].get(key)
        if a is None or a.shape != tuple(shape):
            if a is not None:
                _ffi.pinned_free(a)
            a = self._pin[slot][key] = _ffi.pinned_empty(shape, np.uint8)
        return a

    def get_depth_batch_into(self, pairs, outs):
        """pairs: [(img1, img2)] host uint8 arrays; outs: one (H1,W1) float64 destination per pair (torch CUDA tensors or host
        arrays).  Upload, kernels and the copy into `outs` of different pairs overlap across the handles' streams."""
        S = len(self.handles)
        for i, ((a, b), out) in enumerate(zip(pairs, outs)):
            slot = i % S
            h = self.handles[slot]
            a, b, cn = check_pair(a, b, self.scalars)
            if i >= S:
                h.sync()  # the slot's staging buffers are still being uploaded by its previous call
            if a.ctypes.data not in _ffi._PINNED:
                buf = self._pinned(slot, "a", a.shape)
                np.copyto(buf, a)
                a = buf
            if b.ctypes.data not in _ffi._PINNED:
                buf = self._pinned(slot, "b", b.shape)
                np.copyto(buf, b)
                b = buf
            o = _ffi.DepthOut()
            o.unrectify_depth = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
            h.call("b2s_get_depth_async", _ffi.ptr(a), _ffi.ptr(b), cn, 1, ctypes.byref(o))
        for h in self.handles:
            h.sync()

    def get_depth_into(self, img1, img2, out):
        """One pair: img1/img2 host uint8 arrays; out: (H1,W1) float64 torch CUDA tensor (or host array) for unrectify_depth."""
        self.get_depth_batch_into([(img1, img2)], [out])

    def close(self):
        for pins in self._pin:
            for a in pins.values():
                _ffi.pinned_free(a)
            pins.clear()
        for h in self.handles:
            h.close()


class PendingBatch:
    """A batch whose depth all-gather is still in flight (`ShardedStereo.get_depth_batch(..., wait=False)`): the collective runs
    on NCCL's stream while the caller already computes the next batch.  `result()` waits and returns the gathered tensor."""

    def __init__(self, gathered, work, world, n, H1, W1):
        self.gathered, self.work, self.shape = gathered, work, (world, n, H1, W1)

    def result(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        world, n, H1, W1 = self.shape
        return self.gathered.transpose(0, 1).reshape(world * n, H1, W1)  # (rank, k) -> global index k*world + rank


class ShardedStereo:
    """Rank-local front end of a pair batch sharded over `torch.distributed` ranks (module docstring)."""

    def __init__(self, stereo=None, engine_factory=None, group=None, device=None):
        """stereo: a configured `Stereo` (with `set_stereo_matching(SemiGlobalBlockMatching(...))`) on rank 0, None elsewhere."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.cuda = dist.get_backend(group) == "nccl"
        self.device = device if device is not None else (torch.cuda.current_device() if self.cuda else None)
        self.engine = (engine_factory or (lambda: CudaEngine(self.device)))()
        self._setup(stereo)

    def _setup(self, stereo):
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        if self.rank == 0:
            if stereo is None:
                raise ValueError("rank 0 must pass the configured Stereo")
            sm = getattr(stereo, "stereo_matching", None)
            host = pack_rig_block(stereo, dict(getattr(sm, "cfg", {}) or {}))
            size = [int(host.size)]
        else:
            host, size = None, [0]
        dist.broadcast_object_list(size, src=0, group=self.group)  # control plane: the block size (one int)
        block = torch.empty(size[0], dtype=torch.uint8, device=dev)
        if self.rank == 0:
            block.copy_(torch.from_numpy(host))
        dist.broadcast(block, src=0, group=self.group)  # THE rig broadcast (NVLink with nccl)
        head = block[:1 << 16].cpu().numpy() if block.numel() > (1 << 16) else block.cpu().numpy()
        self.scalars, self.matcher_cfg, self.table = unpack_rig_block(head)
        self.block = block
        self.rig_bytes = int(size[0])
        self.engine.set_rig_block(block if self.cuda else block.numpy(), self.scalars, self.matcher_cfg, self.table)

    def shard(self, n):
        return shard_indices(n, self.rank, self.world)

    def get_depth_batch(self, local_pairs, wait=True):
        """local_pairs: this rank's [(img1, img2)] (pair k of this rank is global pair k*world + rank).
        Returns the gathered `unrectify_depth` of the whole batch, (world * n_local, H1, W1) float64 in global pair order,
        on every rank (device tensor with nccl, host tensor with gloo).  All ranks must pass the same number of pairs.
        wait=False returns a `PendingBatch` right after the all-gather was enqueued, so that the next batch's kernels overlap it."""
        torch, dist = self.torch, self.dist
        n, H1, W1 = len(local_pairs), self.scalars["H1"], self.scalars["W1"]
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        # (torch.empty launches nothing, so the engine's own streams need no ordering against torch's; every pixel is written)
        local = torch.empty((n, H1, W1), dtype=torch.float64, device=dev)
        outs = [local[k] if self.cuda else local[k].numpy() for k in range(n)]
        if hasattr(self.engine, "get_depth_batch_into"):
            self.engine.get_depth_batch_into(local_pairs, outs)  # returns after the engine's streams are idle
        else:
            for (a, b), o in zip(local_pairs, outs):
                a, b, _ = check_pair(a, b, self.scalars)
                self.engine.get_depth_into(a, b, o)
        gathered = torch.empty((self.world, n, H1, W1), dtype=torch.float64, device=dev)
        work = dist.all_gather_into_tensor(gathered.view(self.world * n, H1, W1), local, group=self.group, async_op=not wait)  # THE depth all-gather
        pending = PendingBatch(gathered, None if wait else work, self.world, n, H1, W1)
        return pending.result() if wait else pending
