"""Host logic of the multi-GPU path (calibrating_b200/sharded.py): pair sharding, the single rig broadcast and the depth
all-gather, on CPU with gloo and world_size 2.  The per-rank compute engine is the oracle here (test infrastructure);
the CUDA engine behind the same interface is covered by the gpu-marked test at the bottom (nccl, world_size 1)."""
import os
import socket

import numpy as np
import pytest

import calibrating_b200 as cb
from calibrating_b200 import sharded, synth

RIG_XY = (320, 240)
CFG = {"max_size": 4000, "num_disparities": 64}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleBlockEngine:
    """Runs the reference chain (cv2, as oracle/chain.py does) from nothing but the broadcast rig block."""

    def set_rig_block(self, block, scalars, matcher_cfg, table):
        from oracle import chain
        self.s = scalars
        self.a = {k: np.frombuffer(block, dt, int(np.prod(shape)), off).reshape(shape) for k, (off, shape, dt) in table.items()}
        kw = dict(numDisparities=matcher_cfg.get("num_disparities", 218))
        self.plugin = chain.SgbmPlugin(max_size=matcher_cfg.get("max_size", 1000), **kw)

    def get_depth_into(self, img1, img2, out):
        import cv2
        a, s = self.a, self.s
        r1 = cv2.remap(img1, a["map1x"], a["map1y"], cv2.INTER_LANCZOS4)
        r2 = cv2.remap(img2, a["map2x"], a["map2y"], cv2.INTER_LANCZOS4)
        md = s["min_disparity"]
        if md > 0:
            r2[:, md:] = r2[:, :-md]
            r2[:, :md] = 0
        d = self.plugin(r1, r2) + md
        d = a["valid_mask1"].astype(bool) * d
        with np.errstate(all="ignore"):
            z = np.float64(s["fx_baseline"]) / d  # (np.float64 scalar: float64 result, like the reference expression)
            z[z > s["max_depth"]] = 0
            z[z < 0] = 0
        ys, xs = np.mgrid[:s["H"], :s["W"]]
        m = s["unrect_m"]
        out[...] = cv2.remap(z * (m[0] * xs + m[1] * ys + m[2]), a["unrect_mapx"], a["unrect_mapy"], cv2.INTER_NEAREST)


def _expected(rig, n):
    from oracle import chain
    ref = chain.RefStereo(rig).set_stereo_matching(chain.SgbmPlugin(max_size=4000, numDisparities=64), max_depth=3.5)
    return np.stack([ref.get_depth(*synth.render_rig(rig, seed=i))["unrectify_depth"] for i in range(n)])


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rig = synth.rig_dict(RIG_XY)
        stereo = None
        if rank == 0:  # only rank 0 builds the rig (host-side map generation); the others receive the block
            stereo = cb.Stereo.load(rig)
            stereo.stereo_matching = cb.MetaStereoMatching(CFG)  # (cfg travels with the block; no GPU on this box)
            stereo.set_stereo_matching(stereo.stereo_matching, max_depth=3.5)
        sh = sharded.ShardedStereo(stereo, engine_factory=OracleBlockEngine)
        mine = sh.shard(n_pairs)
        assert mine == list(range(rank, n_pairs, world))
        got = sh.get_depth_batch([synth.render_rig(rig, seed=i) for i in mine])
        q.put((rank, sh.rig_bytes, got.numpy()))
    finally:
        dist.destroy_process_group()


def test_pack_unpack_and_shards():
    rig = synth.rig_dict(RIG_XY)
    st = cb.Stereo.load(rig)
    st.stereo_matching = cb.MetaStereoMatching(CFG)
    st.set_stereo_matching(st.stereo_matching, max_depth=3.5)
    buf = sharded.pack_rig_block(st, CFG)
    scalars, cfg, table = sharded.unpack_rig_block(buf)
    assert cfg == CFG and scalars["W"] == 320 and scalars["H1"] == 240 and scalars["min_disparity"] == st.min_disparity
    arrays, _ = sharded.rig_arrays(st)
    for k, (off, shape, dt) in table.items():
        assert off % 256 == 0
        assert np.array_equal(np.frombuffer(buf, dt, int(np.prod(shape)), off).reshape(shape), arrays[k]), k
    assert sharded.shard_indices(10, 1, 4) == [1, 5, 9] and sharded.shard_indices(3, 3, 4) == []
    assert sorted(sum((sharded.shard_indices(64, r, 8) for r in range(8)), [])) == list(range(64))
    with pytest.raises(ValueError):
        sharded.unpack_rig_block(np.zeros(64, np.uint8))
    # a maps="device" rig travels as parameters only: ~1 KB instead of the map planes
    st.maps = "device"
    small = sharded.pack_rig_block(st, CFG)
    sc, _, tb = sharded.unpack_rig_block(small)
    assert small.size < 16384 < buf.size and not tb and set(sc["map_params"]) == {"rect1", "rect2", "unrect", "undist"}
    rp = sharded.rig_params_struct(sc)
    assert (rp.W, rp.H, rp.rect1.W, rp.undist.H) == (320, 240, 320, 240) and rp.rect1.fx == st.cam1.K[0, 0]


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    world, n_pairs = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    exp = _expected(synth.rig_dict(RIG_XY), n_pairs)
    for rank, rig_bytes, got in res:
        assert rig_bytes > 4 * 4 * 320 * 240
        assert got.shape == exp.shape and got.dtype == np.float64
        assert np.allclose(got, exp, rtol=1e-9, atol=0), "rank %d: gathered depth differs from the single-process chain" % rank
        assert (got > 0).mean() > 0.3


@pytest.mark.gpu
def test_cuda_engine_nccl_world1():
    """The CUDA engine behind ShardedStereo: rig block broadcast into device memory and handed over as device pointers,
    depth written straight into the gather tensor; same numbers as Stereo.get_depth."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        rig = synth.rig_dict(RIG_XY)
        st = cb.Stereo.load(rig).set_stereo_matching(cb.SemiGlobalBlockMatching(dict(CFG)), max_depth=3.5)
        sh = sharded.ShardedStereo(st)
        pairs = [synth.render_rig(rig, seed=i) for i in range(3)]
        got = sh.get_depth_batch(pairs).cpu().numpy()
        for i, (a, b) in enumerate(pairs):
            assert np.array_equal(got[i], st.get_depth(a, b)["unrectify_depth"])
        st2 = cb.Stereo.load(rig, maps="device").set_stereo_matching(cb.SemiGlobalBlockMatching(dict(CFG)), max_depth=3.5)
        sh2 = sharded.ShardedStereo(st2)  # parameters only in the broadcast block
        assert sh2.rig_bytes < 16384 < sh.rig_bytes
        assert np.array_equal(sh2.get_depth_batch(pairs).cpu().numpy(), got)
        exp = _expected(rig, 3)
        assert np.allclose(got, exp, rtol=1e-9, atol=0)
    finally:
        dist.destroy_process_group()


def _nccl_worker(rank, world, port, n_pairs, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rig = synth.rig_dict(RIG_XY)
        st = None
        if rank == 0:
            st = cb.Stereo.load(rig, device=0).set_stereo_matching(cb.SemiGlobalBlockMatching(dict(CFG), device=0), max_depth=3.5)
        sh = sharded.ShardedStereo(st)
        mine = sh.shard(n_pairs)
        got = sh.get_depth_batch([synth.render_rig(rig, seed=i) for i in mine])
        bad = None
        try:
            sh.engine.get_depth_batch_into([(np.zeros((10, 10, 3), np.uint8),) * 2], [got[0]])
        except ValueError as e:
            bad = str(e)
        q.put((rank, sh.rig_bytes, got.cpu().numpy(), bad))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_cuda_engine_nccl_multi_rank():
    """BASELINE config 3 in small: ShardedStereo on every GPU of the box (>= 2), one process per GPU over NCCL: rig block
    broadcast from rank 0 into device memory, pairs i % world per rank on several streams, depth all-gathered into global pair
    order on every rank; equal to the single-process chain.  Skipped on a one-GPU box (gpurun --gpus N runs it)."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 CUDA devices")
    n_pairs = 2 * world
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    exp = _expected(synth.rig_dict(RIG_XY), n_pairs)
    for rank, rig_bytes, got, bad in res:
        assert got.shape == exp.shape and got.dtype == np.float64
        assert np.allclose(got, exp, rtol=1e-9, atol=0), "rank %d: gathered depth differs from the single-process chain" % rank
        assert np.array_equal(got, res[0][2]), "ranks disagree"
        assert bad and "do not match the rig" in bad, "a wrong-sized image must be rejected before the FFI call"
