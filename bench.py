#!/usr/bin/env python3
"""bench.py -- stereo pairs/s of the `Stereo.get_depth` matcher hot path at BASELINE config 2
(1920x1080 RGB rectified pair, 128 disparities, 8-path SGM = cv2 MODE_HH, block 5), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one batch of `--pairs-per-step` independent synthetic pairs per GPU (pairs shard across ranks with no
data-path collective: weak scaling).  `value` = whole-job pairs/s with the inputs already resident in HBM, timed with
CUDA events on the engine's own streams, max over ranks.  `e2e` = the same metric through the public host API
(`DisparityBatchEngine.compute_batch`, i.e. the plugin's batch call) with pinned HOST buffers, H2D and D2H inside the
timed region.  `roofline` = the aggregation kernel group (horizontal +x scan, fused vertical sweep, horizontal -x scan: 3
launches per pair) timed alone by CUDA events, algorithmic bytes 8 B/voxel per pair.  `cpu_baseline` / `--impl reference` = the reference's own CPU arithmetic
(cv2.StereoSGBM, what calibrating/stereo_matching.py:63 executes) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, D, CN = 1080, 1920, 128, 3
SGBM = dict(min_disparity=0, num_disparities=D, block_size=5, P1=8 * 3 * 25, P2=32 * 3 * 25, disp12_max_diff=1, pre_filter_cap=0,
            uniqueness_ratio=5, speckle_window_size=200, speckle_range=2, mode=1)
WORKLOAD = "BASELINE config 2: 1920x1080 RGB synthetic rectified pair, 128 disparities, 8-path SGM (cv2 MODE_HH), block 5"
VOXELS = H * (W - D) * D  # cost-volume extent H*(W-minD-D)*D
AGG_BYTES_PER_PAIR = 8 * VOXELS  # canonical two-sweep schedule: C read twice, S written once and read once, int16


def cv2_matcher():
    import cv2
    return cv2.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1, uniquenessRatio=5,
                                 speckleWindowSize=200, speckleRange=2, mode=cv2.STEREO_SGBM_MODE_HH)


def host_threads():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        avail_gb = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) / 1e6
    except Exception:
        avail_gb = 16
    return max(1, min(n, int(avail_gb / 2.5), 64))  # one MODE_HH matcher holds ~1.3 GB of cost volumes at 1080p/128


def cpu_round(pairs, threads):
    """`threads` cv2 matchers, one pair each, in parallel (cv2 releases the GIL).  Returns seconds."""
    ms = [cv2_matcher() for _ in range(threads)]
    out = [None] * threads

    def work(i):
        l, r = pairs[i % len(pairs)]
        out[i] = ms[i].compute(l, r)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    return time.perf_counter() - t0


class ClockSampler:
    """nvidia-smi polled in the background; samples are time-stamped on arrival so that only those taken while the GPU
    was under load (warm-up + timed region, marked with begin()/end()) are summarised."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p, self.rows, self.t0, self.t1 = None, [], None, None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line))

    def wait_first(self, timeout=10.0):
        t = time.time()
        while self.p and not self.rows and time.time() - t < timeout:
            time.sleep(0.02)

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in list(self.rows):
            if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or ts) + 0.05):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower() == "active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank):
    """The reference's own CPU path (cv2.StereoSGBM MODE_HH) with all usable host threads; rank 0 only."""
    if rank != 0:
        return
    import cv2
    from calibrating_b200 import synth
    T = host_threads()
    pairs = [synth.rectified_pair(H, W, D, seed=s)[:2] for s in range(min(T, 4))]
    for _ in range(args.warmup):
        cpu_round(pairs, T)
    t = sum(cpu_round(pairs, T) for _ in range(args.steps))
    v = T * args.steps / t
    sample = "%d threads x 1 pair of the workload per step, cv2 %s, getNumThreads=%d" % (T, cv2.__version__, cv2.getNumThreads())
    print(json.dumps({
        "impl": "reference", "metric": "stereo pairs/sec @1080p/128-disp SGM", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": T, "host": "cv2.StereoSGBM MODE_HH on host cores"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": T, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=16)
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="skip the full-chain (config 5) throughput measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist
    from calibrating_b200 import _ffi, synth
    from calibrating_b200.batch import DisparityBatchEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    P, S, K, Wm = args.pairs_per_step, args.streams, args.steps, max(args.warmup, 3)
    pairs = [synth.rectified_pair(H, W, D, seed=rank * P + i)[:2] for i in range(P)]
    eng = DisparityBatchEngine(SGBM, device=local, streams=S)

    # ---- value: inputs resident in HBM -----------------------------------------------------------------------------
    dl = [(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()) for l, r in pairs]
    dout = [torch.empty((H, W), dtype=torch.int16, device="cuda") for _ in range(P)]
    ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in dl]
    optrs = [o.data_ptr() for o in dout]
    torch.cuda.synchronize()

    def step_dev(sync):
        for i in range(P):
            eng.handles[i % S].call("b2s_compute_disparity_dev", ptrs[i][0], ptrs[i][1], H, W, CN, optrs[i], None)
        if sync:
            for h in eng.handles:
                h.sync()

    sampler = ClockSampler(local)
    sampler.wait_first()
    sampler.begin()  # clocks are sampled under load: warm-up + timed region
    for _ in range(Wm):
        step_dev(True)
    barrier()
    l0 = eng.launch_count()
    for h in eng.handles:
        h.event_record(0)
    for _ in range(K):
        step_dev(False)
    for h in eng.handles:
        h.event_record(1)
    ms = max(eng.handles[0].event_elapsed(0, h, 1) for h in eng.handles)
    for h in eng.handles:
        h.sync()
    barrier()
    launches = eng.launch_count() - l0
    sampler.end()
    clocks = sampler.stop()
    # parity guard: the timed outputs are the real thing (checked against cv2 in tests/; here a cheap sanity check)
    d0 = dout[0].cpu().numpy()
    assert (d0[:, :D] == -16).all() and (d0 >= 0).mean() > 0.5, "benchmark output is not a disparity map"

    # ---- e2e: public host API, pinned host buffers, H2D + D2H inside the timed region --------------------------
    hp = []
    for l, r in pairs:
        a, b = _ffi.pinned_empty(l.shape, np.uint8), _ffi.pinned_empty(r.shape, np.uint8)
        a[...] = l; b[...] = r
        hp.append((a, b))
    hout = [_ffi.pinned_empty((H, W), np.float32) for _ in range(P)]
    if world > 1:
        # BASELINE config 3: the results of a step stay on the device, are all-gathered over NVLink (every rank holds the batch in
        # global pair order) and each rank reads its own shard back to pinned host memory -- all inside the timed region
        # Two buffer sets: the all-gather and the device-to-host read of step k overlap the kernels of step k+1 (the collective
        # runs on NCCL's stream, the read on a copy stream); the clock stops only after the last step's gather and read.
        loc = [torch.empty((P, H, W), dtype=torch.float32, device="cuda") for _ in range(2)]
        gat = [torch.empty((world * P, H, W), dtype=torch.float32, device="cuda") for _ in range(2)]
        hloc = [torch.empty((P, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
        louts = [[l[i] for i in range(P)] for l in loc]
        copy_stream = torch.cuda.Stream()
        inflight = [None, None]
        coll_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e2e_k = [0]

        def drain(cur):
            if inflight[cur] is not None:
                work, ev = inflight[cur]
                work.wait()
                ev.synchronize()
                inflight[cur] = None

        def step_e2e():
            cur = e2e_k[0] & 1
            e2e_k[0] += 1
            drain(cur)                                  # buffer set `cur` was handed to the gather / read two steps ago
            eng.compute_batch(hp, out=louts[cur])       # H2D of the pinned inputs, kernels, result into loc[cur]; returns when the streams are idle
            coll_ev[0].record()
            work = dist.all_gather_into_tensor(gat[cur], loc[cur], async_op=True)  # THE all-gather of the step's disparities
            with torch.cuda.stream(copy_stream):
                hloc[cur].copy_(loc[cur], non_blocking=True)                          # D2H of this rank's shard
                ev = torch.cuda.Event()
                ev.record()
            inflight[cur] = (work, ev)

        def finish_e2e():
            drain(0)
            drain(1)
            coll_ev[1].record()
            torch.cuda.synchronize()
    else:
        def step_e2e():
            eng.compute_batch(hp, out=hout)

        def finish_e2e():
            pass
    for _ in range(Wm):
        step_e2e()
    finish_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    finish_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        allgather_ms = coll_ev[0].elapsed_time(coll_ev[1])  # last step: from the end of its kernels to the end of its all-gather and read
        last = (e2e_k[0] - 1) & 1
        assert torch.equal(gat[last][rank * P:(rank + 1) * P], loc[last]) and torch.equal(hloc[last], loc[last].cpu())
    eng.matchers[0].compute(*pairs[0])  # one synchronous call for the per-stage CUDA-event times
    stage = eng.handles[0].timings()

    # ---- full chain (BASELINE config 5): raw distorted 1080p pairs -> unrectify_depth through Stereo.get_depth_batch ----------
    chain = None
    if not args.no_chain:
        import calibrating_b200 as cb
        rig = synth.rig_dict((W, H))
        raw = [synth.render_rig(rig, seed=rank * 2 + i) for i in range(2)]
        cfg = dict(SGBM, max_size=1 << 20)
        cpairs = []
        for i in range(P):  # pinned host images, as in the e2e leg above
            a, b = _ffi.pinned_empty(raw[i % 2][0].shape, np.uint8), _ffi.pinned_empty(raw[i % 2][1].shape, np.uint8)
            a[...] = raw[i % 2][0]; b[...] = raw[i % 2][1]
            cpairs.append((a, b))
        nrep = max(K // 2, 2)
        if world > 1:
            # the product path of config 3: ShardedStereo = ONE rig broadcast from rank 0 (device to device over NVLink), then
            # per step pairs i % world on S streams per rank and ONE all-gather of the float64 depth maps, all timed
            from calibrating_b200 import sharded
            st = None
            if rank == 0:
                st = cb.Stereo.load(rig, device=local).set_stereo_matching(cb.SemiGlobalBlockMatching(cfg, device=local), max_depth=4.0)
            barrier()
            tb = time.perf_counter()
            sh = sharded.ShardedStereo(st, engine_factory=lambda: sharded.CudaEngine(local, streams=S))
            barrier()
            bcast_s = time.perf_counter() - tb
            hres = torch.empty((P, H, W), dtype=torch.float64).pin_memory()

            pend = [None]

            def chain_finish():
                if pend[0] is not None:
                    g = pend[0].result()                               # (world*P, H, W) float64 on the device, global pair order
                    hres.copy_(g[rank::world], non_blocking=True)      # this rank's shard back to the host
                    pend[0] = None

            def chain_step():
                p = sh.get_depth_batch(cpairs, wait=False)             # kernels done, all-gather enqueued on NCCL's stream
                chain_finish()                                         # the previous step's gather overlapped these kernels
                pend[0] = p
            for _ in range(2):
                chain_step()
            chain_finish()
            barrier()
            t0 = time.perf_counter()
            for _ in range(nrep):
                chain_step()
            chain_finish()
            barrier()
            chain_s = (time.perf_counter() - t0) / nrep
            what = ("ShardedStereo.get_depth_batch: undistort+rectify (LANCZOS4) -> SGBM -> depth -> unrectify on %d streams per rank, all-gather of the "
                    "float64 unrectify_depth of all ranks (%.0f MB per rank and step) inside the timed region, own shard read back to the host; "
                    "max over ranks" % (S, P * H * W * 8 / 1e6))
        else:
            st = cb.Stereo.load(rig, device=local).set_stereo_matching(cb.SemiGlobalBlockMatching(cfg, device=local), max_depth=4.0)
            couts = [{"unrectify_depth": _ffi.pinned_empty((H, W), np.float64)} for _ in range(P)]
            for _ in range(2):
                st.get_depth_batch(cpairs, streams=S, out=couts)
            barrier()
            t0 = time.perf_counter()
            for _ in range(nrep):
                st.get_depth_batch(cpairs, streams=S, out=couts)
            barrier()
            chain_s = (time.perf_counter() - t0) / nrep
            bcast_s = None
            what = ("undistort+rectify (LANCZOS4) -> SGBM (same parameters) -> depth -> unrectify, host uint8 raw pairs -> host float64 "
                    "unrectify_depth via Stereo.get_depth_batch")
        chain = {"seconds_per_step": chain_s, "unit": "pairs/s", "what": what, "h2d_bytes_per_step": P * 2 * H * W * CN, "d2h_bytes_per_step": P * H * W * 8}
        if bcast_s is not None:
            chain["rig_broadcast_s_incl_setup"] = bcast_s
            chain["rig_broadcast_bytes"] = sh.rig_bytes

    # ---- roofline: the aggregation group (path aggregation + winner-take-all) alone ---------------------------------------
    # production form: every repetition re-runs the cost stage first (untimed), so the first launch is the scan that also forms C
    # from the cost stage's row sums (agg_hscan_vsum_kernel), as in a real pair
    h0 = eng.handles[0]
    h0.call("b2s_compute_disparity_dev", ptrs[0][0], ptrs[0][1], H, W, CN, optrs[0], None)
    h0.sync()
    agg_parts = h0.bench_aggregate_parts(10)   # one CUDA-event interval per launch of the group
    agg_loop_ms = h0.bench_aggregate(10)       # one interval around the whole group (with the memsets and launch gaps between kernels)
    agg_ms = float(sum(agg_parts))             # kernel time of the group: what the roofline is computed from
    n_launch = len(agg_parts)
    # the other convention (round 1): the cost stage finishes C itself (vsum_kernel) and the first launch is the plain +x scan
    os.environ["B2S_NO_VSUM_FUSION"] = "1"
    try:
        alt_parts = h0.bench_aggregate_parts(10)
    finally:
        del os.environ["B2S_NO_VSUM_FUSION"]
    # several pairs in flight, as in the timed steps above: S handles run the group concurrently (on the finished C)
    for h in eng.handles:
        h.enqueue_aggregate(1)
    for h in eng.handles:
        h.sync()
    for h in eng.handles:
        h.event_record(2)
    for _ in range(8):
        for h in eng.handles:
            h.enqueue_aggregate(1)
    for h in eng.handles:
        h.event_record(3)
    agg_batched_ms = max(eng.handles[0].event_elapsed(2, h, 3) for h in eng.handles) / (8 * S)
    for h in eng.handles:
        h.sync()

    coll = None
    if world > 1:
        vals = [ms, e2e_s, allgather_ms, chain["seconds_per_step"] if chain else 0.0, chain.get("rig_broadcast_s_incl_setup", 0.0) if chain else 0.0]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # every multi-GPU number is the max over ranks
        ms, e2e_s, allgather_ms, chain_max, bcast_max = t.tolist()
        if chain:
            chain["seconds_per_step"] = chain_max
            chain["rig_broadcast_s_incl_setup"] = bcast_max
        coll = {"disparity_allgather_ms_per_step": allgather_ms, "disparity_allgather_bytes_per_rank": P * H * W * 4,
                "note": "inside the timed e2e region, overlapped with the next step's kernels (two buffer sets); the figure is the last step's tail: end of its kernels to the end of its all-gather and host read, max over ranks; the rig broadcast happens once per rig (chain_e2e)"}
    if chain:
        chain["value"] = world * P / chain["seconds_per_step"]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    # DRAM traffic of the group from the ncu capture under profiles/ -- only if it was taken on THIS build of the library
    # (scripts/agg_traffic.py records b2s_build_hash() next to the numbers)
    traffic, traffic_note = None, "no profiles/agg_traffic.json"
    build_hash = _ffi.lib().b2s_build_hash().decode()
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "agg_traffic.json")))
        if tj.get("build_hash") == build_hash:
            traffic, traffic_note = tj["dram_bytes_per_pair"] / n_launch, "ncu dram__bytes_read+write of the group / launches, %s, build %s" % (tj.get("source"), build_hash)
        else:
            traffic_note = "profiles/agg_traffic.json is from build %s, this library is %s: not reported" % (tj.get("build_hash"), build_hash)
    except Exception:
        pass
    achieved = AGG_BYTES_PER_PAIR / (agg_ms * 1e-3) / 1e9  # = (bytes per pair / launches) / (group time / launches)
    if n_launch == 3:
        names = ["agg_hscan_vsum_kernel (+x, forms C from the cost stage's row sums)", "agg_vsweep2_kernel (6 of 8 directions)",
                 "agg_hscan_kernel<ACCUM2, WTA> (-x, folds S2, winner-take-all fused)"]
        dirs = [1, 6, 1]
    elif n_launch == 2:
        names = ["agg_wave_kernel (both sweeps of four paths)", "wta_kernel (adds the two sums, winner-take-all)"]
        dirs = [8, 0]
    else:
        names, dirs = ["agg_scan_kernel"] * n_launch, [1] * n_launch
    out = {
        "metric": "stereo pairs/sec @1080p/128-disp SGM", "value": world * K * P / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world,
        "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": P, "streams_per_gpu": S, "parallelism": "dp%d (pairs sharded, no collective)" % world,
                   "l2": "working set per pair (C, S, S2 volumes, 1.49 GB) exceeds the 126 MB L2; %d distinct pairs rotate" % P},
        "clocks": clocks, "gpu_launches": int(launches) * world,  # (rank 0's count x ranks: every rank runs the same schedule)
        "e2e": {"value": world * K * P / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": P * 2 * H * W * CN, "d2h_bytes_per_step": P * H * W * 4,
                "api": "DisparityBatchEngine.compute_batch (host uint8 pairs -> host float32 disparity), wall clock between synchronisations"
                       + ("; N > 1: results stay on the device, NCCL all-gather of every step's disparities and the read of the own shard into pinned host memory, both overlapped with the next step's kernels" if world > 1 else "")},
        "roofline": {"bound": "hbm", "kernel": "aggregation group: %d launches per pair (dominant: agg_vsweep2_kernel)" % n_launch,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note,
                     "build_hash": build_hash,
                     "first_launch_plain_scan": {"ms_per_launch": [round(t, 4) for t in alt_parts], "group_ms": round(float(sum(alt_parts)), 4),
                                                 "frac": AGG_BYTES_PER_PAIR / (float(sum(alt_parts)) * 1e-3) / 1e9 / peak,
                                                 "note": "round-1 convention: C finished by the cost stage (vsum_kernel), first launch = agg_hscan_kernel<INIT>; the round-1 bench line (frac 0.237) used this form, so this is the figure to compare it with"},
                     "batched": {"handles_in_flight": S, "ms_per_pair": round(agg_batched_ms, 4), "achieved": AGG_BYTES_PER_PAIR / (agg_batched_ms * 1e-3) / 1e9,
                                 "frac": AGG_BYTES_PER_PAIR / (agg_batched_ms * 1e-3) / 1e9 / peak,
                                 "note": "the group of %d handles enqueued round-robin on their streams (on the finished C), wall = first start to last end" % S},
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy), of measured" if peaks else "fallback 6650 GB/s, of fallback",
                     "algorithmic_bytes_per_launch": AGG_BYTES_PER_PAIR // n_launch, "ms_per_launch": agg_ms / n_launch, "group_back_to_back_ms": agg_loop_ms,
                     "launches": [{"kernel": nm, "ms": round(t, 4), "algorithmic_bytes": AGG_BYTES_PER_PAIR * d // sum(dirs),
                                   "achieved_GBps": round(AGG_BYTES_PER_PAIR * d / sum(dirs) / (t * 1e-3) / 1e9, 1)}
                                  for nm, t, d in zip(names, agg_parts, dirs)],
                     "timed": "alone in its production form (cost stage re-run untimed before every repetition), one CUDA-event interval per launch on the engine stream; the group includes the winner-take-all; canonical 8 B/voxel (SURVEY 8(d)), this schedule moves 20 B/voxel"},
        "stage_ms_last_pair": {k: round(v, 3) for k, v in stage.items() if k.endswith("_ms")},
    }
    if coll:
        out["collectives"] = coll
    if chain:
        out["chain_e2e"] = chain
    if world == 1 and not args.no_cpu_baseline:
        import cv2
        T = host_threads()
        cpu_round(pairs, T)  # warm-up round
        rounds = 2
        t = sum(cpu_round(pairs, T) for _ in range(rounds))
        out["cpu_baseline"] = {"value": T * rounds / t, "unit": "pairs/s", "cores": T, "kind": "reference",
                               "sample": "%d rounds of %d threads x 1 pair (cv2 %s StereoSGBM MODE_HH, same pairs and parameters)" % (rounds, T, cv2.__version__)}
        if chain:
            # the reference's whole chain on the host (oracle/chain.py = calibrating's Stereo.get_depth restated with the same cv2 calls;
            # the real package needs boxx, which is not installable): T threads x 1 raw pair each, one round
            from oracle import chain as ochain
            rig = synth.rig_dict((W, H))
            raws = [synth.render_rig(rig, seed=i) for i in range(2)]
            refs = [ochain.RefStereo(rig).set_stereo_matching(
                ochain.SgbmPlugin(max_size=1 << 20, minDisparity=0, numDisparities=D, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1, uniquenessRatio=5,
                                  speckleWindowSize=200, speckleRange=2, mode=cv2.STEREO_SGBM_MODE_HH), max_depth=4.0) for _ in range(T)]
            refs[0].get_depth(*raws[0])  # warm-up (builds the maps)

            def cwork(i):
                with np.errstate(all="ignore"):
                    refs[i].get_depth(*raws[i % 2])
            ts = [threading.Thread(target=cwork, args=(i,)) for i in range(T)]
            t0 = time.perf_counter()
            [x.start() for x in ts]
            [x.join() for x in ts]
            tc = time.perf_counter() - t0
            out["chain_e2e"]["cpu_baseline"] = {"value": T / tc, "unit": "pairs/s", "cores": T, "kind": "port",
                                                "sample": "1 round of %d threads x 1 raw 1080p pair through the cv2 restatement of Stereo.get_depth (rectify LANCZOS4 x2, "
                                                          "StereoSGBM MODE_HH, depth, unrectify, undistort)" % T}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
