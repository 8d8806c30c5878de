"""Aggregation group (path aggregation + winner-take-all) at 1080p / 128 disparities, MODE_HH: alone on one stream and with
several handles in flight (development aid and the source of bench.py's roofline figures; run under gpurun).
  python scripts/agg_timing.py [schedule ...]      schedule = sweep | wave"""
import hashlib
import os
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

import calibrating_b200 as cb  # noqa: E402
from calibrating_b200 import synth  # noqa: E402

V = 1080 * 1792 * 128
PRM = dict(minDisparity=0, numDisparities=128, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1, uniquenessRatio=5,
           speckleWindowSize=200, speckleRange=2, mode=cb.MODE_HH)


def setenv(sched):
    for k in ("B2S_AGG_SCHEDULE", "B2S_WAVE_LPP"):
        os.environ.pop(k, None)
    if sched == "sweep":
        os.environ["B2S_AGG_SCHEDULE"] = "sweep"
    elif sched == "wave":
        os.environ["B2S_AGG_SCHEDULE"] = "wave"


def run(sched, nh=3, iters=10):
    setenv(sched)
    l, r, _ = synth.rectified_pair(1080, 1920, 128, seed=0)
    ms_ = [cb.StereoSGBM_create(**PRM) for _ in range(nh)]
    d = None
    for m in ms_:
        d = m.compute(l, r)
    h0 = ms_[0].handle
    print("[%s] stage times of one pair: %s" % (sched, h0.timings()))
    lone = h0.bench_aggregate(iters)
    parts = h0.bench_aggregate_parts(iters)
    print("[%s] group alone: %.3f ms = %.0f GB/s canonical (8 B/voxel); launches (ms): %s" % (sched, lone, V * 8 / lone / 1e6, ["%.3f" % p for p in parts]))
    for n in range(2, nh + 1):
        hs = [m.handle for m in ms_[:n]]
        for h in hs:
            h.enqueue_aggregate(1)
        for h in hs:
            h.sync()
        for h in hs:
            h.event_record(0)
        for _ in range(iters):  # round-robin so that the streams interleave like a batch of pairs
            for h in hs:
                h.enqueue_aggregate(1)
        for h in hs:
            h.event_record(1)
        t = max(hs[0].event_elapsed(0, h, 1) for h in hs)
        for h in hs:
            h.sync()
        per = t / (iters * n)
        print("[%s] %d handles in flight: %.3f ms per pair = %.0f GB/s canonical" % (sched, n, per, V * 8 / per / 1e6))
    print("[%s] disp md5 %s" % (sched, hashlib.md5(d.tobytes()).hexdigest()))
    for m in ms_:
        m.handle.close()


if __name__ == "__main__":
    for s in (sys.argv[1:] or ["sweep", "wave"]):
        run(s)
