"""Quick per-stage timing of the 1080p/128 HH matcher (development aid, run under gpurun)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import calibrating_b200 as cb
from calibrating_b200 import synth

l, r, _ = synth.rectified_pair(1080, 1920, 128, seed=0)
m = cb.StereoSGBM_create(minDisparity=0, numDisparities=128, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1, uniquenessRatio=5,
                         speckleWindowSize=200, speckleRange=2, mode=cb.MODE_HH)
import os
if os.environ.get('FUSE'):
    m.handle.fuse_wta(os.environ['FUSE'] == '1')
for i in range(3):
    t = time.time(); d = m.compute(l, r); dt = time.time() - t
    print("call %d: %.2f ms wall" % (i, dt * 1e3), json.dumps(m.handle.timings()))
ms = m.handle.bench_aggregate(10)
V = 1080 * 1792 * 128
print("aggregate alone: %.3f ms  -> canonical %.1f GB/s (8 B/voxel)" % (ms, V * 8 / ms / 1e6))
parts = m.handle.bench_aggregate_parts(10)
import hashlib
from calibrating_b200 import _ffi
print("build_hash", _ffi.lib().b2s_build_hash().decode())
print("disp md5", hashlib.md5(d.tobytes()).hexdigest())
print("aggregation launches (ms):", ["%.3f" % p for p in parts], "sum %.3f" % sum(parts))
