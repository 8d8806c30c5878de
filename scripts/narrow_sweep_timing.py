"""agg_vsweep6_kernel against agg_vsweep_kernel on images of 13 / 26 / 130 cost columns x 1080 rows (1, 2, 10 CTAs): cycles per row of
the lock-step sweep without (13 columns) and with CTA boundaries.  Development aid, run under gpurun."""
import sys, os
sys.path.insert(0, ".")
import numpy as np
import calibrating_b200 as cb
from calibrating_b200 import synth
for wcols in (13, 26, 130):
    for env in ({"B2S_SWEEP6": "1"}, {}):
        for k in ("B2S_SWEEP6",):
            os.environ.pop(k, None)
        os.environ.update(env)
        l, r, _ = synth.rectified_pair(1080, 128 + wcols, 128, seed=0)
        m = cb.StereoSGBM_create(minDisparity=0, numDisparities=128, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1, uniquenessRatio=5, speckleWindowSize=200, speckleRange=2, mode=cb.MODE_HH)
        m.compute(l, r)
        parts = m.handle.bench_aggregate_parts(5)
        print(wcols, env, ["%.3f" % p for p in parts], "cycles/row of sweep: %.0f" % (parts[1] * 1e-3 * 1.965e9 / 1080))
        m.handle.close()
