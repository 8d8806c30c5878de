"""Timing of the reference's literal matcher parameters (stereo_matching.py:29-58: MODE_SGBM, D=218, minD=2, block 11) at
1080p and at the reference's default working size (longest side 1000 px), with cv2 on the host beside it (development aid)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import calibrating_b200 as cb
from calibrating_b200 import synth

for (h, w) in ((1080, 1920), (562, 1000)):
    l, r, _ = synth.rectified_pair(h, w, 200, seed=0)
    sm = cb.SemiGlobalBlockMatching({"max_size": 4000})
    for i in range(3):
        t = time.time(); d = sm(l, r); dt = time.time() - t
    tm = sm.stereo_sgbm.handle.timings()
    line = "%dx%d: wall %.2f ms " % (w, h, dt * 1e3) + json.dumps({k: round(v, 3) for k, v in tm.items() if k.endswith("_ms")})
    try:
        import cv2
        m = cv2.StereoSGBM_create(minDisparity=2, numDisparities=218, blockSize=11, uniquenessRatio=5, speckleWindowSize=200, speckleRange=2,
                                  disp12MaxDiff=0, P1=968, P2=3872)
        t = time.time(); ref = m.compute(l, r); ct = time.time() - t
        exp = ref.astype(np.float32).clip(0); exp[exp < 32] = 0
        line += " | cv2 %.0f ms, identical: %s" % (ct * 1e3, np.array_equal(d, exp / 16.0))
    except ImportError:
        pass
    print(line)
