"""MODE_HH4 at 1080p/128: stage times (and equality with cv2 when cv2 is importable)."""
import json
import sys

sys.path.insert(0, ".")

import numpy as np

import calibrating_b200 as cb
from calibrating_b200 import _ffi, synth

h = _ffi.Handle(0)
l, r, _ = synth.rectified_pair(1080, 1920, 128, seed=0)
kw = dict(minDisparity=0, numDisparities=128, blockSize=5, P1=8 * 3 * 25, P2=32 * 3 * 25, disp12MaxDiff=1, uniquenessRatio=5,
          speckleWindowSize=200, speckleRange=2, mode=cb.MODE_HH4)
m = cb.StereoSGBM_create(handle=h, **kw)
for _ in range(3):
    got = m.compute(l, r)
print(json.dumps({k: round(v, 3) for k, v in h.timings().items()}))
try:
    import cv2, time
    t = time.time()
    ref = cv2.StereoSGBM_create(**kw).compute(l, r)
    print("cv2 MODE_HH4 %.0f ms, equal: %s" % ((time.time() - t) * 1e3, np.array_equal(ref, got)))
except ImportError:
    pass
