"""Per-stage timing of the full Stereo.get_depth chain at 1080p (development aid, run under gpurun)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import calibrating_b200 as cb
from calibrating_b200 import synth

rig = synth.rig_dict((1920, 1080))
img1, img2 = synth.render_rig(rig, seed=0)
cfg = dict(max_size=4000, min_disparity=0, num_disparities=128, block_size=5, P1=600, P2=2400, disp12_max_diff=1, uniqueness_ratio=5,
           speckle_window_size=200, speckle_range=2, mode=cb.MODE_HH)
for interp in ("lanczos4", "linear"):
    st = cb.Stereo.load(rig, interp=interp).set_stereo_matching(cb.SemiGlobalBlockMatching(cfg), max_depth=4.0)
    for i in range(3):
        t = time.time(); r = st.get_depth(img1, img2); dt = time.time() - t
    print(interp, "wall %.2f ms" % (dt * 1e3), json.dumps({k: round(v, 3) for k, v in st.handle.timings().items()}))
