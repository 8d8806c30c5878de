"""Small end-to-end run for compute-sanitizer (memcheck / racecheck) under gpurun."""
import sys
sys.path.insert(0, ".")
import numpy as np
import calibrating_b200 as cb
from calibrating_b200 import synth
from oracle import sgbm as osgbm

for (h, w, D, cn, mode, bs) in [(24, 150, 64, 3, 1, 5), (17, 300, 218, 3, 0, 11), (30, 120, 16, 1, 1, 3), (20, 400, 256, 1, 0, 7), (21, 330, 130, 3, 1, 9), (19, 170, 48, 3, 3, 5), (9, 140, 100, 1, 1, 1)]:
    l, r, _ = synth.rectified_pair(h, w, D, 1, cn)
    p = dict(min_disparity=2, num_disparities=D, block_size=bs, P1=8 * cn * bs * bs, P2=32 * cn * bs * bs, disp12_max_diff=1,
             uniqueness_ratio=5, speckle_window_size=30, speckle_range=2, mode=mode)
    m = cb.StereoSGBM(**p)
    m.handle.keep_volumes(True)  # (the default configuration does not store S; it is exercised below)
    got = m.compute(l, r)
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **p)
    m.handle.keep_volumes(False)
    assert np.array_equal(m.compute(l, r), ref["disp"])
    m.handle.fuse_wta(False)
    assert np.array_equal(m.compute(l, r), ref["disp"])
    print((h, w, D, cn, mode, bs), "C", np.array_equal(m.handle.fetch_volume(0), ref["C"]), "S", np.array_equal(m.handle.fetch_volume(1), ref["S"]),
          "raw", np.array_equal(m.handle.fetch_raw(h, w), ref["raw"]), "disp", np.array_equal(got, ref["disp"]))
rig = synth.rig_dict((320, 240))
img1, img2 = synth.render_rig(rig, seed=0)
st = cb.Stereo.load(rig).set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}), max_depth=3.5)
res = st.get_depth(img1, img2)
print("chain ok", {k: v.shape for k, v in res.items()})
# later additions: census cost, device-side maps, distort_depth, project_cam2_depth, fused winner-take-all
l, r, _ = synth.rectified_pair(20, 200, 64, 2, 1)
p = dict(min_disparity=0, num_disparities=64, block_size=5, P1=10, P2=120, disp12_max_diff=1, uniqueness_ratio=5, speckle_window_size=20,
         speckle_range=2, mode=1, cost=cb.COST_CENSUS)
m = cb.StereoSGBM(**p)
print("census", np.array_equal(m.compute(l, r), osgbm.sgbm_compute(l, r, **p)))
m.handle.fuse_wta(False)
print("census separate wta", np.array_equal(m.compute(l, r), osgbm.sgbm_compute(l, r, **p)))
st2 = cb.Stereo.load(rig, maps="device").set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}), max_depth=3.5)
res2 = st2.get_depth(img1, img2, return_distort_depth=True)
print("device maps chain equal", all(np.array_equal(res[k], res2[k]) for k in res), res2["distort_depth"].shape)
cam1, cam2 = cb.Cam.load(rig["cam1"]), cb.Cam.load(rig["cam2"])
T = np.eye(4); T[:3, 3] = [0.05, 0, 0.01]
print("project", (cam1.project_cam2_depth(cam2, res["unrectify_depth"], T=T) > 0).mean())
print("batch", len(st.get_depth_batch([(img1, img2)] * 3, streams=2)))
# round 2: device-side max_size (resize kernels), point clouds, sparse interpolation
st3 = cb.Stereo.load(rig).set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 200, "num_disparities": 48}), max_depth=3.5)
print("max_size chain", st3.get_depth(img1, img2)["unrectify_depth"].shape)
K = np.float64([[210.0, 0, 81.5], [0, 209.0, 58.25], [0, 0, 1]])
dm = np.random.default_rng(0).random((60, 80)) * 2 + 0.5
pc = cb.depth_to_point_cloud(dm, K, interpolation_rate=1.5)
print("cloud", pc.shape, cb.point_cloud_to_depth(pc, K, (80, 60)).shape)
uvz = np.random.default_rng(1).random((300, 3)) * [70, 50, 3] + [5, 5, 1]
for it in ("lstsq", "nearest", "rbf"):
    print("interpolate", it, cb.interpolate_uvzs(uvz[:100], (60, 80), True, it).shape)
