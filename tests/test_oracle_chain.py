"""CPU: the cv2 restatement of the reference chain (oracle/chain.py) reproduces the REAL reference's outputs
(tests/golden/rig320_*.npz, written by tests/golden/make_golden.py from /root/reference)."""
import os

import numpy as np
import pytest

from calibrating_b200 import synth
from oracle import chain

pytest.importorskip("cv2")


def _run(golden, max_depth, **kw):
    rig = synth.rig_dict((320, 240))
    st = chain.RefStereo(rig)
    st.set_stereo_matching(chain.SgbmPlugin(max_size=4000, **kw), max_depth=max_depth)
    return st, st.get_depth(golden["img1"], golden["img2"])


def test_chain_default_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "rig320_default.npz"))
    img1, img2 = synth.render_rig(synth.rig_dict((320, 240)), seed=0)
    assert np.array_equal(img1, g["img1"]) and np.array_equal(img2, g["img2"])  # the generator is deterministic
    st, res = _run(g, 3.5)
    assert st.min_disparity == int(g["min_disparity"])
    assert np.allclose(st.K, g["K"], rtol=0, atol=1e-12) and np.allclose(st.R1, g["R1"], rtol=0, atol=1e-14)
    assert np.array_equal(st.map1[0][::16, ::16], g["map1x"]) and np.array_equal(st.map2[1][::16, ::16], g["map2y"])
    for k in ("rectify_img1", "rectify_img2", "undistort_img1"):
        assert np.array_equal(res[k], g[k]), k
    for k in ("disparity", "rectify_depth", "unrectify_depth"):
        assert np.array_equal(res[k].astype(np.float32), g[k]), k
    assert res["rectify_depth"].dtype == np.float64 and res["disparity"].dtype == np.float32


def test_chain_d64_matches_reference(golden_dir):
    g0 = np.load(os.path.join(golden_dir, "rig320_default.npz"))
    g = np.load(os.path.join(golden_dir, "rig320_d64.npz"))
    st, res = _run(g0, None, numDisparities=64)
    assert int(g["min_disparity"]) == st.min_disparity
    for k in ("disparity", "rectify_depth", "unrectify_depth"):
        assert np.array_equal(res[k].astype(np.float32), g[k]), k
    assert np.array_equal(res["undistort_img1"], g["undistort_img1"])


def test_distort_depth_matches_reference(golden_dir):
    """oracle.chain.RefStereo.distort_depth against the real reference's Stereo.distort_depth (stereo_camera.py:433-464)."""
    from calibrating_b200 import synth
    from oracle import chain
    g = np.load(os.path.join(golden_dir, "rig320_distort.npz"))
    got = chain.RefStereo(synth.rig_dict((320, 240))).distort_depth(g["unrectify_depth"])
    assert got.dtype == np.float64 and np.array_equal(got, g["distort_depth"])


def test_project_cam2_depth_matches_reference(golden_dir):
    """oracle.reproject against the real reference's Cam.project_cam2_depth (camera.py:298-309), bit for bit."""
    from calibrating_b200 import synth
    from oracle import reproject
    g = np.load(os.path.join(golden_dir, "rig320_project.npz"))
    rig = synth.rig_dict((320, 240))
    K = lambda c: np.float64([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]])
    got = reproject.project_cam2_depth(K(rig["cam1"]), rig["cam1"]["xy"], K(rig["cam2"]), g["depth2"], g["T"])
    assert np.array_equal(got, g["depth1"])
    assert abs(reproject.interpolation_rate(K(rig["cam1"]), K(rig["cam2"])) - float(g["rate"])) == 0


def test_point_cloud_restatements_match_reference(golden_dir):
    """oracle.reproject.depth_to_point_cloud / point_cloud_to_depth against the real reference's utils.py:213-317 outputs."""
    from oracle import reproject
    g = np.load(os.path.join(golden_dir, "cloud_small.npz"))
    assert np.array_equal(reproject.depth_to_point_cloud(g["depth"], g["K"]), g["cloud_rate1"])
    assert np.array_equal(reproject.depth_to_point_cloud(g["depth16"], g["K"], 1.5, True), g["xyzuv_rate15"])
    assert np.array_equal(reproject.point_cloud_to_depth(g["moved"], g["K"], (160, 120)), g["depth_back"])


def test_sparse_interpolation_restatement_matches_reference(golden_dir):
    """oracle.sparse against the real reference's interpolate_uvzs / interpolate_sparse2d (utils.py:347-411): plane fit with and
    without the convex-hull mask, default hw, nearest sample within 2 and 6 px, and MatchingByBoard's 1/interp(1/sparse)."""
    from oracle import sparse
    g = np.load(os.path.join(golden_dir, "sparse_small.npz"))
    hw = tuple(int(v) for v in g["hw"])
    assert np.array_equal(sparse.interpolate_uvzs(g["uvzs"], hw), g["lstsq"])
    assert np.array_equal(sparse.interpolate_uvzs(g["uvzs"], hw, "convex_hull"), g["lstsq_hull"])
    assert np.array_equal(sparse.interpolate_uvzs(g["uvzs"]), g["lstsq_nohw"])
    assert np.array_equal(sparse.interpolate_uvzs(g["uvzs"], hw, None, "nearest"), g["nearest2"])
    assert np.array_equal(sparse.interpolate_uvzs(g["uvzs"], hw, True, "nearest", 6), g["nearest6_hull"])
    assert np.array_equal(sparse.interpolate_sparse2d(g["sparse"], "convex_hull"), g["sparse2d_hull"])
    with np.errstate(divide="ignore"):
        assert np.array_equal(1 / sparse.interpolate_sparse2d(1 / g["sparse"], "convex_hull"), g["board_dense"])
    assert sparse.interpolate_uvzs(np.zeros((0, 3)), (4, 5)).shape == (4, 5)
    # thin-plate "rbf" (scipy.interpolate.Rbf in the reference): float64, two channels (the reference's own output shape)
    for key, hull in (("rbf", None), ("rbf_hull", True)):
        got = sparse.interpolate_uvzs(g["uvzs"][:120], hw, hull, "rbf")
        assert got.shape == g[key].shape == hw + (2,) and got.dtype == np.float64
        assert np.allclose(got, g[key], rtol=0, atol=1e-9 * np.abs(g[key]).max())
