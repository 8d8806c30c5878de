"""GPU parity of the full `Stereo.get_depth` chain against the reference's own outputs (golden) and the cv2 restatement."""
import os

import numpy as np
import pytest

import calibrating_b200 as cb
from calibrating_b200 import synth

pytestmark = pytest.mark.gpu


def _close_depth(a, b, rtol=1e-3):
    """north_star tolerance: <= 1e-3 relative on the float depth map (we expect ~1e-15)."""
    both = (a > 0) & (b > 0)
    assert ((a > 0) == (b > 0)).all(), "validity masks differ"
    return np.abs(a[both] - b[both]) <= rtol * np.abs(b[both])


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "rig320_default.npz")), np.load(os.path.join(golden_dir, "rig320_d64.npz"))


def test_get_depth_default_matches_reference_golden(golden):
    g, _ = golden
    st = cb.Stereo.load(synth.rig_dict((320, 240)))
    st.set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000}), max_depth=3.5)
    assert st.min_disparity == int(g["min_disparity"])
    assert np.allclose(st.K, g["K"], rtol=0, atol=1e-12)
    res = st.get_depth(g["img1"], g["img2"])
    assert set(res) == {"rectify_img1", "rectify_img2", "disparity", "rectify_depth", "unrectify_depth", "undistort_img1"}
    for k in ("rectify_img1", "rectify_img2", "undistort_img1"):
        assert res[k].dtype == np.uint8 and np.array_equal(res[k], g[k]), k
    assert res["disparity"].dtype == np.float32 and np.array_equal(res["disparity"], g["disparity"])
    assert res["rectify_depth"].dtype == np.float64 and res["unrectify_depth"].dtype == np.float64
    assert np.array_equal(res["rectify_depth"].astype(np.float32), g["rectify_depth"])
    assert _close_depth(res["unrectify_depth"], g["unrectify_depth"].astype(np.float64), 1e-6).all()


def test_get_depth_d64_no_translation(golden):
    g0, g = golden
    st = cb.Stereo.load(synth.rig_dict((320, 240)))
    st.set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}))
    res = st.get_depth(g0["img1"], g0["img2"])
    assert np.array_equal(res["disparity"], g["disparity"])
    assert np.array_equal(res["rectify_depth"].astype(np.float32), g["rectify_depth"])
    assert _close_depth(res["unrectify_depth"], g["unrectify_depth"].astype(np.float64), 1e-6).all()
    assert np.array_equal(res["undistort_img1"], g["undistort_img1"])


def test_foreign_plugin_and_stage_methods(golden):
    """A user-defined MetaStereoMatching (host code) still plugs in; the stage methods match the cv2 restatement."""
    cv2 = pytest.importorskip("cv2")
    from oracle import chain
    g, _ = golden
    rig = synth.rig_dict((320, 240))

    class Cv2Plugin(cb.MetaStereoMatching):
        def __call__(self, a, b):
            return dict(disparity=chain.SgbmPlugin(max_size=4000)(a, b), extra=1)

    st = cb.Stereo.load(rig).set_stereo_matching(Cv2Plugin(), max_depth=3.5)
    res = st.get_depth(g["img1"], g["img2"])
    assert res["extra"] == 1
    assert np.array_equal(res["disparity"], g["disparity"])
    ref = chain.RefStereo(rig).set_stereo_matching(chain.SgbmPlugin(max_size=4000), max_depth=3.5)
    r1, r2 = st.rectify(g["img1"], g["img2"])
    e1, e2 = ref.rectify(g["img1"], g["img2"])
    assert np.array_equal(r1, e1) and np.array_equal(r2, e2)
    d = np.abs(np.random.default_rng(0).normal(20, 10, (240, 320))).astype(np.float32)
    d[::7] = 0
    assert np.array_equal(st.disparity_to_depth(d), ref.disparity_to_depth(d.copy()))
    z = ref.disparity_to_depth(d.copy())
    a, b = st.unrectify_depth(z), ref.unrectify_depth(z)
    assert ((a > 0) == (b > 0)).all() and np.allclose(a, b, rtol=1e-12, atol=0)
    assert np.array_equal(st.undistort_img(g["img1"]), ref.undistort_img(g["img1"]))
    gray = g["img1"][..., 0].copy()
    assert np.array_equal(st.undistort_img(gray), ref.undistort_img(gray))
    gray2 = g["img2"][..., 1].copy()  # single-channel LANCZOS4 rectification (its own tap-pair path in remap_u8_kernel)
    r1, r2 = st.rectify(gray, gray2)
    e1, e2 = ref.rectify(gray, gray2)
    assert r1.shape == e1.shape and np.array_equal(r1, e1) and np.array_equal(r2, e2)


@pytest.mark.parametrize("interp", ["lanczos4", "linear"])
def test_chain_1080p_vs_cv2_restatement(interp):
    """BASELINE config 5 stand-in at 1080p: whole chain vs the cv2 restatement, and depth vs analytic ground truth."""
    cv2 = pytest.importorskip("cv2")
    from oracle import chain
    rig = synth.rig_dict((1920, 1080))
    img1, img2 = synth.render_rig(rig, seed=1)
    cfg = {"max_size": 4000, "num_disparities": 128, "min_disparity": 0, "block_size": 5, "P1": 600, "P2": 2400, "disp12_max_diff": 1,
           "mode": cb.MODE_HH}
    st = cb.Stereo.load(rig, interp=interp).set_stereo_matching(cb.SemiGlobalBlockMatching(cfg), max_depth=4.0)
    res = st.get_depth(img1, img2)
    ref = chain.RefStereo(rig).set_stereo_matching(
        chain.SgbmPlugin(max_size=4000, numDisparities=128, minDisparity=0, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1, mode=1), max_depth=4.0)
    exp = ref.get_depth(img1, img2, cv2.INTER_LANCZOS4 if interp == "lanczos4" else cv2.INTER_LINEAR)
    for k in ("rectify_img1", "rectify_img2", "undistort_img1", "disparity"):
        assert np.array_equal(res[k], exp[k]), k
    assert np.array_equal(res["rectify_depth"], exp["rectify_depth"])
    assert _close_depth(res["unrectify_depth"], exp["unrectify_depth"], 1e-9).all()
    gt = synth.gt_depth_cam1(rig)
    v = res["unrectify_depth"] > 0
    assert v.mean() > 0.5
    assert np.median(np.abs(res["unrectify_depth"][v] - gt[v]) / gt[v]) < 0.01


@pytest.mark.parametrize("size", [(320, 240), (1920, 1080)])
def test_device_side_map_generation(size, golden):
    """the rig described by parameters only (`maps="device"`, b2s_set_rig_params).  The
    generated maps equal cv2.initUndistortRectifyMap's bit for bit (float64 in cv2's operation order, no FMA), so the
    whole chain gives the same arrays as with uploaded maps."""
    import cv2
    rig = synth.rig_dict(size)
    host = cb.Stereo.load(rig).set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}), max_depth=3.5)
    dev = cb.Stereo.load(rig, maps="device").set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}),
                                                                 max_depth=3.5)
    img1, img2 = synth.render_rig(rig, seed=3)
    a, b = host.get_depth(img1, img2), dev.get_depth(img1, img2)
    w, h = dev.xy
    w1, h1 = dev.cam1.xy
    exp = dict(zip(["map1x", "map1y"], dev.undistort_rectify_map1))
    exp.update(zip(["map2x", "map2y"], dev.undistort_rectify_map2))
    exp.update(zip(["unrect_mapx", "unrect_mapy"], dev._unrectify_maps()))
    for name, e in exp.items():
        got = dev.handle.fetch_rig(name, e.shape, np.float32)
        # cv2's AVX path contracts some multiply-adds: at 1080p ONE pixel of map1x (row 974, column 1309 of the synthetic rig)
        # lands on the other side of a float32 rounding boundary (1314.9327 vs 1314.9329, 1 ulp); everything else is identical
        diff = np.argwhere(got != e)
        ulp = np.abs(got.view(np.int32).astype(np.int64) - e.view(np.int32).astype(np.int64))
        assert len(diff) <= 2 and ulp.max() <= 1, "%s: %d pixels differ from cv2.initUndistortRectifyMap, first %s, max %d ulp" % (
            name, len(diff), diff[:3].tolist(), ulp.max())
        if size == (320, 240):
            assert len(diff) == 0, name
    und_xy, und_fxy = cv2.initUndistortRectifyMap(dev.cam1.K, dev.cam1.D, None, dev.cam1.K, (w1, h1), cv2.CV_16SC2)
    assert np.array_equal(dev.handle.fetch_rig("undist_xy", (h1, w1, 2), np.int16), und_xy)
    assert np.array_equal(dev.handle.fetch_rig("undist_fxy", (h1, w1), np.uint16), und_fxy)
    assert np.array_equal(dev.handle.fetch_rig("valid_mask1", (h, w), np.uint8).astype(bool), dev.rectify_valid_mask1)
    for k in a:  # identical maps, identical kernels: identical results
        if a[k].dtype == np.uint8 or "depth" not in k:
            assert np.array_equal(a[k], b[k]), k
        else:
            assert _close_depth(np.float64(b[k]), np.float64(a[k]), 1e-12).all(), k


def test_distort_depth(golden_dir):
    """Stereo.distort_depth on the device, bit-exact against the real reference's output
    (golden), stand-alone and inside get_depth(return_distort_depth=True); 1080p against the oracle restatement."""
    from oracle import chain
    g = np.load(os.path.join(golden_dir, "rig320_distort.npz"))
    rig = synth.rig_dict((320, 240))
    for maps in ("host", "device"):
        st = cb.Stereo.load(rig, maps=maps).set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}), max_depth=3.5)
        got = st.distort_depth(g["unrectify_depth"])
        assert got.dtype == np.float64 and np.array_equal(got, g["distort_depth"]), maps
    img1, img2 = synth.render_rig(rig, seed=1)
    res = st.get_depth(img1, img2, return_distort_depth=True)
    assert set(res) >= {"distort_img1", "distort_depth", "unrectify_depth", "undistort_img1"} and res["distort_img1"] is not None
    assert np.array_equal(res["distort_depth"], chain.RefStereo(rig).distort_depth(res["unrectify_depth"]))
    with pytest.raises(ValueError):
        st.distort_depth(np.zeros((10, 10)))
    rig = synth.rig_dict((1920, 1080))
    depth = np.random.default_rng(5).random((1080, 1920)) * 4
    depth[depth < 0.5] = 0
    assert np.array_equal(cb.Stereo.load(rig).distort_depth(depth), chain.RefStereo(rig).distort_depth(depth))


@pytest.mark.parametrize("maps", ["host", "device"])
def test_get_depth_batch(maps):
    """Stereo.get_depth_batch: several pairs in flight on several handles give exactly get_depth's arrays, in order."""
    rig = synth.rig_dict((320, 240))
    st = cb.Stereo.load(rig, maps=maps).set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}), max_depth=3.5)
    pairs = [synth.render_rig(rig, seed=s) for s in range(5)]
    keys = ("disparity", "unrectify_depth", "undistort_img1", "rectify_img2", "distort_depth")
    got = st.get_depth_batch(pairs, streams=3, keys=keys)
    assert len(got) == 5
    for (a, b), g in zip(pairs, got):
        ref = st.get_depth(a, b, return_distort_depth=True)
        for k in keys:
            assert g[k].dtype == ref[k].dtype and np.array_equal(g[k], ref[k]), k
    st.set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 48}), max_depth=3.0)  # new rig constants
    g2 = st.get_depth_batch(pairs[:2], streams=3, keys=("unrectify_depth",))
    assert np.array_equal(g2[1]["unrectify_depth"], st.get_depth(*pairs[1])["unrectify_depth"])
    with pytest.raises(ValueError):
        st.get_depth_batch(pairs[:1], keys=("nope",))
    from calibrating_b200 import _ffi  # pinned inputs in place, results straight into caller-owned pinned buffers
    pp = []
    for a, b in pairs[:3]:
        x, y = _ffi.pinned_empty(a.shape, np.uint8), _ffi.pinned_empty(b.shape, np.uint8)
        x[...] = a
        y[...] = b
        pp.append((x, y))
    outs = [{"unrectify_depth": _ffi.pinned_empty((240, 320), np.float64)} for _ in range(3)]
    g3 = st.get_depth_batch(pp, streams=2, out=outs)
    for i in range(3):
        assert g3[i]["unrectify_depth"] is outs[i]["unrectify_depth"]
        assert np.array_equal(outs[i]["unrectify_depth"], st.get_depth(*pairs[i])["unrectify_depth"])


def test_project_cam2_depth(golden_dir):
    """Cam.project_cam2_depth on the device against the real reference's output (golden) and,
    at other sizes / rates / dtypes, the oracle restatement.  Same pixels hit, z within 1e-12 relative (the reference's
    matrix products go through BLAS)."""
    from oracle import reproject
    g = np.load(os.path.join(golden_dir, "rig320_project.npz"))
    rig = synth.rig_dict((320, 240))
    cam1, cam2 = cb.Cam.load(rig["cam1"]), cb.Cam.load(rig["cam2"])
    got = cam1.project_cam2_depth(cam2, g["depth2"], T=g["T"])
    exp = g["depth1"]
    assert got.dtype == np.float64 and got.shape == exp.shape
    assert ((got != 0) == (exp != 0)).all() and np.allclose(got, exp, rtol=1e-12, atol=0)
    with pytest.raises(NotImplementedError):
        cam1.project_cam2_depth(cam2, g["depth2"])
    rng = np.random.default_rng(3)
    big1 = cb.Cam(np.float64([[900, 0, 640], [0, 905, 360], [0, 0, 1]]), None, (1280, 720))
    small2 = cb.Cam(np.float64([[420, 0, 330], [0, 418, 236], [0, 0, 1]]), None, (640, 480))
    T = np.eye(4)
    T[:3, :3] = __import__("cv2").Rodrigues(np.float64([0.02, -0.05, 0.01]))[0]
    T[:3, 3] = [0.06, -0.01, 0.02]
    d16 = (rng.random((480, 640)) * 3000 + 500).astype(np.uint16)
    d16[rng.random((480, 640)) < 0.2] = 0
    for interp in (1.5, 0, 0.5):
        for d2, (ca, cb2) in ((d16, (big1, small2)), (np.float64(d16) / 1000.0, (small2, small2))):
            got = ca.project_cam2_depth(cb2, d2, T=T, interpolation=interp)
            exp = reproject.project_cam2_depth(ca.K, ca.xy, cb2.K, d2, T, interpolation=interp)
            same = ((got != 0) == (exp != 0)).mean()
            assert same > 0.99999, (interp, same)
            both = (got != 0) & (exp != 0)
            assert np.allclose(got[both], exp[both], rtol=1e-9, atol=0) or (np.abs(got[both] - exp[both]) <= 1e-9 * exp[both]).mean() > 0.9999


def test_two_rigs_share_one_matcher():
    """example/test_different_stereo.py and example/test_rotate_stereo.py of the reference hang ONE SemiGlobalBlockMatching on
    two Stereo objects.  The engine handle (and the rig uploaded into it) is shared with the matcher, so every Stereo must
    re-upload its rig when the other one used the handle in between."""
    rig_a, rig_b = synth.rig_dict((320, 240)), synth.rig_dict((256, 192))
    sm = cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64})
    A = cb.Stereo.load(rig_a).set_stereo_matching(sm, max_depth=3.5)
    B = cb.Stereo.load(rig_b).set_stereo_matching(sm, max_depth=2.5)
    assert A.handle is B.handle
    ia, ib = synth.render_rig(rig_a, seed=1), synth.render_rig(rig_b, seed=2)
    a1 = A.get_depth(*ia)
    b1 = B.get_depth(*ib)
    a2 = A.get_depth(*ia)  # B's rig is in the handle now: A has to notice
    b2 = B.get_depth(*ib)
    for k in a1:
        assert np.array_equal(a1[k], a2[k]), k
        assert np.array_equal(b1[k], b2[k]), k
    alone = cb.Stereo.load(rig_a).set_stereo_matching(cb.SemiGlobalBlockMatching({"max_size": 4000, "num_disparities": 64}), max_depth=3.5)
    ref = alone.get_depth(*ia)
    for k in a1:
        assert np.array_equal(a1[k], ref[k]), k
    r1, r2 = A.rectify(*ia)  # the stage methods follow the same rule
    B.rectify(*ib)
    assert np.array_equal(A.undistort_img(ia[0]), ref["undistort_img1"])
    assert np.array_equal(r1, ref["rectify_img1"])


def test_plugin_disparity_shape_is_checked():
    rig = synth.rig_dict((320, 240))

    class Half(cb.MetaStereoMatching):
        def __call__(self, a, b):
            return np.zeros((a.shape[0] // 2, a.shape[1] // 2), np.float32)

    st = cb.Stereo.load(rig).set_stereo_matching(Half(), max_depth=3.5)
    with pytest.raises(ValueError, match="shape"):
        st.get_depth(*synth.render_rig(rig, seed=0))


def test_default_max_size_runs_on_the_device():
    """`calibrating.SemiGlobalBlockMatching()` keeps the reference's default max_size = 1000 (stereo_matching.py:26,60-70): on a
    1080p pair the matcher works at 1000 x 562.  Down-scale, matcher and up-scale x w/sw are one C-ABI call; against the
    restatement (oracle/resize.py = cv2.resize INTER_LINEAR for the unpinned boxx.resize; cv2.StereoSGBM live) the float
    disparity is identical, and against cv2.resize's own float up-scale it is within 1e-6 of the range."""
    cv2 = pytest.importorskip("cv2")
    from oracle import chain, resize
    l, r, _ = synth.rectified_pair(1080, 1920, 200, seed=5)
    sm = cb.SemiGlobalBlockMatching()
    assert sm.max_size == 1000
    n0 = sm.stereo_sgbm.handle.launch_count()
    got = sm(l, r)
    assert got.shape == (1080, 1920) and got.dtype == np.float32
    assert sm.stereo_sgbm.handle.launch_count() - n0 >= 12, "the resize kernels and the matcher run on the device"
    ref = chain.SgbmPlugin()  # reference defaults, max_size 1000
    exp = ref(l, r)
    assert np.array_equal(got, exp)
    sl, sr = (cv2.resize(a, (1000, 562), interpolation=cv2.INTER_LINEAR) for a in (l, r))
    via_cv2 = cv2.resize(ref._compute_float(sl, sr), (1920, 1080), interpolation=cv2.INTER_LINEAR) * np.float32(1920) / np.float32(1000)
    assert np.abs(got - via_cv2).max() <= 1e-6 * 1920 * 220 / 1000
    with pytest.raises(Exception, match="int16 disparity is not defined"):
        sm.stereo_sgbm.handle.call("b2s_set_option", 4, 1000)
        try:
            sm.stereo_sgbm.compute(l, r)
        finally:
            sm.stereo_sgbm.handle.call("b2s_set_option", 4, 0)
    assert np.array_equal(sm.stereo_sgbm.compute(l, r).shape, (1080, 1920))  # the cv2-style object never scales


def test_get_depth_with_default_matcher_is_one_call():
    """Stereo.get_depth with the reference's default matcher (max_size 1000) on a 1280 x 720 rig: one b2s_get_depth call, results
    equal to the cv2 restatement of the reference chain."""
    pytest.importorskip("cv2")
    from oracle import chain
    rig = synth.rig_dict((1280, 720))
    img1, img2 = synth.render_rig(rig, seed=4)
    st = cb.Stereo.load(rig).set_stereo_matching(cb.SemiGlobalBlockMatching(), max_depth=3.5)
    res = st.get_depth(img1, img2)
    exp = chain.RefStereo(rig).set_stereo_matching(chain.SgbmPlugin(), max_depth=3.5).get_depth(img1, img2)
    assert np.array_equal(res["rectify_img1"], exp["rectify_img1"])
    assert np.array_equal(res["disparity"], exp["disparity"])
    assert np.array_equal(res["rectify_depth"], exp["rectify_depth"])
    assert _close_depth(res["unrectify_depth"], exp["unrectify_depth"], 1e-9).all()
    got = st.get_depth_batch([(img1, img2)] * 2, streams=2, keys=("disparity",))
    assert np.array_equal(got[1]["disparity"], exp["disparity"])


def test_point_cloud_functions(golden_dir):
    """SURVEY.md section 8(f) rank 3, stand-alone: depth_to_point_cloud and point_cloud_to_depth (utils.py:213-317) on the device
    against the real reference's outputs: same points in the same (row-major) order, same pixels hit; values within 1e-12
    relative (the reference's 3x3 products go through BLAS)."""
    from oracle import reproject
    g = np.load(os.path.join(golden_dir, "cloud_small.npz"))
    pc = cb.depth_to_point_cloud(g["depth"], g["K"])
    assert pc.shape == g["cloud_rate1"].shape and np.allclose(pc, g["cloud_rate1"], rtol=1e-12, atol=1e-15)
    xyzuv = cb.depth_to_point_cloud(g["depth16"], g["K"], interpolation_rate=1.5, return_xyzuv=True)
    assert xyzuv.shape == g["xyzuv_rate15"].shape and np.allclose(xyzuv, g["xyzuv_rate15"], rtol=1e-12, atol=1e-15)
    assert np.array_equal(xyzuv[:, 3:], g["xyzuv_rate15"][:, 3:]), "u, v of the up-sampled grid"
    back = cb.point_cloud_to_depth(g["moved"], g["K"], (160, 120))
    assert back.shape == (120, 160) and ((back != 0) == (g["depth_back"] != 0)).all()
    assert np.allclose(back, g["depth_back"], rtol=1e-12, atol=0)
    assert cb.depth_to_point_cloud(np.zeros((8, 9)), g["K"]).shape == (0, 3)
    assert (cb.point_cloud_to_depth(np.zeros((0, 3)), g["K"], (5, 4)) == 0).all()
    # a 1080p image with three quarters of the pixels valid, float32 metres, against the restatement
    rng = np.random.default_rng(2)
    d = (rng.random((1080, 1920)) * 3 + 0.5).astype(np.float32)
    d[rng.random(d.shape) < 0.25] = 0
    K = np.float64([[1000, 0, 960], [0, 1001, 540], [0, 0, 1]])
    got, exp = cb.depth_to_point_cloud(d, K), reproject.depth_to_point_cloud(d, K)
    assert got.shape == exp.shape and np.allclose(got, exp, rtol=1e-12, atol=1e-15)
    z = cb.point_cloud_to_depth(got, K, (1920, 1080))
    assert np.allclose(z, np.float64(d), rtol=1e-12, atol=0), "un-project and project back is the identity on the pixel grid"


def test_sparse_interpolation_and_sparse_matchers(golden_dir):
    """interpolate_uvzs / interpolate_sparse2d (utils.py:347-411) with the dense half on the device,
    bit-equal to the real reference's outputs, and the two sparse matcher plugins built on them (stereo_matching.py:73-142)."""
    from oracle import sparse
    g = np.load(os.path.join(golden_dir, "sparse_small.npz"))
    hw = tuple(int(v) for v in g["hw"])
    assert np.array_equal(cb.interpolate_uvzs(g["uvzs"], hw), g["lstsq"])
    assert np.array_equal(cb.interpolate_uvzs(g["uvzs"], hw, "convex_hull"), g["lstsq_hull"])
    assert np.array_equal(cb.interpolate_uvzs(g["uvzs"]), g["lstsq_nohw"])
    assert np.array_equal(cb.interpolate_uvzs(g["uvzs"], hw, None, "nearest"), g["nearest2"])
    assert np.array_equal(cb.interpolate_uvzs(g["uvzs"], hw, True, "nearest", distance=6), g["nearest6_hull"])
    assert np.array_equal(cb.interpolate_sparse2d(g["sparse"], "convex_hull"), g["sparse2d_hull"])
    out = cb.interpolate_uvzs(np.zeros((0, 3)), (4, 5))
    assert out.shape == (4, 5) and not out.any()
    with pytest.raises(NotImplementedError):
        cb.interpolate_uvzs(g["uvzs"], hw, inter_type="cubic")
    for key, hull in (("rbf", None), ("rbf_hull", True)):  # thin plate: host solve, dense evaluation on the device, float64
        got = cb.interpolate_uvzs(g["uvzs"][:120], hw, hull, "rbf")
        assert got.shape == g[key].shape and got.dtype == np.float64
        assert np.allclose(got, g[key], rtol=0, atol=1e-9 * np.abs(g[key]).max()), np.abs(got - g[key]).max()
    # 1080p, 2000 samples (more than one shared-memory tile of the nearest search), against the restatement
    rng = np.random.default_rng(4)
    uvz = np.stack([rng.random(2000) * 1900 + 5, rng.random(2000) * 1060 + 5, rng.random(2000) * 50 + 1], 1)
    assert np.array_equal(cb.interpolate_uvzs(uvz, (1080, 1920), True, "nearest", distance=25), sparse.interpolate_uvzs(uvz, (1080, 1920), True, "nearest", 25))
    assert np.array_equal(cb.interpolate_uvzs(uvz, (1080, 1920), True), sparse.interpolate_uvzs(uvz, (1080, 1920), True))

    # MatchingByBoard: a stand-in board whose "detector" returns seeded corners with a known disparity plane
    class Board:
        def __init__(self):
            gy, gx = np.mgrid[20:80:7j, 25:120:9j]
            self.p1 = np.float32(np.stack([gx.ravel(), gy.ravel()], 1))
            self.p2 = self.p1 - np.float32(np.stack([8 + 0.02 * self.p1[:, 0], 0.01 * np.sin(self.p1[:, 1])], 1))
        def find_image_points(self, d):
            d["image_points"] = self.p1 if d["img"][0, 0, 0] == 1 else self.p2
    b = Board()
    img1, img2 = np.ones((90, 140, 3), np.uint8), np.zeros((90, 140, 3), np.uint8)
    got = cb.MatchingByBoard(b)(img1, img2)
    sp = np.zeros((90, 140), np.float32)
    sp[np.int32(b.p1[:, 1].round()), np.int32(b.p1[:, 0].round())] = (b.p1 - b.p2)[:, 0]
    with np.errstate(divide="ignore"):
        exp = 1 / sparse.interpolate_sparse2d(1 / sp, "convex_hull")
    assert np.array_equal(got["disparity"], exp) and got["rectify_std"] > 0
    assert np.count_nonzero(cb.MatchingByBoard(b, dense_predict=False)(img1, img2)["disparity"]) == len(b.p1)

    # FeatureMatchingAsStereoMatching: a stand-in matcher returning normalised matches; 1/8-resolution nearest fill, then up-scale
    class FM:
        cfg = dict(shape=(240, 320))
        def __call__(self, a, b):
            r = np.random.default_rng(9)
            uv1 = np.float32(r.random((400, 2)))
            uv2 = uv1.copy()
            uv2[:, 0] -= np.float32(0.02 + 0.05 * uv1[:, 1])
            return dict(uvs1=uv1, uvs2=uv2)
    fm = FM()
    a = np.zeros((240, 320, 3), np.uint8)
    got = cb.FeatureMatchingAsStereoMatching(fm)(a, a)
    m = fm(a, a)
    small = (30, 40)
    uvs1, uvs2 = m["uvs1"] * small[::-1], m["uvs2"] * small[::-1]
    uvds = np.concatenate((uvs1, (uvs1 - uvs2)[:, :1]), 1)
    d = sparse.interpolate_uvzs(uvds, small, None, "nearest")
    d = d * 320 / 40
    import cv2
    exp = cv2.resize(d, (320, 240), interpolation=cv2.INTER_NEAREST)
    assert got["disparity"].shape == (240, 320) and np.array_equal(got["disparity"], exp)
