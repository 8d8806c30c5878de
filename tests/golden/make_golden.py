"""Generate tests/golden/*.npz by running the REAL reference package (imported from /root/reference through a
throw-away `boxx` shim: boxx is a hard import of the reference, requirements.txt:1, and is not installable
offline) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box; the tests read only the committed .npz files.
"""
import os
import sys
import tempfile
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SHIM = textwrap.dedent('''
    """Minimal stand-in for the `boxx` names the reference's hot path touches (throw-away, never shipped)."""
    import contextlib, glob as _glob
    import numpy as np
    import cv2
    from numpy import pi, arctan, tan
    glob = _glob.glob
    @contextlib.contextmanager
    def inpkg():
        yield
    def npa(x):
        return np.array(x)
    def imread(p):
        return cv2.imread(p)[..., ::-1]
    def imsave(p, a):
        cv2.imwrite(p, a[..., ::-1])
    def resize(img, arg2, interpolation=None):
        # boxx.resize semantics are not available offline; the golden vectors only use ratio == 1 paths
        hw = img.shape[:2]
        if isinstance(arg2, (int, float)):
            if arg2 == 1:
                return img
            new = (int(round(hw[0] * arg2)), int(round(hw[1] * arg2)))
        else:
            new = tuple(arg2)
        if new == tuple(hw):
            return img
        raise NotImplementedError("boxx.resize with a real size change is unpinned offline")
    def increase(*a, **k): return 0
    def strnum(x, *a, **k): return str(x)
    def _noop(*a, **k): return None
    shows = show = showb = tree = loga = mg = timeit = _noop
''')


def import_reference():
    d = tempfile.mkdtemp(prefix="boxx_shim_")
    os.makedirs(os.path.join(d, "boxx"))
    with open(os.path.join(d, "boxx", "__init__.py"), "w") as f:
        f.write(SHIM)
    sys.path.insert(0, d)
    sys.path.insert(0, "/root/reference")
    import calibrating  # noqa
    return calibrating


def sparse_case(calibrating):
    """case G: utils.interpolate_uvzs / interpolate_sparse2d (utils.py:347-411) of the real reference on seeded sparse samples.
    `np.bool8` (utils.py:324) left NumPy in 2.0; it is aliased to np.bool_ here so the convex-hull branch runs unmodified."""
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    rng = np.random.default_rng(23)
    hw = (90, 140)
    n = 300
    uv = np.stack([rng.random(n) * (hw[1] - 21) + 10, rng.random(n) * (hw[0] - 17) + 8], 1)
    z = 0.004 * uv[:, 0] - 0.007 * uv[:, 1] + 2.5 + rng.normal(0, 0.01, n)
    uvzs = np.concatenate([uv, z[:, None]], 1)
    U = calibrating.utils
    sparse = np.zeros(hw, np.float32)
    sel = rng.random(hw) < 0.02
    sel[:12] = sel[-9:] = False
    sel[:, :15] = sel[:, -11:] = False
    sparse[sel] = (40 + 0.05 * np.mgrid[:hw[0], :hw[1]][1] + rng.normal(0, 0.2, hw))[sel]
    with np.errstate(divide="ignore"):
        board = 1 / U.interpolate_sparse2d(1 / sparse, "convex_hull")  # MatchingByBoard's dense_predict (stereo_matching.py:105)
    np.savez_compressed(os.path.join(HERE, "sparse_small.npz"), uvzs=uvzs, hw=np.int32(hw), sparse=sparse,
                        lstsq=U.interpolate_uvzs(uvzs.copy(), hw), lstsq_hull=U.interpolate_uvzs(uvzs.copy(), hw, "convex_hull"),
                        lstsq_nohw=U.interpolate_uvzs(uvzs.copy()),
                        nearest2=U.interpolate_uvzs(uvzs.copy(), hw, None, "nearest"),
                        nearest6_hull=U.interpolate_uvzs(uvzs.copy(), hw, True, "nearest", distance=6),
                        sparse2d_hull=U.interpolate_sparse2d(sparse.copy(), "convex_hull"), board_dense=board,
                        rbf=U.interpolate_uvzs(uvzs[:120].copy(), hw, None, "rbf"), rbf_hull=U.interpolate_uvzs(uvzs[:120].copy(), hw, True, "rbf"))


def main():
    import cv2
    from calibrating_b200 import synth
    calibrating = import_reference()
    if sys.argv[1:] == ["sparse"]:
        sparse_case(calibrating)
        return

    class Param64(calibrating.MetaStereoMatching):
        """BASELINE config 1 matcher: the reference's parameters with numDisparities=64."""
        def __init__(self):
            self.sgbm = cv2.StereoSGBM_create(minDisparity=2, numDisparities=64, blockSize=11, uniquenessRatio=5,
                                              speckleWindowSize=200, speckleRange=2, disp12MaxDiff=0, P1=968, P2=3872)
        def __call__(self, a, b):
            d = self.sgbm.compute(a, b).astype(np.float32).clip(0)
            d[d < 2 * 16] = 0
            return d / 16.0

    out = {}
    # --- case A: 320x240 rig, reference default SemiGlobalBlockMatching (D=218, minD=2, block 11, 5-path)
    rig = synth.rig_dict((320, 240))
    img1, img2 = synth.render_rig(rig, seed=0)
    st = calibrating.Stereo.load(rig)
    st.set_stereo_matching(calibrating.SemiGlobalBlockMatching({"max_size": 4000}), max_depth=3.5)
    with np.errstate(all="ignore"):
        res = st.get_depth(img1, img2)
    np.savez_compressed(os.path.join(HERE, "rig320_default.npz"), img1=img1, img2=img2,
                        min_disparity=st.min_disparity, K=st.K, R1=st.R1, R2=st.R2,
                        map1x=st.undistort_rectify_map1[0][::16, ::16], map2y=st.undistort_rectify_map2[1][::16, ::16],
                        **{k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in res.items()})
    out["rig320_default"] = {k: (v.shape, str(v.dtype)) for k, v in res.items()}
    # --- case B: 320x240 rig, 64 disparities, no translation (max_depth=None)
    st = calibrating.Stereo.load(rig)
    st.set_stereo_matching(Param64())
    with np.errstate(all="ignore"):
        res = st.get_depth(img1, img2)
    np.savez_compressed(os.path.join(HERE, "rig320_d64.npz"), min_disparity=st.min_disparity,
                        **{k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in res.items()
                           if k in ("disparity", "rectify_depth", "unrectify_depth", "undistort_img1")})
    # --- case D: distort_depth (stereo_camera.py:433-464) of case B's unrectify_depth, from the real reference
    with np.errstate(all="ignore"):
        res = st.get_depth(img1, img2, return_distort_depth=True)
    np.savez_compressed(os.path.join(HERE, "rig320_distort.npz"), unrectify_depth=res["unrectify_depth"],
                        distort_depth=res["distort_depth"])
    # --- case E: Cam.project_cam2_depth (camera.py:298-309): a seeded depth image of cam2 (20 % holes) seen from cam1 through
    # the rig's own T = [R | t], default interpolation 1.5 (pins oracle/reproject.py and b2s_project_depth)
    rng = np.random.default_rng(7)
    depth2 = rng.random((240, 320)) * 1.5 + 0.5
    depth2[rng.random((240, 320)) < 0.2] = 0
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = np.float64(rig["R"]), np.float64(rig["t"]).ravel()
    cam1, cam2 = calibrating.Cam.load(rig["cam1"]), calibrating.Cam.load(rig["cam2"])
    with np.errstate(all="ignore"):
        depth1 = cam1.project_cam2_depth(cam2, depth2, T=T)
    np.savez_compressed(os.path.join(HERE, "rig320_project.npz"), depth2=depth2, T=T, depth1=depth1,
                        rate=calibrating.utils._get_appropriate_interpolation_rate(cam1, cam2, 1.5))
    # --- case F: utils.depth_to_point_cloud / point_cloud_to_depth (utils.py:213-317) of the real reference on a small seeded
    # depth image: rate 1 on float64 metres, rate 1.5 on uint16 millimetres with (u, v) appended, and the z-buffered way back
    rng = np.random.default_rng(11)
    Kc = np.float64([[210.0, 0, 81.5], [0, 209.0, 58.25], [0, 0, 1]])
    dm = rng.random((120, 160)) * 2.5 + 0.4
    dm[rng.random((120, 160)) < 0.25] = 0
    d16 = np.uint16(dm * 1000)
    pc1 = calibrating.utils.depth_to_point_cloud(dm, Kc)
    pc2 = calibrating.utils.depth_to_point_cloud(d16, Kc, interpolation_rate=1.5, return_xyzuv=True)
    Rc = cv2.Rodrigues(np.float64([0.03, -0.08, 0.02]))[0]
    moved = pc1 @ Rc.T + np.float64([0.05, -0.02, 0.1])
    back = calibrating.utils.point_cloud_to_depth(moved.copy(), Kc, (160, 120))
    np.savez_compressed(os.path.join(HERE, "cloud_small.npz"), K=Kc, depth=dm, depth16=d16, cloud_rate1=pc1, xyzuv_rate15=pc2, moved=moved, depth_back=back)
    # --- case C: raw cv2.StereoSGBM outputs on a small rectified pair, MODE_SGBM / MODE_HH / MODE_HH4 (pins oracle/sgbm_ref.c)
    l, r, _ = synth.rectified_pair(96, 200, 48, seed=3)
    for mode, name in ((0, "sgbm"), (1, "hh"), (3, "hh4")):
        m = cv2.StereoSGBM_create(minDisparity=0, numDisparities=48, blockSize=5, P1=8 * 3 * 25, P2=32 * 3 * 25,
                                  disp12MaxDiff=1, uniquenessRatio=5, speckleWindowSize=50, speckleRange=2, mode=mode)
        out[name] = m.compute(l, r)
    m = cv2.StereoSGBM_create(minDisparity=2, numDisparities=40, blockSize=11, P1=968, P2=3872, disp12MaxDiff=0,
                              uniquenessRatio=5, speckleWindowSize=200, speckleRange=2)
    np.savez_compressed(os.path.join(HERE, "sgbm_small.npz"), left=l, right=r, disp_sgbm=out["sgbm"], disp_hh=out["hh"], disp_hh4=out["hh4"],
                        disp_refparams=m.compute(l, r), cv2_version=cv2.__version__)
    sparse_case(calibrating)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")), "cv2", cv2.__version__)


if __name__ == "__main__":
    main()
