"""Seeded synthetic inputs for tests and bench.py (there is no example data offline:
the reference clones `calibrating_example_data` at run time, calibrating/utils.py:722-739).

* `rectified_pair`  -- BASELINE config 2: an already-rectified textured pair with known disparity.
* `rig_dict` / `render_rig` -- BASELINE configs 1/5 stand-in: a physically consistent distorted two-camera
  rig looking at a textured plane, in the dict schema `Stereo.load` accepts
  (calibrating/stereo_camera.py:264-297).
Host-side helper only: numpy + cv2, nothing here is on the device path.
"""
import cv2
import numpy as np


def _texture(h, w, rng, cell=4):
    base = rng.integers(0, 256, (h // cell + 2, w // cell + 2, 3), dtype=np.uint8)
    tex = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.int32)
    tex += rng.integers(0, 20, (h, w, 3))
    return tex.clip(0, 255).astype(np.uint8)


def gt_disparity(h, w, D):
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    d = 0.25 * D + 0.5 * D * (0.5 + 0.5 * np.sin(2 * np.pi * xs / w) * np.cos(np.pi * ys / h))
    d[h // 3:2 * h // 3, w // 3:2 * w // 3] = 0.8 * D
    return d.astype(np.float32)


def rectified_pair(h=1080, w=1920, D=128, seed=0, cn=3):
    """left(x) = right(x - d(x,y)).  Returns (left, right, gt_disparity)."""
    rng = np.random.default_rng(seed)
    right = _texture(h, w, rng)
    d = gt_disparity(h, w, D)
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    left = cv2.remap(right, xs - d, ys, cv2.INTER_LINEAR)
    if cn == 1:
        left = cv2.cvtColor(left, cv2.COLOR_RGB2GRAY)
        right = cv2.cvtColor(right, cv2.COLOR_RGB2GRAY)
    return np.ascontiguousarray(left), np.ascontiguousarray(right), d


def rig_dict(xy=(640, 480)):
    """The survey's 640x480 rig, scaled to `xy` (fx scales with width)."""
    s = xy[0] / 640.0
    return dict(
        R=cv2.Rodrigues(np.array([0.004, -0.012, 0.003]))[0].tolist(),
        t=[[-0.12], [0.001], [0.002]],
        cam1=dict(fx=520 * s, fy=521 * s, cx=322 * s, cy=238 * s, D=[[-0.10, 0.03, 8e-4, -5e-4, 0.0]], xy=list(xy), name="cam1"),
        cam2=dict(fx=518 * s, fy=519 * s, cx=318 * s, cy=242 * s, D=[[-0.09, 0.02, -4e-4, 6e-4, 0.0]], xy=list(xy), name="cam2"),
    )


def _K(c):
    return np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1.0]])


PLANE = dict(z0=1.6, ax=0.12, ay=-0.08)  # plane  ax*X + ay*Y + Z = z0  in cam1's frame (metres)


def _render(cam, R, t, tex, plane, tex_scale):
    w, h = cam["xy"]
    us, vs = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    pts = np.stack([us.ravel(), vs.ravel()], 1).reshape(-1, 1, 2)
    nrm = cv2.undistortPoints(pts, _K(cam), np.float64(cam["D"])).reshape(-1, 2)
    rays = np.concatenate([nrm, np.ones((len(nrm), 1))], 1)
    # camera frame X_c = R X_1 + t  ->  centre o = -R^T t, direction R^T ray, both in cam1's frame
    o = -R.T @ t.reshape(3)
    dirs = rays @ R
    n = np.array([plane["ax"], plane["ay"], 1.0])
    lam = (plane["z0"] - o @ n) / (dirs @ n)
    hit = o[None] + dirs * lam[:, None]
    tu = (hit[:, 0] * tex_scale + tex.shape[1] / 2).astype(np.float32).reshape(h, w)
    tv = (hit[:, 1] * tex_scale + tex.shape[0] / 2).astype(np.float32).reshape(h, w)
    return cv2.remap(tex, tu, tv, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)


def render_rig(rig, seed=0):
    """Ray-cast the textured plane through both distorted cameras.  Returns (img1, img2) RGB uint8."""
    rng = np.random.default_rng(seed)
    w, h = rig["cam1"]["xy"]
    tex = _texture(3 * h, 3 * w, rng, cell=6)
    tex_scale = rig["cam1"]["fx"] / PLANE["z0"] * 1.2
    R = np.float64(rig["R"])
    t = np.float64(rig["t"])
    img1 = _render(rig["cam1"], np.eye(3), np.zeros(3), tex, PLANE, tex_scale)
    img2 = _render(rig["cam2"], R, t, tex, PLANE, tex_scale)
    return np.ascontiguousarray(img1), np.ascontiguousarray(img2)


def gt_depth_cam1(rig):
    """Analytic depth of the plane in cam1's UNDISTORTED pinhole frame (what `unrectify_depth` returns)."""
    w, h = rig["cam1"]["xy"]
    us, vs = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    x = (us - rig["cam1"]["cx"]) / rig["cam1"]["fx"]
    y = (vs - rig["cam1"]["cy"]) / rig["cam1"]["fy"]
    return PLANE["z0"] / (PLANE["ax"] * x + PLANE["ay"] * y + 1.0)
