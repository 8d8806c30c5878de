"""CPU: pin the oracle (oracle/sgbm_ref.c, oracle/remap.py, oracle/chain.py) against the committed golden vectors
(generated from the REAL reference + cv2 by tests/golden/make_golden.py) and against the installed cv2 live."""
import os

import numpy as np
import pytest

from calibrating_b200 import synth
from oracle import remap as oremap
from oracle import sgbm as osgbm

cv2 = pytest.importorskip("cv2")


def test_sgbm_oracle_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "sgbm_small.npz"))
    l, r = g["left"], g["right"]
    common = dict(min_disparity=0, num_disparities=48, block_size=5, P1=8 * 3 * 25, P2=32 * 3 * 25, disp12_max_diff=1,
                  uniqueness_ratio=5, speckle_window_size=50, speckle_range=2)
    assert np.array_equal(osgbm.sgbm_compute(l, r, mode=0, **common), g["disp_sgbm"])
    assert np.array_equal(osgbm.sgbm_compute(l, r, mode=1, **common), g["disp_hh"])
    assert np.array_equal(osgbm.sgbm_compute(l, r, mode=3, **common), g["disp_hh4"])
    ref = osgbm.sgbm_compute(l, r, min_disparity=2, num_disparities=40, block_size=11, P1=968, P2=3872, disp12_max_diff=0,
                             uniqueness_ratio=5, speckle_window_size=200, speckle_range=2)
    assert np.array_equal(ref, g["disp_refparams"])


@pytest.mark.parametrize("seed", range(20))
def test_sgbm_oracle_vs_cv2_random(seed):
    rng = np.random.default_rng(100 + seed)
    h = int(rng.integers(20, 60)); D = int(rng.choice([16, 24, 32, 48, 70])); minD = int(rng.choice([0, 2, 3, 5]))
    w = int(rng.integers(D + minD + 12, D + minD + 90)); cn = int(rng.choice([1, 3])); mode = int(rng.choice([0, 1]))
    if seed >= 12:
        mode = 3  # MODE_HH4
    bs = int(rng.choice([1, 3, 5, 7, 9, 11])); uniq = int(rng.choice([0, 1, 5, 10, 15])); d12 = int(rng.choice([-1, 0, 1, 2, 100]))
    spk = int(rng.choice([0, 20, 200]))
    P1, P2 = 8 * cn * bs * bs, 32 * cn * bs * bs
    if seed % 3 == 0:
        l = rng.integers(0, 256, (h, w, cn), dtype=np.uint8).squeeze()
        r = rng.integers(0, 256, (h, w, cn), dtype=np.uint8).squeeze()
    else:
        l, r, _ = synth.rectified_pair(h, w, D, seed, cn)
    ref = cv2.StereoSGBM_create(minDisparity=minD, numDisparities=D, blockSize=bs, P1=P1, P2=P2, disp12MaxDiff=d12,
                                uniquenessRatio=uniq, speckleWindowSize=spk, speckleRange=2, mode=mode).compute(l, r)
    got = osgbm.sgbm_compute(l, r, min_disparity=minD, num_disparities=D, block_size=bs, P1=P1, P2=P2, disp12_max_diff=d12,
                             uniqueness_ratio=uniq, speckle_window_size=spk, speckle_range=2, mode=mode)
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("uniq", [50, 99, 100, 150])
def test_sgbm_oracle_extreme_uniqueness(uniq):
    l, r, _ = synth.rectified_pair(40, 120, 32, 5, 1)
    ref = cv2.StereoSGBM_create(0, 32, 3, 72, 288, 1, 0, uniq, 0, 0, 1).compute(l, r)
    got = osgbm.sgbm_compute(l, r, num_disparities=32, block_size=3, P1=72, P2=288, disp12_max_diff=1, uniqueness_ratio=uniq, mode=1)
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("P1,P2", [(100, 8000), (4000, 12000), (11000, 12000)])
def test_sgbm_oracle_large_penalties(P1, P2):
    """Inside max C + 2 * P2 <= 32767 the int16 sums of cv2 never saturate and the restatement (int arithmetic) is exact.
    Beyond that bound (e.g. P2 = 30000) cv2's saturating SIMD adds and its scalar tail disagree with each other, so its
    result depends on the SIMD width of the build: not part of the parity claim."""
    l, r, _ = synth.rectified_pair(36, 150, 48, 41, 3)
    r = np.random.default_rng(3).integers(0, 256, r.shape, dtype=np.uint8)
    for mode in (0, 1, 3):
        ref = cv2.StereoSGBM_create(0, 48, 5, P1, P2, 1, 0, 5, 0, 0, mode).compute(l, r)
        got = osgbm.sgbm_compute(l, r, num_disparities=48, block_size=5, P1=P1, P2=P2, disp12_max_diff=1, uniqueness_ratio=5, mode=mode,
                                 want_volumes=True)
        assert int(got["C"].max()) + 2 * P2 <= 32767
        assert np.array_equal(ref, got["disp"]), mode


@pytest.mark.parametrize("h", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("bs", [3, 5, 11])
def test_sgbm_oracle_hh4_few_rows(h, bs):
    """MODE_HH4 on images shorter than the block: cv2 keeps a constant cost for every row y > 0 with y + bs/2 >= H
    (oracle/sgbm_ref.c, vertical half of A.3), i.e. for all rows but the first one here."""
    rng = np.random.default_rng(10 * h + bs)
    r = cv2.GaussianBlur(rng.integers(0, 255, (h, 60), dtype=np.uint8), (5, 1), 0)
    l = (np.roll(r, 6, axis=1).astype(int) + rng.integers(0, 9, (h, 60))).clip(0, 255).astype(np.uint8)
    ref = cv2.StereoSGBM_create(0, 16, bs, 20, 80, -1, 0, 0, 0, 0, 3).compute(l, r)
    got = osgbm.sgbm_compute(l, r, num_disparities=16, block_size=bs, P1=20, P2=80, disp12_max_diff=-1, uniqueness_ratio=0, mode=3,
                             want_volumes=True)
    assert np.array_equal(ref, got["disp"])
    if h > 1:
        assert not got["C"][max(1, h - bs // 2):].any()


def test_sgbm_oracle_precondition():
    l = np.zeros((20, 30), np.uint8)
    with pytest.raises(ValueError):
        osgbm.sgbm_compute(l, l, num_disparities=32, block_size=5)
    with pytest.raises(cv2.error):
        cv2.StereoSGBM_create(numDisparities=32, blockSize=5).compute(l, l)


def test_remap_oracle_vs_cv2():
    rng = np.random.default_rng(1)
    H, W = 90, 130
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    mx = (xs * 1.05 - 6 + 3 * np.sin(ys / 17)).astype(np.float32) + rng.random((H, W), dtype=np.float32)
    my = (ys * 0.97 + 3 + 2 * np.cos(xs / 23)).astype(np.float32) + rng.random((H, W), dtype=np.float32)
    assert np.array_equal(oremap.remap_lanczos4_u8(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LANCZOS4))
    assert np.array_equal(oremap.remap_linear_u8(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR))
    g = src[..., 1].copy()
    assert np.array_equal(oremap.remap_lanczos4_u8(g, mx, my), cv2.remap(g, mx, my, cv2.INTER_LANCZOS4))
    z = rng.random((H, W)) * 5
    assert np.array_equal(oremap.remap_nearest(z, mx, my), cv2.remap(z, mx, my, cv2.INTER_NEAREST))


def test_undistort_oracle_vs_cv2():
    rng = np.random.default_rng(2)
    H, W = 120, 160
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    K = np.array([[0.8 * W, 0, W / 2 + 2.3], [0, 0.81 * W, H / 2 - 1.7], [0, 0, 1]])
    D = np.array([[-0.10, 0.03, 8e-4, -5e-4, 0.0]])
    m1, m2 = cv2.initUndistortRectifyMap(K, D, None, K, (W, H), cv2.CV_16SC2)
    assert np.array_equal(oremap.remap_linear_u8_fixed(src, m1[..., 0], m1[..., 1], m2), cv2.undistort(src, K, D))


@pytest.mark.parametrize("hw", [(1080, 1920), (720, 1280), (1200, 1600), (1081, 1923), (333, 2001)])
def test_resize_restatement_vs_cv2(hw):
    """oracle/resize.py (the stand-in for the reference's unpinned boxx.resize calls, stereo_matching.py:61-69) against
    cv2.resize(INTER_LINEAR): uint8 down-scale bit-exact, float32 up-scale within 1e-6 of the value range (cv2's IPP path
    keeps float64 coefficients; its accumulation order is not public)."""
    from oracle import resize
    rng = np.random.default_rng(1)
    h, w = hw
    nh, nw = resize.scaled_size(h, w, 1000)
    assert max(nh, nw) == 1000
    for cn in (1, 3):
        img = rng.integers(0, 256, (h, w, cn), dtype=np.uint8).squeeze()
        assert np.array_equal(resize.resize_u8(img, nh, nw), cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR))
    d = (rng.random((nh, nw)) * 200).astype(np.float32)
    d[rng.random((nh, nw)) < 0.2] = 0
    assert np.abs(resize.resize_f32(d, h, w) - cv2.resize(d, (w, h), interpolation=cv2.INTER_LINEAR)).max() <= 200 * 1e-6
    assert resize.scaled_size(480, 640, 1000) == (480, 640)


@pytest.mark.parametrize("minD", [-1, -3, -16, -40, -64])
def test_negative_min_disparity_vs_cv2(minD):
    """SURVEY.md Appendix A.1 left negative minDisparity unverified (minX1 = max(maxD, 0), maxX1 = W + min(minD, 0)); the reference
    has the matching negative-shift branch (stereo_camera.py:236-240).  The restatement equals cv2 there too, including
    maxD < 0 (the whole disparity range to the right)."""
    for mode in (0, 1, 3):
        for D, cn in ((16, 1), (48, 3), (70, 1)):
            l, r, _ = synth.rectified_pair(30, 180, 48, seed=3 + D, cn=cn)
            if (minD + D) % 2:
                l = np.random.default_rng(1).integers(0, 256, l.shape, dtype=np.uint8)
            kw = dict(minDisparity=minD, numDisparities=D, blockSize=5, P1=8 * cn * 25, P2=32 * cn * 25, disp12MaxDiff=1, uniquenessRatio=5,
                      speckleWindowSize=20, speckleRange=2, mode=mode)
            ref = cv2.StereoSGBM_create(**kw).compute(l, r)
            got = osgbm.sgbm_compute(l, r, min_disparity=minD, num_disparities=D, block_size=5, P1=8 * cn * 25, P2=32 * cn * 25, disp12_max_diff=1,
                                     uniqueness_ratio=5, speckle_window_size=20, speckle_range=2, mode=mode)
            assert np.array_equal(got, ref), (minD, mode, D, cn)
