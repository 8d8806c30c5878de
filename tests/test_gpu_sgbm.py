"""GPU parity: the CUDA matcher (through the C-ABI) against the C oracle (oracle/sgbm_ref.c), the golden vectors and,
when cv2 is importable on the box, cv2.StereoSGBM live.  Integer work: bit-exact."""
import os

import numpy as np
import pytest

import calibrating_b200 as cb
from calibrating_b200 import synth
from oracle import sgbm as osgbm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from calibrating_b200 import _ffi
    h = _ffi.Handle(0)
    h.keep_volumes(True)  # the stage-level checks fetch S (by default the last pass fuses the WTA and does not store it)
    yield h
    h.close()


def _case(rng, D=None, cn=None, mode=None, minD=None, bs=None):
    D = D or int(rng.choice([16, 24, 32, 48, 64, 70, 100, 128, 130, 200, 218, 256]))
    minD = int(rng.choice([0, 2, 3, 5])) if minD is None else minD
    cn = cn or int(rng.choice([1, 3]))
    mode = int(rng.choice([0, 1, 3])) if mode is None else mode
    bs = bs or int(rng.choice([1, 3, 5, 7, 9, 11]))
    h = int(rng.integers(12, 70))
    w = int(rng.integers(D + minD + 12, D + minD + 150))
    return dict(h=h, w=w, cn=cn, p=dict(min_disparity=minD, num_disparities=D, block_size=bs, P1=8 * cn * bs * bs, P2=32 * cn * bs * bs,
                                        disp12_max_diff=int(rng.choice([-1, 0, 1, 2, 100])), uniqueness_ratio=int(rng.choice([0, 1, 5, 10, 15])),
                                        speckle_window_size=int(rng.choice([0, 20, 200])), speckle_range=2, mode=mode))


@pytest.mark.parametrize("seed", range(40))
def test_stage_parity_random(handle, seed):
    rng = np.random.default_rng(seed)
    c = _case(rng)
    if seed % 4 == 0:
        l = rng.integers(0, 256, (c["h"], c["w"], c["cn"]), dtype=np.uint8).squeeze()
        r = rng.integers(0, 256, (c["h"], c["w"], c["cn"]), dtype=np.uint8).squeeze()
    else:
        l, r, _ = synth.rectified_pair(c["h"], c["w"], c["p"]["num_disparities"], seed, c["cn"])
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **c["p"])
    m = cb.StereoSGBM(handle=handle, **c["p"])
    got = m.compute(l, r)
    assert np.array_equal(handle.fetch_volume(0), ref["C"]), "cost volume"
    assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume"
    if seed % 2 == 0:  # the default configuration (S not stored) and the separate winner-take-all kernel
        handle.keep_volumes(False)
        try:
            assert np.array_equal(m.compute(l, r), ref["disp"]), "winner-take-all fused into the last scan, S not stored"
            with pytest.raises(Exception, match="not stored"):
                handle.fetch_volume(1)
            handle.fuse_wta(False)
            assert np.array_equal(m.compute(l, r), ref["disp"]), "separate winner-take-all kernel"
            assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume stored by the unfused pass"
        finally:
            handle.keep_volumes(True)
            handle.fuse_wta(True)
    # (the device keeps the pre-L/R-check WTA map; the L/R check is fused into the median kernel's loads)
    assert np.array_equal(got, ref["disp"]), "WTA / uniqueness / subpixel / L-R check / median / speckle"
    f = m.compute_float(l, r)
    exp = ref["disp"].astype(np.float32).clip(0)
    exp[exp < c["p"]["min_disparity"] * 16] = 0
    assert f.dtype == np.float32 and np.array_equal(f, exp / 16.0)


@pytest.mark.parametrize("seed", range(24))
def test_wavefront_schedule(handle, seed, monkeypatch):
    """The wavefront schedule (sgbm_wave.cu; B2S_AGG_SCHEDULE=wave / B2S_OPT_AGG_SCHEDULE = 1) on images of several bands (hand-over between CTAs through the
    global ring), of fewer rows than a band, and taller than one wave of CTAs (bands wait for SMs)."""
    if seed % 2:
        monkeypatch.setenv("B2S_AGG_SCHEDULE", "wave")
    else:
        handle.call("b2s_set_option", 3, 1)
    rng = np.random.default_rng(300 + seed)
    c = _case(rng, mode=1)
    if seed % 3 == 0:
        c["h"] = int(rng.integers(1, 9))
    if seed % 8 == 7:
        c["h"] = 1500 if c["p"]["num_disparities"] > 128 else 2600  # more bands than CTAs fit on the device at once
        c["w"] = c["p"]["num_disparities"] + c["p"]["min_disparity"] + 40
    l, r, _ = synth.rectified_pair(c["h"], c["w"], c["p"]["num_disparities"], seed, c["cn"])
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **c["p"])
    try:
        got = cb.StereoSGBM(handle=handle, **c["p"]).compute(l, r)
        assert handle.volume_dims()[3] % 128 == 0, "the wavefront schedule (block layout) was not selected"
        assert np.array_equal(handle.fetch_volume(0), ref["C"]), "cost volume (block layout, natural order through the ABI)"
        assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume"
        assert np.array_equal(got, ref["disp"])
    finally:
        handle.call("b2s_set_option", 3, 0)


@pytest.mark.parametrize("grouped", ["1", "0"])
@pytest.mark.parametrize("seed", range(8))
def test_cost_kernel_lane_mappings(handle, seed, grouped, monkeypatch):
    """pixcost_hsum_kernel with the grouped lane mapping (four columns of one parity per warp, two adjacent words per lane; the default
    from 65 disparities on) and with the round-1 mapping (B2S_COST_GROUPED=0): both give the oracle's cost volume, for 65..256
    disparities incl. padded ranges, 1 and 3 channels, blocks 1..11, tiles cut by the image border."""
    monkeypatch.setenv("B2S_COST_GROUPED", grouped)
    rng = np.random.default_rng(4200 + seed)
    D = [65, 96, 128, 130, 192, 218, 250, 256][seed]
    c = _case(rng, D=D, mode=seed % 2)
    c["w"] = D + c["p"]["min_disparity"] + [12, 70, 129, 64, 200, 131, 66, 90][seed]
    l, r, _ = synth.rectified_pair(c["h"], c["w"], D, seed, c["cn"])
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **c["p"])
    got = cb.StereoSGBM(handle=handle, **c["p"]).compute(l, r)
    assert np.array_equal(handle.fetch_volume(0), ref["C"]), "cost volume"
    assert np.array_equal(got, ref["disp"])


@pytest.mark.parametrize("split", ["1", "0"])
@pytest.mark.parametrize("cols", ["1", "2", "3", "5", "13", "14", "32", "legacy"])
@pytest.mark.parametrize("seed", [1, 2, 3, 5, 6])
def test_aggregation_strip_handover(handle, seed, cols, split, monkeypatch):
    """The round-1 schedule (B2S_AGG_SCHEDULE=sweep): the fused vertical sweep cut into strips of 1..32 columns (several CTAs
    exchanging diagonal states through the global hand-over rings) and the legacy one-kernel-per-direction path all give the
    oracle's S volume.  split = 1: MODE_HH strips of 3..14 columns run agg_vsweep2_kernel (boundary columns shared with a helper
    warp); split = 0: the same cases through agg_vsweep_kernel."""
    monkeypatch.setenv("B2S_AGG_SCHEDULE", "sweep")
    monkeypatch.setenv("B2S_VSWEEP2", split)
    if cols == "legacy":
        monkeypatch.setenv("B2S_AGG_LEGACY", "1")
    else:
        monkeypatch.setenv("B2S_VSWEEP_COLS", cols)
    rng = np.random.default_rng(100 + seed)
    c = _case(rng, mode=seed % 2)
    l, r, _ = synth.rectified_pair(c["h"], c["w"], c["p"]["num_disparities"], seed, c["cn"])
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **c["p"])
    got = cb.StereoSGBM(handle=handle, **c["p"]).compute(l, r)
    assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume"
    assert np.array_equal(got, ref["disp"])


def test_golden_vectors(handle, golden_dir):
    g = np.load(os.path.join(golden_dir, "sgbm_small.npz"))
    l, r = g["left"], g["right"]
    common = dict(min_disparity=0, num_disparities=48, block_size=5, P1=8 * 3 * 25, P2=32 * 3 * 25, disp12_max_diff=1,
                  uniqueness_ratio=5, speckle_window_size=50, speckle_range=2)
    assert np.array_equal(cb.StereoSGBM(handle=handle, mode=0, **common).compute(l, r), g["disp_sgbm"])
    assert np.array_equal(cb.StereoSGBM(handle=handle, mode=1, **common).compute(l, r), g["disp_hh"])
    assert np.array_equal(cb.StereoSGBM(handle=handle, mode=cb.MODE_HH4, **common).compute(l, r), g["disp_hh4"])
    m = cb.StereoSGBM_create(minDisparity=2, numDisparities=40, blockSize=11, P1=968, P2=3872, disp12MaxDiff=0, uniquenessRatio=5,
                             speckleWindowSize=200, speckleRange=2, handle=handle)
    assert np.array_equal(m.compute(l, r), g["disp_refparams"])


@pytest.mark.parametrize("mode", [0, 1, 3])
def test_vs_cv2_live_640(handle, mode):
    cv2 = pytest.importorskip("cv2")
    l, r, _ = synth.rectified_pair(480, 640, 64, seed=7)
    kw = dict(minDisparity=0, numDisparities=64, blockSize=5, P1=8 * 3 * 25, P2=32 * 3 * 25, disp12MaxDiff=1, uniquenessRatio=5,
              speckleWindowSize=200, speckleRange=2, mode=mode)
    ref = cv2.StereoSGBM_create(**kw).compute(l, r)
    got = cb.StereoSGBM_create(handle=handle, **kw).compute(l, r)
    assert np.array_equal(ref, got)


def test_reference_default_matcher_vs_cv2(handle):
    """The reference's literal parameter set (stereo_matching.py:29-58): D=218 (not a multiple of 16), minD=2, 5-path."""
    cv2 = pytest.importorskip("cv2")
    l, r, _ = synth.rectified_pair(360, 720, 200, seed=11)
    ref = cv2.StereoSGBM_create(minDisparity=2, numDisparities=218, blockSize=11, uniquenessRatio=5, speckleWindowSize=200,
                                speckleRange=2, disp12MaxDiff=0, P1=968, P2=3872).compute(l, r)
    sm = cb.SemiGlobalBlockMatching({"max_size": 4000}, handle=handle)
    got16 = sm.stereo_sgbm.compute(l, r)
    assert np.array_equal(ref, got16)
    exp = ref.astype(np.float32).clip(0)
    exp[exp < 2 * 16] = 0
    assert np.array_equal(sm(l, r), exp / 16.0)


def test_full_size_1080p_hh(handle):
    """BASELINE config 2 at full size: bit-exact against cv2 when available, plus size-independent properties."""
    l, r, gt = synth.rectified_pair(1080, 1920, 128, seed=0)
    kw = dict(minDisparity=0, numDisparities=128, blockSize=5, P1=8 * 3 * 25, P2=32 * 3 * 25, disp12MaxDiff=1, uniquenessRatio=5,
              speckleWindowSize=200, speckleRange=2, mode=1)
    m = cb.StereoSGBM_create(handle=handle, **kw)
    got = m.compute(l, r)
    assert np.array_equal(got, m.compute(l, r)), "deterministic"
    assert (got[:, :128] == -16).all(), "columns x < minD+D are always invalid"
    valid = got >= 0
    assert valid.mean() > 0.7
    err = np.abs(got[valid] / 16.0 - gt[valid])
    assert (err <= 1).mean() > 0.97, "accuracy against the synthetic ground truth"
    # vertical flip equivariance: the algorithm is symmetric under y -> H-1-y
    assert np.array_equal(m.compute(l[::-1].copy(), r[::-1].copy()), got[::-1])
    try:
        import cv2
    except ImportError:
        return
    assert np.array_equal(cv2.StereoSGBM_create(**kw).compute(l, r), got)


def test_errors(handle):
    l = np.zeros((20, 30, 3), np.uint8)
    m = cb.StereoSGBM(handle=handle, num_disparities=32, block_size=5)
    with pytest.raises(ValueError, match="too small"):
        m.compute(l, l)
    with pytest.raises(ValueError):
        m.compute(l, l[:, :20])
    with pytest.raises(ValueError):
        m.compute(l.astype(np.float32), l.astype(np.float32))
    with pytest.raises(ValueError):
        cb.StereoSGBM(handle=handle, num_disparities=0)
    with pytest.raises(ValueError):
        cb.StereoSGBM(handle=handle, num_disparities=16, mode=2)


@pytest.mark.parametrize("uniq", [1, 50, 90, 99, 100, 150])
@pytest.mark.parametrize("D", [32, 70, 128])
def test_extreme_uniqueness(handle, uniq, D):
    """The fused WTA turns S(d) * (100 - uniq) < minS * 100 into S(d) < ceil(minS * 100 / (100 - uniq)) with a multiply-high
    division; ratios >= 100 fall back to the separate kernel.  Noise and textured pairs, gray and RGB."""
    rng = np.random.default_rng(uniq * 7 + D)
    for k in range(2):
        if k == 0:
            l, r, _ = synth.rectified_pair(40, D + 90, D, uniq, 1)
        else:
            l = rng.integers(0, 256, (33, D + 50, 3), dtype=np.uint8)
            r = rng.integers(0, 256, (33, D + 50, 3), dtype=np.uint8)
        p = dict(num_disparities=D, block_size=3, P1=72, P2=288, disp12_max_diff=1, uniqueness_ratio=uniq, mode=k)
        assert np.array_equal(cb.StereoSGBM(handle=handle, **p).compute(l, r), osgbm.sgbm_compute(l, r, **p)), (uniq, D, k)


@pytest.mark.parametrize("P1,P2", [(100, 8000), (4000, 12000), (11000, 12000), (1, 2)])
def test_large_penalties(handle, P1, P2):
    """Penalties far above the reference's, inside the domain where cv2 itself is well defined: max C + 2 * P2 <= 32767, so that
    no int16 sum inside cv2's SIMD adds saturates (outside it cv2's result depends on its SIMD width; tests/test_oracle.py)."""
    for k, (cn, bs) in enumerate([(1, 3), (3, 5)]):
        l, r, _ = synth.rectified_pair(36, 150, 48, 40 + k, cn)
        if k:
            r = np.random.default_rng(3).integers(0, 256, r.shape, dtype=np.uint8)
        for mode in (0, 1, 3):
            p = dict(num_disparities=48, block_size=bs, P1=P1, P2=P2, disp12_max_diff=1, uniqueness_ratio=5, mode=mode)
            ref = osgbm.sgbm_compute(l, r, want_volumes=True, **p)
            assert int(ref["C"].max()) + 2 * P2 <= 32767
            got = cb.StereoSGBM(handle=handle, **p).compute(l, r)
            assert np.array_equal(handle.fetch_volume(1), ref["S"]), (cn, bs, mode, "S")
            assert np.array_equal(got, ref["disp"]), (cn, bs, mode)


def test_textureless_and_saturated(handle):
    """Constant images (all costs tie) and maximal-contrast noise (S saturates at 32767 with the reference's P2)."""
    for l, r in [(np.full((40, 300, 3), 128, np.uint8),) * 2,
                 tuple(np.random.default_rng(s).integers(0, 2, (40, 300, 3), dtype=np.uint8) * 255 for s in (1, 2))]:
        for mode in (0, 1):
            p = dict(min_disparity=2, num_disparities=218, block_size=11, P1=968, P2=3872, uniqueness_ratio=0, disp12_max_diff=0,
                     speckle_window_size=200, speckle_range=2, mode=mode)
            ref = osgbm.sgbm_compute(l, r, **p)
            assert np.array_equal(cb.StereoSGBM(handle=handle, **p).compute(l, r), ref)


def test_full_size_4k_256(handle):
    """BASELINE config 4's geometry (3840x2160 gray, 256 disparities, 8 paths) with the BT cost (the census cost named there
    has no oracle in this OpenCV build): strips of 25 columns, one sweep per launch, NP = 4 registers per lane.
    Size-independent properties, and bit-exact against cv2 when it is available on the box (~25 s of CPU)."""
    l, r, gt = synth.rectified_pair(2160, 3840, 256, seed=1, cn=1)
    kw = dict(minDisparity=0, numDisparities=256, blockSize=5, P1=8 * 25, P2=32 * 25, disp12MaxDiff=1, uniquenessRatio=5,
              speckleWindowSize=200, speckleRange=2, mode=1)
    m = cb.StereoSGBM_create(handle=handle, **kw)
    got = m.compute(l, r)
    assert (got[:, :256] == -16).all()
    valid = got >= 0
    assert valid.mean() > 0.6
    assert (np.abs(got[valid] / 16.0 - gt[valid]) <= 1).mean() > 0.95
    assert np.array_equal(m.compute(l[::-1].copy(), r[::-1].copy()), got[::-1]), "vertical flip equivariance"
    try:
        import cv2
    except ImportError:
        return
    assert np.array_equal(cv2.StereoSGBM_create(**kw).compute(l, r), got)


def test_batch_engine_streams():
    """DisparityBatchEngine: several pairs in flight on several streams (pinned caller buffers used in place, others staged)
    give exactly the per-pair results, in order."""
    from calibrating_b200 import _ffi
    from calibrating_b200.batch import DisparityBatchEngine
    p = dict(min_disparity=0, num_disparities=64, block_size=5, P1=600, P2=2400, disp12_max_diff=1, uniqueness_ratio=5,
             speckle_window_size=100, speckle_range=2, mode=1)
    pairs = [synth.rectified_pair(120, 320, 64, seed=s)[:2] for s in range(7)]
    ref = [osgbm.sgbm_compute(l, r, **p) for l, r in pairs]
    eng = DisparityBatchEngine(p, device=0, streams=3)
    try:
        got = eng.compute_batch(pairs, as_float=False)  # pageable inputs, engine-owned outputs
        assert all(np.array_equal(g, e) for g, e in zip(got, ref))
        pinned, outs = [], []
        for l, r in pairs:
            a, b = _ffi.pinned_empty(l.shape, np.uint8), _ffi.pinned_empty(r.shape, np.uint8)
            a[...] = l
            b[...] = r
            pinned.append((a, b))
            outs.append(_ffi.pinned_empty((120, 320), np.float32))
        got = eng.compute_batch(pinned, out=outs)       # everything pinned: no synchronisation until the end
        for g, e in zip(got, ref):
            f = e.astype(np.float32).clip(0) / 16.0
            assert np.array_equal(g, f)
    finally:
        eng.close()


@pytest.mark.parametrize("seed", range(6))
def test_census_cost_vs_own_restatement(handle, seed):
    """cost=COST_CENSUS (BASELINE config 4's cost function).  cv2 has no census matcher, so parity is UNPINNED: the CUDA path is
    compared bit for bit with this repo's own C restatement (oracle/sgbm_ref.c, cost = 1), and with the ground truth."""
    rng = np.random.default_rng(500 + seed)
    D = int(rng.choice([16, 48, 64, 100, 128, 256]))
    minD = int(rng.choice([0, 3]))
    h, w = int(rng.integers(12, 60)), int(rng.integers(D + minD + 12, D + minD + 120))
    p = dict(min_disparity=minD, num_disparities=D, block_size=5, P1=10, P2=120, disp12_max_diff=1, uniqueness_ratio=int(rng.choice([0, 5, 10])),
             speckle_window_size=int(rng.choice([0, 50])), speckle_range=2, mode=seed % 2, cost=cb.COST_CENSUS)
    l, r, _ = synth.rectified_pair(h, w, D, seed, 1)
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **p)
    got = cb.StereoSGBM(handle=handle, **p).compute(l, r)
    assert np.array_equal(handle.fetch_volume(0), ref["C"]), "census cost volume"
    assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume"
    assert np.array_equal(got, ref["disp"])
    with pytest.raises(ValueError, match="gray"):
        cb.StereoSGBM(handle=handle, **p).compute(np.zeros((h, w, 3), np.uint8), np.zeros((h, w, 3), np.uint8))


def test_census_4k_256_accuracy(handle):
    """BASELINE config 4 as named: 3840x2160 gray, 256 disparities, census cost, 8 paths; accuracy against the synthetic
    ground truth (parity unpinned, see above) and the size-independent properties."""
    l, r, gt = synth.rectified_pair(2160, 3840, 256, seed=2, cn=1)
    m = cb.StereoSGBM_create(minDisparity=0, numDisparities=256, blockSize=5, P1=10, P2=120, disp12MaxDiff=1, uniquenessRatio=5,
                             speckleWindowSize=200, speckleRange=2, mode=cb.MODE_HH, cost=cb.COST_CENSUS, handle=handle)
    got = m.compute(l, r)
    assert (got[:, :256] == -16).all()
    valid = got >= 0
    assert valid.mean() > 0.6
    assert (np.abs(got[valid] / 16.0 - gt[valid]) <= 1).mean() > 0.97
    assert np.array_equal(m.compute(l[::-1].copy(), r[::-1].copy()), got[::-1]), "vertical flip equivariance"
    t = handle.timings()
    print("census 4K/256 stage ms:", {k: round(v, 3) for k, v in t.items() if k.endswith("_ms")})


@pytest.mark.parametrize("shape", [(1, 40), (2, 41), (3, 60), (5, 37), (9, 300)])
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_degenerate_shapes(handle, shape, mode):
    """Images of one to a few rows and cost volumes a few columns wide (fewer rows than the cp.async ring is deep, strips of a
    single column pair, block windows larger than the image) against the oracle."""
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w + mode)
    l = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    r = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for D, bs in ((16, 11), (32, 3)):
        if w - D <= bs // 2:
            continue
        p = dict(min_disparity=0, num_disparities=D, block_size=bs, P1=8 * 3 * bs * bs, P2=32 * 3 * bs * bs, disp12_max_diff=1,
                 uniqueness_ratio=5, speckle_window_size=10, speckle_range=2, mode=mode)
        ref = osgbm.sgbm_compute(l, r, want_volumes=True, **p)
        got = cb.StereoSGBM(handle=handle, **p).compute(l, r)
        assert np.array_equal(handle.fetch_volume(0), ref["C"]), (D, bs, "C")
        assert np.array_equal(handle.fetch_volume(1), ref["S"]), (D, bs, "S")
        assert np.array_equal(got, ref["disp"]), (D, bs)


@pytest.mark.parametrize("schedule", ["sweep", "legacy"])
def test_cost_domain_is_checked(handle, schedule, monkeypatch):
    """The packed unsigned arithmetic of the aggregation needs C >= 0.  A block sum wraps past 32767 only with large windows on
    adversarial input (block 11 x 3 channels: at most 121 * 279 = 33759): the cost stage flags it on the device and the call
    fails instead of returning a silently different result.  The reference's block 11 on an adversarial RGB pair (inverted
    binary noise) stays inside the domain and is exact."""
    monkeypatch.setenv("B2S_AGG_SCHEDULE", schedule)
    rng = np.random.default_rng(5)
    l = rng.integers(0, 2, (40, 260, 3), dtype=np.uint8) * 255
    r = 255 - l
    bad = dict(min_disparity=0, num_disparities=64, block_size=21, P1=968, P2=3872, mode=1)
    assert osgbm.sgbm_compute(l, r, want_volumes=True, **bad)["C"].min() < 0, "the test input does not wrap C"
    with pytest.raises(Exception, match="wrapped past 32767"):
        cb.StereoSGBM(handle=handle, **bad).compute(l, r)
    ok = dict(bad, block_size=11)
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **ok)
    assert ref["C"].min() >= 0
    assert np.array_equal(cb.StereoSGBM(handle=handle, **ok).compute(l, r), ref["disp"])  # (the failed call cleared the sticky flag)


def test_cooperative_sweep_launch(handle, monkeypatch):
    """B2S_SWEEP_COOPERATIVE=1 (automatic under MPS): the lock-step sweep is launched cooperatively, so the driver guarantees the
    co-residency its spin-waits need; same results."""
    monkeypatch.setenv("B2S_SWEEP_COOPERATIVE", "1")
    l, r, _ = synth.rectified_pair(120, 700, 128, seed=9)
    p = dict(min_disparity=0, num_disparities=128, block_size=5, P1=600, P2=2400, disp12_max_diff=1, uniqueness_ratio=5,
             speckle_window_size=100, speckle_range=2, mode=1)
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **p)
    got = cb.StereoSGBM(handle=handle, **p).compute(l, r)
    assert np.array_equal(handle.fetch_volume(1), ref["S"]) and np.array_equal(got, ref["disp"])


@pytest.mark.parametrize("seed", range(16))
def test_negative_min_disparity(handle, seed):
    """cv2 accepts a negative minDisparity (the search range reaches to the right of the left pixel); so does the engine: C, S and
    the disparity equal the oracle's, which is pinned against cv2 for these parameters (tests/test_oracle.py)."""
    rng = np.random.default_rng(700 + seed)
    c = _case(rng, minD=int(rng.choice([-1, -3, -16, -40, -64, -130])))
    l, r, _ = synth.rectified_pair(c["h"], c["w"] + 140, c["p"]["num_disparities"], seed, c["cn"])
    if seed % 3 == 0:
        r = np.roll(r, -int(rng.integers(1, 12)), axis=1)  # true disparities on both sides of zero
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **c["p"])
    got = cb.StereoSGBM(handle=handle, **c["p"]).compute(l, r)
    assert np.array_equal(handle.fetch_volume(0), ref["C"]), "cost volume"
    assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume"
    assert np.array_equal(got, ref["disp"])


@pytest.mark.parametrize("D,mode,cn,bs", [(300, 1, 1, 5), (384, 0, 3, 3), (512, 1, 3, 7), (260, 3, 1, 9), (512, 0, 1, 11)])
def test_more_than_256_disparities(handle, D, mode, cn, bs):
    """cv2 takes any numDisparities; above 256 the engine runs the generic one-scan-per-direction schedule (NP = 5..8 packed
    registers per lane).  Bit-exact against the oracle and, live, against cv2."""
    cv2 = pytest.importorskip("cv2")
    l, r, _ = synth.rectified_pair(40, D + 120, D, seed=D, cn=cn)
    p = dict(min_disparity=2, num_disparities=D, block_size=bs, P1=8 * cn * bs * bs, P2=32 * cn * bs * bs, disp12_max_diff=1, uniqueness_ratio=5,
             speckle_window_size=30, speckle_range=2, mode=mode)
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **p)
    got = cb.StereoSGBM(handle=handle, **p).compute(l, r)
    assert np.array_equal(handle.fetch_volume(0), ref["C"]), "cost volume"
    assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume"
    assert np.array_equal(got, ref["disp"])
    live = cv2.StereoSGBM_create(minDisparity=2, numDisparities=D, blockSize=bs, P1=p["P1"], P2=p["P2"], disp12MaxDiff=1, uniquenessRatio=5,
                                 speckleWindowSize=30, speckleRange=2, mode=mode).compute(l, r)
    assert np.array_equal(got, live)
    with pytest.raises(ValueError, match="numDisparities > 512"):
        cb.StereoSGBM(handle=handle, **dict(p, num_disparities=528)).compute(np.zeros((8, 700), np.uint8), np.zeros((8, 700), np.uint8))


@pytest.mark.parametrize("cols", ["2", "3", "5", "16", "default", "old"])
@pytest.mark.parametrize("seed", range(6))
def test_six_path_sweep(handle, seed, cols, monkeypatch):
    """agg_vsweep6_kernel (sgbm_sweep6.cu: MODE_HH, 65..128 disparities, block layout, all six row-crossing paths of a column in
    one warp): strips of 2..16 columns (several CTAs exchanging diagonal states through the global hand-over rings), padded
    disparity ranges, images narrower than a strip; `old` = the same cases through the default agg_vsweep_kernel."""
    if cols != "old":
        monkeypatch.setenv("B2S_SWEEP6", "1")  # (opt-in: the kernel is exact but slower than agg_vsweep_kernel)
    if cols not in ("old", "default"):
        monkeypatch.setenv("B2S_VSWEEP_COLS", cols)
    rng = np.random.default_rng(900 + seed)
    D = [70, 100, 128, 65, 127, 96][seed]
    c = _case(rng, D=D, mode=1)
    if seed == 4:
        c["w"] = D + c["p"]["min_disparity"] + 7  # 7 cost columns
    l, r, _ = synth.rectified_pair(c["h"], c["w"], D, seed, c["cn"])
    ref = osgbm.sgbm_compute(l, r, want_volumes=True, **c["p"])
    got = cb.StereoSGBM(handle=handle, **c["p"]).compute(l, r)
    assert (handle.volume_dims()[3] == 128) == (cols != "old" or D > 64), "block layout (Dp = 128) <-> six-path sweep"
    assert np.array_equal(handle.fetch_volume(0), ref["C"]), "cost volume"
    assert np.array_equal(handle.fetch_volume(1), ref["S"]), "aggregated volume"
    assert np.array_equal(got, ref["disp"])
