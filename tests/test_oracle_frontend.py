"""CPU tests: pin the numpy spec (oracle/frontend_spec.py) against the real
OpenCV 4.13 that the reference links, the std::sort restatement against the
real libstdc++, and check that the C-ABI library exports every declared symbol.
(The reference ships no tests or golden vectors -- SURVEY.md section 4 -- so cv2
in this image *is* the pin for the third-party arithmetic.)"""
import ctypes
import re
import os

import cv2
import numpy as np
import pytest

from oracle import frontend_spec as S
from oracle import frontend_ref as FR
from vrf_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tex(h, w, seed):
    return synth.band_limited_texture(h, w, seed)


@pytest.mark.parametrize("shape", [(480, 640), (121, 77), (60, 85), (720, 1280)])
def test_pyrdown_bit_exact(shape):
    img = tex(*shape, seed=1)
    assert np.array_equal(S.pyr_down(img), cv2.pyrDown(img))
    _, cvp = cv2.buildOpticalFlowPyramid(img, (21, 21), 3, withDerivatives=False)
    mine = S.build_pyramid(img, len(cvp) - 1)
    for a, b in zip(mine, cvp):
        assert np.array_equal(a, b)


def test_scharr_bit_exact():
    img = tex(100, 130, 2)
    dx, dy = S.scharr_deriv(img)
    assert np.array_equal(dx, cv2.Scharr(img, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101))
    assert np.array_equal(dy, cv2.Scharr(img, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101))


def test_rgb2gray_bit_exact():
    r = np.random.default_rng(0)
    rgb = r.integers(0, 256, (96, 128, 3)).astype(np.uint8)
    assert np.array_equal(synth.rgb_to_gray(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))


@pytest.mark.parametrize("rect", [(0, 0, 83, 71), (77, 65, 86, 74), (557, 405, 83, 75), (0, 0, 640, 480)])
def test_fast_bit_exact(rect):
    det = cv2.FastFeatureDetector_create()
    assert (det.getThreshold(), det.getNonmaxSuppression(), det.getType()) == (10, True, 2)
    img = tex(480, 640, 3)
    x, y, w, h = rect
    roi = img[y:y + h, x:x + w]
    mask = np.full((h, w), 255, np.uint8)
    cv2.circle(mask, (w // 2, h // 2), 25, 0, -1)
    for m in (None, mask):
        kps = det.detect(roi, m)
        got = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps]
        assert got == S.fast_detect(roi, mask=m)


@pytest.mark.parametrize("shape,seed", [((480, 640), 1), ((240, 320), 2), ((720, 1280), 3), ((480, 640), 4), ((64, 64), 5),
                                        ((250, 336), 6), ((483, 640), 7), ((480, 644), 8), ((77, 85), 9)])
def test_clahe_bit_exact(shape, seed):
    """EQUALIZE: the numpy spec of cv::createCLAHE(3.0, Size(8, 8)) equals the real OpenCV, incl. a dark low-contrast frame
    (heavy clipping + residual redistribution), a frame of constant blocks, and sizes that are not multiples of the 8x8 tile
    grid (OpenCV's REFLECT_101 extension, which adds a full tile row/column to a dimension that already divides)."""
    img = tex(*shape, seed=seed)
    if seed == 4:
        img = (img.astype(np.float32) * 0.3 + 20).astype(np.uint8)
    if seed == 5:
        img = np.kron(np.random.default_rng(5).integers(0, 256, (8, 8)), np.ones((8, 8))).astype(np.uint8)
    assert np.array_equal(S.clahe(img), cv2.createCLAHE(3.0, (8, 8)).apply(img))


def test_fast_flat_and_tiny():
    assert S.fast_detect(np.full((40, 40), 128, np.uint8)) == []
    assert S.fast_detect(np.zeros((6, 6), np.uint8)) == []


@pytest.mark.parametrize("r", [1, 10, 20, 25, 30, 60])
def test_filled_circle_is_disc(r):
    m = np.full((200, 200), 255, np.uint8)
    cv2.circle(m, (90, 110), r, 0, -1)
    yy, xx = np.mgrid[0:200, 0:200]
    assert np.array_equal(m == 0, S.circle_covers(90, 110, r, xx, yy))


@pytest.mark.parametrize("ml,use_init", [(1, True), (3, False), (2, True)])
def test_lk_spec_matches_cv2(ml, use_init):
    rng = np.random.default_rng(7)
    h, w = 480, 640
    img0 = tex(h, w, 5)
    M = np.array([[1.002, 0.004, 2.3], [-0.003, 0.999, -1.7]])
    img1 = cv2.warpAffine(img0, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    pts = rng.uniform([0, 0], [w, h], (40, 2)).astype(np.float32)
    pts[:6] = [[0.5, 0.7], [639, 479], [3, 470], [635, 2], [10.2, 11.9], [320, 0]]
    init = pts + rng.normal(0, 1.0, pts.shape).astype(np.float32)
    a, st, _ = cv2.calcOpticalFlowPyrLK(
        img0, img1, pts, init.copy() if use_init else None, winSize=(21, 21), maxLevel=ml,
        criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
        flags=cv2.OPTFLOW_USE_INITIAL_FLOW if use_init else 0)
    b, sb = S.lk_track(S.build_pyramid(img0, ml), S.build_pyramid(img1, ml), pts, init, ml, use_initial_flow=use_init)
    assert np.array_equal(st.ravel(), sb)
    assert np.abs(a - b).max() <= 1e-3          # float summation order only


def test_stdsort_restatement_matches_libstdcxx(built_lib):
    """csrc/introsort.h (product) == real std::sort (oracle/stdsort.cpp), incl. heavy ties,
    sizes around the 16-element insertion threshold, and a median-of-3 killer that
    drives introsort into its heapsort fallback."""
    from vrf_b200 import binding
    rng = np.random.default_rng(3)
    cases = []
    for n in [0, 1, 2, 15, 16, 17, 31, 33, 64, 150, 151, 300, 500, 1000]:
        cases.append(rng.integers(1, 4, n))
        cases.append(rng.integers(1, 40, n))
        cases.append(np.arange(n))
        cases.append(np.arange(n)[::-1])
        cases.append(np.ones(n, np.int64))
    # median-of-3 killer (descending comparator => negate)
    for n in [128, 512, 1000]:
        k = n // 2
        a = np.zeros(n, np.int64)
        for i in range(1, k + 1):
            if i % 2 == 1:
                a[i - 1] = i
                a[i] = k + i
            a[k + i - 1] = 2 * i
        cases.append(-a)
    for c in cases:
        c = np.asarray(c, np.int32)
        assert np.array_equal(binding.sort_desc_perm(c), FR.stdsort_desc_perm(c)), len(c)


def test_abi_exports_every_declared_symbol(built_lib):
    from vrf_b200 import binding
    declared = set()
    for hdr in ("vrf.h", "vrf_ba.h", "vrf_fm.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        declared |= set(re.findall(r"\b(vrf_[a-z0-9_]+)\s*\(", txt))
    declared.discard("vrf_handle")
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)
    for name in declared:
        assert hasattr(built_lib, name), name


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device vrf_create must fail loudly (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vrf_b200 import binding
    cfg = binding.default_config()
    hp = ctypes.c_void_p()
    rc = built_lib.vrf_create(ctypes.byref(cfg), 1, 0, ctypes.byref(hp))
    assert rc == -5 and not hp.value


def test_oracle_tracker_runs_and_ids_are_unique():
    seq, frames, rels = synth.render_gray_frames(99, 7)
    ft = FR.FeatureTrackerRef(FR.FrontendConfig())
    for k in range(7):
        ft.read_image(frames[k], seq.time(k), rels[k], pub_this_frame=(k % 3 == 0))
        assert len(set(ft.ids)) == len(ft.ids)
        assert len(ft.ids) == len(ft.cur_pts) == len(ft.track_cnt) == len(ft.pts_velocity)
    assert max(ft.track_cnt) == 7
    # min-distance invariant established by setMask/addPoints on publish frames
    p = np.rint(ft.cur_pts)
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1) + np.eye(len(p)) * 1e9
    assert len(ft.ids) > 100


def test_depth_decode_32fc1_spec_matches_opencv_float_pipeline():
    """estimator_nodelet.cpp:523-527 (convertTo(CV_16UC1, 1000)): the numpy spec the kernel follows equals OpenCV's
    float32 scale + saturate_cast<ushort> pipeline, incl. NaN / inf / negative / out-of-range inputs."""
    from oracle.frontend_ref import decode_depth, decode_depth_numpy
    rng = np.random.default_rng(5)
    d = (rng.random((480, 640)).astype(np.float32) * 12.0 - 1.0)
    d[0, :8] = [np.nan, np.inf, -np.inf, 1e8, -5.0, 70.0, 3e6, 65.535]
    d[1, :4] = [0.0005, 0.0015, 0.0025, 65.5355]           # half-way cases (round half to even)
    assert np.array_equal(decode_depth(d, 480, 640), decode_depth_numpy(d))
    assert decode_depth(None, 4, 6).shape == (4, 6) and not decode_depth(None, 4, 6).any()
    u16 = rng.integers(0, 65535, (4, 6)).astype(np.uint16)
    assert decode_depth(u16, 4, 6) is u16


def test_depth_lookup_truncates_and_culls():
    """feature_manager.cpp:71-80: (int) truncation of (v, u); 0 < d < DEPTH_MIN_DIST is erased, 0 (invalid) is kept."""
    from oracle.frontend_ref import depth_lookup
    dep = np.zeros((10, 12), np.uint16)
    dep[3, 5] = 299; dep[3, 6] = 300; dep[4, 5] = 0; dep[4, 6] = 4000
    pts = np.array([[5.99, 3.99], [6.0, 3.5], [5.5, 4.2], [6.9, 4.9]], np.float32)
    mm, keep = depth_lookup(dep, pts, 0.3)
    assert mm.tolist() == [299, 300, 0, 4000]
    assert keep.tolist() == [0, 1, 1, 1]


def test_vectorised_glue_equals_per_point_evaluation():
    """The tracker restatement evaluates liftProjective / spaceToPlane / R*p / inBorder on whole point arrays (so that the
    timed CPU baseline is not interpreter time); the array forms must be bit-identical to the per-point statements of
    PinholeCamera.cc:450-543 and feature_tracker.cpp:96-103,595-608."""
    from oracle.frontend_ref import FeatureTrackerRef, FrontendConfig, PinholeCamera, cv_round
    cfg = FrontendConfig()
    cam = PinholeCamera(cfg)
    r = np.random.default_rng(5)
    pts = np.stack([r.uniform(-5, 645, 400), r.uniform(-5, 485, 400)], 1).astype(np.float32)
    pts[:8] = [[0.5, 0.5], [1.5, 2.5], [638.5, 478.5], [637.5, 477.5], [639.49, 10], [10, 479.49], [0.49, 3], [3, 0.51]]
    x, y = cam.lift_projective_pts(pts)
    for i in range(len(pts)):
        xs, ys, _ = cam.lift_projective(float(pts[i, 0]), float(pts[i, 1]))
        assert xs == x[i] and ys == y[i]
    # R * [x, y, 1] and the re-projection
    th = 0.01
    R = np.array([[np.cos(th), -np.sin(th), 0.001], [np.sin(th), np.cos(th), -0.002], [-0.001, 0.002, 1.0]])
    ft = FeatureTrackerRef(cfg)
    ft.cur_pts = pts
    ft.predict_pts_in_next_frame(R)
    for i in range(len(pts)):
        xs, ys, zs = cam.lift_projective(float(pts[i, 0]), float(pts[i, 1]))
        P = R @ np.array([xs, ys, zs], np.float64)
        u, v = cam.space_to_plane(P[0], P[1], P[2])
        assert np.float32(u) == ft.predict_pts[i, 0] and np.float32(v) == ft.predict_pts[i, 1]
    ib = ft.in_border_pts(pts)
    for i in range(len(pts)):
        assert bool(ib[i]) == ft.in_border(pts[i])
