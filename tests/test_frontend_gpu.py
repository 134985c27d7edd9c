"""GPU parity tests of the front end: the CUDA path called through the C ABI
(vrf_tracker_read_image*) against the cv2-backed oracle (oracle/frontend_ref.py)
on identical seeded synthetic frames.

Bars (BASELINE.json north_star): feature IDs, track counts and grid-cell
occupancy bit-exact; LK status identical; tracked (u,v) within 1e-3 px;
undistorted points / velocities within float tolerance (1e-6 / 1e-4)."""
import numpy as np
import pytest

from oracle.frontend_ref import FeatureTrackerRef, FrontendConfig
from vrf_b200 import binding, synth

pytestmark = pytest.mark.gpu

PX_TOL = 1e-3       # px, north_star
UN_TOL = 2e-6       # normalised-plane units (float32 rounding of a double computation)


def run_pair(cfg_over, n_frames, seed, pub_every=3, cam=None, ransac=0, rgb=False):
    cam = cam or synth.CamModel()
    seq = synth.Sequence(seed, cam)
    cfg = binding.default_config(row=cam.height, col=cam.width, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy,
                                 k1=cam.k1, k2=cam.k2, p1=cam.p1, p2=cam.p2, use_ransac=ransac, **cfg_over)
    h = binding.Handle(cfg, 1, 0)
    ref = FeatureTrackerRef(FrontendConfig(
        row=cam.height, col=cam.width, max_cnt=cfg.max_cnt, min_dist=cfg.min_dist, use_imu=cfg.use_imu,
        num_grid_rows=cfg.num_grid_rows, num_grid_cols=cfg.num_grid_cols, lk_max_level=cfg.lk_max_level,
        fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, k1=cam.k1, k2=cam.k2, p1=cam.p1, p2=cam.p2,
        use_ransac=ransac, f_threshold=cfg.f_threshold))
    for k in range(n_frames):
        rgbf, gray, _ = seq.frame(k)
        R = seq.relative_R(k)
        pub = (k % pub_every == 0)
        out = h.read_image(0, rgbf if rgb else gray, seq.time(k), R, pub=pub)
        ref.read_image(gray, seq.time(k), R, pub_this_frame=pub)
        yield k, out, ref
    h.close()


def check_frame(k, out, ref):
    if ref.last_lk_pts is not None:
        assert out.n_predict == len(ref.last_lk_pts), k
        if len(ref.predict_pts) == len(out.predict_pts):      # predict_pts only exists with USE_IMU
            assert np.abs(out.predict_pts - ref.predict_pts).max() <= 1e-4, k
        assert np.array_equal(out.lk_status, ref.last_lk_status), (k, np.nonzero(out.lk_status != ref.last_lk_status))
        ok = ref.last_lk_status.astype(bool)
        assert np.abs(out.lk_pts[ok] - ref.last_lk_pts[ok]).max() <= PX_TOL, k
    assert out.n == len(ref.ids), (k, out.n, len(ref.ids))
    assert np.array_equal(out.ids, np.asarray(ref.ids, np.int32)), k            # bit-exact feature IDs
    assert np.array_equal(out.track_cnt, np.asarray(ref.track_cnt, np.int32)), k
    assert out.n_id == ref.n_id, k
    assert np.abs(out.cur_pts - ref.cur_pts).max() <= PX_TOL, k
    assert np.abs(out.cur_un_pts - ref.cur_un_pts).max() <= UN_TOL, k
    assert np.abs(out.pts_velocity - ref.pts_velocity).max() <= 1e-3, k
    assert np.array_equal(out.grids_track_num, np.asarray(ref.grids_track_num, np.int32)), k   # occupancy
    assert np.array_equal(out.grids_texture_status.astype(bool), np.asarray(ref.grids_texture_status)), k
    assert out.n_unstable == len(ref.unstable_pts), k


def test_sequence_parity_reference_default():
    """640x480, 150 feats, IMU-predicted LK maxLevel 1 (reference default), publish every 3rd frame."""
    n = 0
    for k, out, ref in run_pair({}, 12, 1234):
        check_frame(k, out, ref)
        n += 1
    assert n == 12


def test_sequence_parity_with_ransac_long():
    """Reference behaviour incl. rejectWithF (cv::findFundamentalMat RANSAC) on publish frames,
    24 frames: IDs stay bit-exact over the whole run."""
    n = 0
    for k, out, ref in run_pair({}, 24, 4242, ransac=1):
        check_frame(k, out, ref)
        if ref.last_ransac_status is not None:
            n += 1
    assert n >= 6


def test_lk_tracks_bit_exact():
    """The LK kernel accumulates in OpenCV's SIMD lane order => raw tracks are bit-identical to
    cv2.calcOpticalFlowPyrLK (no drift between the two pipelines)."""
    for k, out, ref in run_pair({"lk_max_level": 2}, 10, 2024):
        if ref.last_lk_pts is not None:
            assert np.array_equal(out.lk_status, ref.last_lk_status), k
            assert np.array_equal(out.lk_pts, ref.last_lk_pts), (k, np.abs(out.lk_pts - ref.last_lk_pts).max())
        assert np.array_equal(out.cur_pts, ref.cur_pts), k


def test_sequence_parity_three_level_pyramid():
    """BASELINE config 2: 3-level pyramid (lk_max_level=2)."""
    for k, out, ref in run_pair({"lk_max_level": 2}, 8, 77):
        check_frame(k, out, ref)


def test_sequence_parity_no_imu_four_levels():
    """USE_IMU=0 branch: maxLevel 3, no initial flow (feature_tracker.cpp:309-310)."""
    for k, out, ref in run_pair({"use_imu": 0}, 7, 5, pub_every=1):
        check_frame(k, out, ref)


def test_sequence_parity_with_clahe():
    """EQUALIZE = 1: cv::createCLAHE(3.0, Size(8, 8)) on every frame before tracking/detection; a dim, low-contrast
    sequence (the case the option exists for), IDs / occupancy bit-exact, tracks within 1e-3 px."""
    cam = synth.CamModel()
    seq = synth.Sequence(2718, cam)
    cfg = binding.default_config(use_ransac=1, equalize=1)
    h = binding.Handle(cfg, 1, 0)
    ref = FeatureTrackerRef(FrontendConfig(use_ransac=1, equalize=1))
    for k in range(9):
        gray = (seq.frame(k)[1].astype(np.float32) * 0.35 + 15).astype(np.uint8)
        R = seq.relative_R(k)
        out = h.read_image(0, gray, seq.time(k), R, pub=(k % 3 == 0))
        ref.read_image(gray, seq.time(k), R, pub_this_frame=(k % 3 == 0))
        check_frame(k, out, ref)
    assert out.n > 60
    h.close()


def test_sequence_parity_with_fisheye_mask():
    """FISHEYE = 1: setMask starts from fisheye_mask (255 inside a disc, a grey ring where FAST may detect but nothing is
    kept, 0 outside): tracked points and new corners only survive where the mask is 255."""
    cam = synth.CamModel()
    seq = synth.Sequence(1618, cam)
    yy, xx = np.mgrid[0:cam.height, 0:cam.width]
    rr = np.hypot(xx - cam.width / 2 + 0.5, yy - cam.height / 2 + 0.5)
    fm = np.where(rr < 230, 255, np.where(rr < 260, 128, 0)).astype(np.uint8)
    cfg = binding.default_config(use_ransac=1, fisheye=1)
    h = binding.Handle(cfg, 1, 0)
    with pytest.raises(RuntimeError):                       # no mask yet: refused, not ignored
        h.read_image(0, seq.frame(0)[1], seq.time(0), np.eye(3), pub=True)
    h.set_fisheye_mask(fm)
    ref = FeatureTrackerRef(FrontendConfig(use_ransac=1, fisheye=1))
    ref.fisheye_mask = fm
    for k in range(9):
        gray = seq.frame(k)[1]
        R = seq.relative_R(k)
        out = h.read_image(0, gray, seq.time(k), R, pub=(k % 3 == 0))
        ref.read_image(gray, seq.time(k), R, pub_this_frame=(k % 3 == 0))
        check_frame(k, out, ref)
        if k % 3 == 0:
            p = np.rint(out.cur_pts).astype(int)
            assert np.all(fm[p[:, 1], p[:, 0]] == 255)
    assert 40 < out.n < 150
    h.close()


def test_odd_frame_size_three_levels():
    """336x250 (width a multiple of 16 only, odd pyramid sizes 168x125 and 84x63): the vectorised pyrDown's partial
    segments and reflected borders, LK and FAST on cells that do not tile the frame evenly."""
    cam = synth.CamModel(fx=320.0, fy=320.0, cx=168.0, cy=125.0, width=336, height=250)
    n = 0
    for k, out, ref in run_pair({"lk_max_level": 2, "max_cnt": 120, "min_dist": 15}, 7, 4711, cam=cam, ransac=1):
        check_frame(k, out, ref)
        n += 1
    assert n == 7 and out.n > 60


def test_clahe_on_a_frame_size_that_needs_padding():
    """EQUALIZE at 336x250: rows are not a multiple of the 8x8 tile grid, so OpenCV extends the frame (REFLECT_101,
    8 extra columns and 6 extra rows) before cutting tiles; the device path must follow bit-exactly."""
    cam = synth.CamModel(fx=320.0, fy=320.0, cx=168.0, cy=125.0, width=336, height=250)
    seq = synth.Sequence(31, cam)
    cfg = binding.default_config(row=250, col=336, fx=320.0, fy=320.0, cx=168.0, cy=125.0, use_ransac=0, equalize=1,
                                 lk_max_level=2, max_cnt=120, min_dist=15)
    h = binding.Handle(cfg, 1, 0)
    ref = FeatureTrackerRef(FrontendConfig(row=250, col=336, fx=320.0, fy=320.0, cx=168.0, cy=125.0, use_ransac=0, equalize=1,
                                           lk_max_level=2, max_cnt=120, min_dist=15))
    for k in range(5):
        gray = (seq.frame(k)[1].astype(np.float32) * 0.4 + 10).astype(np.uint8)
        R = seq.relative_R(k)
        out = h.read_image(0, gray, seq.time(k), R, pub=(k % 3 == 0))
        ref.read_image(gray, seq.time(k), R, pub_this_frame=(k % 3 == 0))
        check_frame(k, out, ref)
    assert out.n > 50
    h.close()


def test_rgb_ingest_equals_gray_path():
    """RGB8 payload: device-side cvtColor(RGB2GRAY) then the same pipeline."""
    for k, out, ref in run_pair({}, 5, 31, rgb=True):
        check_frame(k, out, ref)


def test_1280x720_500_features():
    """BASELINE config 4 geometry: 1280x720, 500 feats, 4 levels, min_dist 30."""
    cam = synth.CamModel(fx=900.0, fy=900.0, cx=640.0, cy=360.0, width=1280, height=720)
    for k, out, ref in run_pair({"max_cnt": 500, "min_dist": 30, "lk_max_level": 3}, 5, 11, cam=cam):
        check_frame(k, out, ref)


def test_1280x720_500_features_with_ransac():
    """BASELINE configs[3] incl. rejectWithF on the publish frames (feature_tracker.cpp:441-473): 1280x720, 500 feats,
    4 levels, min_dist 30; IDs / occupancy bit-exact over 7 frames (3 publish frames)."""
    cam = synth.CamModel(fx=900.0, fy=900.0, cx=640.0, cy=360.0, width=1280, height=720)
    n = 0
    for k, out, ref in run_pair({"max_cnt": 500, "min_dist": 30, "lk_max_level": 3}, 7, 17, cam=cam, ransac=1):
        check_frame(k, out, ref)
        n += ref.last_ransac_status is not None
    assert n >= 2 and out.n > 250


def test_300_features_full_front_end():
    """BASELINE configs[4]: 640x480, max_cnt 300 (grid threshold 5), reference-default 2-level LK with IMU prediction,
    RANSAC on; 10 frames."""
    n = 0
    for k, out, ref in run_pair({"max_cnt": 300}, 10, 303, ransac=1):
        check_frame(k, out, ref)
        n += ref.last_ransac_status is not None
    assert n >= 3 and out.n > 150


def test_batch_equals_single():
    """Batched call over 3 independent sequences == three single-sequence calls."""
    cam = synth.CamModel()
    cfg = binding.default_config(use_ransac=0)
    hb = binding.Handle(cfg, 3, 0)
    hs = [binding.Handle(cfg, 1, 0) for _ in range(3)]
    seqs = [synth.Sequence(100 + i, cam) for i in range(3)]
    for k in range(5):
        imgs = [s.frame(k)[1] for s in seqs]
        Rs = [s.relative_R(k) for s in seqs]
        ts = [s.time(k) for s in seqs]
        pubs = [1 if (k + i) % 2 == 0 else 0 for i in range(3)]
        outs = hb.read_image_batch([2, 0, 1], [imgs[2], imgs[0], imgs[1]], [ts[2], ts[0], ts[1]],
                                   [Rs[2], Rs[0], Rs[1]], [pubs[2], pubs[0], pubs[1]])
        for slot, i in enumerate([2, 0, 1]):
            single = hs[i].read_image(0, imgs[i], ts[i], Rs[i], pub=pubs[i], debug=False)
            assert outs[slot].n == single.n
            assert np.array_equal(outs[slot].ids, single.ids)
            assert np.array_equal(outs[slot].cur_pts, single.cur_pts)
            assert np.array_equal(outs[slot].pts_velocity, single.pts_velocity)
    hb.close()
    for x in hs:
        x.close()


def test_chunked_host_batch_equals_single():
    """A host-frame batch large enough to be cut into copy/compute chunks (100 sequences -> 2 chunks,
    api.cu vrf_tracker_read_image_batch) returns, per batch position, what single-sequence handles return."""
    cam = synth.CamModel()
    cfg = binding.default_config(use_ransac=1)
    n = 100
    hb = binding.Handle(cfg, n, 0)
    base = [synth.Sequence(300 + i, cam) for i in range(4)]
    probe = [0, 49, 50, 51, 99]                     # both sides of the chunk boundary
    hs = {i: binding.Handle(cfg, 1, 0) for i in probe}
    order = list(range(n))[::-1]                    # batch position != sequence slot
    for k in range(5):
        fr = [b.frame(k)[1] for b in base]
        Rs = [b.relative_R(k) for b in base]
        pubs = [1 if (k + i) % 3 == 0 else 0 for i in range(n)]
        outs = hb.read_image_batch(order, [fr[i % 4] for i in order], [1.0 + k / 30.0] * n,
                                   [Rs[i % 4] for i in order], [pubs[i] for i in order])
        for pos, i in enumerate(order):
            if i not in hs:
                continue
            single = hs[i].read_image(0, fr[i % 4], 1.0 + k / 30.0, Rs[i % 4], pub=pubs[i], debug=False)
            assert outs[pos].n == single.n, (k, i, outs[pos].n, single.n)
            if k >= 3:
                assert single.n > 50          # every probe has published at least once by then
            assert np.array_equal(outs[pos].ids, single.ids)
            assert np.array_equal(outs[pos].cur_pts, single.cur_pts)
            assert np.array_equal(outs[pos].cur_un_pts, single.cur_un_pts)
            assert np.array_equal(outs[pos].pts_velocity, single.pts_velocity)
            assert np.array_equal(outs[pos].track_cnt, single.track_cnt)
    hb.close()
    for x in hs.values():
        x.close()


def test_textureless_and_reset():
    """Flat frames: no corners -> texture flags drop, n == 0; then reset_sequence restarts ids at 0."""
    cfg = binding.default_config(use_ransac=0)
    h = binding.Handle(cfg, 1, 0)
    flat = np.full((480, 640), 127, np.uint8)
    out = h.read_image(0, flat, 1.0, None, pub=True)
    assert out.n == 0 and not out.grids_texture_status.any()
    seq = synth.Sequence(8)
    out = h.read_image(0, seq.frame(0)[1], 1.1, None, pub=True)
    # cells were flagged textureless => this publish frame only re-arms them (feature_tracker.cpp:383-392)
    assert out.n == 0 and out.grids_texture_status.all()
    out = h.read_image(0, seq.frame(1)[1], 1.2, None, pub=True)
    assert out.n > 100
    h.reset(0)
    out = h.read_image(0, seq.frame(0)[1], 2.0, None, pub=True)
    assert out.n > 100 and out.ids.min() == 0 and out.ids.max() == out.n - 1
    h.close()


@pytest.mark.parametrize("fmt", ["16UC1", "32FC1"])
def test_rgbd_depth_lookup_bit_exact(fmt):
    """depth decode (estimator_nodelet.cpp:512-533) + per-feature lookup and DEPTH_MIN_DIST test
    (feature_manager.cpp:71-80) on the device vs the oracle: depth_mm and the keep flags bit-exact, on publish
    frames; zeros on non-publish frames (no depth frame is consumed there)."""
    from oracle.frontend_ref import decode_depth, depth_lookup
    cam = synth.CamModel()
    seq = synth.Sequence(77, cam)
    cfg = binding.default_config(use_ransac=1, depth_min_dist=2.2)      # scene depth 1.5..4 m: both branches taken
    h = binding.Handle(cfg, 1, 0)
    ref = FeatureTrackerRef(FrontendConfig(use_ransac=1))
    culled = kept = 0
    for k in range(7):
        rgbf, gray, dep = seq.frame(k)
        if fmt == "32FC1":
            depm = dep.astype(np.float32) / np.float32(1000.0)
            if k == 3:
                depm[::7, ::5] = np.nan
                depm[1::7, ::5] = 1e7
        else:
            depm = dep
        R = seq.relative_R(k)
        pub = (k % 3 == 0)
        out = h.read_rgbd(0, rgbf, depm, seq.time(k), R, pub=pub)
        ref.read_image(gray, seq.time(k), R, pub_this_frame=pub)
        check_frame(k, out, ref)
        if pub:
            mm, keep = depth_lookup(decode_depth(depm, cam.height, cam.width), ref.cur_pts, cfg.depth_min_dist)
            assert np.array_equal(out.depth_mm, mm), k
            assert np.array_equal(out.depth_keep, keep), k
            culled += int((keep == 0).sum()); kept += int(keep.sum())
        else:
            assert not out.depth_mm.any() and out.depth_keep.all(), k
    assert culled > 0 and kept > 0
    h.close()


def test_rgbd_batch_dev_depth():
    """Device-resident RGB-D batch (vrf_tracker_enqueue_batch_dev with a depth batch) == per-sequence host calls."""
    import torch
    from oracle.frontend_ref import depth_lookup
    cam = synth.CamModel()
    S = 3
    seqs = [synth.Sequence(200 + s, cam) for s in range(S)]
    cfg = binding.default_config(use_ransac=1, depth_min_dist=2.0)
    h = binding.Handle(cfg, S, 0)
    refs = [FeatureTrackerRef(FrontendConfig(use_ransac=1)) for _ in range(S)]
    for k in range(4):
        fr = [s.frame(k) for s in seqs]
        rgb = torch.from_numpy(np.stack([f[0] for f in fr])).cuda()
        dep = torch.from_numpy(np.stack([f[2] for f in fr]).view(np.int16)).cuda()
        Rs = np.stack([s.relative_R(k) for s in seqs])
        pubs = [1 if (k + s) % 2 == 0 else 0 for s in range(S)]
        h.enqueue_dev(list(range(S)), rgb.data_ptr(), binding.FMT_RGB8, [seqs[0].time(k)] * S, Rs, pubs,
                      d_depth=dep.data_ptr(), depth_fmt=binding.DEPTH_16UC1)
        outs = h.fetch(list(range(S)))
        for s in range(S):
            refs[s].read_image(fr[s][1], seqs[0].time(k), Rs[s], pub_this_frame=bool(pubs[s]))
            assert np.array_equal(outs[s].ids, np.asarray(refs[s].ids, np.int32)), (k, s)
            if pubs[s]:
                mm, keep = depth_lookup(fr[s][2], refs[s].cur_pts, cfg.depth_min_dist)
                assert np.array_equal(outs[s].depth_mm, mm) and np.array_equal(outs[s].depth_keep, keep), (k, s)
            else:
                assert not outs[s].depth_mm.any(), (k, s)
    h.close()


def test_pipelined_submit_collect_equals_synchronous_calls():
    """vrf_tracker_submit_rgbd_batch / vrf_tracker_collect_batch with two batches in flight return exactly what the
    synchronous batched call returns (same ids, tracks, depths), and a third submit without a collect is refused."""
    import ctypes as C
    cam = synth.CamModel()
    S, T = 3, 6
    seqs = [synth.Sequence(300 + s, cam) for s in range(S)]
    frames = [[sq.frame(k) for k in range(T)] for sq in seqs]
    cfg = binding.default_config(use_ransac=1, depth_min_dist=2.0)
    h_sync = binding.Handle(cfg, S, 0)
    h_pipe = binding.Handle(cfg, S, 0)
    seq_a = np.arange(S, dtype=np.int32)

    def args(k):
        ptrs = (C.c_void_p * S)(*[frames[s][k][0].ctypes.data for s in range(S)])
        dptrs = (C.c_void_p * S)(*[frames[s][k][2].ctypes.data for s in range(S)])
        t_a = np.full(S, seqs[0].time(k), np.float64)
        R_a = np.ascontiguousarray(np.stack([seqs[s].relative_R(k) for s in range(S)]).reshape(S, 9))
        p_a = np.asarray([1 if (k + s) % 3 == 0 else 0 for s in range(S)], np.int32)
        return ptrs, dptrs, t_a, R_a, p_a

    expect = []
    for k in range(T):
        ptrs, dptrs, t_a, R_a, p_a = args(k)
        outs, res = h_sync.make_track_batch(S)
        h_sync.read_image_batch_into(seq_a, ptrs, binding.FMT_RGB8, t_a, R_a, p_a, outs, dptrs=dptrs, dfmt=binding.DEPTH_16UC1)
        expect.append([(r.finish().ids.copy(), r.cur_pts.copy(), r.cur_un_pts.copy(), r.pts_velocity.copy(),
                        r.track_cnt.copy(), r.depth_mm.copy(), r.depth_keep.copy()) for r in res])
    keep_alive = []
    DEPTH = 3                                   # batches in flight (VRF_PIPE_DEPTH): submit k+2, then collect k

    def submit(k):
        a = args(k); keep_alive.append(a)
        h_pipe.submit_batch_into(seq_a, a[0], binding.FMT_RGB8, a[2], a[3], a[4], dptrs=a[1], dfmt=binding.DEPTH_16UC1)
        return a

    for k in range(min(DEPTH - 1, T)):
        submit(k)
    for k in range(T):
        if k + DEPTH - 1 < T:
            a = submit(k + DEPTH - 1)
            if k == 0:      # three batches pending: a fourth submit must be refused, not queued
                rc = h_pipe.lib.vrf_tracker_submit_rgbd_batch(h_pipe.h, S, seq_a.ctypes.data, C.cast(a[0], C.c_void_p), 0,
                                                              binding.FMT_RGB8, None, 0, 0, a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data)
                assert rc == -4
        outs, res = h_pipe.make_track_batch(S)
        h_pipe.collect_batch_into(seq_a, outs)
        for s in range(S):
            r = res[s].finish()
            e = expect[k][s]
            for got, want in zip((r.ids, r.cur_pts, r.cur_un_pts, r.pts_velocity, r.track_cnt, r.depth_mm, r.depth_keep), e):
                assert np.array_equal(got, want), (k, s)
    h_sync.close(); h_pipe.close()


def _two_view_pixels(rng, n, out_frac, noise):
    """Distorted pixel coordinates of n scene points in two nearby views (what cur_pts / forw_pts hold)."""
    import cv2
    from oracle.frontend_ref import PinholeCamera
    cam = PinholeCamera(FrontendConfig())
    P = rng.uniform([-1.6, -1.2, 2.0], [1.6, 1.2, 6.0], (n, 3))
    R, _ = cv2.Rodrigues(rng.normal(0, 0.02, 3))
    P2 = P @ R.T + rng.normal(0, 0.05, 3)
    u1, v1 = cam.space_to_plane(P[:, 0], P[:, 1], P[:, 2])
    u2, v2 = cam.space_to_plane(P2[:, 0], P2[:, 1], P2[:, 2])
    a = np.stack([u1, v1], 1)
    b = np.stack([u2, v2], 1) + rng.normal(0, noise, (n, 2))
    no = int(out_frac * n)
    b[:no] += rng.normal(0, 8, (no, 2))
    return cam, a.astype(np.float32), b.astype(np.float32)


def _cv2_reject_with_f(cam, cur, forw, cfg):
    """feature_tracker.cpp:441-473 on the real OpenCV."""
    import cv2
    x, y = cam.lift_projective_pts(cur)
    un_cur = np.stack([cfg.focal_length * x + cfg.col / 2.0, cfg.focal_length * y + cfg.row / 2.0], 1).astype(np.float32)
    x, y = cam.lift_projective_pts(forw)
    un_forw = np.stack([cfg.focal_length * x + cfg.col / 2.0, cfg.focal_length * y + cfg.row / 2.0], 1).astype(np.float32)
    _, st = cv2.findFundamentalMat(un_cur, un_forw, cv2.FM_RANSAC, cfg.f_threshold, 0.99)
    return None if st is None else st.ravel()


def test_reject_with_f_lmeds_regime():
    """rejectWithF for 8 <= n < 15, where cv::findFundamentalMat(FM_RANSAC) runs LMedS (feature_tracker.cpp:445 enters at
    forw_pts.size() >= 8).  n = 14 is pinned: identical masks to cv2.  For n <= 13 cv2's own mask is decided by rounding noise
    (tests/test_oracle_ransac.py), so the kernel is held to the algorithm's semantics: LMedS ran, a minimal-sample model's
    consensus (>= 7 points, <= n) survives, and fewer than 8 points are left untouched."""
    cfg = binding.default_config(use_ransac=1)
    h = binding.Handle(cfg, 1, 0)
    rng = np.random.default_rng(1414)
    for n, out_frac in [(14, 0.0), (14, 0.15), (14, 0.3), (14, 0.15), (14, 0.3), (14, 0.0)]:
        cam, cur, forw = _two_view_pixels(rng, n, out_frac, rng.choice([0.05, 0.3, 0.8]))
        want = _cv2_reject_with_f(cam, cur, forw, cfg)
        got = h.reject_with_f(0, cur, forw)
        assert want is not None and np.array_equal(got, want), (n, out_frac, got, want)
    for n in (8, 9, 10, 11, 12, 13):
        cam, cur, forw = _two_view_pixels(rng, n, 0.2, 0.5)
        got = h.reject_with_f(0, cur, forw)
        assert 7 <= int(got.sum()) <= n
    cam, cur, forw = _two_view_pixels(rng, 7, 0.0, 0.1)
    assert h.reject_with_f(0, cur, forw).all()              # n < 8: rejectWithF does nothing
    # and the RANSAC regime through the same entry (n >= 15)
    for n in (15, 40, 150):
        cam, cur, forw = _two_view_pixels(rng, n, 0.2, 0.3)
        want = _cv2_reject_with_f(cam, cur, forw, cfg)
        assert np.array_equal(h.reject_with_f(0, cur, forw), want), n
    h.close()


@pytest.mark.parametrize("shape,levels", [((480, 640), 3), ((720, 1280), 4), ((250, 336), 3), ((357, 400), 3), ((123, 208), 2)])
@pytest.mark.parametrize("rgb", [False, True])
def test_pyramid_levels_bit_exact_vs_cv2(shape, levels, rgb):
    """k_pyr (ingest fused with the first pyrDown, then the deeper levels): every level of every sequence of a batch is
    bit-identical to cv2.cvtColor(RGB2GRAY) + cv2.pyrDown -- random frames (every rounding case), odd level sizes
    (reflected borders, partial 8-output groups), strips that end below the image, two sequences per launch."""
    import cv2
    rows, cols = shape
    rng = np.random.default_rng(rows * 7 + cols + int(rgb))
    cfg = binding.default_config(row=rows, col=cols, fx=300.0, fy=300.0, cx=cols / 2, cy=rows / 2, use_ransac=0,
                                 lk_max_level=levels - 1, max_cnt=60, min_dist=12)
    h = binding.Handle(cfg, 2, 0)
    for k in range(2):
        frames = [rng.integers(0, 256, (rows, cols, 3) if rgb else (rows, cols), dtype=np.uint8) for _ in range(2)]
        h.read_image_batch([0, 1], frames, [0.1 * k, 0.1 * k], pubs=[1, 1])
        for s in range(2):
            lvl = cv2.cvtColor(frames[s], cv2.COLOR_RGB2GRAY) if rgb else frames[s]
            for l in range(levels):
                got = h.pyramid_level(s, l)
                assert got.shape == lvl.shape, (s, l)
                assert np.array_equal(got, lvl), (k, s, l, np.argwhere(got != lvl)[:4])
                lvl = cv2.pyrDown(lvl)
    h.close()
