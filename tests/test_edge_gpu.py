"""GPU edge cases and size-independent properties, through the C ABI:
VO mode (USE_IMU = 0: para_Pose[0] constant, estimator.cpp:1182-1185), a window without landmarks, a window at
the NUM_OF_F = 1000 landmark capacity, idempotence of a converged solve, empty / invalid inputs of every entry point
(the reference aborts with ROS_ASSERT / ROS_BREAK there; the C ABI returns an error code and never falls back)."""
import ctypes as C

import numpy as np
import pytest

import fm_cases as FC
from oracle import ba_ref
from test_ba_gpu import compare, compare_prior, make_cfg
from vrf_b200 import ba_problem as BP
from vrf_b200 import binding as B

pytestmark = pytest.mark.gpu


def test_ba_vo_mode_without_imu():
    cfg = make_cfg(use_imu=0)
    h = B.Handle(cfg, 1, 0)
    sim = BP.WindowSimulator(15, cfg, n_landmarks=120)
    for a in range(2):
        pb = sim.window(a)
        pb.c.use_imu = 0
        so = ba_ref.solve(cfg, pb)
        sg = h.ba_solve(0, pb)
        compare(sg, so, pb)
        assert np.array_equal(sg.pose[0], pb.pose[0])            # pose 0 is a constant block in VO mode
        compare_prior(sg.new_prior, so.new_prior)
        sim.commit(a, so)
    h.close()


def test_ba_window_without_landmarks():
    cfg = make_cfg()
    h = B.Handle(cfg, 1, 0)
    sim = BP.WindowSimulator(7, cfg, n_landmarks=60)
    pb = sim.window(0)
    pb.set_landmarks(np.zeros(0), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros((0, 2)))
    pb.finalize()
    so = ba_ref.solve(cfg, pb)
    sg = h.ba_solve(0, pb)
    assert (sg.c.iterations, sg.c.successful_steps) == (so.c.iterations, so.c.successful_steps)
    assert np.abs(sg.Ps - so.Ps).max() <= 1e-6 and sg.new_prior.n == so.new_prior.n
    h.close()


def big_window(seed, cfg, copies=4):
    """A window at the landmark capacity: the simulator's landmarks replicated with perturbed inverse depths
    (landmarks hosted at frame 0 only once, the marginalization drops at most BA_MAX_M0 = 384 of them)."""
    sim = BP.WindowSimulator(seed, cfg, n_landmarks=3300, spawn_per_frame=40)
    pb = sim.window(0)
    rng = np.random.default_rng(seed)
    lam, start, flag, ptr, obs = [pb.lam], [pb.start], [pb.flag], list(pb.obs_ptr), [pb.obs_pts]
    keep = np.nonzero(pb.start > 0)[0]
    for c in range(copies - 1):
        for l in keep:
            if sum(len(x) for x in lam) >= 1000:
                break
            o0, o1 = pb.obs_ptr[l], pb.obs_ptr[l + 1]
            lam.append(np.array([pb.lam[l] * (1 + rng.normal(0, 0.02))])); start.append(np.array([pb.start[l]], np.int32))
            flag.append(np.array([pb.flag[l]], np.int32)); obs.append(pb.obs_pts[o0:o1] + rng.normal(0, 1e-4, (o1 - o0, 2)))
            ptr.append(ptr[-1] + (o1 - o0))
    pb.set_landmarks(np.concatenate(lam), np.concatenate(start), np.concatenate(flag), np.array(ptr, np.int32), np.concatenate(obs))
    return pb.finalize()


def test_ba_at_landmark_capacity_and_idempotence():
    cfg = make_cfg()
    h = B.Handle(cfg, 1, 0)
    pb = big_window(21, cfg)
    assert pb.M == 1000 and len(pb.obs_pts) <= 8192
    so = ba_ref.solve(cfg, pb)
    sg = h.ba_solve(0, pb)
    compare(sg, so, pb)
    compare_prior(sg.new_prior, so.new_prior)
    # size-independent property: restarting from the solver's own result cannot increase the cost and moves nothing
    # measurable (a converged trust-region solve is a fixed point up to its tolerances)
    pb2 = big_window(21, cfg)
    pb2.pose, pb2.sb = sg.pose.copy(), sg.sb.copy()
    pb2.set_landmarks(sg.lam[:pb.M].copy(), pb.start, pb.flag, pb.obs_ptr, pb.obs_pts)
    pb2.finalize()
    s2 = h.ba_solve(0, pb2)
    assert s2.c.initial_cost <= sg.c.final_cost * (1 + 1e-9)
    assert s2.c.final_cost <= s2.c.initial_cost * (1 + 1e-12)
    assert np.abs(s2.pose[:, :3] - sg.pose[:, :3]).max() < 5e-3
    # more landmarks than the library's capacity (1024 >= NUM_OF_F) are refused, never truncated
    pb3 = big_window(21, cfg)
    k = 25
    pb3.set_landmarks(np.append(pb3.lam, np.full(k, 0.5)), np.append(pb3.start, np.ones(k)).astype(np.int32),
                      np.append(pb3.flag, np.zeros(k)).astype(np.int32),
                      np.append(pb3.obs_ptr, pb3.obs_ptr[-1] + 2 * np.arange(1, k + 1)).astype(np.int32), np.vstack([pb3.obs_pts, np.zeros((2 * k, 2))]))
    pb3.finalize()
    assert pb3.M == 1025
    with pytest.raises(RuntimeError):
        h.ba_solve(0, pb3)
    h.close()


def test_invalid_and_empty_inputs_fail_loudly():
    lib = B.load()
    cfg = B.default_config()
    hp = C.c_void_p()
    bad = B.default_config(); bad.col = 642                                  # rows must allow 16-byte accesses
    assert lib.vrf_create(C.byref(bad), 1, 0, C.byref(hp)) == -1 and not hp.value
    bad = B.default_config(); bad.max_cnt = 10                               # grids_threshold = 0: ROS_ASSERT in the reference (:89-93)
    assert lib.vrf_create(C.byref(bad), 1, 0, C.byref(hp)) == -1
    assert lib.vrf_create(C.byref(cfg), 0, 0, C.byref(hp)) == -1             # no sequences
    assert lib.vrf_create(C.byref(cfg), 1, 99, C.byref(hp)) == -1            # no such device
    h = B.Handle(cfg, 2, 0)
    img = np.zeros((cfg.row, cfg.col), np.uint8)
    with pytest.raises(RuntimeError):
        h.read_image(5, img, 0.0, np.eye(3), pub=True)                       # sequence out of range
    out = h.read_image(0, img, 0.0, np.eye(3), pub=True)                     # texture-less frame: no features, no error
    assert out.n == 0
    # stateless calls: empty batches and empty landmark lists are fine, inconsistent CSR offsets are not
    h.fm_triangulate_with_depth([])
    c = FC.make_case(1, M=4)
    p = B.FmProblem(c["Ps"], c["Rs"], c["tic"], c["ric"], c["start"][:0], c["obs_ptr"][:1], c["obs_pts"][:0], c["obs_depth"][:0], c["est_depth"][:0])
    h.fm_triangulate_with_depth([p]); h.fm_moving_consistency_check([p])
    start = c["start"].copy(); start[0] = 10                                 # start_frame + track length beyond the window
    p = B.FmProblem(c["Ps"], c["Rs"], c["tic"], c["ric"], start, np.array([0, 4, 5, 6, 7], np.int32), np.zeros((7, 2)), np.ones(7), c["est_depth"])
    with pytest.raises(RuntimeError):
        h.fm_triangulate_with_depth([p])
    assert len(h.imu_preintegrate([])) == 0
    h.close()


def test_handles_on_two_devices_in_one_process():
    """The > 48 KB dynamic shared-memory opt-in of the BA kernels is a per-device attribute: a process that opens handles on
    two GPUs must be able to solve on both (regression: a process-wide `configured` flag only configured the first)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = make_cfg()
    sim = BP.WindowSimulator(15, cfg, n_landmarks=60)
    pb = sim.window(0)
    so = ba_ref.solve(cfg, pb)
    for dev in (0, 1):
        h = B.Handle(cfg, 1, dev)
        sg = h.ba_solve(0, pb)
        assert sg.c.iterations == so.c.iterations and np.abs(sg.Ps - so.Ps).max() <= 1e-6
        h.close()


def test_flag2_landmark_starting_above_its_bound_is_projected_first():
    """estimate_flag == 2 landmarks carry inv_depth <= 2 / DEPTH_MAX_DIST (estimator.cpp:1293-1298).  Ceres'
    TrustRegionMinimizer::Init projects an infeasible start point onto the bounds before the first evaluation; oracle and
    kernel must agree on the initial cost and on the whole iteration sequence when several landmarks start infeasible."""
    cfg = make_cfg()
    h = B.Handle(cfg, 1, 0)
    sim = BP.WindowSimulator(77, cfg, n_landmarks=120, flag2_frac=0.3)
    pb = sim.window(0)
    ub = 2.0 / cfg.depth_max_dist
    flag2 = np.nonzero(pb.flag == 2)[0]
    assert len(flag2) >= 10
    lam = pb.lam.copy()
    lam[flag2[:6]] = ub * np.array([1.05, 1.5, 2.0, 3.0, 1.01, 1.2])         # depth 1.7 .. 4.9 m: infeasible starts
    pb.set_landmarks(lam, pb.start, pb.flag, pb.obs_ptr, pb.obs_pts)
    pb.finalize()
    so = ba_ref.solve(cfg, pb)
    sg = h.ba_solve(0, pb)
    compare(sg, so, pb)
    assert (sg.lam[flag2] <= ub * (1 + 1e-12)).all()
    h.close()
