#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ (run once, in the build container).

The reference ships no tests, fixtures or golden vectors, and its C++ cannot be
built here (no Eigen / OpenCV C++ / Ceres / ROS).  What CAN be run here is the third-party arithmetic
the reference's front end links: OpenCV 4.13 through cv2.  The front-end vectors below are therefore
outputs of the REAL library calls the reference makes
    cv::pyrDown / buildOpticalFlowPyramid, cv::FastFeatureDetector (thr 10, NMS, TYPE_9_16),
    cv::calcOpticalFlowPyrLK (21x21, eps 0.01, 30 it), cv::findFundamentalMat(FM_RANSAC, 1.0, 0.99)
(call sites feature_tracker.cpp:29,109,302-310,462), plus the per-frame outputs of the control-flow
restatement of FeatureTracker::readImage (oracle/frontend_ref.py) on stored frames.
The back-end vectors come from oracle/ba_ref.c -- a restatement, NOT real Ceres ("parity unpinned",
DESIGN.md section 2); they freeze the oracle so that oracle drift is caught, nothing more.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vins-rgbd-fast_b200"))

FRONT_CAM = dict(width=320, height=240, fx=300.0, fy=300.0, cx=160.0, cy=120.0, k1=0.1, k2=-0.2, p1=1e-3, p2=1e-3)
FRONT_CFG = dict(max_cnt=200, min_dist=12, num_grid_rows=4, num_grid_cols=5, use_imu=1, lk_max_level=2)
FRONT_SEED, FRONT_FRAMES, PUB_EVERY = 777, 8, 3


def front_end():
    import cv2
    from oracle.frontend_ref import FeatureTrackerRef, FrontendConfig
    from vrf_b200 import synth
    cam = synth.CamModel(**FRONT_CAM)
    seq = synth.Sequence(FRONT_SEED, cam)
    frames = np.stack([seq.frame(k)[1] for k in range(FRONT_FRAMES)])
    rgb0 = seq.frame(0)[0]
    Rs = np.stack([seq.relative_R(k) for k in range(FRONT_FRAMES)])
    times = np.array([seq.time(k) for k in range(FRONT_FRAMES)])
    out = dict(frames=frames, rgb0=rgb0, Rs=Rs, times=times, pub_every=np.int32(PUB_EVERY))
    for ransac in (0, 1):
        ref = FeatureTrackerRef(FrontendConfig(
            row=cam.height, col=cam.width, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, k1=cam.k1, k2=cam.k2, p1=cam.p1, p2=cam.p2,
            use_ransac=ransac, f_threshold=1.0, **FRONT_CFG))
        for k in range(FRONT_FRAMES):
            ref.read_image(frames[k], times[k], Rs[k], pub_this_frame=(k % PUB_EVERY == 0))
            p = f"r{ransac}_f{k}_"
            out[p + "ids"] = np.asarray(ref.ids, np.int32)
            out[p + "track_cnt"] = np.asarray(ref.track_cnt, np.int32)
            out[p + "cur_pts"] = np.asarray(ref.cur_pts, np.float32).reshape(-1, 2)
            out[p + "cur_un_pts"] = np.asarray(ref.cur_un_pts, np.float32).reshape(-1, 2)
            out[p + "pts_velocity"] = np.asarray(ref.pts_velocity, np.float32).reshape(-1, 2)
            out[p + "grids_track_num"] = np.asarray(ref.grids_track_num, np.int32)
            out[p + "n_id"] = np.int32(ref.n_id)
            if ref.last_lk_pts is not None:
                out[p + "lk_pts"] = np.asarray(ref.last_lk_pts, np.float32).reshape(-1, 2)
                out[p + "lk_status"] = np.asarray(ref.last_lk_status, np.uint8)
    # ---- primitive-level vectors straight from cv2 (the third-party arithmetic itself) ----
    f0, f1 = frames[0], frames[1]
    out["cv_gray0"] = cv2.cvtColor(rgb0, cv2.COLOR_RGB2GRAY)
    out["cv_pyr1"] = cv2.pyrDown(f0)
    out["cv_pyr2"] = cv2.pyrDown(out["cv_pyr1"])
    det = cv2.FastFeatureDetector_create()
    rect = (61, 57, 70, 66)                                     # an interior grid cell with the +3 px overlap
    x, y, w, h = rect
    mask = np.full((h, w), 255, np.uint8)
    cv2.circle(mask, (w // 3, h // 2), 15, 0, -1)
    kps = det.detect(f0[y:y + h, x:x + w], mask)
    out["cv_fast_rect"] = np.asarray(rect, np.int32)
    out["cv_fast_mask"] = mask
    out["cv_fast_kps"] = np.asarray([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps], np.int32).reshape(-1, 3)
    r = np.random.default_rng(5)
    pts = np.stack([r.uniform(-5, cam.width + 5, 120), r.uniform(-5, cam.height + 5, 120)], 1).astype(np.float32)
    init = (pts + r.normal(0, 1.5, pts.shape)).astype(np.float32)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts.reshape(-1, 1, 2), init.reshape(-1, 1, 2).copy(), winSize=(21, 21), maxLevel=2,
                                            criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                            flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    out["cv_lk_prev"] = pts; out["cv_lk_init"] = init
    out["cv_lk_next"] = nxt.reshape(-1, 2); out["cv_lk_status"] = st.reshape(-1)
    # RANSAC: a planar-ish correspondence set with outliers (float32, as rejectWithF passes them)
    r = np.random.default_rng(9)
    n = 90
    X = np.stack([r.uniform(-2, 2, n), r.uniform(-1.5, 1.5, n), r.uniform(2, 6, n)], 1)
    Rr = synth.so3_exp(np.array([0.01, -0.02, 0.015])); t = np.array([0.05, 0.01, -0.02])
    X2 = X @ Rr.T + t
    a = (460.0 * X[:, :2] / X[:, 2:3] + np.array([160.0, 120.0])).astype(np.float32)
    b = (460.0 * X2[:, :2] / X2[:, 2:3] + np.array([160.0, 120.0])).astype(np.float32)
    b[::9] += r.normal(0, 6, b[::9].shape).astype(np.float32)
    F, m = cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, 1.0, 0.99)
    out["cv_ransac_a"] = a; out["cv_ransac_b"] = b; out["cv_ransac_mask"] = m.reshape(-1).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "frontend_320x240.npz"), **out)
    print("frontend_320x240.npz:", len(out), "arrays")


# ---- back end: problem (inputs) + oracle solution -------------------------------------------------
def pack_problem(pb):
    """numpy view of a vrf_b200.ba_problem.BaProblem (everything a VrfBaProblem points to)."""
    from vrf_b200 import binding as B
    d = dict(frame_count=np.int32(pb.c.frame_count), use_imu=np.int32(pb.c.use_imu), ex_constant=np.int32(pb.c.ex_constant),
             td_constant=np.int32(pb.c.td_constant), marg_flag=np.int32(pb.c.marginalization_flag),
             max_iterations=np.int32(pb.c.max_iterations), pose=pb.pose, sb=pb.sb, ex=pb.ex, td=np.float64(pb.td),
             lam=pb.lam, start=pb.start, flag=pb.flag, obs_ptr=pb.obs_ptr, obs_pts=pb.obs_pts,
             imu=np.frombuffer(bytes(pb.imu), np.uint8).copy())
    if pb.prior is not None:
        d["prior"] = np.frombuffer(bytes(pb.prior), np.uint8).copy()
    if pb.obs_vel is not None:
        d.update(obs_vel=pb.obs_vel, obs_cur_td=pb.obs_cur_td, obs_row=pb.obs_row)
    return d


def unpack_problem(d, prefix=""):
    from vrf_b200 import binding as B
    from vrf_b200.ba_problem import BaProblem
    g = lambda k: d[prefix + k]
    pb = BaProblem(int(g("frame_count")), int(g("use_imu")))
    pb.c.ex_constant, pb.c.td_constant = int(g("ex_constant")), int(g("td_constant"))
    pb.c.marginalization_flag, pb.c.max_iterations = int(g("marg_flag")), int(g("max_iterations"))
    pb.pose, pb.sb, pb.ex, pb.td = g("pose").copy(), g("sb").copy(), g("ex").copy(), float(g("td"))
    pb.set_landmarks(g("lam"), g("start"), g("flag"), g("obs_ptr"), g("obs_pts"))
    C.memmove(pb.imu, g("imu").tobytes(), C.sizeof(pb.imu))
    if prefix + "prior" in d:
        pb.prior = B.VrfPrior()
        C.memmove(C.byref(pb.prior), g("prior").tobytes(), C.sizeof(B.VrfPrior))
    if prefix + "obs_vel" in d:
        pb.set_td_observations(g("obs_vel"), g("obs_cur_td"), g("obs_row"))
    return pb.finalize()


def golden_cfg(**over):
    from vrf_b200 import binding as B
    cfg = B.default_config()
    cfg.num_iterations, cfg.fix_depth, cfg.depth_max_dist, cfg.g_norm = 8, 0, 10.0, 9.81
    cfg.acc_n, cfg.acc_w, cfg.gyr_n, cfg.gyr_w = 0.1, 0.001, 0.01, 0.0001
    cfg.focal_length = 460.0
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def back_end():
    from oracle import ba_ref
    from vrf_b200 import ba_problem as BP
    cfg = golden_cfg()
    BP.WindowSimulator.FLAG2_DEPTH = (0.52, 1.2)        # bounded landmarks next to their bound, like tests/conftest.py
    sim = BP.WindowSimulator(11, cfg, n_landmarks=80, preintegrate=ba_ref.preintegrate)
    out = {}
    for a in range(3):              # window 0: no prior; 1, 2: with the prior of the previous window
        pb = sim.window(a)
        sol = ba_ref.solve(cfg, pb)
        sim.commit(a, sol)
        for k, v in pack_problem(pb).items():
            out[f"w{a}_{k}"] = v
        A, b = BP.prior_normal_equations(sol.new_prior)
        out.update({f"w{a}_o_iterations": np.int32(sol.c.iterations), f"w{a}_o_successful": np.int32(sol.c.successful_steps),
                    f"w{a}_o_termination": np.int32(sol.c.termination), f"w{a}_o_initial_cost": np.float64(sol.c.initial_cost),
                    f"w{a}_o_final_cost": np.float64(sol.c.final_cost), f"w{a}_o_pose": sol.pose, f"w{a}_o_sb": sol.sb,
                    f"w{a}_o_lam": sol.lam[: pb.M].copy(), f"w{a}_o_Ps": sol.Ps, f"w{a}_o_Rs": sol.Rs, f"w{a}_o_Vs": sol.Vs,
                    f"w{a}_o_prior_n": np.int32(sol.new_prior.n), f"w{a}_o_prior_A": A, f"w{a}_o_prior_b": b})
    np.savez_compressed(os.path.join(HERE, "ba_windows.npz"), **out)
    print("ba_windows.npz:", len(out), "arrays")


if __name__ == "__main__":
    front_end()
    back_end()
