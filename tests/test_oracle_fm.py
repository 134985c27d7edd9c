"""CPU tests of the feature-manager oracle (oracle/fm_ref.py): the reference has no tests for these functions, so the
pins are geometric -- noise-free measured depths are recovered exactly, the linear-triangulation (SVD) branch recovers
the true depth of an exactly observed static point, far points take the "rough" path, moving points are flagged by
movingConsistencyCheck, and the skip conditions of feature_manager.cpp:390-398 hold."""
import numpy as np

import fm_cases as FC
from oracle import fm_ref


def run_tri(c, dmin=0.3, dmax=10.0):
    est, flag = c["est_depth"].copy(), c["est_flag"].copy()
    fm_ref.triangulate_with_depth(c["Ps"], c["Rs"], c["tic"], c["ric"], c["start"], c["obs_ptr"], c["obs_pts"], c["obs_depth"],
                                  est, flag, c["is_dynamic"], dmin, dmax)
    return est, flag


def eligible(c):
    n = np.diff(c["obs_ptr"])
    return (n >= 2) & (c["start"] < 8)


def test_noise_free_depths_are_recovered():
    c = FC.make_case(3, M=300)
    est, flag = run_tri(c)
    el, kd, zt = eligible(c), c["kind"], c["true_depth"]
    m = el & (kd == 0)
    # a landmark whose only measured depths sit on observations that were dropped keeps estimated_depth < 0
    got = m & (est > 0)
    assert got.sum() > 50
    assert np.all(flag[got] == 1) and np.abs(est[got] - zt[got]).max() < 1e-9
    m = el & (kd == 1)
    assert m.sum() > 5 and np.all(flag[m] == 0) and np.abs(est[m] - zt[m]).max() < 1e-9          # rough depths only
    m = el & (kd == 2)
    assert m.sum() > 5 and np.all(flag[m] == 2) and np.abs(est[m] / zt[m] - 1).max() < 1e-6       # SVD triangulation
    m = (kd == 4)
    assert np.all(est[m] == 3.3) and np.all(flag[m] == 0)                                        # already initialised: untouched
    m = (kd == 5) | ~el
    assert np.all(est[m & (kd != 4)] == -1.0)                                                    # dynamic / short tracks: skipped


def test_svd_branch_clamps_and_init_depth():
    c = FC.make_case(4, M=200)
    est, flag = run_tri(c, dmin=100.0)               # every triangulated depth is "too close" -> DEPTH_MAX_DIST
    m = eligible(c) & (c["kind"] == 2)
    assert m.sum() > 3 and np.all(est[m] == 10.0) and np.all(flag[m] == 2)
    # a measured depth below 0.1 m falls back to INIT_DEPTH with flag 0 (:537-541)
    c = FC.make_case(5, M=50)
    c["obs_depth"] = np.where(c["obs_depth"] > 0, 1e-9, 0.0)
    l = int(np.nonzero(eligible(c) & (c["kind"] == 0))[0][0])
    o0 = c["obs_ptr"][l]
    c["obs_pts"][o0 + 1] = c["obs_pts"][o0]          # make the tiny depth verify: identical rays, negligible parallax at 1e-9 m
    est, flag = run_tri(c)
    assert np.all(est[(est > 0) & (c["kind"] == 0)] == fm_ref.INIT_DEPTH)


def test_moving_points_are_flagged():
    c = FC.make_case(6, M=300)
    est, flag = run_tri(c)
    est_all = np.where(est > 0, est, c["true_depth"])
    dyn = np.zeros(len(est), np.uint8)
    rem = fm_ref.moving_consistency_check(c["Ps"], c["Rs"], c["tic"], c["ric"], c["start"], c["obs_ptr"], c["obs_pts"], est_all, dyn)
    el, kd = eligible(c), c["kind"]
    n = np.diff(c["obs_ptr"])
    static = el & np.isin(kd, [0, 1, 2, 4, 5]) & (np.abs(est_all - c["true_depth"]) < 1e-6)
    assert static.sum() > 100 and not rem[static].any() and not dyn[static].any()
    moving = el & (kd == 3) & (n >= 4)
    assert moving.sum() > 5 and rem[moving].mean() > 0.8
    assert np.array_equal(rem, dyn)                  # removeIndex <=> is_dynamic for the landmarks that were evaluated
    assert not rem[~el].any()
