"""CPU restatement of the steps either side of Estimator::optimization() -- TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import it.

    triangulate_with_depth      FeatureManager::triangulateWithDepth
                                (vins_estimator/src/feature_manager/feature_manager.cpp:386-543)
    moving_consistency_check    Estimator::movingConsistencyCheck + reprojectionError{,3D}
                                (vins_estimator/src/estimator/estimator.cpp:1944-2009)

Plain numpy / Python loops in the reference's statement order (the lists are small: <= 1000 landmarks x <= 11
observations).  Eigen::JacobiSVD -> numpy.linalg.svd (LAPACK); both return the right singular vector of the smallest
singular value up to sign, and only the ratio V[2]/V[3] is used.
PARITY UNPINNED against the compiled reference (it cannot be built here): the pins are the
geometric properties in tests/test_oracle_fm.py (noise-free depths are recovered exactly, the SVD branch agrees with
the closed-form two-view depth, moving points are flagged).
IMU pre-integration's oracle is oracle/ba_ref.c (oracle_preint_*), pinned against ground-truth motion there."""
import numpy as np

WINDOW_SIZE = 10
INIT_DEPTH = 5.0          # parameters.cpp:215


def triangulate_with_depth(Ps, Rs, tic, ric, start, obs_ptr, obs_pts, obs_depth, est_depth, est_flag, is_dynamic,
                           depth_min_dist, depth_max_dist):
    """In-place on est_depth / est_flag (numpy arrays), like the reference mutates FeaturePerId."""
    Ps = np.asarray(Ps, float); Rs = np.asarray(Rs, float).reshape(-1, 3, 3)
    tic = np.asarray(tic, float); ric = np.asarray(ric, float).reshape(3, 3)
    for l in range(len(start)):
        if est_depth[l] > 0:                                   # :390
            continue
        if is_dynamic[l]:                                      # :392
            continue
        o0, o1 = obs_ptr[l], obs_ptr[l + 1]
        n = o1 - o0
        if not (n >= 2 and start[l] < WINDOW_SIZE - 2):        # :396-398
            continue
        i = int(start[l])
        tr = Ps[i] + Rs[i] @ tic
        Rr = Rs[i] @ ric
        verified, rough = [], []
        no_depth = 0
        for k in range(n):
            dk = obs_depth[o0 + k]
            if dk == 0:
                no_depth += 1
                continue
            t0 = Ps[i + k] + Rs[i + k] @ tic
            R0 = Rs[i + k] @ ric
            point0 = np.array([obs_pts[o0 + k][0], obs_pts[o0 + k][1], 1.0]) * dk
            t2r = Rr.T @ (t0 - tr)
            R2r = Rr.T @ R0
            for j in range(n):
                if k == j:
                    continue
                t1 = Ps[i + j] + Rs[i + j] @ tic
                R1 = Rs[i + j] @ ric
                t20 = R0.T @ (t1 - t0)
                R20 = R0.T @ R1
                pp = R20.T @ point0 - R20.T @ t20
                res = np.array([obs_pts[o0 + j][0] - pp[0] / pp[2], obs_pts[o0 + j][1] - pp[1] / pp[2]])
                if np.sqrt(res[0] ** 2 + res[1] ** 2) < 10.0 / 460:                     # :444
                    z = (R2r @ point0 + t2r)[2]
                    (rough if dk > depth_max_dist else verified).append(z)
        if not verified:
            if not rough:
                if no_depth == n:                              # :464-513
                    t0 = Ps[i] + Rs[i] @ tic
                    R0 = Rs[i] @ ric
                    A = np.zeros((2 * n, 4))
                    for k in range(n):
                        t1 = Ps[i + k] + Rs[i + k] @ tic
                        R1 = Rs[i + k] @ ric
                        t = R0.T @ (t1 - t0)
                        R = R0.T @ R1
                        P = np.hstack([R.T, (-R.T @ t)[:, None]])
                        f = np.array([obs_pts[o0 + k][0], obs_pts[o0 + k][1], 1.0])
                        f = f / np.linalg.norm(f)
                        A[2 * k] = f[0] * P[2] - f[2] * P[0]
                        A[2 * k + 1] = f[1] * P[2] - f[2] * P[1]
                    v = np.linalg.svd(A, full_matrices=False)[2][-1]
                    svd_method = v[2] / v[3]
                    est_depth[l] = depth_max_dist if svd_method < depth_min_dist else svd_method
                    est_flag[l] = 2
                else:
                    continue
            else:
                est_depth[l] = sum(rough, 0.0) / len(rough)
                est_flag[l] = 0
        else:
            est_depth[l] = sum(verified, 0.0) / len(verified)
            est_flag[l] = 1
        if est_depth[l] < 0.1:                                 # :537-541
            est_depth[l] = INIT_DEPTH
            est_flag[l] = 0


def _reproj(Ri, Pi, ric, tic, Rj, Pj, depth, uvi, uvj):
    pts_w = Ri @ (ric @ (depth * uvi) + tic) + Pi
    pts_cj = ric.T @ (Rj.T @ (pts_w - Pj) - tic)
    r = (pts_cj / pts_cj[2])[:2] - uvj[:2]
    return np.sqrt(r[0] ** 2 + r[1] ** 2), np.linalg.norm(pts_cj - uvj) / depth


def moving_consistency_check(Ps, Rs, tic, ric, start, obs_ptr, obs_pts, est_depth, is_dynamic, focal_length=460.0):
    """Returns the `remove` flags (feature ids inserted into removeIndex); is_dynamic is updated in place."""
    Ps = np.asarray(Ps, float); Rs = np.asarray(Rs, float).reshape(-1, 3, 3)
    tic = np.asarray(tic, float); ric = np.asarray(ric, float).reshape(3, 3)
    remove = np.zeros(len(start), np.uint8)
    for l in range(len(start)):
        o0, o1 = obs_ptr[l], obs_ptr[l + 1]
        n = o1 - o0
        if not (n >= 2 and start[l] < WINDOW_SIZE - 2):
            continue
        depth = est_depth[l]
        if depth < 0:
            continue
        i = int(start[l])
        uvi = np.array([obs_pts[o0][0], obs_pts[o0][1], 1.0])
        err = err3 = 0.0
        cnt = 0
        for k in range(1, n):
            uvj = np.array([obs_pts[o0 + k][0], obs_pts[o0 + k][1], 1.0])
            e, e3 = _reproj(Rs[i], Ps[i], ric, tic, Rs[i + k], Ps[i + k], depth, uvi, uvj)
            err += e; err3 += e3; cnt += 1
        if cnt > 0:
            if focal_length * err / cnt > 10 or err3 / cnt > 2.0:
                remove[l] = 1
                is_dynamic[l] = 1
            else:
                is_dynamic[l] = 0
    return remove
