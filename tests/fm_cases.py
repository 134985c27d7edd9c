"""Seeded synthetic inputs for the feature-manager steps (triangulateWithDepth / movingConsistencyCheck) and for
IMU pre-integration: a keyframe window on a smooth trajectory, landmarks with measured depths (some missing, some
beyond DEPTH_MAX_DIST, some landmarks without any depth, some moving).  Shared by the CPU and the GPU tests."""
import numpy as np

from vrf_b200 import synth

NF = 11


def window_poses(seed, kf_dt=0.1):
    tr = synth.Trajectory(seed)
    Ps = np.stack([tr.p_w(1.0 + kf_dt * k) for k in range(NF)])
    Rs = np.stack([tr.R_wb(1.0 + kf_dt * k) for k in range(NF)])
    rng = np.random.default_rng(seed + 77)
    ric = synth.so3_exp(rng.normal(0, 0.02, 3))
    tic = rng.normal(0, 0.03, 3)
    return Ps, Rs, tic, ric


def make_case(seed, M=160, noise=0.0, depth_max=10.0, kf_dt=0.1):
    """Returns a dict of arrays; `truth` holds the true host-frame depth and the kind of every landmark:
    0 measured depth in range, 1 all measured depths beyond depth_max (rough), 2 no depth at all (SVD),
    3 moving point, 4 pre-set estimated_depth (skipped), 5 pre-set dynamic (skipped by triangulate)."""
    rng = np.random.default_rng(seed)
    Ps, Rs, tic, ric = window_poses(seed, kf_dt)
    start, obs_ptr, pts, dep, kind, true_depth = [], [0], [], [], [], []
    for l in range(M):
        i = int(rng.integers(0, 10))
        n = int(rng.integers(1, NF - i + 1))
        kd = int(rng.choice([0, 0, 0, 0, 1, 2, 3, 4, 5]))
        z = rng.uniform(12.0, 20.0) if kd == 1 else rng.uniform(0.6, 8.0)
        xy = rng.uniform(-0.5, 0.5, 2)
        Rc_i, tc_i = Rs[i] @ ric, Ps[i] + Rs[i] @ tic
        Xw = Rc_i @ (np.array([xy[0], xy[1], 1.0]) * z) + tc_i
        vel = rng.normal(0, 1.5, 3) if kd == 3 else np.zeros(3)
        for k in range(n):
            f = i + k
            Rc, tc = Rs[f] @ ric, Ps[f] + Rs[f] @ tic
            Xc = Rc.T @ (Xw + vel * kf_dt * k - tc)
            p = Xc[:2] / Xc[2] + rng.normal(0, noise, 2)
            d = Xc[2] * (1 + rng.normal(0, noise))
            if kd == 2 or (kd != 1 and rng.random() < 0.15 and k > 0):
                d = 0.0
            pts.append(p); dep.append(d)
        start.append(i); obs_ptr.append(len(dep)); kind.append(kd); true_depth.append(z)
    kind = np.array(kind)
    est = np.where(kind == 4, 3.3, -1.0)
    dyn = (kind == 5).astype(np.uint8)
    return dict(Ps=Ps, Rs=Rs, tic=tic, ric=ric, start=np.array(start, np.int32), obs_ptr=np.array(obs_ptr, np.int32),
                obs_pts=np.array(pts).reshape(-1, 2), obs_depth=np.array(dep), est_depth=est, est_flag=np.zeros(M, np.int32),
                is_dynamic=dyn, kind=kind, true_depth=np.array(true_depth))


def imu_segment(seed, n_samples=20, rate=200.0):
    """(acc0, gyr0, ba, bg, dt[n], acc[n,3], gyr[n,3]) on a smooth trajectory with noise and bias."""
    rng = np.random.default_rng(seed)
    tr = synth.Trajectory(seed)
    ba, bg = rng.uniform(-0.05, 0.05, 3), rng.uniform(-0.01, 0.01, 3)
    t0 = 1.0 + rng.uniform(0, 1)
    ts = t0 + np.arange(n_samples + 1) / rate
    acc = np.stack([tr.acc_body(t) for t in ts]) + ba + rng.normal(0, 0.05, (n_samples + 1, 3))
    gyr = np.stack([tr.gyro_body(t) for t in ts]) + bg + rng.normal(0, 0.005, (n_samples + 1, 3))
    dt = np.diff(ts) * (1 + rng.normal(0, 1e-3, n_samples))
    return acc[0], gyr[0], ba + rng.normal(0, 0.01, 3), bg + rng.normal(0, 0.002, 3), dt, acc[1:], gyr[1:]
