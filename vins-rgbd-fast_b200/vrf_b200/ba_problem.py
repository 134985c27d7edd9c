"""Harness-side containers for one Estimator::optimization() call (include/vrf_ba.h)
and a seeded synthetic sliding-window generator (no datasets in the reference,
SURVEY.md section 4).  Not product code."""
import ctypes as C

import numpy as np

from . import binding as B
from . import synth


class BaProblem:
    """Owns the numpy buffers referenced by a VrfBaProblem."""

    def __init__(self, frame_count=10, use_imu=1):
        self.c = B.VrfBaProblem()
        self.c.frame_count = frame_count
        self.c.use_imu = use_imu
        self.c.ex_constant = 1
        self.c.td_constant = 1
        self.c.marginalization_flag = B.MARGIN_OLD
        self.c.max_iterations = 0
        self.pose = np.zeros((B.NUM_FRAMES, 7)); self.pose[:, 6] = 1
        self.sb = np.zeros((B.NUM_FRAMES, 9))
        self.ex = np.array([0, 0, 0, 0, 0, 0, 1.0])
        self.imu = (B.VrfImuPreint * B.NUM_FRAMES)()
        self.prior = None
        self.M = 0
        self.td = 0.0
        self.obs_vel = self.obs_cur_td = self.obs_row = None      # ProjectionTdFactor inputs (estimate_td)

    def set_td_observations(self, vel, cur_td, row):
        self.obs_vel = np.ascontiguousarray(vel, np.float64).reshape(-1, 2)
        self.obs_cur_td = np.ascontiguousarray(cur_td, np.float64)
        self.obs_row = np.ascontiguousarray(row, np.float64)

    def set_landmarks(self, lam, start, flag, obs_ptr, obs_pts):
        self.lam = np.ascontiguousarray(lam, np.float64)
        self.start = np.ascontiguousarray(start, np.int32)
        self.flag = np.ascontiguousarray(flag, np.int32)
        self.obs_ptr = np.ascontiguousarray(obs_ptr, np.int32)
        self.obs_pts = np.ascontiguousarray(obs_pts, np.float64).reshape(-1, 2)
        self.M = len(self.lam)

    def finalize(self):
        c = self.c
        for i in range(B.NUM_FRAMES):
            for k in range(7):
                c.para_Pose[i][k] = self.pose[i, k]
            for k in range(9):
                c.para_SpeedBias[i][k] = self.sb[i, k]
        for k in range(7):
            c.para_Ex_Pose[k] = self.ex[k]
        c.para_Td = float(self.td)
        c.obs_velocity = None if self.obs_vel is None else self.obs_vel.ctypes.data
        c.obs_cur_td = None if self.obs_cur_td is None else self.obs_cur_td.ctypes.data
        c.obs_row = None if self.obs_row is None else self.obs_row.ctypes.data
        c.n_landmarks = self.M
        c.n_obs = len(self.obs_pts)
        c.para_Feature = self.lam.ctypes.data
        c.lm_start_frame = self.start.ctypes.data
        c.lm_estimate_flag = self.flag.ctypes.data
        c.lm_obs_ptr = self.obs_ptr.ctypes.data
        c.obs_pts = self.obs_pts.ctypes.data
        c.imu = C.cast(self.imu, C.POINTER(B.VrfImuPreint))
        c.prior = C.pointer(self.prior) if self.prior is not None else None
        return self


class BaSolution:
    def __init__(self, M):
        self.c = B.VrfBaResult()
        self.lam = np.zeros(max(M, 1))
        self.c.para_Feature = self.lam.ctypes.data
        self.new_prior = B.VrfPrior()
        self.c.new_prior = C.pointer(self.new_prior)
        self.M = M
        self.rc = None

    def arr(self, name, shape):
        return np.ctypeslib.as_array(getattr(self.c, name)).reshape(shape).copy()

    @property
    def pose(self): return self.arr("para_Pose", (B.NUM_FRAMES, 7))
    @property
    def sb(self): return self.arr("para_SpeedBias", (B.NUM_FRAMES, 9))
    @property
    def ex(self): return np.array(list(self.c.para_Ex_Pose))
    @property
    def td(self): return float(self.c.para_Td)
    @property
    def Ps(self): return self.arr("Ps", (B.NUM_FRAMES, 3))
    @property
    def Rs(self): return self.arr("Rs", (B.NUM_FRAMES, 3, 3))
    @property
    def Vs(self): return self.arr("Vs", (B.NUM_FRAMES, 3))
    @property
    def Bas(self): return self.arr("Bas", (B.NUM_FRAMES, 3))
    @property
    def Bgs(self): return self.arr("Bgs", (B.NUM_FRAMES, 3))


def prior_normal_equations(prior):
    """(J0^T J0, J0^T r0): the sign/rotation-invariant content of a prior."""
    n = prior.n
    J = np.ctypeslib.as_array(prior.linearized_jacobians)[: n * n].reshape(n, n)
    r = np.ctypeslib.as_array(prior.linearized_residuals)[:n]
    return J.T @ J, J.T @ r


def prior_blocks(prior):
    return [(b.kind, b.index, b.size, b.idx, tuple(b.x0[: b.size])) for b in prior.blocks[: prior.n_blocks]]


def R_to_quat(R):
    """Eigen Quaterniond(Matrix3d) (x,y,z,w)."""
    t = np.trace(R)
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t; q[1] = (R[0, 2] - R[2, 0]) * t; q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]: i = 1
        if R[2, 2] > R[i, i]: i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0); q[i] = 0.5 * t; t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t; q[j] = (R[j, i] + R[i, j]) * t; q[k] = (R[k, i] + R[i, k]) * t
    return q


# IntegrationBase implementation used when a WindowSimulator is built without an explicit `preintegrate` callback.  The
# package itself never imports the checker: tests/conftest.py, the CPU arm of bench.py and __graft_entry__.smoke() install
# the C oracle's here; the GPU arm passes the library's own vrf_imu_preintegrate_batch.
DEFAULT_PREINTEGRATE = None


class WindowSimulator:
    """Ground-truth trajectory + landmark pool -> a chain of 11-frame windows
    (what FeatureManager / processIMU would hand to optimization())."""

    FLAG2_DEPTH = (1.02, 2.2)      # depth range of the estimate_flag == 2 landmarks in units of DEPTH_MAX_DIST (see _spawn)

    def __init__(self, seed, cfg, n_landmarks=150, kf_dt=0.1, imu_rate=200.0, flag2_frac=0.1,
                 pix_noise=0.5, ric=None, tic=None, td_true=0.0, ex_constant=1, td_constant=1, ex_perturb=0.0, preintegrate=None,
                 spawn_per_frame=None):
        self.cfg = cfg
        # IntegrationBase for the generated IMU samples: preintegrate(samples, acc0, gyr0, ba, bg, cfg) -> VrfImuPreint.
        # None = ba_problem.DEFAULT_PREINTEGRATE (installed by the checker legs); bench.py's GPU arm passes the library's
        # own vrf_imu_preintegrate_batch so that nothing under oracle/ is touched outside the checker legs.
        self._preintegrate = preintegrate
        self.rng = np.random.default_rng(seed)
        self.traj = synth.Trajectory(seed, fps=1.0 / kf_dt, trans_per_frame=0.06, rot_deg_per_frame=1.5)
        self.kf_dt = kf_dt
        self.imu_rate = imu_rate
        self.ric = np.eye(3) if ric is None else ric
        self.tic = np.zeros(3) if tic is None else tic
        self.pix_noise = pix_noise / 460.0
        self.ba_true = self.rng.uniform(-0.02, 0.02, 3)
        self.bg_true = self.rng.uniform(-0.005, 0.005, 3)
        self.n_landmarks = n_landmarks
        self.spawn_per_frame = spawn_per_frame
        self.flag2_frac = flag2_frac
        self.pool = []          # dicts: P (world), first, last (absolute frame idx), eps, flag
        self.est = {}           # absolute frame -> (P, R, V, Ba, Bg) current estimate
        self.prior = None
        self.seed = seed
        # time offset camera <-> IMU: images are stamped td_true too early, i.e. the feature seen in the image
        # stamped t(k) was really taken at t(k) + td_true (estimate_td); the estimate starts at 0
        self.td_true = td_true
        self.td_est = 0.0
        self.ex_constant, self.td_constant = ex_constant, td_constant
        self.ex_est = None
        if ex_perturb > 0:
            r = np.random.default_rng(seed + 991)
            self.ex_est = (self.tic + r.normal(0, ex_perturb, 3), self.ric @ synth.so3_exp(r.normal(0, ex_perturb, 3)))

    def t(self, k):
        return 2.0 + k * self.kf_dt

    def true_state(self, k):
        t = self.t(k)
        return self.traj.p_w(t), self.traj.R_wb(t), self.traj.v_w(t)

    def _cam(self, k, dt=0.0):
        t = self.t(k) + dt
        p, R = self.traj.p_w(t), self.traj.R_wb(t)
        return p + R @ self.tic, R @ self.ric

    def _spawn(self, k):
        """new landmarks first seen at absolute frame k"""
        pc, Rc = self._cam(k)
        out = []
        # 40 new landmarks per keyframe fill windows of up to ~320 landmarks (the reference's 150-feature configs); larger
        # windows (BASELINE configs[3]: 500 feats) spawn proportionally more
        per_frame = self.spawn_per_frame or (40 if self.n_landmarks <= 300 else -(-self.n_landmarks * 40 // 280))
        for _ in range(25 * per_frame):
            if len(out) >= per_frame:
                break
            d = self.rng.uniform(1.5, 6.0)
            xy = self.rng.uniform([-0.5, -0.38], [0.5, 0.38])
            P = pc + Rc @ (np.array([xy[0], xy[1], 1.0]) * d)
            length = int(self.rng.integers(2, 14))
            eps = self.rng.normal(0, 0.01)
            flag = 2 if self.rng.random() < self.flag2_frac else 1
            if flag == 2:
                # estimate_flag 2 = no depth measurement, depth from triangulation: points beyond the sensor range
                # (DEPTH_MAX_DIST, 10 m as shipped).  The reference bounds their inverse depth by 2 / DEPTH_MAX_DIST
                # (estimator.cpp:1293-1298), i.e. depth >= DEPTH_MAX_DIST / 2.  FLAG2_DEPTH = (lo, hi) in units of DEPTH_MAX_DIST:
                # the default spreads them over 1.02 .. 2.2 x the range; the parity tests use (0.52, 1.2) = 5.2 .. 12 m, next to
                # the bound (a few start just above it through `eps`), so that the projection of the start point, the projection
                # inside Plus() and Ceres' projected line search are exercised by every window chain.
                lo, hi = (v * float(self.cfg.depth_max_dist) for v in self.FLAG2_DEPTH)
                d = lo + (d - 1.5) / 4.5 * (hi - lo)
                P = pc + Rc @ (np.array([xy[0], xy[1], 1.0]) * d)
            out.append({"P": P, "first": k, "last": k + length - 1, "eps": eps, "flag": flag, "id": len(self.pool) + len(out)})
        return out

    def _observe(self, lm, k):
        pc, Rc = self._cam(k, self.td_true)
        q = Rc.T @ (lm["P"] - pc)
        if q[2] < 0.3:
            return None
        r = np.random.default_rng((self.seed * 1000003 + lm["id"] * 7919 + k) % (2 ** 32))
        return q[:2] / q[2] + r.normal(0, self.pix_noise, 2)

    def _velocity(self, lm, k, h=1e-3):
        """normalised-plane velocity of the feature (what undistortedPoints differences between frames)"""
        out = []
        for dt in (-h, h):
            pc, Rc = self._cam(k, self.td_true + dt)
            q = Rc.T @ (lm["P"] - pc)
            out.append(q[:2] / q[2])
        return (out[1] - out[0]) / (2 * h)

    def _preint(self, k0, k1, ba, bg):
        """IMU between absolute frames k0 -> k1 (processIMU: first sample initialises acc_0/gyr_0)."""
        pre_fn = self._preintegrate or DEFAULT_PREINTEGRATE
        if pre_fn is None:
            raise RuntimeError("WindowSimulator needs a preintegrate callback (the library's vrf_imu_preintegrate_batch, or the "
                               "checker's, installed by tests/conftest.py through ba_problem.DEFAULT_PREINTEGRATE)")
        t0, t1 = self.t(k0), self.t(k1)
        n = max(2, int(round((t1 - t0) * self.imu_rate)))
        ts = np.linspace(t0, t1, n + 1)
        rng = np.random.default_rng((self.seed * 31 + k0) % (2 ** 32))
        meas = []
        for tt in ts:
            g = self.traj.gyro_body(tt) + self.bg_true + rng.normal(0, self.cfg.gyr_n, 3)
            a = self.traj.acc_body(tt) + self.ba_true + rng.normal(0, self.cfg.acc_n, 3)
            meas.append((a, g))
        samples = [(ts[i] - ts[i - 1], meas[i][0], meas[i][1]) for i in range(1, n + 1)]
        return pre_fn(samples, meas[0][0], meas[0][1], ba, bg, self.cfg)

    def window(self, a, marg_flag=B.MARGIN_OLD, perturb=True):
        """Problem for absolute frames a..a+10."""
        while len(self.pool) < 1 or max(l["first"] for l in self.pool) < a + 10:
            k = (max(l["first"] for l in self.pool) + 1) if self.pool else max(0, a - 6)
            self.pool.extend(self._spawn(k))
        pb = BaProblem(10, 1)
        rng = np.random.default_rng(self.seed * 977 + a)
        for i in range(B.NUM_FRAMES):
            k = a + i
            if k not in self.est:
                p, R, v = self.true_state(k)
                if perturb:
                    p = p + rng.normal(0, 0.02, 3)
                    R = R @ synth.so3_exp(rng.normal(0, 0.008, 3))
                    v = v + rng.normal(0, 0.05, 3)
                self.est[k] = (p, R, v, self.ba_true + rng.normal(0, 0.005, 3), self.bg_true + rng.normal(0, 0.001, 3))
            p, R, v, ba, bg = self.est[k]
            pb.pose[i, :3] = p
            pb.pose[i, 3:] = R_to_quat(R)
            pb.sb[i] = np.concatenate([v, ba, bg])
        tic_e, ric_e = (self.tic, self.ric) if self.ex_est is None else self.ex_est
        pb.ex[:3] = tic_e
        pb.ex[3:] = R_to_quat(ric_e)
        pb.c.ex_constant, pb.c.td_constant = self.ex_constant, self.td_constant
        pb.td = self.td_est
        for j in range(1, B.NUM_FRAMES):
            _, _, _, ba, bg = self.est[a + j - 1]
            pb.imu[j - 1] = self._preint(a + j - 1, a + j, ba, bg)
        lam, start, flag, ptr, pts = [], [], [], [0], []
        vel, ctd, row = [], [], []
        for lm in self.pool:
            f0, f1 = max(lm["first"], a), min(lm["last"], a + 10)
            if f1 - f0 + 1 < 2 or f0 - a >= 8:        # used_num >= 2 && start_frame < WINDOW_SIZE - 2
                continue
            obs = [self._observe(lm, k) for k in range(f0, f1 + 1)]
            if any(o is None for o in obs):
                continue
            pc, Rc = self._cam(f0)
            depth = (Rc.T @ (lm["P"] - pc))[2] * (1.0 + lm["eps"])
            lam.append(1.0 / depth); start.append(f0 - a); flag.append(lm["flag"])
            pts.extend(obs); ptr.append(len(pts))
            if self.cfg.estimate_td:
                for k, o in zip(range(f0, f1 + 1), obs):
                    vel.append(self._velocity(lm, k))
                    ctd.append(0.0)                              # FeaturePerFrame::cur_td: td when the frame came in
                    row.append(460.0 * o[1] + 0.5 * self.cfg.row)   # uv.y of the observation
            if len(lam) >= self.n_landmarks:
                break
        pb.set_landmarks(lam, start, flag, ptr, np.array(pts))
        if self.cfg.estimate_td:
            pb.set_td_observations(vel, ctd, row)
        pb.prior = self.prior
        pb.c.marginalization_flag = marg_flag
        return pb.finalize()

    def commit(self, a, sol, marg_flag=B.MARGIN_OLD):
        """slideWindow bookkeeping after optimization(): keep the solution as the next window's
        initial estimate and adopt the new prior."""
        Ps, Rs, Vs, Bas, Bgs = sol.Ps, sol.Rs, sol.Vs, sol.Bas, sol.Bgs
        for i in range(B.NUM_FRAMES):
            self.est[a + i] = (Ps[i], Rs[i], Vs[i], Bas[i], Bgs[i])
        if not self.td_constant:
            self.td_est = sol.td
        if not self.ex_constant:
            q = sol.ex[3:] / np.linalg.norm(sol.ex[3:])
            x, y, z, w = q
            Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                           [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                           [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            self.ex_est = (sol.ex[:3].copy(), Rm)
        if sol.c.has_new_prior:
            p = B.VrfPrior()
            C.memmove(C.byref(p), C.byref(sol.new_prior), C.sizeof(B.VrfPrior))
            self.prior = p
