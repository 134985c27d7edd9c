"""CPU tests of the back-end oracle (oracle/ba_ref.c).  The reference has no tests for
this path; its only aids are the never-called finite-difference
printers ProjectionFactor::check (projection_factor.cpp:132-234) and the commented
invariants J^T J ~ A, J^T r ~ b of marginalize() (marginalization_factor.cpp:312-314).
Both are turned into assertions here, plus an independent minimiser (scipy) that any
correct solver must agree with at convergence."""
import ctypes as C

import numpy as np
import pytest

from oracle import ba_ref
from vrf_b200 import ba_problem as BP
from vrf_b200 import binding as B
from vrf_b200 import synth


def make_cfg():
    cfg = B.VrfConfig()
    cfg.num_iterations = 8; cfg.fix_depth = 0; cfg.depth_max_dist = 10.0; cfg.g_norm = 9.81
    cfg.acc_n = 0.1; cfg.acc_w = 0.001; cfg.gyr_n = 0.01; cfg.gyr_w = 0.0001
    return cfg


def pose_plus(x, d):
    dq = np.array([d[3] / 2, d[4] / 2, d[5] / 2, 1.0])
    q = x[3:]
    w = q[3] * dq[3] - q[:3] @ dq[:3]
    v = q[3] * dq[:3] + dq[3] * q[:3] + np.cross(q[:3], dq[:3])
    qq = np.concatenate([v, [w]]); qq /= np.linalg.norm(qq)
    return np.concatenate([x[:3] + d[:3], qq])


def rand_pose(rng, scale=1.0):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return np.concatenate([rng.normal(0, scale, 3), q])


def test_projection_jacobian_finite_difference():
    rng = np.random.default_rng(0)
    for _ in range(20):
        pi = rand_pose(rng, 0.3); pj = pose_plus(pi, rng.normal(0, 0.05, 6))
        ex = pose_plus(np.array([0.05, -0.02, 0.01, 0, 0, 0, 1.0]), rng.normal(0, 0.05, 6))
        lam = rng.uniform(0.2, 0.8)
        pts_i = rng.uniform(-0.4, 0.4, 2); pts_j = pts_i + rng.normal(0, 0.02, 2)
        r, Ji, Jj, Je, Jf = ba_ref.projection_eval(pi, pj, ex, lam, pts_i, pts_j)
        eps = 1e-6
        for blk, J in ((0, Ji), (1, Jj), (2, Je)):
            for c in range(6):
                d = np.zeros(6); d[c] = eps
                args = [pi, pj, ex]
                args[blk] = pose_plus(args[blk], d)
                r2 = ba_ref.projection_eval(args[0], args[1], args[2], lam, pts_i, pts_j, jac=False)[0]
                assert np.allclose((r2 - r) / eps, J[:, c], rtol=2e-4, atol=2e-3)
            assert np.all(J[:, 6] == 0)
        r2 = ba_ref.projection_eval(pi, pj, ex, lam + eps, pts_i, pts_j, jac=False)[0]
        assert np.allclose((r2 - r) / eps, Jf, rtol=2e-4, atol=2e-3)


def sim_preint(cfg, seed, noise=True):
    sim = BP.WindowSimulator(seed, cfg)
    if not noise:
        cfg2 = make_cfg(); cfg2.acc_n = 0; cfg2.gyr_n = 0
        sim.cfg = cfg2
    return sim


def test_preintegration_reproduces_motion():
    """Noise-free IMU through IntegrationBase::propagate must reproduce the true relative motion."""
    cfg = make_cfg()
    sim = sim_preint(cfg, 3)
    cfg0 = make_cfg(); cfg0.acc_n = 0.0; cfg0.gyr_n = 0.0
    sim.cfg = cfg0
    sim.imu_rate = 1000.0
    pre = sim._preint(0, 1, sim.ba_true, sim.bg_true)
    p0, R0, v0 = sim.true_state(0); p1, R1, v1 = sim.true_state(1)
    dt = pre.sum_dt
    G = np.array([0, 0, 9.81])
    dp = R0.T @ (p1 - p0 - v0 * dt + 0.5 * G * dt * dt)
    dv = R0.T @ (v1 - v0 + G * dt)
    assert abs(dt - sim.kf_dt) < 1e-12
    assert np.allclose(np.array(pre.delta_p), dp, atol=2e-5)
    assert np.allclose(np.array(pre.delta_v), dv, atol=2e-4)
    dq = np.array(pre.delta_q)
    Rd = R0.T @ R1
    q_true = BP.R_to_quat(Rd)
    assert min(np.abs(dq - q_true).max(), np.abs(dq + q_true).max()) < 1e-5
    cov = np.array(pre.covariance).reshape(15, 15)
    assert np.allclose(cov, cov.T, atol=1e-12)


def test_imu_jacobian_finite_difference():
    cfg = make_cfg()
    sim = BP.WindowSimulator(7, cfg)
    pre = sim._preint(0, 1, sim.ba_true + 0.01, sim.bg_true - 0.002)
    rng = np.random.default_rng(1)
    p0, R0, v0 = sim.true_state(0); p1, R1, v1 = sim.true_state(1)
    pi = np.concatenate([p0, BP.R_to_quat(R0)]); pj = np.concatenate([p1, BP.R_to_quat(R1)])
    pi = pose_plus(pi, rng.normal(0, 0.01, 6)); pj = pose_plus(pj, rng.normal(0, 0.01, 6))
    sbi = np.concatenate([v0, sim.ba_true + 0.01, sim.bg_true]) + rng.normal(0, 0.01, 9) * np.array([1, 1, 1, .1, .1, .1, .05, .05, .05])
    sbj = np.concatenate([v1, sim.ba_true, sim.bg_true + 0.001])
    r, Jpi, Jsi, Jpj, Jsj = ba_ref.imu_eval(pre, pi, sbi, pj, sbj)
    eps = 1e-7
    for blk, J, n in ((0, Jpi, 6), (1, Jsi, 9), (2, Jpj, 6), (3, Jsj, 9)):
        for c in range(n):
            args = [pi.copy(), sbi.copy(), pj.copy(), sbj.copy()]
            if n == 6:
                d = np.zeros(6); d[c] = eps
                args[blk] = pose_plus(args[blk], d)
            else:
                args[blk][c] += eps
            r2 = ba_ref.imu_eval(pre, *args, jac=False)[0]
            num = (r2 - r) / eps
            assert np.allclose(num, J[:, c], rtol=1e-4, atol=1e-5 * (1 + np.abs(J).max())), (blk, c)
    assert np.all(Jpi[:, 6] == 0) and np.all(Jpj[:, 6] == 0)


def numpy_marginalize(cfg, pb, x_pose, x_sb, x_ex, x_lam, prior):
    """Independent numpy assembly of MarginalizationInfo::marginalize (MARGIN_OLD) from the
    single-factor oracle evaluations; returns (A', b') of the kept blocks in the canonical order."""
    M = pb.M
    hosted = [l for l in range(M) if pb.start[l] == 0]
    idx = {("pose", 0): 0, ("sb", 0): 6}
    pos = 15
    for l in hosted:
        idx[("lm", l)] = pos; pos += 1
    m = pos
    idx[("ex", 0)] = pos; pos += 6
    present_pose = {1}
    for l in hosted:
        for k in range(1, pb.obs_ptr[l + 1] - pb.obs_ptr[l]):
            present_pose.add(k)
    if prior is not None:
        for b in prior.blocks[: prior.n_blocks]:
            if b.kind == B.BLK_POSE and b.index > 0: present_pose.add(b.index)
    for f in sorted(present_pose):
        idx[("pose", f)] = pos; pos += 6
    idx[("sb", 1)] = pos; pos += 9
    A = np.zeros((pos, pos)); bv = np.zeros(pos)

    def add(cols, J, r):
        A[np.ix_(cols, cols)] += J.T @ J
        bv[cols] += J.T @ r
    if prior is not None:
        n = prior.n
        J0 = np.ctypeslib.as_array(prior.linearized_jacobians)[: n * n].reshape(n, n)
        r = ba_ref.prior_residual(prior, x_pose, x_sb, x_ex)
        cols = np.zeros(n, int)
        for b in prior.blocks[: prior.n_blocks]:
            key = {B.BLK_POSE: "pose", B.BLK_SPEEDBIAS: "sb", B.BLK_EXPOSE: "ex"}[b.kind]
            ls = 6 if b.size == 7 else b.size
            cols[b.idx:b.idx + ls] = idx[(key, b.index)] + np.arange(ls)
        add(cols, J0, r)
    r, Jpi, Jsi, Jpj, Jsj = ba_ref.imu_eval(pb.imu[0], x_pose[0], x_sb[0], x_pose[1], x_sb[1])
    cols = np.concatenate([idx[("pose", 0)] + np.arange(6), idx[("sb", 0)] + np.arange(9), idx[("pose", 1)] + np.arange(6), idx[("sb", 1)] + np.arange(9)])
    add(cols, np.hstack([Jpi[:, :6], Jsi, Jpj[:, :6], Jsj]), r)
    for l in hosted:
        o0 = pb.obs_ptr[l]
        for k in range(1, pb.obs_ptr[l + 1] - o0):
            r, Ji, Jj, Je, Jf = ba_ref.projection_eval(x_pose[0], x_pose[k], x_ex, x_lam[l], pb.obs_pts[o0], pb.obs_pts[o0 + k])
            s = r @ r
            rho1 = 1.0 / (1.0 + s)           # Cauchy: rho'' < 0 => plain sqrt(rho') scaling
            J = np.sqrt(rho1) * np.hstack([Ji[:, :6], Jj[:, :6], Je[:, :6], Jf.reshape(2, 1)])
            cols = np.concatenate([idx[("pose", 0)] + np.arange(6), idx[("pose", k)] + np.arange(6), idx[("ex", 0)] + np.arange(6), [idx[("lm", l)]]])
            add(cols, J, np.sqrt(rho1) * r)
    Amm = 0.5 * (A[:m, :m] + A[:m, :m].T)
    w, V = np.linalg.eigh(Amm)
    winv = np.where(w > 1e-8, 1.0 / np.where(w > 1e-8, w, 1.0), 0.0)
    Ainv = (V * winv) @ V.T
    Ar = A[m:, m:] - A[m:, :m] @ Ainv @ A[:m, m:]
    br = bv[m:] - A[m:, :m] @ Ainv @ bv[:m]
    return Ar, br


def test_marginalization_matches_independent_numpy_schur():
    cfg = make_cfg()
    sim = BP.WindowSimulator(11, cfg, n_landmarks=120)
    prior_in = None
    for a in range(3):
        pb = sim.window(a)
        sol = ba_ref.solve(cfg, pb)
        assert sol.c.has_new_prior == 1
        # states at which the reference linearises the new prior = vector2double after double2vector
        x_pose = np.zeros((11, 7)); x_sb = np.zeros((11, 9))
        Ps, Rs, Vs, Bas, Bgs = sol.Ps, sol.Rs, sol.Vs, sol.Bas, sol.Bgs
        for i in range(11):
            x_pose[i, :3] = Ps[i]; x_pose[i, 3:] = BP.R_to_quat(Rs[i])
            x_sb[i] = np.concatenate([Vs[i], Bas[i], Bgs[i]])
        x_lam = 1.0 / (1.0 / sol.lam[: pb.M])
        Ar, br = numpy_marginalize(cfg, pb, x_pose, x_sb, pb.ex, x_lam, pb.prior)
        JtJ, Jtr = BP.prior_normal_equations(sol.new_prior)
        scale = np.abs(Ar).max()
        assert JtJ.shape == Ar.shape
        assert np.abs(JtJ - Ar).max() <= 1e-7 * scale          # the commented invariant J^T J ~ A
        assert np.abs(Jtr - br).max() <= 1e-7 * max(1.0, np.abs(br).max())
        kinds = [(k, i) for (k, i, s, ix, x0) in BP.prior_blocks(sol.new_prior)]
        assert kinds[0] == (B.BLK_EXPOSE, 0) and (B.BLK_POSE, 0) in kinds and (B.BLK_SPEEDBIAS, 0) in kinds
        sim.commit(a, sol)


def test_window_chain_converges_and_improves():
    cfg = make_cfg()
    sim = BP.WindowSimulator(5, cfg, n_landmarks=150)
    for a in range(4):
        pb = sim.window(a)
        sol = ba_ref.solve(cfg, pb)
        c = sol.c
        assert sol.rc == 0 and c.final_cost < c.initial_cost and c.final_cost < 1e3
        ptrue = np.array([sim.true_state(a + i)[0] for i in range(11)])
        e0 = np.abs((pb.pose[:, :3] - pb.pose[0, :3]) - (ptrue - ptrue[0])).max()
        e1 = np.abs((sol.Ps - sol.Ps[0]) - (ptrue - ptrue[0])).max()
        assert e1 < e0 + 1e-3
        # gauge fix: frame 0 keeps its position and yaw (estimator.cpp:998-1031)
        assert np.allclose(sol.Ps[0], pb.pose[0, :3], atol=1e-12)
        sim.commit(a, sol)


def test_solver_agrees_with_scipy_at_convergence():
    """Run-to-convergence: the oracle's dogleg/Schur solver and scipy's trust-region
    least_squares (cauchy loss, numerical Jacobian) must reach the same minimum."""
    from scipy.optimize import least_squares
    cfg = make_cfg()
    sim = BP.WindowSimulator(21, cfg, n_landmarks=24, flag2_frac=0.0)
    pb0 = sim.window(0)
    sol = ba_ref.solve(cfg, pb0)
    sim.commit(0, sol)
    pb = sim.window(1)                      # has a prior => no gauge freedom
    pb.c.max_iterations = 8
    sol = ba_ref.solve(cfg, pb)
    M = pb.M

    def unpack(d):
        pose = np.array([pose_plus(pb.pose[i], d[6 * i:6 * i + 6]) for i in range(11)])
        sb = pb.sb + d[66:165].reshape(11, 9)
        lam = pb.lam + d[165:165 + M]
        return pose, sb, lam

    def residuals(d):
        pose, sb, lam = unpack(d)
        out = []
        for l in range(M):
            o0 = pb.obs_ptr[l]; i = pb.start[l]
            for k in range(1, pb.obs_ptr[l + 1] - o0):
                r = ba_ref.projection_eval(pose[i], pose[i + k], pb.ex, lam[l], pb.obs_pts[o0], pb.obs_pts[o0 + k], jac=False)[0]
                s = r @ r
                out.append(r * np.sqrt(np.log1p(s) / s) if s > 0 else r)     # rho(s) = log(1+s)
        for j in range(1, 11):
            out.append(ba_ref.imu_eval(pb.imu[j - 1], pose[j - 1], sb[j - 1], pose[j], sb[j], jac=False)[0])
        out.append(ba_ref.prior_residual(pb.prior, pose, sb, pb.ex))
        return np.concatenate(out)

    res = least_squares(residuals, np.zeros(165 + M), method="trf", xtol=1e-13, ftol=1e-13, gtol=1e-11, max_nfev=400)
    cost_scipy = 0.5 * np.sum(res.fun ** 2)
    # (a) 8 Ceres-style iterations get close to the true minimum from above.  (Convergence is
    # only linear: with CauchyLoss Ceres drops the rho'' term of the Hessian, marginalization_
    # factor.cpp:39-72 restates the same Corrector, so the GN model over-estimates curvature.)
    assert cost_scipy - 1e-9 <= sol.c.final_cost <= 1.05 * cost_scipy
    # (b) the true minimiser is a fixed point of the oracle's solver: restarted there it
    # neither lowers the cost nor moves the states.
    pose_s, sb_s, lam_s = unpack(res.x)
    pb.pose[:] = pose_s; pb.sb[:] = sb_s; pb.lam[:] = lam_s
    pb.finalize()
    pb.c.max_iterations = 20
    sol2 = ba_ref.solve(cfg, pb)
    assert abs(sol2.c.initial_cost - cost_scipy) <= 1e-9 * cost_scipy
    assert cost_scipy * (1 - 1e-7) <= sol2.c.final_cost <= cost_scipy * (1 + 1e-12)
    assert np.abs(sol2.pose[:, :3] - pose_s[:, :3]).max() < 1e-5
    assert np.abs(sol2.lam[:M] - lam_s).max() < 1e-4


def test_projection_td_jacobian_finite_difference():
    """ProjectionTdFactor (projection_td_factor.cpp:34-150): the check the reference leaves unused
    (ProjectionTdFactor::check, :152-267), incl. the td column and the rolling-shutter row term."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        pi = rand_pose(rng, 0.3); pj = pose_plus(pi, rng.normal(0, 0.05, 6))
        ex = pose_plus(np.array([0.05, -0.02, 0.01, 0, 0, 0, 1.0]), rng.normal(0, 0.05, 6))
        lam = rng.uniform(0.2, 0.8); td = rng.normal(0, 0.01)
        pts_i = rng.uniform(-0.4, 0.4, 2); pts_j = pts_i + rng.normal(0, 0.02, 2)
        vi, vj = rng.normal(0, 0.3, 2), rng.normal(0, 0.3, 2)
        tdi, tdj = rng.normal(0, 0.01, 2); rows = rng.uniform(0, 480, 2); tor = 0.033 / 480
        args = dict(td_i=tdi, td_j=tdj, row_i=rows[0], row_j=rows[1], tr_over_row=tor)
        r, Ji, Jj, Je, Jf, Jt = ba_ref.projection_td_eval(pi, pj, ex, lam, td, pts_i, pts_j, vi, vj, **args)
        # with zero velocity the factor reduces to ProjectionFactor
        r0 = ba_ref.projection_td_eval(pi, pj, ex, lam, td, pts_i, pts_j, 0 * vi, 0 * vj, jac=False, **args)[0]
        assert np.allclose(r0, ba_ref.projection_eval(pi, pj, ex, lam, pts_i, pts_j, jac=False)[0], atol=1e-12)
        eps = 1e-6
        for blk, J in ((0, Ji), (1, Jj), (2, Je)):
            for c in range(6):
                d = np.zeros(6); d[c] = eps
                a3 = [pi, pj, ex]
                a3[blk] = pose_plus(a3[blk], d)
                r2 = ba_ref.projection_td_eval(a3[0], a3[1], a3[2], lam, td, pts_i, pts_j, vi, vj, jac=False, **args)[0]
                assert np.allclose((r2 - r) / eps, J[:, c], rtol=2e-4, atol=2e-3)
        r2 = ba_ref.projection_td_eval(pi, pj, ex, lam + eps, td, pts_i, pts_j, vi, vj, jac=False, **args)[0]
        assert np.allclose((r2 - r) / eps, Jf, rtol=2e-4, atol=2e-3)
        r2 = ba_ref.projection_td_eval(pi, pj, ex, lam, td + eps, pts_i, pts_j, vi, vj, jac=False, **args)[0]
        assert np.allclose((r2 - r) / eps, Jt, rtol=2e-4, atol=2e-3)


def test_td_and_extrinsic_estimation_recover_truth():
    """estimate_td / estimate_extrinsic (estimator.cpp:1186-1212,1270-1285): with para_Td and para_Ex_Pose variable the
    window chain tracks the simulated camera-IMU time offset (the estimate moves one-to-one with the true offset; the
    simulator's IMU sampling leaves a constant bias, so two offsets are compared) and pulls the perturbed extrinsic
    rotation towards the true one; the new prior carries a td block (kind 3, size 1) next to the ex block."""
    est = {}
    for td_true in (0.03, -0.03):
        cfg = make_cfg()
        cfg.estimate_td = 1; cfg.tr = 0.0; cfg.row = 480
        sim = BP.WindowSimulator(9, cfg, n_landmarks=120, td_true=td_true, td_constant=0, ex_constant=0, ex_perturb=0.03,
                                 tic=np.array([0.05, -0.03, 0.02]), pix_noise=0.1)
        ang0 = np.linalg.norm(synth.so3_log(sim.ric.T @ sim.ex_est[1]))
        for a in range(3):
            pb = sim.window(a)
            sol = ba_ref.solve(cfg, pb)
            assert sol.rc == 0 and sol.c.final_cost < sol.c.initial_cost
            sim.commit(a, sol)
            kinds = [(k, s) for (k, i, s, ix, x0) in BP.prior_blocks(sol.new_prior)]
            assert (B.BLK_EXPOSE, 7) in kinds and (B.BLK_TD, 1) in kinds
        est[td_true] = sim.td_est
        # (the extrinsic translation is barely observable under this gentle rotation; the rotation is)
        assert np.linalg.norm(synth.so3_log(sim.ric.T @ sim.ex_est[1])) < 0.75 * ang0
    assert abs((est[0.03] - est[-0.03]) - 0.06) < 0.2 * 0.06


def _pivoted_cholesky_c0(A, b, eps=1e-8):
    """numpy restatement of the c0 computation in k_ba_marg: diagonally pivoted Cholesky of the (semi-definite) prior
    information matrix that stops at pivots <= eps, forward substitution folded in; returns 1/2 b^T A^+ b."""
    M = A.copy(); y = b.copy()
    done = np.zeros(len(b), bool)
    c = 0.0
    for _ in range(len(b)):
        d = np.where(done, -1.0, np.diag(M))
        p = int(np.argmax(d))
        if not d[p] > eps:
            break
        l = np.sqrt(d[p]); yk = y[p] / l
        c += yk * yk
        col = np.where(done, 0.0, M[:, p] / l); col[p] = 0.0
        y -= col * yk
        done[p] = True
        M -= np.outer(col, col)
    return 0.5 * c


def test_information_form_of_the_prior_is_equivalent():
    """The GPU keeps the prior as (HP, gp, c0) = (J0^T J0, J0^T r0, r0^T r0 / 2) and evaluates
    c0 + dx.(gp + HP dx / 2) instead of 1/2 |r0 + J0 dx|^2.  On the oracle's priors (factor form,
    eigenvalues <= 1e-8 truncated like the reference): (i) the two cost expressions agree for random dx, (ii) the
    pivoted-Cholesky constant equals r0^T r0 / 2 although the matrix is rank deficient."""
    cfg = make_cfg()
    sim = BP.WindowSimulator(12, cfg, n_landmarks=120)
    rng = np.random.default_rng(0)
    for a in range(3):
        pb = sim.window(a)
        sol = ba_ref.solve(cfg, pb)
        sim.commit(a, sol)
        P = sol.new_prior
        n = P.n
        J = np.ctypeslib.as_array(P.linearized_jacobians)[: n * n].reshape(n, n).copy()
        r0 = np.ctypeslib.as_array(P.linearized_residuals)[:n].copy()
        HP, gp, c0 = J.T @ J, J.T @ r0, 0.5 * r0 @ r0
        for _ in range(5):
            dx = rng.normal(0, 1e-2, n)
            lhs = 0.5 * np.sum((r0 + J @ dx) ** 2)
            rhs = c0 + dx @ (gp + 0.5 * (HP @ dx))
            assert abs(lhs - rhs) <= 1e-9 * max(1.0, lhs)
        assert np.linalg.matrix_rank(J) < n or a > 0          # the first prior has truncated (zero) rows
        c_piv = _pivoted_cholesky_c0(HP, gp)
        assert abs(c_piv - c0) <= 1e-6 * max(c0, 1e-12), (a, c_piv, c0)


def _ceres_interpolating_step(xs, vals, grads, lo, hi):
    """Ceres' FindInterpolatingPolynomial + MinimizePolynomial written with numpy's own tools: np.linalg.solve for the
    Vandermonde-type system and np.roots -- the eigenvalues of the companion matrix, which is how Ceres' FindPolynomialRoots
    obtains the roots -- taking the real part of every root like MinimizePolynomial does."""
    n = 2 * len(xs)
    deg = n - 1
    A, b = [], []
    for x, v, g in zip(xs, vals, grads):
        A.append([x ** (deg - j) for j in range(n)]); b.append(v)
        A.append([(deg - j) * x ** (deg - j - 1) if j < deg else 0.0 for j in range(n)]); b.append(g)
    c = np.linalg.solve(np.array(A), np.array(b))
    cand = [(lo + hi) / 2.0, lo, hi]
    cand += [r.real for r in np.roots(np.polyder(c)) if lo <= r.real <= hi]
    vals_ = [np.polyval(c, t) for t in cand]
    return cand[int(np.argmin(vals_))], c


def test_line_search_step_matches_companion_matrix_minimiser():
    """The step-size rule of the projected Armijo line search (oracle/ba_ref.c::ls_interpolating_step: only the real
    critical points inside the interval, bracketed through the derivative's roots and bisected)
    against an independent statement with numpy's companion-matrix roots: cubic (two samples) and quintic (three samples)
    interpolants, contraction interval [1e-3 t, 0.6 t] as ArmijoLineSearch::DoSearch passes it."""
    rng = np.random.default_rng(5)
    n_checked = 0
    for trial in range(400):
        ns = 2 + trial % 2
        t_cur = float(rng.uniform(0.05, 1.0))
        xs = [0.0, t_cur] + ([float(rng.uniform(t_cur * 1.5, t_cur * 5.0))] if ns == 3 else [])
        f0 = float(rng.uniform(10, 100))
        g0 = -float(rng.uniform(0.1, 5.0))
        vals = [f0] + [f0 + float(rng.uniform(-0.2, 2.0)) * x for x in xs[1:]]
        grads = [g0] + [float(rng.uniform(-3.0, 6.0)) for _ in xs[1:]]
        lo, hi = 1e-3 * t_cur, 0.6 * t_cur
        want, c = _ceres_interpolating_step(xs, vals, grads, lo, hi)
        got = ba_ref.ls_interpolating_step(xs, vals, grads, lo, hi)
        # equal minimisers, or -- when two candidates tie to rounding -- equal polynomial values
        if abs(got - want) > 1e-9 * max(1.0, abs(want)):
            assert abs(np.polyval(c, got) - np.polyval(c, want)) <= 1e-9 * max(1.0, abs(np.polyval(c, want))), (trial, got, want)
        n_checked += 1
    assert n_checked == 400
    # a sample that could not be evaluated: bisection inside the interval
    assert ba_ref.ls_interpolating_step([0.0, 0.8], [1.0, np.nan], [-1.0, np.nan], 1e-3 * 0.8, 0.6 * 0.8) == 0.4


class DenseCeresRestatement:
    """A second, structurally independent statement of the solver the reference configures (ceres::Solve with DENSE_SCHUR + DOGLEG,
    Ceres defaults otherwise, estimator.cpp:1348-1363) -- dense numpy linear algebra on the full Jacobian, no Schur complement, no
    packed storage, its own line search -- used to cross-check oracle/ba_ref.c::solve.  Only the single-factor evaluations
    (projection / IMU / prior residuals and Jacobians, pinned by the finite-difference tests above) are shared with the C oracle.
    Tangent layout: pose f -> 6 f, speed-bias f -> 66 + 9 f, ex-pose -> 165, td -> 171, landmark l -> 172 + l; x = (pose, sb, lam, ex, td)."""

    NC = 172

    def __init__(self, cfg, pb, use_imu=True):
        self.cfg, self.pb, self.M = cfg, pb, pb.M
        self.ub = np.where(pb.flag == 2, 2.0 / cfg.depth_max_dist, np.inf)
        self.NT = self.NC + self.M
        self.td_factor = bool(cfg.estimate_td)
        # Constant blocks are not part of the Ceres program -- their columns do not exist: the ex-pose / td unless they are estimated
        # (estimator.cpp:1186-1212); in VO mode (USE_IMU = 0: no IMU factors) the speed-bias blocks and para_Pose[0] (:1174-1185)
        self.use_imu = use_imu
        self.active = np.ones(self.NT, bool)
        self.active[165:171] = not pb.c.ex_constant
        self.active[171] = self.td_factor and not (use_imu and pb.c.td_constant)
        if not use_imu:
            self.active[0:6] = False; self.active[66:165] = False
        P = pb.prior
        self.pcols = None
        if P is not None:
            n = P.n
            self.J0 = np.ctypeslib.as_array(P.linearized_jacobians)[: n * n].reshape(n, n).copy()
            cols = np.full(n, -1)
            for b in P.blocks[: P.n_blocks]:
                ls = 6 if b.size == 7 else b.size
                base = {B.BLK_POSE: 6 * b.index, B.BLK_SPEEDBIAS: 66 + 9 * b.index, B.BLK_EXPOSE: 165, B.BLK_TD: 171}[b.kind]
                cols[b.idx:b.idx + ls] = base + np.arange(ls)
            cols[cols >= 0] = np.where(self.active[cols[cols >= 0]], cols[cols >= 0], -1)
            self.pcols = cols

    def start(self):
        pb = self.pb
        return (pb.pose.copy(), pb.sb.copy(), np.minimum(pb.lam.copy(), self.ub), np.array(pb.ex, float), float(pb.td))

    def plus(self, x, d):
        pose, sb, lam, ex, td = x
        d = np.where(self.active, d, 0.0)
        pose2 = np.array([pose_plus(pose[f], d[6 * f:6 * f + 6]) for f in range(11)])
        sb2 = sb + d[66:165].reshape(11, 9)
        lam2 = np.minimum(lam + d[172:], self.ub)                       # ParameterBlock::Plus projects onto the bounds
        return pose2, sb2, lam2, pose_plus(ex, d[165:171]), td + d[171]

    def evaluate(self, x, jac=True):
        pose, sb, lam, ex, td = x
        pb, M = self.pb, self.M
        rows_r, rows_J, cost = [], [], 0.0
        for l in range(M):
            o0, i = pb.obs_ptr[l], pb.start[l]
            for k in range(1, pb.obs_ptr[l + 1] - o0):
                Jt = np.zeros(2)
                if self.td_factor:
                    r, Ji, Jj, Je, Jf, Jt = ba_ref.projection_td_eval(
                        pose[i], pose[i + k], ex, lam[l], td, pb.obs_pts[o0], pb.obs_pts[o0 + k], pb.obs_vel[o0], pb.obs_vel[o0 + k],
                        pb.obs_cur_td[o0], pb.obs_cur_td[o0 + k], pb.obs_row[o0], pb.obs_row[o0 + k], self.cfg.tr / self.cfg.row, jac=jac)
                else:
                    r, Ji, Jj, Je, Jf = ba_ref.projection_eval(pose[i], pose[i + k], ex, lam[l], pb.obs_pts[o0], pb.obs_pts[o0 + k], jac=jac)
                s = r @ r
                cost += 0.5 * np.log1p(s)
                if jac:
                    w = np.sqrt(1.0 / (1.0 + s))          # CauchyLoss: rho'' < 0 => the Corrector is the plain sqrt(rho') scaling
                    J = np.zeros((2, self.NT))
                    J[:, 6 * i:6 * i + 6] = w * Ji[:, :6]; J[:, 6 * (i + k):6 * (i + k) + 6] = w * Jj[:, :6]; J[:, 172 + l] = w * Jf
                    J[:, 165:171] = w * Je[:, :6]; J[:, 171] = w * Jt
                    rows_J.append(J); rows_r.append(w * r)
        for j in range(1, 11 if self.use_imu else 0):
            r, Jpi, Jsi, Jpj, Jsj = ba_ref.imu_eval(pb.imu[j - 1], pose[j - 1], sb[j - 1], pose[j], sb[j], jac=jac)
            cost += 0.5 * r @ r
            if jac:
                J = np.zeros((15, self.NT))
                J[:, 6 * (j - 1):6 * j] = Jpi[:, :6]; J[:, 66 + 9 * (j - 1):66 + 9 * j] = Jsi
                J[:, 6 * j:6 * j + 6] = Jpj[:, :6]; J[:, 66 + 9 * j:66 + 9 * j + 9] = Jsj
                rows_J.append(J); rows_r.append(r)
        if self.pcols is not None:
            r = ba_ref.prior_residual(pb.prior, pose, sb, ex, td)
            cost += 0.5 * r @ r
            if jac:
                J = np.zeros((len(r), self.NT))
                keep = self.pcols >= 0
                J[:, self.pcols[keep]] = self.J0[:, keep]
                rows_J.append(J); rows_r.append(r)
        if not jac:
            return cost
        J = np.vstack(rows_J)
        J[:, ~self.active] = 0.0
        return cost, J, np.concatenate(rows_r)

    def ambient(self, x):
        """the non-constant parameter blocks in the ambient parameterisation (what Ceres' step / parameter norms run over)"""
        pose = x[0] if self.use_imu else x[0][1:]
        parts = [pose.ravel(), x[1].ravel() if self.use_imu else np.zeros(0), x[2]]
        if self.active[165]: parts.append(x[3])
        if self.active[171]: parts.append(np.array([x[4]]))
        return np.concatenate(parts)

    def line_search(self, x, delta, x_cost, gts, cand_cost):
        """TrustRegionMinimizer::DoLineSearch / ArmijoLineSearch::DoSearch with CUBIC interpolation (numpy polynomial tools)."""
        dmax = np.abs(delta).max()
        samples = [(0.0, x_cost, gts)]
        cur = prev = None
        t = 1.0
        for it in range(20):
            if it > 0:
                pts = [samples[0], cur] + ([prev] if prev is not None else [])
                t, _ = _ceres_interpolating_step([p[0] for p in pts], [p[1] for p in pts], [p[2] for p in pts], 1e-3 * cur[0], 0.6 * cur[0])
                if t * dmax < 1e-9:
                    return None
            prev = cur
            c, J, r = self.evaluate(self.plus(x, t * delta))
            cur = (t, c, float(delta @ (J.T @ r)))
            if c <= x_cost + 1e-4 * gts * t:
                return (t, c) if it > 0 else None
        return None

    def solve(self, max_iter=8):
        x = self.start()                                  # TrustRegionMinimizer::Init projects the start point onto the bounds
        x_cost, J, r = self.evaluate(x)
        jscale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
        constrained = bool(np.isfinite(self.ub).any())
        radius, mu, reuse, invalid = 1e4, 1e-8, False, 0
        it = succ = 0
        ls_runs = ls_short = 0
        x_norm = np.linalg.norm(self.ambient(x))
        trace = []
        need_lin = True
        while True:
            if need_lin:
                Js = J * jscale
                g = Js.T @ r
                gmax = np.abs(self.ambient(x) - self.ambient(self.plus(x, -g / jscale))).max()
                need_lin = False
            if it >= max_iter: term = 0; break
            if gmax <= 1e-10: term = 2; break
            if radius <= 1e-32: term = 4; break
            it += 1
            if not reuse:
                reuse = True
                H = Js.T @ Js
                D = np.sqrt(np.clip(np.diag(H), 1e-6, 1e32))
                gd = g / D
                alpha = (gd @ gd) / np.sum((Js @ (gd / D)) ** 2)
                while mu < 1.0:
                    try:
                        L = np.linalg.cholesky(H + mu * np.diag(D * D))
                        y = np.linalg.solve(L.T, np.linalg.solve(L, g))
                        if np.isfinite(y).all():
                            break
                    except np.linalg.LinAlgError:
                        pass
                    mu *= 10.0
                gn = -D * y
            gnorm, gnn = np.linalg.norm(gd), np.linalg.norm(gn)
            if gnn <= radius:
                sd, dn = gn, gnn
            elif gnorm * alpha >= radius:
                sd, dn = -(radius / gnorm) * gd, radius
            else:
                b_dot_a = -alpha * (gd @ gn); a_sq = (alpha * gnorm) ** 2
                bma = a_sq - 2 * b_dot_a + gnn ** 2; cc = b_dot_a - a_sq
                dd = np.sqrt(cc * cc + bma * (radius ** 2 - a_sq))
                beta = (dd - cc) / bma if cc <= 0 else (radius ** 2 - a_sq) / (dd + cc)
                sd = (-alpha * (1 - beta)) * gd + beta * gn; dn = np.linalg.norm(sd)
            step = sd / D
            Jstep = Js @ step
            mcc = -Jstep @ (r + Jstep / 2.0)
            if not mcc > 0:
                invalid += 1
                if invalid >= 5: term = 5; break
                mu *= 10.0; reuse = False
                continue
            invalid = 0
            delta = step * jscale
            cand = self.plus(x, delta)
            cand_cost = self.evaluate(cand, jac=False)
            if constrained:
                gts = float(g @ step)
                if cand_cost > x_cost + 1e-4 * gts:
                    ls_runs += 1
                    res = self.line_search(x, delta, x_cost, gts, cand_cost)
                    if res is not None:
                        cand = self.plus(x, res[0] * delta); cand_cost = res[1]; ls_short += 1
            step_norm = np.linalg.norm(self.ambient(x) - self.ambient(cand))
            if step_norm <= 1e-8 * (x_norm + 1e-8): term = 3; break
            cost_change = x_cost - cand_cost
            if abs(cost_change) <= 1e-6 * x_cost: term = 1; break
            rho = cost_change / mcc
            trace.append((x_cost, cand_cost, rho, radius))
            if rho > 1e-3:
                x, x_cost = cand, cand_cost
                x_norm = np.linalg.norm(self.ambient(x))
                _, J, r = self.evaluate(x)
                need_lin = True; succ += 1
                if rho < 0.25: radius *= 0.5
                if rho > 0.75: radius = max(radius, 3.0 * dn)
                mu = max(1e-8, 2.0 * mu / 10.0); reuse = False
            else:
                radius *= 0.5
        return dict(x=x, cost=x_cost, iterations=it, successful=succ, termination=term, line_searches=ls_runs, shortened=ls_short, trace=trace)


@pytest.mark.parametrize("seed,nlm,min_ls,min_short", [(2, 24, 8, 4), (7, 24, 6, 1), (13, 24, 4, 0), (21, 24, 0, 0)])
def test_oracle_solver_matches_an_independent_dense_restatement(seed, nlm, min_ls, min_short):
    """oracle/ba_ref.c::solve (Jacobi scaling, per-landmark Schur complement, Cholesky of the reduced camera system, traditional
    dogleg, Ceres' acceptance / radius / mu rules, bounds projection and projected line search) against DenseCeresRestatement on
    windows with a prior and bounded landmarks next to their bound: same iteration / acceptance / termination sequence, same number
    of line searches, costs and states equal to the conditioning of the normal equations.  The chains cover full steps that pass
    the Armijo test, overshooting steps the search shortens, and searches that fail because the step pushes blocked landmarks
    (a scan over 30 chains found no disagreement)."""
    cfg = make_cfg()
    sim = BP.WindowSimulator(seed, cfg, n_landmarks=nlm, flag2_frac=0.25)
    sol0 = ba_ref.solve(cfg, sim.window(0)); sim.commit(0, sol0)
    n_ls = n_short = 0
    for a in (1, 2):
        pb = sim.window(a)
        so = ba_ref.solve(cfg, pb)
        dn = DenseCeresRestatement(cfg, pb).solve(max_iter=8)
        assert (so.c.iterations, so.c.successful_steps, so.c.termination) == (dn["iterations"], dn["successful"], dn["termination"]), (a, dn["trace"])
        assert so.c.armijo_failures == dn["line_searches"], a
        assert abs(so.c.final_cost - dn["cost"]) <= 1e-7 * dn["cost"], (a, so.c.final_cost, dn["cost"])
        assert np.abs(so.pose - dn["x"][0]).max() <= 1e-6 and np.abs(so.sb - dn["x"][1]).max() <= 1e-5, a
        assert np.abs(so.lam[:pb.M] - dn["x"][2]).max() <= 1e-6, a
        n_ls += dn["line_searches"]; n_short += dn["shortened"]
        sim.commit(a, so)
    assert n_ls >= min_ls and n_short >= min_short, (n_ls, n_short)


def test_oracle_solver_matches_the_dense_restatement_in_vo_mode():
    """USE_IMU = 0: projection factors + prior only, para_Pose[0] constant (estimator.cpp:1182-1185) -- its Jacobian columns must
    not exist in the normal equations (r1 kept them and only ignored the update, which changes the step of every other block)."""
    cfg = make_cfg()
    cfg.use_imu = 0
    sim = BP.WindowSimulator(15, cfg, n_landmarks=30, flag2_frac=0.2)
    for a in range(3):
        pb = sim.window(a)
        pb.c.use_imu = 0
        so = ba_ref.solve(cfg, pb)
        if a > 0:
            dn = DenseCeresRestatement(cfg, pb, use_imu=False).solve(max_iter=8)
            assert (so.c.iterations, so.c.successful_steps, so.c.termination) == (dn["iterations"], dn["successful"], dn["termination"]), (a, dn["trace"])
            assert so.c.armijo_failures == dn["line_searches"], a
            assert abs(so.c.final_cost - dn["cost"]) <= 1e-7 * dn["cost"], (a, so.c.final_cost, dn["cost"])
            assert np.abs(so.pose - dn["x"][0]).max() <= 1e-6 and np.abs(so.lam[:pb.M] - dn["x"][2]).max() <= 1e-6, a
            assert np.array_equal(so.pose[0], pb.pose[0])
        sim.commit(a, so)


@pytest.mark.parametrize("ex_constant,td_constant", [(0, 0), (1, 0), (0, 1)])
def test_oracle_solver_matches_the_dense_restatement_with_td_and_extrinsic(ex_constant, td_constant):
    """ESTIMATE_TD (ProjectionTdFactor) with para_Td and / or para_Ex_Pose variable: the extra tangent columns (165..171), their
    prior blocks and the constant-block handling of the C solver against the dense statement."""
    cfg = make_cfg()
    cfg.estimate_td = 1; cfg.tr = 0.02; cfg.row = 480
    sim = BP.WindowSimulator(33, cfg, n_landmarks=30, td_true=0.02, td_constant=td_constant, ex_constant=ex_constant,
                             ex_perturb=0.0 if ex_constant else 0.02, tic=np.array([0.05, -0.03, 0.02]), flag2_frac=0.2)
    sol0 = ba_ref.solve(cfg, sim.window(0)); sim.commit(0, sol0)
    for a in (1, 2):
        pb = sim.window(a)
        so = ba_ref.solve(cfg, pb)
        dn = DenseCeresRestatement(cfg, pb).solve(max_iter=8)
        assert (so.c.iterations, so.c.successful_steps, so.c.termination) == (dn["iterations"], dn["successful"], dn["termination"]), (a, dn["trace"])
        assert so.c.armijo_failures == dn["line_searches"], a
        assert abs(so.c.final_cost - dn["cost"]) <= 1e-7 * dn["cost"], (a, so.c.final_cost, dn["cost"])
        assert np.abs(so.pose - dn["x"][0]).max() <= 1e-6 and np.abs(so.lam[:pb.M] - dn["x"][2]).max() <= 1e-6, a
        assert np.abs(so.ex - dn["x"][3]).max() <= 1e-6 and abs(so.td - dn["x"][4]) <= 1e-7, a
        if not ex_constant: assert np.abs(so.ex - pb.ex).max() > 1e-6
        if not td_constant: assert abs(so.td - pb.td) > 1e-7
        sim.commit(a, so)


def test_margin_second_new_matches_independent_numpy_schur():
    """MARGIN_SECOND_NEW (estimator.cpp:1504-1575): the new prior is the old prior with para_Pose[WINDOW_SIZE - 1] marginalised out
    -- only the MarginalizationFactor enters, linearised at the states after the gauge fix.  Independent numpy statement: A = J0^T J0,
    b = J0^T r(x), pseudo-inverse Schur complement of the dropped 6 columns (marginalization_factor.cpp:273-296)."""
    cfg = make_cfg()
    sim = BP.WindowSimulator(23, cfg, n_landmarks=60)
    sol = ba_ref.solve(cfg, sim.window(0)); sim.commit(0, sol)
    pb = sim.window(1, marg_flag=B.MARGIN_SECOND_NEW)
    P = pb.prior
    assert any(b.kind == B.BLK_POSE and b.index == 9 for b in P.blocks[: P.n_blocks])
    sol = ba_ref.solve(cfg, pb)
    assert sol.c.has_new_prior == 1
    x_pose = np.zeros((11, 7)); x_sb = np.zeros((11, 9))
    for i in range(11):
        x_pose[i, :3] = sol.Ps[i]; x_pose[i, 3:] = BP.R_to_quat(sol.Rs[i])
        x_sb[i] = np.concatenate([sol.Vs[i], sol.Bas[i], sol.Bgs[i]])
    n = P.n
    J0 = np.ctypeslib.as_array(P.linearized_jacobians)[: n * n].reshape(n, n)
    r = ba_ref.prior_residual(P, x_pose, x_sb, pb.ex)
    A, bv = J0.T @ J0, J0.T @ r
    drop = np.zeros(n, bool)
    old_cols = {}
    for b in P.blocks[: P.n_blocks]:
        ls = 6 if b.size == 7 else b.size
        if b.kind == B.BLK_POSE and b.index == 9: drop[b.idx:b.idx + ls] = True
        old_cols[(b.kind, b.index)] = np.arange(b.idx, b.idx + ls)
    # kept columns in the order of the new prior's block table (indices unchanged: only frames >= WINDOW_SIZE - 1 shift, :1546-1560)
    keep = np.concatenate([old_cols[(k, i)] for (k, i, s_, ix, x0) in BP.prior_blocks(sol.new_prior)])
    assert sorted(keep.tolist()) == np.nonzero(~drop)[0].tolist()
    Amm = 0.5 * (A[np.ix_(drop, drop)] + A[np.ix_(drop, drop)].T)
    w, V = np.linalg.eigh(Amm)
    Ainv = (V * np.where(w > 1e-8, 1.0 / np.where(w > 1e-8, w, 1.0), 0.0)) @ V.T
    Arm = A[np.ix_(keep, np.nonzero(drop)[0])]
    Ar = A[np.ix_(keep, keep)] - Arm @ Ainv @ Arm.T
    br = bv[keep] - Arm @ Ainv @ bv[drop]
    JtJ, Jtr = BP.prior_normal_equations(sol.new_prior)
    assert JtJ.shape == Ar.shape
    assert np.abs(JtJ - Ar).max() <= 1e-7 * np.abs(Ar).max()
    assert np.abs(Jtr - br).max() <= 1e-7 * max(1.0, np.abs(br).max())


def test_preintegration_bias_jacobians_match_repropagation():
    """IntegrationBase keeps d(delta_p, delta_q, delta_v) / d(ba, bg) (`jacobian`, integration_base.h:130-162) so that IMUFactor can
    correct the pre-integrated terms to first order instead of re-propagating (imu_factor.h:44-58).  Independent check: re-propagate
    the same samples with perturbed linearisation biases and compare with the first-order prediction."""
    cfg = make_cfg()
    sim = BP.WindowSimulator(9, cfg)
    ba0, bg0 = sim.ba_true + 0.01, sim.bg_true - 0.002
    pre = sim._preint(2, 3, ba0, bg0)
    J = np.array(pre.jacobian).reshape(15, 15)
    dp0, dv0, dq0 = np.array(pre.delta_p), np.array(pre.delta_v), np.array(pre.delta_q)

    def qmul(a, b):
        return np.concatenate([a[3] * b[:3] + b[3] * a[:3] + np.cross(a[:3], b[:3]), [a[3] * b[3] - a[:3] @ b[:3]]])
    for k in range(6):
        d = np.zeros(6); d[k] = 1e-5
        pre2 = sim._preint(2, 3, ba0 + d[:3], bg0 + d[3:])
        dba, dbg = d[:3], d[3:]
        dp_pred = dp0 + J[0:3, 9:12] @ dba + J[0:3, 12:15] @ dbg
        dv_pred = dv0 + J[6:9, 9:12] @ dba + J[6:9, 12:15] @ dbg
        th = J[3:6, 12:15] @ dbg
        dq_pred = qmul(dq0, np.concatenate([0.5 * th, [1.0]])); dq_pred /= np.linalg.norm(dq_pred)
        # first-order predictions: the error is O(|d|^2) ~ 1e-10, the change itself ~ 1e-6 .. 1e-7
        assert np.abs(np.array(pre2.delta_p) - dp_pred).max() <= 2e-9, k
        assert np.abs(np.array(pre2.delta_v) - dv_pred).max() <= 2e-8, k
        dq2 = np.array(pre2.delta_q)
        assert min(np.abs(dq2 - dq_pred).max(), np.abs(dq2 + dq_pred).max()) <= 2e-9, k
        assert np.abs(np.array(pre2.delta_p) - dp0).max() > 1e-8 or k >= 3       # the perturbation really moved the result


def test_preintegration_covariance_matches_numerical_error_propagation():
    """IntegrationBase::midPointIntegration propagates the 15 x 15 error-state covariance with analytic F and V
    (integration_base.h:96-162): P' = F P F^T + V N V^T, N = diag(ACC_N^2, GYR_N^2, ACC_N^2, GYR_N^2, ACC_W^2, GYR_W^2).
    Independent check: an own numpy midpoint integrator, F and V of every step by central differences on the error state
    [d_alpha, d_theta (right-multiplied), d_beta, d_ba, d_bg] and on the four measurement noises, same recursion."""
    cfg = make_cfg()
    rng = np.random.default_rng(12)
    n, dt = 20, 0.005
    acc = [np.array([0.3, -0.2, 9.7]) + rng.normal(0, 0.5, 3) for _ in range(n + 1)]
    gyr = [np.array([0.1, 0.3, -0.2]) + rng.normal(0, 0.2, 3) for _ in range(n + 1)]
    ba, bg = np.array([0.02, -0.01, 0.015]), np.array([0.003, -0.002, 0.001])
    pre = ba_ref.preintegrate([(dt, acc[i], gyr[i]) for i in range(1, n + 1)], acc[0], gyr[0], ba, bg, cfg)

    def qmul(a, b):
        return np.concatenate([a[3] * b[:3] + b[3] * a[:3] + np.cross(a[:3], b[:3]), [a[3] * b[3] - a[:3] @ b[:3]]])

    def qrot(q, v):
        return v + 2.0 * np.cross(q[:3], np.cross(q[:3], v) + q[3] * v)

    def step(state, a0, w0, a1, w1):
        al, q, be, ba_, bg_ = state
        ua0 = qrot(q, a0 - ba_)
        ug = 0.5 * (w0 + w1) - bg_
        q2 = qmul(q, np.concatenate([ug * dt / 2.0, [1.0]]))
        ua = 0.5 * (ua0 + qrot(q2, a1 - ba_))          # (the reference rotates with the not yet normalised product, :74-76,190)
        return (al + be * dt + 0.5 * ua * dt * dt, q2 / np.linalg.norm(q2), be + ua * dt, ba_, bg_)

    def perturb(state, e):          # error state -> state
        al, q, be, ba_, bg_ = state
        q2 = qmul(q, np.concatenate([0.5 * e[3:6], [1.0]])); q2 /= np.linalg.norm(q2)
        return (al + e[0:3], q2, be + e[6:9], ba_ + e[9:12], bg_ + e[12:15])

    def err(nom, st):               # state -> error state around nom
        qi = np.concatenate([-nom[1][:3], [nom[1][3]]])
        dq = qmul(qi, st[1])
        return np.concatenate([st[0] - nom[0], 2.0 * dq[:3] * np.sign(dq[3]), st[2] - nom[2], st[3] - nom[3], st[4] - nom[4]])

    N = np.diag(np.repeat([cfg.acc_n ** 2, cfg.gyr_n ** 2, cfg.acc_n ** 2, cfg.gyr_n ** 2, cfg.acc_w ** 2, cfg.gyr_w ** 2], 3))
    state = (np.zeros(3), np.array([0, 0, 0, 1.0]), np.zeros(3), ba, bg)
    Pc = np.zeros((15, 15))
    h = 1e-6
    for k in range(1, n + 1):
        a0, w0, a1, w1 = acc[k - 1], gyr[k - 1], acc[k], gyr[k]
        nxt = step(state, a0, w0, a1, w1)
        F = np.zeros((15, 15)); V = np.zeros((15, 18))
        for c in range(15):
            e = np.zeros(15); e[c] = h
            F[:, c] = (err(nxt, step(perturb(state, e), a0, w0, a1, w1)) - err(nxt, step(perturb(state, -e), a0, w0, a1, w1))) / (2 * h)
        for c in range(12):
            d = np.zeros(12); d[c] = h
            V[:, c] = (err(nxt, step(state, a0 + d[0:3], w0 + d[3:6], a1 + d[6:9], w1 + d[9:12])) -
                       err(nxt, step(state, a0 - d[0:3], w0 - d[3:6], a1 - d[6:9], w1 - d[9:12]))) / (2 * h)
        V[9:12, 12:15] = np.eye(3) * dt; V[12:15, 15:18] = np.eye(3) * dt          # bias random walks
        Pc = F @ Pc @ F.T + V @ N @ V.T
        state = nxt
    # the integrators agree ...
    assert np.abs(np.array(pre.delta_p) - state[0]).max() <= 1e-12 and np.abs(np.array(pre.delta_v) - state[2]).max() <= 1e-12
    dq = np.array(pre.delta_q)
    assert min(np.abs(dq - state[1]).max(), np.abs(dq + state[1]).max()) <= 1e-12
    # ... and so do the covariances.  The reference's F is first order in dt in its rotation blocks (I - [w]x dt for exp(-[w]x dt),
    # integration_base.h:104-127), the central differences are exact derivatives of the discrete step: they differ at
    # O((|w| dt)^2) per step, a few 1e-5 of sqrt(P_ii P_jj) after 20 steps; a wrong block or sign would show at O(1).
    cov = np.array(pre.covariance).reshape(15, 15)
    scale = np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    assert np.abs(cov - Pc).max() <= 1e-6 * np.abs(cov).max()
    assert np.abs((cov - Pc) / scale).max() <= 2e-4
    assert np.abs(np.diag(cov) / np.diag(Pc) - 1.0).max() <= 1e-4
