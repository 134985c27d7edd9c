"""ctypes wrapper of oracle/ba_ref.c.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

from vrf_b200 import binding as B   # struct layouts of include/vrf_ba.h only

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbaref.so")
_lib = None


class OraclePreint(C.Structure):
    _fields_ = [("acc_0", C.c_double * 3), ("gyr_0", C.c_double * 3), ("s", B.VrfImuPreint), ("noise", C.c_double * 18)]


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = C.CDLL(_SO)
        _lib.oracle_ba_solve.argtypes = [C.POINTER(B.VrfConfig), C.POINTER(B.VrfBaProblem), C.POINTER(B.VrfBaResult)]
        _lib.oracle_projection_eval.argtypes = [C.c_void_p] * 3 + [C.c_double] + [C.c_void_p] * 7
        _lib.oracle_projection_eval.restype = None
        _lib.oracle_projection_td_eval.argtypes = [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_void_p] * 4 + [C.c_double] * 5 + [C.c_void_p] * 6
        _lib.oracle_projection_td_eval.restype = None
        _lib.oracle_imu_eval.argtypes = [C.POINTER(B.VrfImuPreint)] + [C.c_void_p] * 4 + [C.c_double] + [C.c_void_p] * 5
        _lib.oracle_preint_init.argtypes = [C.POINTER(OraclePreint)] + [C.c_void_p] * 4 + [C.c_double] * 4
        _lib.oracle_preint_init.restype = None
        _lib.oracle_preint_propagate.argtypes = [C.POINTER(OraclePreint), C.c_double, C.c_void_p, C.c_void_p]
        _lib.oracle_preint_propagate.restype = None
        _lib.oracle_prior_residual.argtypes = [C.POINTER(B.VrfPrior)] + [C.c_void_p] * 3 + [C.c_double, C.c_void_p]
        _lib.oracle_prior_residual.restype = None
        _lib.oracle_ls_interpolating_step.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_double, C.c_double]
        _lib.oracle_ls_interpolating_step.restype = C.c_double
    return _lib


def ls_interpolating_step(xs, vals, grads, min_step, max_step):
    """Ceres' InterpolatingPolynomialMinimizingStepSize (CUBIC) on samples [lower bound, current(, previous)]."""
    a = [np.ascontiguousarray(v, np.float64) for v in (xs, vals, grads)]
    return float(lib().oracle_ls_interpolating_step(_p(a[0]), _p(a[1]), _p(a[2]), len(a[0]), float(min_step), float(max_step)))


def _p(a):
    return None if a is None else a.ctypes.data


def projection_eval(pose_i, pose_j, ex, inv_dep, pts_i, pts_j, jac=True):
    r = np.zeros(2)
    Ji = np.zeros((2, 7)); Jj = np.zeros((2, 7)); Je = np.zeros((2, 7)); Jf = np.zeros(2)
    a = [np.ascontiguousarray(v, np.float64) for v in (pose_i, pose_j, ex)]
    pi = np.ascontiguousarray(pts_i, np.float64); pj = np.ascontiguousarray(pts_j, np.float64)
    lib().oracle_projection_eval(_p(a[0]), _p(a[1]), _p(a[2]), float(inv_dep), _p(pi), _p(pj), _p(r),
                                 _p(Ji) if jac else None, _p(Jj) if jac else None, _p(Je) if jac else None, _p(Jf) if jac else None)
    return r, Ji, Jj, Je, Jf


def projection_td_eval(pose_i, pose_j, ex, inv_dep, td, pts_i, pts_j, vel_i, vel_j, td_i, td_j, row_i, row_j, tr_over_row, jac=True):
    r = np.zeros(2)
    Ji = np.zeros((2, 7)); Jj = np.zeros((2, 7)); Je = np.zeros((2, 7)); Jf = np.zeros(2); Jt = np.zeros(2)
    a = [np.ascontiguousarray(v, np.float64) for v in (pose_i, pose_j, ex, pts_i, pts_j, vel_i, vel_j)]
    lib().oracle_projection_td_eval(_p(a[0]), _p(a[1]), _p(a[2]), float(inv_dep), float(td), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]),
                                    float(td_i), float(td_j), float(row_i), float(row_j), float(tr_over_row), _p(r),
                                    _p(Ji) if jac else None, _p(Jj) if jac else None, _p(Je) if jac else None,
                                    _p(Jf) if jac else None, _p(Jt) if jac else None)
    return r, Ji, Jj, Je, Jf, Jt


def imu_eval(pre, pose_i, sb_i, pose_j, sb_j, g_norm=9.81, jac=True):
    r = np.zeros(15)
    Jpi = np.zeros((15, 7)); Jsi = np.zeros((15, 9)); Jpj = np.zeros((15, 7)); Jsj = np.zeros((15, 9))
    a = [np.ascontiguousarray(v, np.float64) for v in (pose_i, sb_i, pose_j, sb_j)]
    rc = lib().oracle_imu_eval(C.byref(pre), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), g_norm, _p(r),
                               _p(Jpi) if jac else None, _p(Jsi) if jac else None, _p(Jpj) if jac else None, _p(Jsj) if jac else None)
    assert rc == 0
    return r, Jpi, Jsi, Jpj, Jsj


def preintegrate(samples_dt_acc_gyr, acc0, gyr0, ba, bg, cfg):
    """IntegrationBase: ctor with (acc_0, gyr_0, linearized_ba, linearized_bg), then push_back(dt, acc, gyr)..."""
    p = OraclePreint()
    v = [np.ascontiguousarray(x, np.float64) for x in (acc0, gyr0, ba, bg)]
    lib().oracle_preint_init(C.byref(p), _p(v[0]), _p(v[1]), _p(v[2]), _p(v[3]), cfg.acc_n, cfg.gyr_n, cfg.acc_w, cfg.gyr_w)
    for dt, acc, gyr in samples_dt_acc_gyr:
        a = np.ascontiguousarray(acc, np.float64); g = np.ascontiguousarray(gyr, np.float64)
        lib().oracle_preint_propagate(C.byref(p), float(dt), _p(a), _p(g))
    out = B.VrfImuPreint()
    C.memmove(C.byref(out), C.byref(p.s), C.sizeof(B.VrfImuPreint))
    return out


def prior_residual(prior, pose, sb, ex, td=0.0):
    r = np.zeros(B.PRIOR_MAX_DIM)
    a = [np.ascontiguousarray(v, np.float64) for v in (pose, sb, ex)]
    lib().oracle_prior_residual(C.byref(prior), _p(a[0]), _p(a[1]), _p(a[2]), float(td), _p(r))
    return r[: prior.n]


def solve(cfg, prob_holder):
    """prob_holder: vrf_b200.ba_problem.BaProblem (owns the numpy buffers). Returns BaSolution."""
    from vrf_b200.ba_problem import BaSolution
    sol = BaSolution(prob_holder.M)
    rc = lib().oracle_ba_solve(C.byref(cfg), C.byref(prob_holder.c), C.byref(sol.c))
    sol.rc = rc
    return sol
