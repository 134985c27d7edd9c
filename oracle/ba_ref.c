/*
 * ba_ref.c -- CPU oracle of the back-end hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build or call this file.
 *
 * Double-precision restatement of Estimator::optimization()
 * (reference vins_estimator/src/estimator/estimator.cpp:1161-1578):
 *   - ProjectionFactor::Evaluate        factor/projection_factor.cpp:22-130
 *   - IMUFactor::Evaluate               factor/imu_factor.h:20-205
 *   - IntegrationBase::evaluate/propagate factor/integration_base.h:56-195
 *   - MarginalizationFactor::Evaluate   factor/marginalization_factor.cpp:353-415
 *   - ResidualBlockInfo::Evaluate (loss) factor/marginalization_factor.cpp:3-73
 *   - MarginalizationInfo::marginalize  factor/marginalization_factor.cpp:181-315
 *   - PoseLocalParameterization         factor/pose_local_parameterization.cpp:3-28
 *   - Utility::deltaQ/Qleft/Qright/R2ypr/ypr2R utility/utility.h:12-139
 *   - double2vector gauge fix           estimator.cpp:985-1111
 *
 * ceres::Solve itself is a third-party dependency that is NOT vendored in the
 * reference (Ceres 2.1.0 in README.md:96-99, 2.0.0 in doc/INSTALL.md:24-26) and is
 * not installed in this image, so the solver loop below restates Ceres' published
 * trust-region algorithm (trust_region_minimizer.cc, dogleg_strategy.cc
 * TRADITIONAL_DOGLEG, DENSE_SCHUR, Jacobi scaling, CauchyLoss + Corrector) with the
 * options the reference sets (estimator.cpp:1348-1360) and Ceres defaults otherwise.
 * PARITY UNPINNED against real Ceres: the reference ships no tests/golden vectors for
 * this path.  Known, documented deviations:
 *   - max_solver_time_in_seconds is not applied (wall-clock termination makes the
 *     reference itself non-deterministic); only max_num_iterations stops the loop;
 *   - for bound-constrained problems (estimate_flag == 2 landmarks) Ceres runs a
 *     projected Armijo line search on the trust-region step; here the bound is
 *     enforced by the projection inside Plus() only (no step contraction);
 *   - unordered_map iteration order of marginalize() replaced by a canonical order
 *     (any order yields the same J0^T J0 / J0^T r0 up to rounding).
 */
#include <float.h>
#include <stdio.h>
#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#include "../include/vrf.h"

#define NF VRF_NUM_FRAMES
#define NC 172                       /* 11*6 poses + 11*9 speed-bias + 6 ex-pose + 1 td */
#define COL_POSE(f) (6 * (f))
#define COL_SB(f) (66 + 9 * (f))
#define COL_EX 165
#define COL_TD 171

/* ------------------------------------------------------------------ */
/* quaternion / rotation helpers (Eigen semantics, storage x,y,z,w)    */
/* ------------------------------------------------------------------ */
static void q_mul(const double a[4], const double b[4], double o[4])
{
    double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
static void q_inv(const double a[4], double o[4])
{   /* Eigen::Quaternion::inverse(): conjugate / squaredNorm */
    double n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    o[0] = -a[0] / n2; o[1] = -a[1] / n2; o[2] = -a[2] / n2; o[3] = a[3] / n2;
}
static void q_normalize(double a[4])
{
    double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
    a[0] /= n; a[1] /= n; a[2] /= n; a[3] /= n;
}
static void q_rot(const double q[4], const double v[3], double o[3])
{   /* Eigen: uv = 2 * q.vec x v ; v + w*uv + q.vec x uv */
    double ux = 2 * (q[1] * v[2] - q[2] * v[1]), uy = 2 * (q[2] * v[0] - q[0] * v[2]), uz = 2 * (q[0] * v[1] - q[1] * v[0]);
    o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
static void q_to_R(const double q[4], double R[9])
{   /* Eigen::Quaternion::toRotationMatrix */
    double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static void R_to_q(const double m[9], double q[4])
{   /* Eigen: Quaternion(Matrix3) */
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}
static void m3_mul(const double A[9], const double B[9], double C[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
static void m3_T(const double A[9], double T[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T[i * 3 + j] = A[j * 3 + i];
}
static void m3_v(const double A[9], const double v[3], double o[3])
{
    for (int i = 0; i < 3; i++) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
static void skew(const double v[3], double S[9])
{
    S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
/* Utility::Qleft / Qright bottom-right 3x3 corners (utility.h:47-64) */
static void qleft_br(const double q[4], double M[9])
{   /* w*I + skew(vec) */
    double S[9]; skew(q, S);
    for (int i = 0; i < 9; i++) M[i] = S[i];
    M[0] += q[3]; M[4] += q[3]; M[8] += q[3];
}
static void qleft4(const double q[4], double M[16])
{   /* row/col order (w, x, y, z) */
    M[0] = q[3]; M[1] = -q[0]; M[2] = -q[1]; M[3] = -q[2];
    double B[9]; qleft_br(q, B);
    for (int i = 0; i < 3; i++) { M[(i + 1) * 4] = q[i]; for (int j = 0; j < 3; j++) M[(i + 1) * 4 + j + 1] = B[i * 3 + j]; }
}
static void qright4(const double p[4], double M[16])
{
    M[0] = p[3]; M[1] = -p[0]; M[2] = -p[1]; M[3] = -p[2];
    double S[9]; skew(p, S);
    for (int i = 0; i < 3; i++) {
        M[(i + 1) * 4] = p[i];
        for (int j = 0; j < 3; j++) M[(i + 1) * 4 + j + 1] = (i == j ? p[3] : 0.0) - S[i * 3 + j];
    }
}

/* ------------------------------------------------------------------ */
/* ProjectionFactor::Evaluate (projection_factor.cpp:22-130)            */
/* J blocks row-major 2x7 (7th column zero), Jf 2x1.                    */
/* ------------------------------------------------------------------ */
void oracle_projection_eval(const double *pose_i, const double *pose_j, const double *ex, double inv_dep,
                            const double *pts_i2, const double *pts_j2, double *res, double *Ji, double *Jj,
                            double *Jex, double *Jf)
{
    const double sqrt_info = 460.0 / 1.5;          /* estimator.cpp:23 */
    const double *Pi = pose_i, *Qi = pose_i + 3, *Pj = pose_j, *Qj = pose_j + 3, *tic = ex, *qic = ex + 3;
    double pts_i[3] = {pts_i2[0], pts_i2[1], 1.0};
    double pc_i[3] = {pts_i[0] / inv_dep, pts_i[1] / inv_dep, pts_i[2] / inv_dep};
    double t[3], p_imu_i[3], pw[3], p_imu_j[3], pc_j[3], qinv[4];
    q_rot(qic, pc_i, t);
    for (int k = 0; k < 3; k++) p_imu_i[k] = t[k] + tic[k];
    q_rot(Qi, p_imu_i, t);
    for (int k = 0; k < 3; k++) pw[k] = t[k] + Pi[k];
    double d[3] = {pw[0] - Pj[0], pw[1] - Pj[1], pw[2] - Pj[2]};
    q_inv(Qj, qinv); q_rot(qinv, d, p_imu_j);
    double e[3] = {p_imu_j[0] - tic[0], p_imu_j[1] - tic[1], p_imu_j[2] - tic[2]};
    q_inv(qic, qinv); q_rot(qinv, e, pc_j);
    double dep_j = pc_j[2];
    res[0] = sqrt_info * (pc_j[0] / dep_j - pts_j2[0]);
    res[1] = sqrt_info * (pc_j[1] / dep_j - pts_j2[1]);
    if (!Ji && !Jj && !Jex && !Jf) return;
    double Ri[9], Rj[9], ric[9], RjT[9], ricT[9];
    q_to_R(Qi, Ri); q_to_R(Qj, Rj); q_to_R(qic, ric);
    m3_T(Rj, RjT); m3_T(ric, ricT);
    double red[6] = {sqrt_info * (1. / dep_j), 0, sqrt_info * (-pc_j[0] / (dep_j * dep_j)),
                     0, sqrt_info * (1. / dep_j), sqrt_info * (-pc_j[1] / (dep_j * dep_j))};
    double A[9], B[9], C[9], S[9];
    m3_mul(ricT, RjT, A);                          /* ric^T Rj^T */
    if (Ji) {
        double J36[18];
        m3_mul(A, Ri, B);                          /* ric^T Rj^T Ri */
        skew(p_imu_i, S);
        m3_mul(B, S, C);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) { J36[r * 6 + c] = A[r * 3 + c]; J36[r * 6 + 3 + c] = -C[r * 3 + c]; }
        for (int r = 0; r < 2; r++) {
            for (int c = 0; c < 6; c++) Ji[r * 7 + c] = red[r * 3] * J36[c] + red[r * 3 + 1] * J36[6 + c] + red[r * 3 + 2] * J36[12 + c];
            Ji[r * 7 + 6] = 0;
        }
    }
    if (Jj) {
        double J36[18];
        skew(p_imu_j, S);
        m3_mul(ricT, S, C);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) { J36[r * 6 + c] = -A[r * 3 + c]; J36[r * 6 + 3 + c] = C[r * 3 + c]; }
        for (int r = 0; r < 2; r++) {
            for (int c = 0; c < 6; c++) Jj[r * 7 + c] = red[r * 3] * J36[c] + red[r * 3 + 1] * J36[6 + c] + red[r * 3 + 2] * J36[12 + c];
            Jj[r * 7 + 6] = 0;
        }
    }
    if (Jex) {
        double J36[18], RjTRi[9], M[9], tmp_r[9], v[3], w[3], S1[9], S2[9], T1[9];
        m3_mul(RjT, Ri, RjTRi);
        for (int k = 0; k < 9; k++) M[k] = RjTRi[k] - ((k % 4 == 0) ? 1.0 : 0.0);
        m3_mul(ricT, M, B);                        /* left cols */
        m3_mul(A, Ri, C); m3_mul(C, ric, tmp_r);   /* ric^T Rj^T Ri ric */
        skew(pc_i, S); m3_mul(tmp_r, S, T1);
        m3_v(tmp_r, pc_i, v); skew(v, S1);
        double u[3];
        m3_v(Ri, tic, u);
        for (int k = 0; k < 3; k++) u[k] += Pi[k] - Pj[k];
        m3_v(RjT, u, w);
        for (int k = 0; k < 3; k++) w[k] -= tic[k];
        m3_v(ricT, w, v); skew(v, S2);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                J36[r * 6 + c] = B[r * 3 + c];
                J36[r * 6 + 3 + c] = -T1[r * 3 + c] + S1[r * 3 + c] + S2[r * 3 + c];
            }
        for (int r = 0; r < 2; r++) {
            for (int c = 0; c < 6; c++) Jex[r * 7 + c] = red[r * 3] * J36[c] + red[r * 3 + 1] * J36[6 + c] + red[r * 3 + 2] * J36[12 + c];
            Jex[r * 7 + 6] = 0;
        }
    }
    if (Jf) {
        double tmp[9], tr[9], v[3];
        m3_mul(A, Ri, tmp); m3_mul(tmp, ric, tr);
        m3_v(tr, pts_i, v);
        for (int r = 0; r < 2; r++)
            Jf[r] = (red[r * 3] * v[0] + red[r * 3 + 1] * v[1] + red[r * 3 + 2] * v[2]) * -1.0 / (inv_dep * inv_dep);
    }
}

/* ------------------------------------------------------------------ */
/* ProjectionTdFactor::Evaluate (projection_td_factor.cpp:34-150): the  */
/* projection factor on the time-shifted points                         */
/*   pts_td = pts - (td - td_obs + TR / ROW * row) * velocity  (:52-53) */
/* plus the td column (:139-144).  vel = (vx, vy), z = 0.               */
/* ------------------------------------------------------------------ */
void oracle_projection_td_eval(const double *pose_i, const double *pose_j, const double *ex, double inv_dep, double td,
                               const double *pts_i2, const double *pts_j2, const double *vel_i, const double *vel_j,
                               double td_i, double td_j, double row_i, double row_j, double tr_over_row,
                               double *res, double *Ji, double *Jj, double *Jex, double *Jf, double *Jtd)
{
    const double sqrt_info = 460.0 / 1.5;
    const double si = td - td_i + tr_over_row * row_i, sj = td - td_j + tr_over_row * row_j;
    double pi_td[2] = {pts_i2[0] - si * vel_i[0], pts_i2[1] - si * vel_i[1]};
    double pj_td[2] = {pts_j2[0] - sj * vel_j[0], pts_j2[1] - sj * vel_j[1]};
    oracle_projection_eval(pose_i, pose_j, ex, inv_dep, pi_td, pj_td, res, Ji, Jj, Jex, Jf);
    if (!Jtd) return;
    /* jacobian_td = reduce * ric^T Rj^T Ri ric * velocity_i / inv_dep * -1 + sqrt_info * velocity_j.head(2) */
    const double *Pi = pose_i, *Qi = pose_i + 3, *Pj = pose_j, *Qj = pose_j + 3, *tic = ex, *qic = ex + 3;
    double pc_i[3] = {pi_td[0] / inv_dep, pi_td[1] / inv_dep, 1.0 / inv_dep};
    double t[3], p_imu_i[3], pw[3], p_imu_j[3], pc_j[3], qinv[4];
    q_rot(qic, pc_i, t);
    for (int k = 0; k < 3; k++) p_imu_i[k] = t[k] + tic[k];
    q_rot(Qi, p_imu_i, t);
    for (int k = 0; k < 3; k++) pw[k] = t[k] + Pi[k];
    double d[3] = {pw[0] - Pj[0], pw[1] - Pj[1], pw[2] - Pj[2]};
    q_inv(Qj, qinv); q_rot(qinv, d, p_imu_j);
    double e[3] = {p_imu_j[0] - tic[0], p_imu_j[1] - tic[1], p_imu_j[2] - tic[2]};
    q_inv(qic, qinv); q_rot(qinv, e, pc_j);
    double dep_j = pc_j[2];
    double red[6] = {sqrt_info * (1. / dep_j), 0, sqrt_info * (-pc_j[0] / (dep_j * dep_j)),
                     0, sqrt_info * (1. / dep_j), sqrt_info * (-pc_j[1] / (dep_j * dep_j))};
    double Ri[9], Rj[9], ric[9], RjT[9], ricT[9], A[9], C[9], tmp_r[9], v[3];
    q_to_R(Qi, Ri); q_to_R(Qj, Rj); q_to_R(qic, ric);
    m3_T(Rj, RjT); m3_T(ric, ricT);
    m3_mul(ricT, RjT, A); m3_mul(A, Ri, C); m3_mul(C, ric, tmp_r);
    double v3[3] = {vel_i[0], vel_i[1], 0.0};
    m3_v(tmp_r, v3, v);
    for (int r = 0; r < 2; r++)
        Jtd[r] = (red[r * 3] * v[0] + red[r * 3 + 1] * v[1] + red[r * 3 + 2] * v[2]) / inv_dep * -1.0 + sqrt_info * vel_j[r];
}

/* ------------------------------------------------------------------ */
/* dense helpers                                                        */
/* ------------------------------------------------------------------ */
static int chol_lower(double *A, int n, int lda)
{   /* in-place lower Cholesky; returns 0 ok */
    for (int j = 0; j < n; j++) {
        double s = A[j * lda + j];
        for (int k = 0; k < j; k++) s -= A[j * lda + k] * A[j * lda + k];
        if (!(s > 0.0)) return -1;
        double l = sqrt(s);
        A[j * lda + j] = l;
        for (int i = j + 1; i < n; i++) {
            double t = A[i * lda + j];
            for (int k = 0; k < j; k++) t -= A[i * lda + k] * A[j * lda + k];
            A[i * lda + j] = t / l;
        }
    }
    return 0;
}
static void chol_solve(const double *L, int n, int lda, double *b)
{
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i * lda + k] * b[k];
        b[i] = s / L[i * lda + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= L[k * lda + i] * b[k];
        b[i] = s / L[i * lda + i];
    }
}
/* inverse of an SPD matrix via Cholesky (n <= 15) */
static int spd_inverse15(const double *A, double *Ainv)
{
    double L[225];
    memcpy(L, A, sizeof(L));
    if (chol_lower(L, 15, 15)) return -1;
    for (int c = 0; c < 15; c++) {
        double e[15] = {0};
        e[c] = 1.0;
        chol_solve(L, 15, 15, e);
        for (int r = 0; r < 15; r++) Ainv[r * 15 + c] = e[r];
    }
    /* symmetrise (the exact inverse is symmetric) */
    for (int r = 0; r < 15; r++)
        for (int c = r + 1; c < 15; c++) { double v = 0.5 * (Ainv[r * 15 + c] + Ainv[c * 15 + r]); Ainv[r * 15 + c] = Ainv[c * 15 + r] = v; }
    return 0;
}
/* cyclic Jacobi eigen-decomposition of a symmetric matrix: A = V diag(w) V^T, V column eigenvectors */
static void sym_eig_jacobi(double *A, int n, double *V, double *w)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0, diag = 0;
        for (int i = 0; i < n; i++) {
            diag += A[i * n + i] * A[i * n + i];
            for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double app = A[p * n + p], aqq = A[q * n + q];
                double tau = (aqq - app) / (2.0 * apq);
                double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

/* ------------------------------------------------------------------ */
/* IntegrationBase (integration_base.h)                                 */
/* ------------------------------------------------------------------ */
typedef struct {
    double acc_0[3], gyr_0[3];
    VrfImuPreint s;           /* sum_dt, delta_p/q/v, linearized_ba/bg, jacobian, covariance */
    double noise[18];         /* diagonal of the 18x18 noise matrix */
} OraclePreint;

void oracle_preint_init(OraclePreint *p, const double acc0[3], const double gyr0[3], const double ba[3],
                        const double bg[3], double acc_n, double gyr_n, double acc_w, double gyr_w)
{
    memset(p, 0, sizeof(*p));
    for (int k = 0; k < 3; k++) { p->acc_0[k] = acc0[k]; p->gyr_0[k] = gyr0[k]; p->s.linearized_ba[k] = ba[k]; p->s.linearized_bg[k] = bg[k]; }
    p->s.delta_q[3] = 1.0;
    for (int i = 0; i < 15; i++) p->s.jacobian[i * 15 + i] = 1.0;
    for (int k = 0; k < 3; k++) {
        p->noise[k] = acc_n * acc_n; p->noise[3 + k] = gyr_n * gyr_n; p->noise[6 + k] = acc_n * acc_n;
        p->noise[9 + k] = gyr_n * gyr_n; p->noise[12 + k] = acc_w * acc_w; p->noise[15 + k] = gyr_w * gyr_w;
    }
}

/* IntegrationBase::propagate -> midPointIntegration (integration_base.h:56-162) */
void oracle_preint_propagate(OraclePreint *p, double dt, const double acc1[3], const double gyr1[3])
{
    VrfImuPreint *s = &p->s;
    double a0[3], a1[3], un_gyr[3], un_acc_0[3], un_acc_1[3], un_acc[3], rq[4], dq[4];
    for (int k = 0; k < 3; k++) { a0[k] = p->acc_0[k] - s->linearized_ba[k]; a1[k] = acc1[k] - s->linearized_ba[k]; }
    q_rot(s->delta_q, a0, un_acc_0);
    for (int k = 0; k < 3; k++) un_gyr[k] = 0.5 * (p->gyr_0[k] + gyr1[k]) - s->linearized_bg[k];
    dq[0] = un_gyr[0] * dt / 2; dq[1] = un_gyr[1] * dt / 2; dq[2] = un_gyr[2] * dt / 2; dq[3] = 1;
    q_mul(s->delta_q, dq, rq);
    q_rot(rq, a1, un_acc_1);
    double rp[3], rv[3];
    for (int k = 0; k < 3; k++) {
        un_acc[k] = 0.5 * (un_acc_0[k] + un_acc_1[k]);
        rp[k] = s->delta_p[k] + s->delta_v[k] * dt + 0.5 * un_acc[k] * dt * dt;
        rv[k] = s->delta_v[k] + un_acc[k] * dt;
    }
    /* jacobian / covariance propagation */
    double Rw[9], Ra0[9], Ra1[9], Rd[9], Rr[9];
    skew(un_gyr, Rw); skew(a0, Ra0); skew(a1, Ra1);
    q_to_R(s->delta_q, Rd); q_to_R(rq, Rr);
    double ImRw[9];
    for (int k = 0; k < 9; k++) ImRw[k] = ((k % 4 == 0) ? 1.0 : 0.0) - Rw[k] * dt;
    double F[225] = {0}, V[15 * 18] = {0};
    double T1[9], T2[9], T3[9];
    m3_mul(Rd, Ra0, T1);                       /* delta_q.R * R_a_0_x */
    m3_mul(Rr, Ra1, T2);                       /* result.R * R_a_1_x */
    m3_mul(T2, ImRw, T3);                      /* result.R * R_a_1_x * (I - R_w_x dt) */
#define FB(r, c) (F + (r) * 15 + (c))
#define VB(r, c) (V + (r) * 18 + (c))
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double I = (i == j);
            FB(0, 0)[i * 15 + j] = I;
            FB(0, 3)[i * 15 + j] = -0.25 * T1[i * 3 + j] * dt * dt + -0.25 * T3[i * 3 + j] * dt * dt;
            FB(0, 6)[i * 15 + j] = I * dt;
            FB(0, 9)[i * 15 + j] = -0.25 * (Rd[i * 3 + j] + Rr[i * 3 + j]) * dt * dt;
            FB(0, 12)[i * 15 + j] = -0.25 * T2[i * 3 + j] * dt * dt * -dt;
            FB(3, 3)[i * 15 + j] = ImRw[i * 3 + j];
            FB(3, 12)[i * 15 + j] = -1.0 * I * dt;
            FB(6, 3)[i * 15 + j] = -0.5 * T1[i * 3 + j] * dt + -0.5 * T3[i * 3 + j] * dt;
            FB(6, 6)[i * 15 + j] = I;
            FB(6, 9)[i * 15 + j] = -0.5 * (Rd[i * 3 + j] + Rr[i * 3 + j]) * dt;
            FB(6, 12)[i * 15 + j] = -0.5 * T2[i * 3 + j] * dt * -dt;
            FB(9, 9)[i * 15 + j] = I;
            FB(12, 12)[i * 15 + j] = I;
            VB(0, 0)[i * 18 + j] = 0.25 * Rd[i * 3 + j] * dt * dt;
            VB(0, 3)[i * 18 + j] = 0.25 * -T2[i * 3 + j] * dt * dt * 0.5 * dt;
            VB(0, 6)[i * 18 + j] = 0.25 * Rr[i * 3 + j] * dt * dt;
            VB(0, 9)[i * 18 + j] = VB(0, 3)[i * 18 + j];
            VB(3, 3)[i * 18 + j] = 0.5 * I * dt;
            VB(3, 9)[i * 18 + j] = 0.5 * I * dt;
            VB(6, 0)[i * 18 + j] = 0.5 * Rd[i * 3 + j] * dt;
            VB(6, 3)[i * 18 + j] = 0.5 * -T2[i * 3 + j] * dt * 0.5 * dt;
            VB(6, 6)[i * 18 + j] = 0.5 * Rr[i * 3 + j] * dt;
            VB(6, 9)[i * 18 + j] = VB(6, 3)[i * 18 + j];
            VB(9, 12)[i * 18 + j] = I * dt;
            VB(12, 15)[i * 18 + j] = I * dt;
        }
#undef FB
#undef VB
    double nj[225], FC[225], nc[225];
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 15; j++) {
            double a = 0, b = 0;
            for (int k = 0; k < 15; k++) { a += F[i * 15 + k] * s->jacobian[k * 15 + j]; b += F[i * 15 + k] * s->covariance[k * 15 + j]; }
            nj[i * 15 + j] = a; FC[i * 15 + j] = b;
        }
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 15; j++) {
            double a = 0;
            for (int k = 0; k < 15; k++) a += FC[i * 15 + k] * F[j * 15 + k];
            for (int k = 0; k < 18; k++) a += V[i * 18 + k] * p->noise[k] * V[j * 18 + k];
            nc[i * 15 + j] = a;
        }
    memcpy(s->jacobian, nj, sizeof(nj));
    memcpy(s->covariance, nc, sizeof(nc));
    for (int k = 0; k < 3; k++) { s->delta_p[k] = rp[k]; s->delta_v[k] = rv[k]; }
    memcpy(s->delta_q, rq, sizeof(rq));
    q_normalize(s->delta_q);
    s->sum_dt += dt;
    for (int k = 0; k < 3; k++) { p->acc_0[k] = acc1[k]; p->gyr_0[k] = gyr1[k]; }
}

/* ------------------------------------------------------------------ */
/* IMUFactor::Evaluate (imu_factor.h:20-205); J row-major 15x7 / 15x9    */
/* ------------------------------------------------------------------ */
#define O_P 0
#define O_R 3
#define O_V 6
#define O_BA 9
#define O_BG 12
int oracle_imu_eval(const VrfImuPreint *pre, const double *pose_i, const double *sb_i, const double *pose_j,
                    const double *sb_j, double g_norm, double *res, double *Jpi, double *Jsi, double *Jpj, double *Jsj)
{
    const double *Pi = pose_i, *Qi = pose_i + 3, *Vi = sb_i, *Bai = sb_i + 3, *Bgi = sb_i + 6;
    const double *Pj = pose_j, *Qj = pose_j + 3, *Vj = sb_j, *Baj = sb_j + 3, *Bgj = sb_j + 6;
    const double G[3] = {0, 0, g_norm};
    const double sum_dt = pre->sum_dt;
    double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            dp_dba[i * 3 + j] = pre->jacobian[(O_P + i) * 15 + O_BA + j];
            dp_dbg[i * 3 + j] = pre->jacobian[(O_P + i) * 15 + O_BG + j];
            dq_dbg[i * 3 + j] = pre->jacobian[(O_R + i) * 15 + O_BG + j];
            dv_dba[i * 3 + j] = pre->jacobian[(O_V + i) * 15 + O_BA + j];
            dv_dbg[i * 3 + j] = pre->jacobian[(O_V + i) * 15 + O_BG + j];
        }
    double dba[3], dbg[3];
    for (int k = 0; k < 3; k++) { dba[k] = Bai[k] - pre->linearized_ba[k]; dbg[k] = Bgi[k] - pre->linearized_bg[k]; }
    /* IntegrationBase::evaluate */
    double th[3], dq[4], cdq[4], cdv[3], cdp[3], t1[3], t2[3];
    m3_v(dq_dbg, dbg, th);
    dq[0] = th[0] / 2; dq[1] = th[1] / 2; dq[2] = th[2] / 2; dq[3] = 1;      /* Utility::deltaQ */
    q_mul(pre->delta_q, dq, cdq);
    m3_v(dv_dba, dba, t1); m3_v(dv_dbg, dbg, t2);
    for (int k = 0; k < 3; k++) cdv[k] = pre->delta_v[k] + t1[k] + t2[k];
    m3_v(dp_dba, dba, t1); m3_v(dp_dbg, dbg, t2);
    for (int k = 0; k < 3; k++) cdp[k] = pre->delta_p[k] + t1[k] + t2[k];
    double Qi_inv[4], v[3], w[3], r[15];
    q_inv(Qi, Qi_inv);
    for (int k = 0; k < 3; k++) v[k] = 0.5 * G[k] * sum_dt * sum_dt + Pj[k] - Pi[k] - Vi[k] * sum_dt;
    q_rot(Qi_inv, v, w);
    for (int k = 0; k < 3; k++) r[O_P + k] = w[k] - cdp[k];
    double cdq_inv[4], qij[4], qe[4];
    q_inv(cdq, cdq_inv); q_mul(Qi_inv, Qj, qij); q_mul(cdq_inv, qij, qe);
    for (int k = 0; k < 3; k++) r[O_R + k] = 2 * qe[k];
    for (int k = 0; k < 3; k++) v[k] = G[k] * sum_dt + Vj[k] - Vi[k];
    q_rot(Qi_inv, v, w);
    for (int k = 0; k < 3; k++) { r[O_V + k] = w[k] - cdv[k]; r[O_BA + k] = Baj[k] - Bai[k]; r[O_BG + k] = Bgj[k] - Bgi[k]; }
    /* sqrt_info = LLT(covariance^-1).matrixL().transpose() */
    double cinv[225], L[225], S[225];
    if (spd_inverse15(pre->covariance, cinv)) return -1;
    memcpy(L, cinv, sizeof(L));
    if (chol_lower(L, 15, 15)) return -1;
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 15; j++) S[i * 15 + j] = (j >= i) ? L[j * 15 + i] : 0.0;   /* upper = L^T */
    for (int i = 0; i < 15; i++) {
        double a = 0;
        for (int k = i; k < 15; k++) a += S[i * 15 + k] * r[k];
        res[i] = a;
    }
    if (!Jpi && !Jsi && !Jpj && !Jsj) return 0;
    double RiT[9], Rtmp[9];
    q_to_R(Qi_inv, RiT);                   /* Qi.inverse().toRotationMatrix() */
    double J[15 * 9];
#define JSET(J, cols, r0, c0, M, sgn) for (int i_ = 0; i_ < 3; i_++) for (int j_ = 0; j_ < 3; j_++) (J)[((r0) + i_) * (cols) + (c0) + j_] = (sgn) * (M)[i_ * 3 + j_]
    if (Jpi) {
        memset(J, 0, sizeof(double) * 15 * 7);
        JSET(J, 7, O_P, O_P, RiT, -1.0);
        for (int k = 0; k < 3; k++) v[k] = 0.5 * G[k] * sum_dt * sum_dt + Pj[k] - Pi[k] - Vi[k] * sum_dt;
        q_rot(Qi_inv, v, w); skew(w, Rtmp);
        JSET(J, 7, O_P, O_R, Rtmp, 1.0);
        double Qj_inv[4], qji[4], QL[16], QR[16], M4[16];
        q_inv(Qj, Qj_inv); q_mul(Qj_inv, Qi, qji);
        qleft4(qji, QL); qright4(cdq, QR);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) { double a = 0; for (int k = 0; k < 4; k++) a += QL[i * 4 + k] * QR[k * 4 + j]; M4[i * 4 + j] = a; }
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rtmp[i * 3 + j] = M4[(i + 1) * 4 + j + 1];
        JSET(J, 7, O_R, O_R, Rtmp, -1.0);
        for (int k = 0; k < 3; k++) v[k] = G[k] * sum_dt + Vj[k] - Vi[k];
        q_rot(Qi_inv, v, w); skew(w, Rtmp);
        JSET(J, 7, O_V, O_R, Rtmp, 1.0);
        for (int i = 0; i < 15; i++)
            for (int j = 0; j < 7; j++) { double a = 0; for (int k = i; k < 15; k++) a += S[i * 15 + k] * J[k * 7 + j]; Jpi[i * 7 + j] = a; }
    }
    if (Jsi) {
        memset(J, 0, sizeof(double) * 15 * 9);
        for (int k = 0; k < 9; k++) Rtmp[k] = RiT[k] * sum_dt;
        JSET(J, 9, O_P, O_V - O_V, Rtmp, -1.0);
        JSET(J, 9, O_P, O_BA - O_V, dp_dba, -1.0);
        JSET(J, 9, O_P, O_BG - O_V, dp_dbg, -1.0);
        double Qj_inv[4], q1[4], q2[4], B[9], C[9];
        q_inv(Qj, Qj_inv); q_mul(Qj_inv, Qi, q1); q_mul(q1, pre->delta_q, q2);
        qleft_br(q2, B); m3_mul(B, dq_dbg, C);
        JSET(J, 9, O_R, O_BG - O_V, C, -1.0);
        JSET(J, 9, O_V, O_V - O_V, RiT, -1.0);
        JSET(J, 9, O_V, O_BA - O_V, dv_dba, -1.0);
        JSET(J, 9, O_V, O_BG - O_V, dv_dbg, -1.0);
        for (int k = 0; k < 3; k++) { J[(O_BA + k) * 9 + O_BA - O_V + k] = -1.0; J[(O_BG + k) * 9 + O_BG - O_V + k] = -1.0; }
        for (int i = 0; i < 15; i++)
            for (int j = 0; j < 9; j++) { double a = 0; for (int k = i; k < 15; k++) a += S[i * 15 + k] * J[k * 9 + j]; Jsi[i * 9 + j] = a; }
    }
    if (Jpj) {
        memset(J, 0, sizeof(double) * 15 * 7);
        JSET(J, 7, O_P, O_P, RiT, 1.0);
        double q1[4], q2[4], B[9];
        q_mul(cdq_inv, Qi_inv, q1); q_mul(q1, Qj, q2);
        qleft_br(q2, B);
        JSET(J, 7, O_R, O_R, B, 1.0);
        for (int i = 0; i < 15; i++)
            for (int j = 0; j < 7; j++) { double a = 0; for (int k = i; k < 15; k++) a += S[i * 15 + k] * J[k * 7 + j]; Jpj[i * 7 + j] = a; }
    }
    if (Jsj) {
        memset(J, 0, sizeof(double) * 15 * 9);
        JSET(J, 9, O_V, O_V - O_V, RiT, 1.0);
        for (int k = 0; k < 3; k++) { J[(O_BA + k) * 9 + O_BA - O_V + k] = 1.0; J[(O_BG + k) * 9 + O_BG - O_V + k] = 1.0; }
        for (int i = 0; i < 15; i++)
            for (int j = 0; j < 9; j++) { double a = 0; for (int k = i; k < 15; k++) a += S[i * 15 + k] * J[k * 9 + j]; Jsj[i * 9 + j] = a; }
    }
#undef JSET
    return 0;
}

/* ------------------------------------------------------------------ */
/* problem state / linearisation                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    double pose[NF][7], sb[NF][9], ex[7], td;
    double *lam;              /* [M] */
} State;

typedef struct {
    int M, O, nimu, np;       /* np = prior n (0 if none) */
    /* per non-host observation (factor) */
    int nfac;
    int *f_lm, *f_i, *f_j;
    double *f_pi, *f_pj;      /* [nfac][2] */
    double *f_vi, *f_vj;      /* [nfac][2] velocities (td factor) */
    double *f_tdi, *f_tdj, *f_rowi, *f_rowj;    /* [nfac] cur_td and pixel row of both observations (td factor) */
    double *f_r, *f_Ji, *f_Jj, *f_Jex, *f_Jl, *f_Jtd;   /* 2, 12, 12, 12, 2, 2 per factor (local, corrected) */
    /* imu */
    int imu_j[NF];            /* frame j of each used IMU factor */
    const VrfImuPreint *imu_pre[NF];
    double imu_r[NF][15], imu_J[NF][15 * 30];
    /* prior */
    double *pr_r;             /* [np] */
    int pr_col[VRF_PRIOR_MAX_BLOCKS];   /* tangent column of each kept block (-1: landmark not allowed) */
    /* flags */
    int ex_active, pose0_const, use_imu, nframes;
    int td_factor, td_active; /* ProjectionTdFactor in use (ESTIMATE_TD) / para_Td variable */
    double tr_over_row;       /* TR / ROW */
    unsigned char *lm_const;
    double *lm_ub;            /* upper bound or +inf */
} Lin;

static void pose_plus(const double *x, const double *d, double *o)
{   /* PoseLocalParameterization::Plus */
    double dq[4] = {d[3] / 2, d[4] / 2, d[5] / 2, 1.0}, q[4];
    o[0] = x[0] + d[0]; o[1] = x[1] + d[1]; o[2] = x[2] + d[2];
    q_mul(x + 3, dq, q); q_normalize(q);
    o[3] = q[0]; o[4] = q[1]; o[5] = q[2]; o[6] = q[3];
}

/* CauchyLoss(1.0) + Corrector: returns rho0, scales r and the given Jacobian blocks */
static double cauchy_correct(double *r, double *J[], const int *ncols, int nblk)
{
    double s = r[0] * r[0] + r[1] * r[1];
    double sum = 1.0 + s, inv = 1.0 / sum;
    double rho0 = log(sum), rho1 = inv > DBL_MIN ? inv : DBL_MIN, rho2 = -(inv * inv);
    double sqrt_rho1 = sqrt(rho1), residual_scaling, alpha_sq_norm;
    if (s == 0.0 || rho2 <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
    else {
        double D = 1.0 + 2.0 * s * rho2 / rho1;
        double alpha = 1.0 - sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / s;
    }
    for (int b = 0; b < nblk; b++) {
        if (!J[b]) continue;
        int nc = ncols[b];
        for (int c = 0; c < nc; c++) {
            double j0 = J[b][c], j1 = J[b][nc + c];
            double rtj = r[0] * j0 + r[1] * j1;
            J[b][c] = sqrt_rho1 * (j0 - alpha_sq_norm * r[0] * rtj);
            J[b][nc + c] = sqrt_rho1 * (j1 - alpha_sq_norm * r[1] * rtj);
        }
    }
    r[0] *= residual_scaling; r[1] *= residual_scaling;
    return rho0;
}

/* MarginalizationFactor::Evaluate residual (marginalization_factor.cpp:353-400) */
static void prior_residual(const VrfPrior *P, const State *x, double *r)
{
    int n = P->n;
    double dx[VRF_PRIOR_MAX_DIM];
    for (int b = 0; b < P->n_blocks; b++) {
        const VrfPriorBlock *B = &P->blocks[b];
        const double *cur = B->kind == VRF_BLK_POSE ? x->pose[B->index] : B->kind == VRF_BLK_SPEEDBIAS ? x->sb[B->index]
                            : B->kind == VRF_BLK_EXPOSE ? x->ex : &x->td;
        if (B->size != 7) {
            for (int k = 0; k < B->size; k++) dx[B->idx + k] = cur[k] - B->x0[k];
        } else {
            for (int k = 0; k < 3; k++) dx[B->idx + k] = cur[k] - B->x0[k];
            double q0i[4], qe[4];
            q_inv(B->x0 + 3, q0i); q_mul(q0i, cur + 3, qe);
            double sgn = (qe[3] >= 0) ? 1.0 : -1.0;
            for (int k = 0; k < 3; k++) dx[B->idx + 3 + k] = sgn * 2.0 * qe[k];
        }
    }
    for (int i = 0; i < n; i++) {
        double a = P->linearized_residuals[i];
        const double *row = P->linearized_jacobians + (size_t)i * n;
        for (int k = 0; k < n; k++) a += row[k] * dx[k];
        r[i] = a;
    }
}

/* evaluate all residual blocks at x; with_jac fills the linearisation. returns cost */
static double evaluate(const VrfBaProblem *pb, const VrfConfig *cfg, Lin *L, const State *x, int with_jac)
{
    double cost = 0;
    for (int k = 0; k < L->nfac; k++) {
        int i = L->f_i[k], j = L->f_j[k], l = L->f_lm[k];
        double r[2], Ji[14], Jj[14], Jex[14], Jf[2], Jt[2] = {0, 0};
        if (L->td_factor)
            oracle_projection_td_eval(x->pose[i], x->pose[j], x->ex, x->lam[l], x->td, L->f_pi + 2 * k, L->f_pj + 2 * k,
                                      L->f_vi + 2 * k, L->f_vj + 2 * k, L->f_tdi[k], L->f_tdj[k], L->f_rowi[k], L->f_rowj[k],
                                      L->tr_over_row, r, with_jac ? Ji : NULL, with_jac ? Jj : NULL,
                                      (with_jac && L->ex_active) ? Jex : NULL, (with_jac && !L->lm_const[l]) ? Jf : NULL,
                                      (with_jac && L->td_active) ? Jt : NULL);
        else
            oracle_projection_eval(x->pose[i], x->pose[j], x->ex, x->lam[l], L->f_pi + 2 * k, L->f_pj + 2 * k, r,
                                   with_jac ? Ji : NULL, with_jac ? Jj : NULL, (with_jac && L->ex_active) ? Jex : NULL,
                                   (with_jac && !L->lm_const[l]) ? Jf : NULL);
        if (with_jac) {
            /* local parameterisation: first 6 columns of the 2x7 blocks */
            double li[12], lj[12], le[12], lf[2] = {0, 0};
            /* a constant parameter block is not part of the Ceres program: its Jacobian columns do not exist (para_Pose[0] in
             * VO mode, estimator.cpp:1182-1185; para_Ex_Pose / para_Td when held constant) */
            const int host_const = (i == 0 && L->pose0_const);
            for (int rr = 0; rr < 2; rr++)
                for (int c = 0; c < 6; c++) {
                    li[rr * 6 + c] = host_const ? 0.0 : Ji[rr * 7 + c]; lj[rr * 6 + c] = Jj[rr * 7 + c];
                    le[rr * 6 + c] = L->ex_active ? Jex[rr * 7 + c] : 0.0;
                }
            if (!L->lm_const[l]) { lf[0] = Jf[0]; lf[1] = Jf[1]; }
            double *Jb[5] = {li, lj, le, lf, Jt};
            int ncl[5] = {6, 6, 6, 1, 1};
            cost += 0.5 * cauchy_correct(r, Jb, ncl, 5);
            L->f_Jtd[2 * k] = Jt[0]; L->f_Jtd[2 * k + 1] = Jt[1];
            memcpy(L->f_Ji + 12 * k, li, sizeof(li)); memcpy(L->f_Jj + 12 * k, lj, sizeof(lj));
            memcpy(L->f_Jex + 12 * k, le, sizeof(le)); L->f_Jl[2 * k] = lf[0]; L->f_Jl[2 * k + 1] = lf[1];
            L->f_r[2 * k] = r[0]; L->f_r[2 * k + 1] = r[1];
        } else {
            double s = r[0] * r[0] + r[1] * r[1];
            cost += 0.5 * log(1.0 + s);
        }
    }
    for (int f = 0; f < L->nimu; f++) {
        int j = L->imu_j[f], i = j - 1;
        double r[15], Jpi[105], Jsi[135], Jpj[105], Jsj[135];
        oracle_imu_eval(L->imu_pre[f], x->pose[i], x->sb[i], x->pose[j], x->sb[j], cfg->g_norm, r,
                        with_jac ? Jpi : NULL, with_jac ? Jsi : NULL, with_jac ? Jpj : NULL, with_jac ? Jsj : NULL);
        double s = 0;
        for (int k = 0; k < 15; k++) s += r[k] * r[k];
        cost += 0.5 * s;
        if (with_jac) {
            memcpy(L->imu_r[f], r, sizeof(r));
            for (int rr = 0; rr < 15; rr++) {
                double *row = L->imu_J[f] + rr * 30;
                for (int c = 0; c < 6; c++) { row[c] = Jpi[rr * 7 + c]; row[15 + c] = Jpj[rr * 7 + c]; }
                for (int c = 0; c < 9; c++) { row[6 + c] = Jsi[rr * 9 + c]; row[21 + c] = Jsj[rr * 9 + c]; }
            }
        }
    }
    if (L->np) {
        double r[VRF_PRIOR_MAX_DIM];
        prior_residual(pb->prior, x, r);
        double s = 0;
        for (int k = 0; k < L->np; k++) s += r[k] * r[k];
        cost += 0.5 * s;
        if (with_jac) memcpy(L->pr_r, r, sizeof(double) * L->np);
    }
    return cost;
}

/* y += Jc * v_c (+ landmark part), for all residual rows; used for J*step products.
   v has NC + M entries (tangent), already includes any column scaling. */
typedef void (*row_fn)(void *ctx, const double *Jrow_vals, const int *cols, int ncols, double r, int row_id);

/* iterate every residual row with its sparse Jacobian (unscaled) */
static void for_each_row(const VrfBaProblem *pb, const Lin *L, row_fn fn, void *ctx)
{
    int rid = 0;
    for (int k = 0; k < L->nfac; k++) {
        int i = L->f_i[k], j = L->f_j[k], l = L->f_lm[k];
        for (int rr = 0; rr < 2; rr++) {
            double vals[20]; int cols[20]; int n = 0;
            for (int c = 0; c < 6; c++) { vals[n] = L->f_Ji[12 * k + rr * 6 + c]; cols[n++] = COL_POSE(i) + c; }
            for (int c = 0; c < 6; c++) { vals[n] = L->f_Jj[12 * k + rr * 6 + c]; cols[n++] = COL_POSE(j) + c; }
            if (L->ex_active) for (int c = 0; c < 6; c++) { vals[n] = L->f_Jex[12 * k + rr * 6 + c]; cols[n++] = COL_EX + c; }
            if (L->td_active) { vals[n] = L->f_Jtd[2 * k + rr]; cols[n++] = COL_TD; }
            if (!L->lm_const[l]) { vals[n] = L->f_Jl[2 * k + rr]; cols[n++] = NC + l; }
            fn(ctx, vals, cols, n, L->f_r[2 * k + rr], rid++);
        }
    }
    for (int f = 0; f < L->nimu; f++) {
        int j = L->imu_j[f], i = j - 1;
        for (int rr = 0; rr < 15; rr++) {
            double vals[30]; int cols[30]; int n = 0;
            const double *row = L->imu_J[f] + rr * 30;
            for (int c = 0; c < 6; c++) { vals[n] = row[c]; cols[n++] = COL_POSE(i) + c; }
            for (int c = 0; c < 9; c++) { vals[n] = row[6 + c]; cols[n++] = COL_SB(i) + c; }
            for (int c = 0; c < 6; c++) { vals[n] = row[15 + c]; cols[n++] = COL_POSE(j) + c; }
            for (int c = 0; c < 9; c++) { vals[n] = row[21 + c]; cols[n++] = COL_SB(j) + c; }
            fn(ctx, vals, cols, n, L->imu_r[f][rr], rid++);
        }
    }
    if (L->np) {
        const VrfPrior *P = pb->prior;
        int n = P->n;
        static __thread double vals[VRF_PRIOR_MAX_DIM];
        static __thread int cols[VRF_PRIOR_MAX_DIM];
        for (int rr = 0; rr < n; rr++) {
            int m = 0;
            for (int b = 0; b < P->n_blocks; b++) {
                const VrfPriorBlock *B = &P->blocks[b];
                int ls = B->size == 7 ? 6 : B->size;
                int col = L->pr_col[b];
                if (col < 0) continue;          /* constant block: column dropped */
                for (int c = 0; c < ls; c++) { vals[m] = P->linearized_jacobians[(size_t)rr * n + B->idx + c]; cols[m++] = col + c; }
            }
            fn(ctx, vals, cols, m, L->pr_r[rr], rid++);
        }
    }
}

static int col_active(const Lin *L, int col)
{
    if (col < 66) { int f = col / 6; if (f >= L->nframes) return 0; if (f == 0 && L->pose0_const) return 0; return 1; }
    if (col < 165) { int f = (col - 66) / 9; return L->use_imu && f < L->nframes; }
    if (col == COL_TD) return L->td_active;
    return L->ex_active;
}

typedef struct { const double *scale; double *colsq; } CtxColNorm;
static void fn_colnorm(void *c_, const double *v, const int *cols, int n, double r, int rid)
{
    CtxColNorm *c = (CtxColNorm *)c_;
    (void)r; (void)rid;
    for (int k = 0; k < n; k++) { double s = c->scale ? c->scale[cols[k]] : 1.0; c->colsq[cols[k]] += v[k] * v[k] * s * s; }
}
typedef struct { const double *scale; const double *x; double *y; } CtxMul;   /* y[row] = J_s x */
static void fn_mul(void *c_, const double *v, const int *cols, int n, double r, int rid)
{
    CtxMul *c = (CtxMul *)c_;
    (void)r;
    double a = 0;
    for (int k = 0; k < n; k++) a += v[k] * c->scale[cols[k]] * c->x[cols[k]];
    c->y[rid] = a;
}
static void fn_grad(void *c_, const double *v, const int *cols, int n, double r, int rid)     /* g += J^T r (unscaled) */
{
    double *g = (double *)c_;
    (void)rid;
    for (int k = 0; k < n; k++) g[cols[k]] += v[k] * r;
}
typedef struct { const double *scale; double *g; double *H; double *hll; double *W; int M; } CtxNormal;
static void fn_normal(void *c_, const double *v, const int *cols, int n, double r, int rid)
{
    CtxNormal *c = (CtxNormal *)c_;
    (void)rid;
    double sv[40]; int lm = -1; double lv = 0;
    int nc = 0, cc[40];
    for (int k = 0; k < n; k++) {
        double s = v[k] * c->scale[cols[k]];
        c->g[cols[k]] += s * r;
        if (cols[k] >= NC) { lm = cols[k] - NC; lv = s; }
        else { sv[nc] = s; cc[nc++] = cols[k]; }
    }
    if (n <= 40) {
        for (int a = 0; a < nc; a++)
            for (int b = 0; b < nc; b++) c->H[cc[a] * NC + cc[b]] += sv[a] * sv[b];
        if (lm >= 0) {
            c->hll[lm] += lv * lv;
            for (int a = 0; a < nc; a++) c->W[(size_t)lm * NC + cc[a]] += lv * sv[a];
        }
    }
}
/* the prior rows can be longer than 40 columns: dedicated accumulation */
static void normal_prior(const VrfBaProblem *pb, const Lin *L, const double *scale, double *g, double *H)
{
    const VrfPrior *P = pb->prior;
    int n = P->n;
    int colmap[VRF_PRIOR_MAX_DIM];
    for (int k = 0; k < n; k++) colmap[k] = -1;
    for (int b = 0; b < P->n_blocks; b++) {
        const VrfPriorBlock *B = &P->blocks[b];
        int ls = B->size == 7 ? 6 : B->size;
        if (L->pr_col[b] < 0) continue;
        for (int c = 0; c < ls; c++) colmap[B->idx + c] = L->pr_col[b] + c;
    }
    for (int a = 0; a < n; a++) {
        if (colmap[a] < 0) continue;
        double ga = 0;
        for (int rr = 0; rr < n; rr++) ga += P->linearized_jacobians[(size_t)rr * n + a] * L->pr_r[rr];
        g[colmap[a]] += ga * scale[colmap[a]];
        for (int b = 0; b < n; b++) {
            if (colmap[b] < 0) continue;
            double h = 0;
            for (int rr = 0; rr < n; rr++) h += P->linearized_jacobians[(size_t)rr * n + a] * P->linearized_jacobians[(size_t)rr * n + b];
            H[colmap[a] * NC + colmap[b]] += h * scale[colmap[a]] * scale[colmap[b]];
        }
    }
}

static void state_plus(const VrfBaProblem *pb, const Lin *L, const State *x, const double *delta, State *o)
{
    (void)pb;
    *o = *x;            /* shares lam pointer: caller provides separate storage */
}


/* ------------------------------------------------------------------ */
/* Ceres' projected Armijo line search of bound-constrained problems     */
/* (third party, restated from the published algorithm: Ceres Solver      */
/*  internal/ceres/trust_region_minimizer.cc DoLineSearch,                */
/*  line_search.cc ArmijoLineSearch::DoSearch +                           */
/*  InterpolatingPolynomialMinimizingStepSize, polynomial.cc).            */
/* Solver::Options defaults, which estimator.cpp:1348-1363 leaves alone:  */
/*   line_search_interpolation_type CUBIC, sufficient decrease 1e-4,      */
/*   max_line_search_step_contraction 1e-3, min_..._contraction 0.6,      */
/*   max_num_line_search_step_size_iterations 20, min step size 1e-9.     */
/* PARITY UNPINNED against Ceres itself (not installed here).             */
/* ------------------------------------------------------------------ */
typedef struct { double x, value, gradient; int value_ok, grad_ok; } LsSample;

static double lsq_poly_eval(const double *c, int n, double x)      /* n coefficients, highest power first (Horner) */
{
    double v = 0;
    for (int i = 0; i < n; i++) v = v * x + c[i];
    return v;
}

/* value and derivative of a polynomial (n coefficients, highest power first) */
static void lsq_poly_eval2(const double *c, int n, double x, double *f, double *df)
{
    double v = c[0], d = 0.0;
    for (int i = 1; i < n; ++i) { d = d * x + v; v = v * x + c[i]; }
    *f = v; *df = d;
}

/* root of c inside the bracket [u, v] (sign change, fu = c(u)): Newton steps, bisection whenever Newton leaves the bracket or */
/* stops halving the step (Numerical Recipes' rtsafe), down to a few ulp */
static double lsq_root_bracketed(const double *c, int n, double u, double v, double fu)
{
    double xl = fu < 0 ? u : v, xh = fu < 0 ? v : u;          /* c(xl) < 0 < c(xh) */
    double x = 0.5 * (u + v), dxold = fabs(v - u), dx = dxold, f, df;
    lsq_poly_eval2(c, n, x, &f, &df);
    for (int it = 0; it < 200; ++it) {
        if (((x - xh) * df - f) * ((x - xl) * df - f) > 0.0 || fabs(2.0 * f) > fabs(dxold * df)) {
            dxold = dx; dx = 0.5 * (xh - xl);
            const double xn = xl + dx;
            if (xn == xl) return xn;
            x = xn;
        } else {
            dxold = dx; dx = f / df;
            const double xn = x - dx;
            if (xn == x) return x;
            x = xn;
        }
        if (fabs(dx) <= 4.4e-16 * fabs(x)) return x;
        lsq_poly_eval2(c, n, x, &f, &df);
        if (f == 0.0) return x;
        if (f < 0) xl = x; else xh = x;
    }
    return x;
}

/* real roots of c (degree <= 2 after stripping leading zeros) inside [a, b], ascending */
static int lsq_roots_low(const double *c, int n, double a, double b, double *roots)
{
    while (n > 0 && c[0] == 0.0) { ++c; --n; }
    const int deg = n - 1;
    int nr = 0;
    if (deg < 1) return 0;
    if (deg == 1) { const double r = -c[1] / c[0]; if (r >= a && r <= b) roots[nr++] = r; return nr; }
    const double D_ = c[1] * c[1] - 4 * c[0] * c[2];
    if (D_ < 0) return 0;
    const double sD = sqrt(D_);
    double r0, r1;
    if (c[1] >= 0) { r0 = (-c[1] - sD) / (2.0 * c[0]); r1 = (2.0 * c[2]) / (-c[1] - sD); }
    else { r0 = (2.0 * c[2]) / (-c[1] + sD); r1 = (-c[1] + sD) / (2.0 * c[0]); }
    if (r0 > r1) { const double t = r0; r0 = r1; r1 = t; }
    if (r0 >= a && r0 <= b) roots[nr++] = r0;
    if (r1 >= a && r1 <= b && r1 != r0) roots[nr++] = r1;
    return nr;
}

/* real roots inside [a, b] of c given the real roots `crit` of its derivative: they cut [a, b] into monotone pieces */
static int lsq_roots_from_crit(const double *c, int n, double a, double b, const double *crit, int ncrit, double *roots)
{
    int nr = 0;
    double u = a, fu = lsq_poly_eval(c, n, a);
    for (int k = 0; k <= ncrit; ++k) {
        const double v = k < ncrit ? crit[k] : b;
        if (!(v > u)) continue;
        const double fv = lsq_poly_eval(c, n, v);
        double r = 0;
        bool have = true;
        if (fu == 0.0) r = u;
        else if (fv == 0.0) r = v;
        else if ((fu < 0) != (fv < 0)) r = lsq_root_bracketed(c, n, u, v, fu);
        else have = false;
        if (have && (nr == 0 || r != roots[nr - 1])) roots[nr++] = r;
        u = v; fu = fv;
    }
    return nr;
}

/* real roots inside [a, b] of a polynomial of degree <= 4, ascending (degree 2 in closed form, 3 and 4 through the derivative) */
static int lsq_roots_in(const double *c, int n, double a, double b, double *roots)
{
    while (n > 0 && c[0] == 0.0) { ++c; --n; }
    if (n - 1 <= 2) return lsq_roots_low(c, n, a, b, roots);
    double d1[4], d2[3], crit2[4], crit1[4];
    for (int i = 0; i < n - 1; ++i) d1[i] = (n - 1 - i) * c[i];
    int n1;
    if (n - 1 == 3) n1 = lsq_roots_low(d1, 3, a, b, crit1);
    else {
        for (int i = 0; i < 3; ++i) d2[i] = (3 - i) * d1[i];
        const int n2 = lsq_roots_low(d2, 3, a, b, crit2);
        n1 = lsq_roots_from_crit(d1, 4, a, b, crit2, n2, crit1);
    }
    return lsq_roots_from_crit(c, n, a, b, crit1, n1, roots);
}

/* LineSearch::InterpolatingPolynomialMinimizingStepSize for CUBIC interpolation: the minimiser over [min_step, max_step] of the
 * polynomial through value and gradient of (lower bound at 0, current[, previous]).  FindInterpolatingPolynomial solves the
 * Vandermonde-type system of all (4 or 6) constraints with Eigen's fullPivLu; the same polynomial is obtained here in the
 * scaled variable u = x / current.x, q(u) = f0 + g0 xc u + b2 u^2 + ... (the two constraints at 0 fix the low coefficients),
 * from a 2 x 2 / 4 x 4 system with entries of order one.  MinimizePolynomial samples the midpoint, the ends and the real part
 * of EVERY root of the derivative (companion-matrix eigenvalues, complex ones included, "a bit of an overkill") inside the
 * interval; the minimum over an interval sits at an end or at a real critical point, so only the real roots inside the
 * interval can win, and those are found exactly (bracketed through the derivative's roots, polished by safeguarded Newton). */
static double ls_interpolating_step(const LsSample *lower, const LsSample *prev, const LsSample *cur, double min_step, double max_step)
{
    if (!cur->value_ok) return fmin(fmax(cur->x * 0.5, min_step), max_step);
    const double xc = cur->x, f0 = lower->value, g0 = lower->gradient * xc;
    double q[6];
    int nq;
    const double rv = cur->value - f0 - g0, rg = cur->gradient * xc - g0;
    if (!prev->value_ok) {
        /* cubic: b3 + b2 = rv, 3 b3 + 2 b2 = rg */
        const double b3 = rg - 2.0 * rv, b2 = 3.0 * rv - rg;
        q[0] = b3; q[1] = b2; q[2] = g0; q[3] = f0; nq = 4;
    } else {
        const double u = prev->x / xc, u2 = u * u, u3 = u2 * u, u4 = u3 * u, u5 = u4 * u;
        double A[4][5] = {{1, 1, 1, 1, rv}, {5, 4, 3, 2, rg},
                          {u5, u4, u3, u2, prev->value - f0 - g0 * u}, {5 * u4, 4 * u3, 3 * u2, 2 * u, prev->gradient * xc - g0}};
        for (int k = 0; k < 4; k++) {                      /* Gaussian elimination, partial pivoting */
            int pr = k;
            for (int r = k + 1; r < 4; r++) if (fabs(A[r][k]) > fabs(A[pr][k])) pr = r;
            if (pr != k) for (int c = 0; c < 5; c++) { const double t = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t; }
            if (A[k][k] == 0.0) continue;
            for (int r = k + 1; r < 4; r++) {
                const double f = A[r][k] / A[k][k];
                for (int c = k; c < 5; c++) A[r][c] -= f * A[k][c];
            }
        }
        double b[4];
        for (int k = 3; k >= 0; k--) {
            double v = A[k][4];
            for (int c = k + 1; c < 4; c++) v -= A[k][c] * b[c];
            b[k] = A[k][k] != 0.0 ? v / A[k][k] : 0.0;
        }
        q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; q[3] = b[3]; q[4] = g0; q[5] = f0; nq = 6;
    }
    /* MinimizePolynomial over [min_step, max_step] in u */
    const double ua = min_step / xc, ub = max_step / xc;
    double best_x = (min_step + max_step) / 2.0, best = lsq_poly_eval(q, nq, (ua + ub) / 2.0);
    double v = lsq_poly_eval(q, nq, ua);
    if (v < best) { best = v; best_x = min_step; }
    v = lsq_poly_eval(q, nq, ub);
    if (v < best) { best = v; best_x = max_step; }
    double d[5], roots[4];
    for (int i = 0; i < nq - 1; i++) d[i] = (nq - 1 - i) * q[i];
    const int nr = lsq_roots_in(d, nq - 1, ua, ub, roots);
    for (int i = 0; i < nr; i++) {
        v = lsq_poly_eval(q, nq, roots[i]);
        if (v < best) { best = v; best_x = roots[i] * xc; }
    }
    return best_x;
}

/* test hook: minimiser of the interpolating polynomial through up to three (x, value, gradient) samples */
double oracle_ls_interpolating_step(const double *xs, const double *vals, const double *grads, int ns, double min_step, double max_step)
{
    const int cur_ok = isfinite(vals[1]) && isfinite(grads[1]);
    LsSample lower = {xs[0], vals[0], grads[0], 1, 1}, cur = {xs[1], vals[1], grads[1], cur_ok, cur_ok}, prev = {0, 0, 0, 0, 0};
    if (ns > 2) { prev.x = xs[2]; prev.value = vals[2]; prev.gradient = grads[2]; prev.value_ok = 1; prev.grad_ok = 1; }
    return ls_interpolating_step(&lower, &prev, &cur, min_step, max_step);
}

/* ------------------------------------------------------------------ */
/* the solver (Ceres trust-region / traditional dogleg / dense Schur)   */
/* ------------------------------------------------------------------ */
typedef struct {
    int iterations, successful, termination;
    int armijo_failures;    /* steps of a bound-constrained problem that fail Ceres' Armijo test at step size 1: the line search ran */
    int line_search_steps;  /* ... and returned a shortened step */
    double initial_cost, final_cost;
} SolveSummary;

static void apply_delta(const Lin *L, const State *x, const double *delta, State *o, double *olam)
{
    memcpy(o, x, sizeof(State));
    o->lam = olam;
    for (int f = 0; f < NF; f++) {
        if (col_active(L, COL_POSE(f))) pose_plus(x->pose[f], delta + COL_POSE(f), o->pose[f]);
        if (col_active(L, COL_SB(f))) for (int k = 0; k < 9; k++) o->sb[f][k] = x->sb[f][k] + delta[COL_SB(f) + k];
    }
    if (L->ex_active) pose_plus(x->ex, delta + COL_EX, o->ex);
    if (L->td_active) o->td = x->td + delta[COL_TD];
    for (int l = 0; l < L->M; l++) {
        double v = x->lam[l];
        if (!L->lm_const[l]) {
            v += delta[NC + l];
            if (v > L->lm_ub[l]) v = L->lm_ub[l];       /* ParameterBlock::Plus projects onto the bounds */
        }
        olam[l] = v;
    }
}

static double active_x_norm2_diff(const Lin *L, const State *a, const State *b)
{   /* squared norm of (a - b) over the non-constant parameter blocks (global parameterisation);
       b == NULL -> squared norm of a */
    double s = 0;
    for (int f = 0; f < NF; f++) {
        if (col_active(L, COL_POSE(f))) for (int k = 0; k < 7; k++) { double d = a->pose[f][k] - (b ? b->pose[f][k] : 0); s += d * d; }
        if (col_active(L, COL_SB(f))) for (int k = 0; k < 9; k++) { double d = a->sb[f][k] - (b ? b->sb[f][k] : 0); s += d * d; }
    }
    if (L->ex_active) for (int k = 0; k < 7; k++) { double d = a->ex[k] - (b ? b->ex[k] : 0); s += d * d; }
    if (L->td_active) { double d = a->td - (b ? b->td : 0); s += d * d; }
    for (int l = 0; l < L->M; l++) if (!L->lm_const[l]) { double d = a->lam[l] - (b ? b->lam[l] : 0); s += d * d; }
    return s;
}

static int solve(const VrfBaProblem *pb, const VrfConfig *cfg, Lin *L, State *x, SolveSummary *sum)
{
    const int M = L->M, NT = NC + M;
    const int nrows = 2 * L->nfac + 15 * L->nimu + L->np;
    const int max_iter = pb->max_iterations > 0 ? pb->max_iterations : cfg->num_iterations;
    double *jscale = (double *)malloc(sizeof(double) * NT);
    double *g = (double *)malloc(sizeof(double) * NT);          /* gradient of scaled problem */
    double *H = (double *)malloc(sizeof(double) * NC * NC);
    double *Sm = (double *)malloc(sizeof(double) * NC * NC);
    double *W = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1) * NC);
    double *hll = (double *)malloc(sizeof(double) * (M > 0 ? M : 1));
    double *diag = (double *)malloc(sizeof(double) * NT);
    double *gd = (double *)malloc(sizeof(double) * NT);          /* gradient_ / diagonal_ (dogleg space) */
    double *gn = (double *)malloc(sizeof(double) * NT);          /* gauss_newton_step_ (dogleg space) */
    double *step = (double *)malloc(sizeof(double) * NT);
    double *delta = (double *)malloc(sizeof(double) * NT);
    double *tmp = (double *)malloc(sizeof(double) * NT);
    double *rowbuf = (double *)malloc(sizeof(double) * (nrows > 0 ? nrows : 1));
    double *lam2 = (double *)malloc(sizeof(double) * (M > 0 ? M : 1));
    double *lam3 = (double *)malloc(sizeof(double) * (M > 0 ? M : 1));
    State cand;
    int rc = VRF_OK;

    /* TrustRegionMinimizer::Init of a bound-constrained problem: the start point is projected onto the bounds
     * (x = Plus(x, 0)) before the first evaluation.  Only the estimate_flag == 2 landmarks carry a bound
     * (estimator.cpp:1293-1298); constant blocks are not part of the program. */
    for (int l = 0; l < M; l++)
        if (!L->lm_const[l] && x->lam[l] > L->lm_ub[l]) x->lam[l] = L->lm_ub[l];
    double x_cost = evaluate(pb, cfg, L, x, 1);
    sum->initial_cost = x_cost; sum->iterations = 0; sum->successful = 0; sum->termination = 0; sum->armijo_failures = 0; sum->line_search_steps = 0;
    /* Solver::Options::is_constrained: some non-constant parameter block carries a bound */
    int constrained = 0;
    for (int l = 0; l < M; l++) if (!L->lm_const[l] && isfinite(L->lm_ub[l])) constrained = 1;
    /* Jacobi scaling, computed once at iteration 0 */
    {
        CtxColNorm cn = {NULL, tmp};
        memset(tmp, 0, sizeof(double) * NT);
        for_each_row(pb, L, fn_colnorm, &cn);
        for (int c = 0; c < NT; c++) jscale[c] = 1.0 / (1.0 + sqrt(tmp[c]));
    }
    double radius = 1e4, mu = 1e-8, alpha = 0, dogleg_step_norm = 0;
    const double min_mu = 1e-8, max_mu = 1.0, mu_inc = 10.0;
    int reuse = 0, invalid_steps = 0, need_lin = 1;
    double x_norm = sqrt(active_x_norm2_diff(L, x, NULL));
    double gradient_max_norm = 0;

    for (;;) {
        if (need_lin) {
            /* gradient / normal equations of the scaled problem */
            memset(g, 0, sizeof(double) * NT); memset(H, 0, sizeof(double) * NC * NC);
            memset(W, 0, sizeof(double) * (size_t)(M > 0 ? M : 1) * NC); memset(hll, 0, sizeof(double) * (M > 0 ? M : 1));
            /* run the generic rows without the prior (handled densely) */
            int np_save = L->np;
            L->np = 0;
            CtxNormal cn = {jscale, g, H, hll, W, M};
            for_each_row(pb, L, fn_normal, &cn);
            L->np = np_save;
            if (L->np) normal_prior(pb, L, jscale, g, H);
            /* projected gradient norm (bounds): x - P(x + (-g_unscaled)) */
            gradient_max_norm = 0;
            {
                for (int c = 0; c < NT; c++) tmp[c] = -g[c] / jscale[c] * 1.0;   /* unscaled gradient = g_s / scale */
                apply_delta(L, x, tmp, &cand, lam2);
                /* (x - projected step) in the ambient space */
                double mx = 0;
                for (int f = 0; f < NF; f++) {
                    if (col_active(L, COL_POSE(f))) for (int k = 0; k < 7; k++) mx = fmax(mx, fabs(x->pose[f][k] - cand.pose[f][k]));
                    if (col_active(L, COL_SB(f))) for (int k = 0; k < 9; k++) mx = fmax(mx, fabs(x->sb[f][k] - cand.sb[f][k]));
                }
                if (L->ex_active) for (int k = 0; k < 7; k++) mx = fmax(mx, fabs(x->ex[k] - cand.ex[k]));
                if (L->td_active) mx = fmax(mx, fabs(x->td - cand.td));
                for (int l = 0; l < M; l++) if (!L->lm_const[l]) mx = fmax(mx, fabs(x->lam[l] - cand.lam[l]));
                gradient_max_norm = mx;
            }
            need_lin = 0;
        }
        /* FinalizeIterationAndCheckIfMinimizerCanContinue */
        if (sum->iterations >= max_iter) { sum->termination = 0; break; }
        if (gradient_max_norm <= 1e-10) { sum->termination = 2; break; }
        if (radius <= 1e-32) { sum->termination = 4; break; }
        sum->iterations++;

        /* ---- DoglegStrategy::ComputeStep ---- */
        int step_ok = 1;
        if (!reuse) {
            reuse = 1;
            for (int c = 0; c < NC; c++) diag[c] = H[c * NC + c];
            for (int l = 0; l < M; l++) diag[NC + l] = hll[l];
            for (int c = 0; c < NT; c++) diag[c] = sqrt(fmin(fmax(diag[c], 1e-6), 1e32));
            for (int c = 0; c < NT; c++) gd[c] = g[c] / diag[c];
            /* Cauchy point: alpha = |gd|^2 / |J (D^-1 gd)|^2 */
            {
                for (int c = 0; c < NT; c++) tmp[c] = gd[c] / diag[c];
                CtxMul cm = {jscale, tmp, rowbuf};
                for_each_row(pb, L, fn_mul, &cm);
                double a = 0, b = 0;
                for (int c = 0; c < NT; c++) a += gd[c] * gd[c];
                for (int r = 0; r < nrows; r++) b += rowbuf[r] * rowbuf[r];
                alpha = a / b;
            }
            /* Gauss-Newton step with increasing regularisation */
            int solved = 0;
            while (mu < max_mu) {
                /* Schur complement onto the camera block */
                memcpy(Sm, H, sizeof(double) * NC * NC);
                for (int c = 0; c < NC; c++) { Sm[c * NC + c] += mu * diag[c] * diag[c]; tmp[c] = g[c]; }
                int bad = 0;
                for (int l = 0; l < M; l++) {
                    if (L->lm_const[l]) continue;
                    double hl = hll[l] + mu * diag[NC + l] * diag[NC + l];
                    if (!(hl > 0)) { bad = 1; break; }
                    const double *w = W + (size_t)l * NC;
                    double gl = g[NC + l] / hl;
                    for (int a = 0; a < NC; a++) {
                        if (w[a] == 0.0) continue;
                        double wa = w[a] / hl;
                        tmp[a] -= w[a] * gl;
                        for (int b = 0; b < NC; b++) if (w[b] != 0.0) Sm[a * NC + b] -= wa * w[b];
                    }
                }
                if (!bad && chol_lower(Sm, NC, NC) == 0) {
                    chol_solve(Sm, NC, NC, tmp);
                    int finite = 1;
                    for (int c = 0; c < NC; c++) if (!isfinite(tmp[c])) finite = 0;
                    if (finite) {
                        for (int c = 0; c < NC; c++) gn[c] = tmp[c];
                        for (int l = 0; l < M; l++) {
                            if (L->lm_const[l]) { gn[NC + l] = 0; continue; }
                            double hl = hll[l] + mu * diag[NC + l] * diag[NC + l];
                            const double *w = W + (size_t)l * NC;
                            double a = g[NC + l];
                            for (int c = 0; c < NC; c++) a -= w[c] * gn[c];
                            gn[NC + l] = a / hl;
                        }
                        solved = 1;
                        break;
                    }
                }
                mu *= mu_inc;
            }
            if (!solved) { rc = VRF_SOFT_NOT_SPD; step_ok = 0; }
            else for (int c = 0; c < NT; c++) gn[c] *= -diag[c];
        }
        if (step_ok) {
            /* ComputeTraditionalDoglegStep */
            double gnorm = 0, gnn = 0;
            for (int c = 0; c < NT; c++) { gnorm += gd[c] * gd[c]; gnn += gn[c] * gn[c]; }
            gnorm = sqrt(gnorm); gnn = sqrt(gnn);
            if (gnn <= radius) {
                for (int c = 0; c < NT; c++) step[c] = gn[c];
                dogleg_step_norm = gnn;
            } else if (gnorm * alpha >= radius) {
                for (int c = 0; c < NT; c++) step[c] = -(radius / gnorm) * gd[c];
                dogleg_step_norm = radius;
            } else {
                double b_dot_a = 0;
                for (int c = 0; c < NT; c++) b_dot_a += gd[c] * gn[c];
                b_dot_a *= -alpha;
                double a_sq = pow(alpha * gnorm, 2.0);
                double bma = a_sq - 2 * b_dot_a + pow(gnn, 2);
                double cc = b_dot_a - a_sq;
                double dd = sqrt(cc * cc + bma * (pow(radius, 2.0) - a_sq));
                double beta = (cc <= 0) ? (dd - cc) / bma : (radius * radius - a_sq) / (dd + cc);
                double nn = 0;
                for (int c = 0; c < NT; c++) { step[c] = (-alpha * (1.0 - beta)) * gd[c] + beta * gn[c]; nn += step[c] * step[c]; }
                dogleg_step_norm = sqrt(nn);
            }
            for (int c = 0; c < NT; c++) step[c] /= diag[c];
            /* model_cost_change = -(J s) . (r + J s / 2) */
            CtxMul cm = {jscale, step, rowbuf};
            for_each_row(pb, L, fn_mul, &cm);
            double mcc = 0;
            {
                int rid = 0;
                for (int k = 0; k < L->nfac; k++) for (int rr = 0; rr < 2; rr++, rid++) mcc -= rowbuf[rid] * (L->f_r[2 * k + rr] + rowbuf[rid] / 2.0);
                for (int f = 0; f < L->nimu; f++) for (int rr = 0; rr < 15; rr++, rid++) mcc -= rowbuf[rid] * (L->imu_r[f][rr] + rowbuf[rid] / 2.0);
                for (int rr = 0; rr < L->np; rr++, rid++) mcc -= rowbuf[rid] * (L->pr_r[rr] + rowbuf[rid] / 2.0);
            }
            if (!(mcc > 0.0)) step_ok = 0;
            else {
                invalid_steps = 0;
                for (int c = 0; c < NT; c++) delta[c] = step[c] * jscale[c];
                apply_delta(L, x, delta, &cand, lam2);
                double cand_cost = evaluate(pb, cfg, L, &cand, 0);
                /* For a bound-constrained problem Ceres runs a projected Armijo line search on the step before evaluating the
                 * candidate (TrustRegionMinimizer::DoLineSearch).  phi(t) = f(x [+] t delta) with the bounds projection inside
                 * Plus, phi'(t) = delta . gradient(x [+] t delta).  When the full step passes the sufficient-decrease test the
                 * search returns t = 1 and nothing changes; otherwise the step is contracted to the minimiser of the cubic /
                 * quintic through (0, current, previous) inside [1e-3 t, 0.6 t] until the test holds (<= 20 iterations), and
                 * delta is scaled by the accepted t.  A failed search leaves delta alone.  The trust-region bookkeeping
                 * (model cost change, dogleg step norm) keeps referring to the unshortened step, as in Ceres. */
                if (constrained) {
                    double gts = 0, dmax = 0;
                    for (int c = 0; c < NT; c++) { gts += g[c] * step[c]; dmax = fmax(dmax, fabs(delta[c])); }   /* unscaled gradient . delta */
                    if (!(isfinite(cand_cost)) || cand_cost > x_cost + 1e-4 * gts) {
                        sum->armijo_failures++;
                        LsSample lower = {0.0, x_cost, gts, 1, 1}, prev = {0, 0, 0, 0, 0}, cur = {1.0, cand_cost, 0, 0, 1};
                        State trial;
                        int ok = 0;
                        /* value and gradient at t = 1 */
                        for (int it = 0;; it++) {
                            if (it > 0) {
                                if (it >= 20) break;                                  /* max_num_line_search_step_size_iterations */
                                const double t = ls_interpolating_step(&lower, &prev, &cur, 1e-3 * cur.x, 0.6 * cur.x);
                                if (t * dmax < 1e-9) break;                           /* min_line_search_step_size */
                                prev = cur;
                                cur.x = t;
                            }
                            for (int c = 0; c < NT; c++) tmp[c] = cur.x * delta[c];
                            apply_delta(L, x, tmp, &trial, lam3);
                            cur.value = evaluate(pb, cfg, L, &trial, 1);
                            memset(tmp, 0, sizeof(double) * NT);
                            for_each_row(pb, L, fn_grad, tmp);
                            cur.gradient = 0;
                            for (int c = 0; c < NT; c++) cur.gradient += tmp[c] * delta[c];
                            cur.value_ok = isfinite(cur.value) && isfinite(cur.gradient);
                            cur.grad_ok = cur.value_ok;
                            if (getenv("ORACLE_BA_DEBUG"))
                                fprintf(stderr, "   ls it %d t %.6g value %.9f dphi %.6g (phi0 %.9f dphi0 %.6g dmax %.3g)\n", it, cur.x, cur.value, cur.gradient, x_cost, gts, dmax);
                            if (cur.value_ok && cur.value <= x_cost + 1e-4 * gts * cur.x) { ok = (it > 0); break; }
                        }
                        if (ok) {
                            for (int c = 0; c < NT; c++) delta[c] *= cur.x;
                            apply_delta(L, x, delta, &cand, lam2);
                            cand_cost = cur.value;
                            sum->line_search_steps++;
                        }
                        evaluate(pb, cfg, L, x, 1);       /* the linearisation at x again (the rows are re-used by a rejected step) */
                    }
                }
                double step_norm = sqrt(active_x_norm2_diff(L, x, &cand));
                if (step_norm <= 1e-8 * (x_norm + 1e-8)) { sum->termination = 3; break; }      /* parameter tolerance */
                double cost_change = x_cost - cand_cost;
                if (fabs(cost_change) <= 1e-6 * x_cost) {      /* function tolerance: terminate, candidate NOT accepted */
                    sum->termination = 1;
                    break;
                }
                double rho = cost_change / mcc;
                if (getenv("ORACLE_BA_DEBUG"))
                    fprintf(stderr, "it %d cost %.6f cand %.6f mcc %.6g rho %.4f radius %.3g |step| %.3g mu %.1e gradmax %.3g gn %.3g\n",
                            sum->iterations, x_cost, cand_cost, mcc, rho, radius, dogleg_step_norm, mu, gradient_max_norm, 0.0);
                if (rho > 1e-3) {
                    double *keep = x->lam;
                    *x = cand; x->lam = keep; memcpy(x->lam, lam2, sizeof(double) * M);
                    x_cost = cand_cost;
                    x_norm = sqrt(active_x_norm2_diff(L, x, NULL));
                    evaluate(pb, cfg, L, x, 1);
                    need_lin = 1;
                    sum->successful++;
                    /* DoglegStrategy::StepAccepted */
                    if (rho < 0.25) radius *= 0.5;
                    if (rho > 0.75) radius = fmax(radius, 3.0 * dogleg_step_norm);
                    mu = fmax(min_mu, 2.0 * mu / mu_inc);
                    reuse = 0;
                } else {
                    radius *= 0.5;          /* StepRejected */
                    reuse = 1;
                }
                continue;
            }
        }
        /* invalid step */
        if (++invalid_steps >= 5) { sum->termination = 5; break; }
        mu *= mu_inc;                        /* StepIsInvalid */
        reuse = 0;
        if (rc == VRF_SOFT_NOT_SPD && mu >= max_mu) { sum->termination = 5; break; }
    }
    sum->final_cost = x_cost;
    free(jscale); free(g); free(H); free(Sm); free(W); free(hll); free(diag); free(gd); free(gn); free(step);
    free(delta); free(tmp); free(rowbuf); free(lam2); free(lam3);
    (void)state_plus;
    return rc;
}

/* ------------------------------------------------------------------ */
/* double2vector gauge fix (estimator.cpp:985-1111, USE_IMU branch)      */
/* ------------------------------------------------------------------ */
static void R2ypr(const double R[9], double ypr[3])
{
    double n0 = R[0], n1 = R[3], n2 = R[6], o0 = R[1], o1 = R[4], a0 = R[2], a1 = R[5];
    double y = atan2(n1, n0);
    double p = atan2(-n2, n0 * cos(y) + n1 * sin(y));
    double r = atan2(a0 * sin(y) - a1 * cos(y), -o0 * sin(y) + o1 * cos(y));
    ypr[0] = y / M_PI * 180.0; ypr[1] = p / M_PI * 180.0; ypr[2] = r / M_PI * 180.0;
}
static void ypr2R(const double ypr[3], double R[9])
{
    double y = ypr[0] / 180.0 * M_PI, p = ypr[1] / 180.0 * M_PI, r = ypr[2] / 180.0 * M_PI;
    double Rz[9] = {cos(y), -sin(y), 0, sin(y), cos(y), 0, 0, 0, 1};
    double Ry[9] = {cos(p), 0., sin(p), 0., 1., 0., -sin(p), 0., cos(p)};
    double Rx[9] = {1., 0., 0., 0., cos(r), -sin(r), 0., sin(r), cos(r)};
    double T[9];
    m3_mul(Rz, Ry, T); m3_mul(T, Rx, R);
}

static void gauge_fix(const VrfBaProblem *pb, const State *x, VrfBaResult *res)
{
    /* origin = pose of frame 0 BEFORE the solve (Rs[0], Ps[0] <- vector2double input) */
    double R0[9], q0[4] = {pb->para_Pose[0][3], pb->para_Pose[0][4], pb->para_Pose[0][5], pb->para_Pose[0][6]};
    q_to_R(q0, R0);
    double origin_R0[3], origin_R00[3], R00[9];
    R2ypr(R0, origin_R0);
    const double origin_P0[3] = {pb->para_Pose[0][0], pb->para_Pose[0][1], pb->para_Pose[0][2]};
    if (pb->use_imu) {
        q_to_R(x->pose[0] + 3, R00);
        R2ypr(R00, origin_R00);
        double y_diff = origin_R0[0] - origin_R00[0];
        double ypr[3] = {y_diff, 0, 0}, rot_diff[9];
        ypr2R(ypr, rot_diff);
        if (fabs(fabs(origin_R0[1]) - 90) < 1.0 || fabs(fabs(origin_R00[1]) - 90) < 1.0) {
            double R00T[9];
            m3_T(R00, R00T); m3_mul(R0, R00T, rot_diff);
        }
        for (int i = 0; i < NF; i++) {
            double q[4] = {x->pose[i][3], x->pose[i][4], x->pose[i][5], x->pose[i][6]}, R[9];
            q_normalize(q); q_to_R(q, R);
            m3_mul(rot_diff, R, res->Rs[i]);
            double d[3] = {x->pose[i][0] - x->pose[0][0], x->pose[i][1] - x->pose[0][1], x->pose[i][2] - x->pose[0][2]}, t[3];
            m3_v(rot_diff, d, t);
            for (int k = 0; k < 3; k++) res->Ps[i][k] = t[k] + origin_P0[k];
            m3_v(rot_diff, x->sb[i], res->Vs[i]);
            for (int k = 0; k < 3; k++) { res->Bas[i][k] = x->sb[i][3 + k]; res->Bgs[i][k] = x->sb[i][6 + k]; }
        }
    } else {
        for (int i = 0; i < NF; i++) {
            double q[4] = {x->pose[i][3], x->pose[i][4], x->pose[i][5], x->pose[i][6]};
            q_normalize(q); q_to_R(q, res->Rs[i]);
            for (int k = 0; k < 3; k++) { res->Ps[i][k] = x->pose[i][k]; res->Vs[i][k] = 0; res->Bas[i][k] = 0; res->Bgs[i][k] = 0; }
        }
    }
}

/* vector2double after the gauge fix: states at which marginalization linearises */
static void repack_state(const VrfBaProblem *pb, const VrfBaResult *res, const State *x, State *o, double *olam)
{
    memcpy(o, x, sizeof(State));
    o->lam = olam;
    for (int i = 0; i < NF; i++) {
        for (int k = 0; k < 3; k++) o->pose[i][k] = res->Ps[i][k];
        R_to_q(res->Rs[i], o->pose[i] + 3);
        if (pb->use_imu) {
            for (int k = 0; k < 3; k++) { o->sb[i][k] = res->Vs[i][k]; o->sb[i][3 + k] = res->Bas[i][k]; o->sb[i][6 + k] = res->Bgs[i][k]; }
        }
    }
    if (pb->use_imu) {
        /* tic/ric <- para_Ex_Pose (normalised) and back through Quaterniond(ric) */
        double q[4] = {x->ex[3], x->ex[4], x->ex[5], x->ex[6]}, R[9];
        q_normalize(q); q_to_R(q, R); R_to_q(R, o->ex + 3);
    }
    for (int l = 0; l < pb->n_landmarks; l++) {
        double depth = 1.0 / x->lam[l];          /* setDepth: estimated_depth = 1/x */
        olam[l] = 1. / depth;                    /* getDepthVector: 1/estimated_depth */
    }
}

/* ------------------------------------------------------------------ */
/* marginalization (estimator.cpp:1376-1574 + marginalization_factor.cpp) */
/* ------------------------------------------------------------------ */
typedef struct { int kind, index, gsize, lsize, idx, present, drop; double x0[9]; } MBlock;

static int mblock_find(MBlock *B, int nb, int kind, int index)
{
    for (int i = 0; i < nb; i++) if (B[i].kind == kind && B[i].index == index) return i;
    return -1;
}

static int marginalize(const VrfBaProblem *pb, const VrfConfig *cfg, const State *x, VrfPrior *out)
{
    const int flag = pb->marginalization_flag;
    const VrfPrior *P = pb->prior;
    const int M = pb->n_landmarks;
    if (flag == VRF_MARGIN_SECOND_NEW) {
        int has = 0;
        if (P) for (int b = 0; b < P->n_blocks; b++) if (P->blocks[b].kind == VRF_BLK_POSE && P->blocks[b].index == VRF_WINDOW_SIZE - 1) has = 1;
        if (!has) return 0;     /* prior unchanged (estimator.cpp:1505-1507) */
    }
    /* canonical block order: dropped first, then kept (ex, td, pose f.., speed-bias f..) */
    MBlock *B = (MBlock *)calloc(64 + M, sizeof(MBlock));
    int nb = 0;
#define ADDB(k_, i_, gs_, dr_, src_) do { if (mblock_find(B, nb, k_, i_) < 0) { B[nb].kind = k_; B[nb].index = i_; B[nb].gsize = gs_; B[nb].lsize = (gs_) == 7 ? 6 : (gs_); B[nb].drop = dr_; memcpy(B[nb].x0, src_, sizeof(double) * (gs_)); nb++; } } while (0)
    const int LM = 100;         /* kind for landmarks */
    if (flag == VRF_MARGIN_OLD) {
        ADDB(VRF_BLK_POSE, 0, 7, 1, x->pose[0]);
        if (pb->use_imu) ADDB(VRF_BLK_SPEEDBIAS, 0, 9, 1, x->sb[0]);
        for (int l = 0; l < M; l++) {
            int nobs = pb->lm_obs_ptr[l + 1] - pb->lm_obs_ptr[l];
            if (pb->lm_start_frame[l] == 0 && nobs >= 2) ADDB(LM, l, 1, 1, &x->lam[l]);
        }
    } else {
        ADDB(VRF_BLK_POSE, VRF_WINDOW_SIZE - 1, 7, 1, x->pose[VRF_WINDOW_SIZE - 1]);
    }
    /* kept candidates in canonical order; presence decided by the factors below */
    int first_kept = nb;
    ADDB(VRF_BLK_EXPOSE, 0, 7, 0, x->ex);
    ADDB(VRF_BLK_TD, 0, 1, 0, &x->td);
    for (int f = 0; f < NF; f++) ADDB(VRF_BLK_POSE, f, 7, 0, x->pose[f]);
    for (int f = 0; f < NF; f++) ADDB(VRF_BLK_SPEEDBIAS, f, 9, 0, x->sb[f]);
#undef ADDB
    /* mark presence */
    if (P) for (int b = 0; b < P->n_blocks; b++) { int i = mblock_find(B, nb, P->blocks[b].kind, P->blocks[b].index); if (i >= 0) B[i].present = 1; }
    int use_imu01 = (flag == VRF_MARGIN_OLD) && pb->use_imu && pb->imu[0].sum_dt < 10.0;
    if (use_imu01) {
        B[mblock_find(B, nb, VRF_BLK_POSE, 0)].present = 1; B[mblock_find(B, nb, VRF_BLK_SPEEDBIAS, 0)].present = 1;
        B[mblock_find(B, nb, VRF_BLK_POSE, 1)].present = 1; B[mblock_find(B, nb, VRF_BLK_SPEEDBIAS, 1)].present = 1;
    }
    if (flag == VRF_MARGIN_OLD) {
        for (int l = 0; l < M; l++) {
            int nobs = pb->lm_obs_ptr[l + 1] - pb->lm_obs_ptr[l];
            if (pb->lm_start_frame[l] != 0 || nobs < 2) continue;
            B[mblock_find(B, nb, LM, l)].present = 1;
            B[mblock_find(B, nb, VRF_BLK_POSE, 0)].present = 1;
            B[mblock_find(B, nb, VRF_BLK_EXPOSE, 0)].present = 1;
            if (cfg->estimate_td) B[mblock_find(B, nb, VRF_BLK_TD, 0)].present = 1;      /* estimator.cpp:1445-1458 */
            for (int k = 1; k < nobs; k++) B[mblock_find(B, nb, VRF_BLK_POSE, k)].present = 1;
        }
    }
    int pos = 0, m = 0;
    for (int i = 0; i < nb; i++) if (B[i].present && B[i].drop) { B[i].idx = pos; pos += B[i].lsize; }
    m = pos;
    for (int i = first_kept; i < nb; i++) if (B[i].present && !B[i].drop) { B[i].idx = pos; pos += B[i].lsize; }
    const int n = pos - m;
    double *A = (double *)calloc((size_t)pos * pos, sizeof(double));
    double *bv = (double *)calloc(pos, sizeof(double));
    /* --- prior factor --- */
    if (P) {
        int np = P->n;
        double r[VRF_PRIOR_MAX_DIM];
        prior_residual(P, x, r);
        int cmap[VRF_PRIOR_MAX_DIM];
        for (int b = 0; b < P->n_blocks; b++) {
            int i = mblock_find(B, nb, P->blocks[b].kind, P->blocks[b].index);
            int ls = P->blocks[b].size == 7 ? 6 : P->blocks[b].size;
            for (int c = 0; c < ls; c++) cmap[P->blocks[b].idx + c] = B[i].idx + c;
        }
        for (int a = 0; a < np; a++) {
            double ga = 0;
            for (int rr = 0; rr < np; rr++) ga += P->linearized_jacobians[(size_t)rr * np + a] * r[rr];
            bv[cmap[a]] += ga;
            for (int c = 0; c < np; c++) {
                double h = 0;
                for (int rr = 0; rr < np; rr++) h += P->linearized_jacobians[(size_t)rr * np + a] * P->linearized_jacobians[(size_t)rr * np + c];
                A[(size_t)cmap[a] * pos + cmap[c]] += h;
            }
        }
    }
    /* --- IMU factor (0,1) --- */
    if (use_imu01) {
        double r[15], Jpi[105], Jsi[135], Jpj[105], Jsj[135];
        oracle_imu_eval(&pb->imu[0], x->pose[0], x->sb[0], x->pose[1], x->sb[1], cfg->g_norm, r, Jpi, Jsi, Jpj, Jsj);
        int cols[30]; double J[15][30];
        int i0 = B[mblock_find(B, nb, VRF_BLK_POSE, 0)].idx, i1 = B[mblock_find(B, nb, VRF_BLK_SPEEDBIAS, 0)].idx;
        int i2 = B[mblock_find(B, nb, VRF_BLK_POSE, 1)].idx, i3 = B[mblock_find(B, nb, VRF_BLK_SPEEDBIAS, 1)].idx;
        for (int c = 0; c < 6; c++) { cols[c] = i0 + c; cols[15 + c] = i2 + c; }
        for (int c = 0; c < 9; c++) { cols[6 + c] = i1 + c; cols[21 + c] = i3 + c; }
        for (int rr = 0; rr < 15; rr++) {
            for (int c = 0; c < 6; c++) { J[rr][c] = Jpi[rr * 7 + c]; J[rr][15 + c] = Jpj[rr * 7 + c]; }
            for (int c = 0; c < 9; c++) { J[rr][6 + c] = Jsi[rr * 9 + c]; J[rr][21 + c] = Jsj[rr * 9 + c]; }
        }
        for (int a = 0; a < 30; a++) {
            double ga = 0;
            for (int rr = 0; rr < 15; rr++) ga += J[rr][a] * r[rr];
            bv[cols[a]] += ga;
            for (int c = 0; c < 30; c++) {
                double h = 0;
                for (int rr = 0; rr < 15; rr++) h += J[rr][a] * J[rr][c];
                A[(size_t)cols[a] * pos + cols[c]] += h;
            }
        }
    }
    /* --- projection factors hosted at frame 0 (with loss, all four blocks) --- */
    if (flag == VRF_MARGIN_OLD) {
        for (int l = 0; l < M; l++) {
            int o0 = pb->lm_obs_ptr[l], nobs = pb->lm_obs_ptr[l + 1] - o0;
            if (pb->lm_start_frame[l] != 0 || nobs < 2) continue;
            for (int k = 1; k < nobs; k++) {
                double r[2], Ji[14], Jj[14], Jex[14], Jf[2], Jt[2] = {0, 0};
                const int ntd = cfg->estimate_td ? 1 : 0;
                if (ntd)
                    oracle_projection_td_eval(x->pose[0], x->pose[k], x->ex, x->lam[l], x->td, pb->obs_pts + 2 * o0, pb->obs_pts + 2 * (o0 + k),
                                              pb->obs_velocity + 2 * o0, pb->obs_velocity + 2 * (o0 + k), pb->obs_cur_td[o0], pb->obs_cur_td[o0 + k],
                                              pb->obs_row[o0], pb->obs_row[o0 + k], cfg->tr / (double)cfg->row, r, Ji, Jj, Jex, Jf, Jt);
                else
                    oracle_projection_eval(x->pose[0], x->pose[k], x->ex, x->lam[l], pb->obs_pts + 2 * o0, pb->obs_pts + 2 * (o0 + k), r, Ji, Jj, Jex, Jf);
                double li[12], lj[12], le[12], lf[2] = {Jf[0], Jf[1]};
                for (int rr = 0; rr < 2; rr++) for (int c = 0; c < 6; c++) { li[rr * 6 + c] = Ji[rr * 7 + c]; lj[rr * 6 + c] = Jj[rr * 7 + c]; le[rr * 6 + c] = Jex[rr * 7 + c]; }
                double *Jb[5] = {li, lj, le, lf, Jt}; int ncl[5] = {6, 6, 6, 1, 1};
                cauchy_correct(r, Jb, ncl, 4 + ntd);
                int cols[20]; double J[2][20];
                int ci = B[mblock_find(B, nb, VRF_BLK_POSE, 0)].idx, cj = B[mblock_find(B, nb, VRF_BLK_POSE, k)].idx;
                int ce = B[mblock_find(B, nb, VRF_BLK_EXPOSE, 0)].idx, cl = B[mblock_find(B, nb, LM, l)].idx;
                for (int c = 0; c < 6; c++) { cols[c] = ci + c; cols[6 + c] = cj + c; cols[12 + c] = ce + c; }
                cols[18] = cl;
                if (ntd) cols[19] = B[mblock_find(B, nb, VRF_BLK_TD, 0)].idx;
                for (int rr = 0; rr < 2; rr++) { for (int c = 0; c < 6; c++) { J[rr][c] = li[rr * 6 + c]; J[rr][6 + c] = lj[rr * 6 + c]; J[rr][12 + c] = le[rr * 6 + c]; } J[rr][18] = lf[rr]; J[rr][19] = Jt[rr]; }
                for (int a = 0; a < 19 + ntd; a++) {
                    bv[cols[a]] += J[0][a] * r[0] + J[1][a] * r[1];
                    for (int c = 0; c < 19 + ntd; c++) A[(size_t)cols[a] * pos + cols[c]] += J[0][a] * J[0][c] + J[1][a] * J[1][c];
                }
            }
        }
    }
    /* --- Schur complement with eigen-decomposition pseudo-inverse (marginalization_factor.cpp:273-308) --- */
    const double eps = 1e-8;
    double *Amm = (double *)malloc(sizeof(double) * (size_t)m * m), *Vm = (double *)malloc(sizeof(double) * (size_t)m * m), *wm = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]);
    sym_eig_jacobi(Amm, m, Vm, wm);
    double *Ainv = (double *)calloc((size_t)m * m, sizeof(double));
    for (int k = 0; k < m; k++) {
        if (!(wm[k] > eps)) continue;
        double iw = 1.0 / wm[k];
        for (int i = 0; i < m; i++) { double vi = Vm[(size_t)i * m + k] * iw; if (vi == 0.0) continue; for (int j = 0; j < m; j++) Ainv[(size_t)i * m + j] += vi * Vm[(size_t)j * m + k]; }
    }
    double *T = (double *)malloc(sizeof(double) * (size_t)n * m);      /* Arm * Amm_inv */
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) { double a = 0; for (int k = 0; k < m; k++) a += A[(size_t)(m + i) * pos + k] * Ainv[(size_t)k * m + j]; T[(size_t)i * m + j] = a; }
    double *Ar = (double *)malloc(sizeof(double) * (size_t)n * n), *br = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) {
        double bb = bv[m + i];
        for (int k = 0; k < m; k++) bb -= T[(size_t)i * m + k] * bv[k];
        br[i] = bb;
        for (int j = 0; j < n; j++) { double a = A[(size_t)(m + i) * pos + m + j]; for (int k = 0; k < m; k++) a -= T[(size_t)i * m + k] * A[(size_t)k * pos + m + j]; Ar[(size_t)i * n + j] = a; }
    }
    double *Vr = (double *)malloc(sizeof(double) * (size_t)n * n), *wr = (double *)malloc(sizeof(double) * n);
    double *Ac = (double *)malloc(sizeof(double) * (size_t)n * n);
    memcpy(Ac, Ar, sizeof(double) * (size_t)n * n);
    sym_eig_jacobi(Ac, n, Vr, wr);
    memset(out, 0, sizeof(*out));
    out->n = n;
    for (int k = 0; k < n; k++) {
        double S = wr[k] > eps ? wr[k] : 0.0, Sinv = wr[k] > eps ? 1.0 / wr[k] : 0.0;
        double ss = sqrt(S), sis = sqrt(Sinv), vb = 0;
        for (int j = 0; j < n; j++) { out->linearized_jacobians[(size_t)k * n + j] = ss * Vr[(size_t)j * n + k]; vb += Vr[(size_t)j * n + k] * br[j]; }
        out->linearized_residuals[k] = sis * vb;
    }
    /* kept blocks with addr_shift (estimator.cpp:1483-1501 / :1548-1570) */
    int nk = 0;
    for (int i = first_kept; i < nb; i++) {
        if (!B[i].present || B[i].drop) continue;
        VrfPriorBlock *K = &out->blocks[nk++];
        K->kind = B[i].kind; K->size = B[i].gsize; K->idx = B[i].idx - m;
        memcpy(K->x0, B[i].x0, sizeof(double) * B[i].gsize);
        if (B[i].kind == VRF_BLK_EXPOSE || B[i].kind == VRF_BLK_TD) K->index = 0;
        else if (flag == VRF_MARGIN_OLD) K->index = B[i].index - 1;
        else K->index = (B[i].index == VRF_WINDOW_SIZE) ? VRF_WINDOW_SIZE - 1 : B[i].index;
    }
    out->n_blocks = nk;
    free(B); free(A); free(bv); free(Amm); free(Vm); free(wm); free(Ainv); free(T); free(Ar); free(br); free(Vr); free(wr); free(Ac);
    return 1;
}

/* ------------------------------------------------------------------ */
/* entry point: same contract as vrf_ba_solve (include/vrf_ba.h)        */
/* ------------------------------------------------------------------ */
int oracle_ba_solve(const VrfConfig *cfg, const VrfBaProblem *pb, VrfBaResult *res)
{
    const int M = pb->n_landmarks;
    Lin L;
    memset(&L, 0, sizeof(L));
    L.M = M; L.use_imu = pb->use_imu; L.nframes = pb->frame_count + 1;
    L.ex_active = !pb->ex_constant; L.pose0_const = !pb->use_imu;
    /* ESTIMATE_TD selects ProjectionTdFactor (estimator.cpp:1270); para_Td is only added (and possibly fixed) as a
       parameter block under USE_IMU (:1203-1212) -- without IMU Ceres adds it implicitly as a variable */
    L.td_factor = cfg->estimate_td != 0;
    L.td_active = L.td_factor && !(pb->use_imu && pb->td_constant);
    L.tr_over_row = cfg->tr / (double)cfg->row;
    int nfac = 0;
    for (int l = 0; l < M; l++) { int no = pb->lm_obs_ptr[l + 1] - pb->lm_obs_ptr[l]; if (no >= 2) nfac += no - 1; }
    L.nfac = nfac;
    L.f_lm = (int *)malloc(sizeof(int) * (nfac + 1)); L.f_i = (int *)malloc(sizeof(int) * (nfac + 1)); L.f_j = (int *)malloc(sizeof(int) * (nfac + 1));
    L.f_pi = (double *)malloc(sizeof(double) * 2 * (nfac + 1)); L.f_pj = (double *)malloc(sizeof(double) * 2 * (nfac + 1));
    L.f_r = (double *)malloc(sizeof(double) * 2 * (nfac + 1)); L.f_Ji = (double *)malloc(sizeof(double) * 12 * (nfac + 1));
    L.f_Jj = (double *)malloc(sizeof(double) * 12 * (nfac + 1)); L.f_Jex = (double *)malloc(sizeof(double) * 12 * (nfac + 1));
    L.f_Jl = (double *)malloc(sizeof(double) * 2 * (nfac + 1)); L.f_Jtd = (double *)calloc(2 * (nfac + 1), sizeof(double));
    L.f_vi = (double *)calloc(2 * (nfac + 1), sizeof(double)); L.f_vj = (double *)calloc(2 * (nfac + 1), sizeof(double));
    L.f_tdi = (double *)calloc(nfac + 1, sizeof(double)); L.f_tdj = (double *)calloc(nfac + 1, sizeof(double));
    L.f_rowi = (double *)calloc(nfac + 1, sizeof(double)); L.f_rowj = (double *)calloc(nfac + 1, sizeof(double));
    L.lm_const = (unsigned char *)malloc(M + 1); L.lm_ub = (double *)malloc(sizeof(double) * (M + 1));
    int k = 0;
    for (int l = 0; l < M; l++) {
        int o0 = pb->lm_obs_ptr[l], no = pb->lm_obs_ptr[l + 1] - o0, i = pb->lm_start_frame[l];
        L.lm_const[l] = (pb->lm_estimate_flag[l] == 1 && cfg->fix_depth) ? 1 : 0;
        L.lm_ub[l] = (pb->lm_estimate_flag[l] == 2) ? 2.0 / cfg->depth_max_dist : INFINITY;
        for (int t = 1; t < no; t++, k++) {
            L.f_lm[k] = l; L.f_i[k] = i; L.f_j[k] = i + t;
            L.f_pi[2 * k] = pb->obs_pts[2 * o0]; L.f_pi[2 * k + 1] = pb->obs_pts[2 * o0 + 1];
            L.f_pj[2 * k] = pb->obs_pts[2 * (o0 + t)]; L.f_pj[2 * k + 1] = pb->obs_pts[2 * (o0 + t) + 1];
            if (L.td_factor) {
                L.f_vi[2 * k] = pb->obs_velocity[2 * o0]; L.f_vi[2 * k + 1] = pb->obs_velocity[2 * o0 + 1];
                L.f_vj[2 * k] = pb->obs_velocity[2 * (o0 + t)]; L.f_vj[2 * k + 1] = pb->obs_velocity[2 * (o0 + t) + 1];
                L.f_tdi[k] = pb->obs_cur_td[o0]; L.f_tdj[k] = pb->obs_cur_td[o0 + t];
                L.f_rowi[k] = pb->obs_row[o0]; L.f_rowj[k] = pb->obs_row[o0 + t];
            }
        }
    }
    L.nimu = 0;
    if (pb->use_imu)
        for (int j = 1; j <= pb->frame_count; j++) {
            if (pb->imu[j - 1].sum_dt > 10.0) continue;          /* estimator.cpp:1231-1233 */
            L.imu_j[L.nimu] = j; L.imu_pre[L.nimu] = &pb->imu[j - 1]; L.nimu++;
        }
    L.np = pb->prior ? pb->prior->n : 0;
    L.pr_r = (double *)malloc(sizeof(double) * (L.np + 1));
    if (pb->prior)
        for (int b = 0; b < pb->prior->n_blocks; b++) {
            const VrfPriorBlock *B = &pb->prior->blocks[b];
            int col = B->kind == VRF_BLK_POSE ? COL_POSE(B->index) : B->kind == VRF_BLK_SPEEDBIAS ? COL_SB(B->index)
                      : B->kind == VRF_BLK_EXPOSE ? COL_EX : B->kind == VRF_BLK_TD ? COL_TD : -1;
            if (col >= 0 && !col_active(&L, col)) col = -1;
            L.pr_col[b] = col;
        }
    State x;
    memcpy(x.pose, pb->para_Pose, sizeof(x.pose)); memcpy(x.sb, pb->para_SpeedBias, sizeof(x.sb));
    memcpy(x.ex, pb->para_Ex_Pose, sizeof(x.ex)); x.td = pb->para_Td;
    x.lam = (double *)malloc(sizeof(double) * (M + 1));
    memcpy(x.lam, pb->para_Feature, sizeof(double) * M);
    SolveSummary sum;
    int rc = solve(pb, cfg, &L, &x, &sum);
    res->status = rc; res->iterations = sum.iterations; res->successful_steps = sum.successful;
    res->armijo_failures = sum.armijo_failures;
    res->termination = sum.termination; res->initial_cost = sum.initial_cost; res->final_cost = sum.final_cost;
    memcpy(res->para_Pose, x.pose, sizeof(x.pose)); memcpy(res->para_SpeedBias, x.sb, sizeof(x.sb));
    memcpy(res->para_Ex_Pose, x.ex, sizeof(x.ex)); res->para_Td = x.td;
    if (res->para_Feature) memcpy(res->para_Feature, x.lam, sizeof(double) * M);
    for (int i = 0; i < NF; i++)
        for (int c = 0; c < 7; c++) if (!isfinite(x.pose[i][c])) res->status = VRF_SOFT_NONFINITE;
    gauge_fix(pb, &x, res);
    res->has_new_prior = 0;
    if (pb->frame_count == VRF_WINDOW_SIZE && res->new_prior) {
        State xm; double *lamm = (double *)malloc(sizeof(double) * (M + 1));
        repack_state(pb, res, &x, &xm, lamm);
        res->has_new_prior = marginalize(pb, cfg, &xm, res->new_prior);
        free(lamm);
    }
    free(x.lam); free(L.f_lm); free(L.f_i); free(L.f_j); free(L.f_pi); free(L.f_pj); free(L.f_r); free(L.f_Ji); free(L.f_Jj);
    free(L.f_Jex); free(L.f_Jl); free(L.lm_const); free(L.lm_ub); free(L.pr_r);
    free(L.f_Jtd); free(L.f_vi); free(L.f_vj); free(L.f_tdi); free(L.f_tdj); free(L.f_rowi); free(L.f_rowj);
    return res->status;
}

/* single prior evaluation (tests) */
void oracle_prior_residual(const VrfPrior *P, const double *pose, const double *sb, const double *ex, double td, double *r)
{
    State x;
    memcpy(x.pose, pose, sizeof(x.pose)); memcpy(x.sb, sb, sizeof(x.sb)); memcpy(x.ex, ex, sizeof(x.ex)); x.td = td; x.lam = NULL;
    prior_residual(P, &x, r);
}
