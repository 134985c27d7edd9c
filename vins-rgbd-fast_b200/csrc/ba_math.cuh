// ba_math.cuh -- device functions shared by the back-end kernels: small FP64 linear algebra,
// ProjectionFactor / IMUFactor / MarginalizationFactor evaluation (see ba_kernels.cu header for
// the reference file:line map).
#pragma once
#include "ba_dev.cuh"

namespace vrf {

// --------------------------------------------------------------------------
// small device linear algebra
// --------------------------------------------------------------------------
__device__ __forceinline__ void d_q2R(const double *q, double *R)
{   // Eigen::Quaternion::toRotationMatrix, q = (x,y,z,w)
    double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ void d_qmul(const double *a, const double *b, double *o)
{
    double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
__device__ __forceinline__ void d_qinv(const double *a, double *o)
{
    double n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    o[0] = -a[0] / n2; o[1] = -a[1] / n2; o[2] = -a[2] / n2; o[3] = a[3] / n2;
}
__device__ __forceinline__ void d_qrot(const double *q, const double *v, double *o)
{
    double ux = 2 * (q[1] * v[2] - q[2] * v[1]), uy = 2 * (q[2] * v[0] - q[0] * v[2]), uz = 2 * (q[0] * v[1] - q[1] * v[0]);
    o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ void d_mv(const double *A, const double *v, double *o)
{
    o[0] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
    o[1] = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
    o[2] = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
}
__device__ __forceinline__ void d_mtv(const double *A, const double *v, double *o)
{   // A^T v
    o[0] = A[0] * v[0] + A[3] * v[1] + A[6] * v[2];
    o[1] = A[1] * v[0] + A[4] * v[1] + A[7] * v[2];
    o[2] = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
}
__device__ __forceinline__ void d_mm(const double *A, const double *B, double *C)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void d_mtm(const double *A, const double *B, double *C)
{   // A^T B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
__device__ __forceinline__ void d_skew(const double *v, double *S)
{
    S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
__device__ __forceinline__ void d_pose_plus(const double *x, const double *d, double *o)
{   // PoseLocalParameterization::Plus
    double dq[4] = {d[3] * 0.5, d[4] * 0.5, d[5] * 0.5, 1.0}, q[4];
    o[0] = x[0] + d[0]; o[1] = x[1] + d[1]; o[2] = x[2] + d[2];
    d_qmul(x + 3, dq, q);
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    o[3] = q[0] / n; o[4] = q[1] / n; o[5] = q[2] / n; o[6] = q[3] / n;
}

static __device__ void d_R2ypr(const double *R, double *ypr)
{   // Utility::R2ypr (utility.h:66-83), degrees
    const double n0 = R[0], n1 = R[3], n2 = R[6], o0 = R[1], o1 = R[4], a0 = R[2], a1 = R[5];
    const double y = atan2(n1, n0);
    const double pp = atan2(-n2, n0 * cos(y) + n1 * sin(y));
    const double r = atan2(a0 * sin(y) - a1 * cos(y), -o0 * sin(y) + o1 * cos(y));
    ypr[0] = y / 3.14159265358979323846 * 180.0; ypr[1] = pp / 3.14159265358979323846 * 180.0; ypr[2] = r / 3.14159265358979323846 * 180.0;
}
static __device__ void d_R2q(const double *m, double *q)
{   // Eigen::Quaterniond(Matrix3d)
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the 16-lane half of a warp the calling lane belongs to
__device__ __forceinline__ double half_sum_d(double v)
{
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum 16 per-lane values over the 32 lanes with 16 shuffles: after the four scatter stages lane L holds the
// total of value (L >> 1) over its 16-lane partner set, the last stage completes it (lanes 2m, 2m+1 agree).
__device__ __forceinline__ double reduce_scatter16(double (&v)[16], int lane)
{
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const int n2 = 8 >> st, mask = 16 >> st;
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int k = 0; k < n2; ++k) {
            const double keep = up ? v[k + n2] : v[k];
            const double send = up ? v[k] : v[k + n2];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// deterministic block sum (BA_THREADS threads), result broadcast to all threads
static __device__ double block_sum(double v, double *s_red)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) s_red[w] = v;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int i = 0; i < BA_THREADS / 32; ++i) t += s_red[i];
    return t;
}
__device__ __forceinline__ int pk(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }
// tangent column of entry k of a landmark coupling row (BA_WS entries: 66 pose columns, then ex-pose, then td)
__device__ __forceinline__ int wcol(int k) { return k < 66 ? k : k + (BA_COL_EX - 66); }
// observation o of a problem, time-shifted for ProjectionTdFactor:
// pts_td = pts - (td - td_obs + TR / ROW * row) * velocity (projection_td_factor.cpp:52-53)
__device__ __forceinline__ void obs_at(const BaMeta &m, const BaProbDev &p, int o, double td, double &x, double &y)
{
    x = p.obs[2 * o]; y = p.obs[2 * o + 1];
    if (m.td_factor) {
        const double s = td - p.obs_td[o] + m.tr_over_row * p.obs_row[o];
        x -= s * p.obs_vel[2 * o]; y -= s * p.obs_vel[2 * o + 1];
    }
}

// --------------------------------------------------------------------------
// ProjectionFactor::Evaluate with precomputed rotations; CauchyLoss + Corrector applied.
// Returns rho0 (cost contribution is 0.5*rho0).  J blocks are 2x6 (local), Jl 2x1.
// --------------------------------------------------------------------------
struct FrameRot { double R[9]; };

__device__ __forceinline__ double proj_eval(const double *Pi, const double *Ri, const double *Pj, const double *Rj,
                                            const double *tic, const double *ric, double lam, double xi, double yi,
                                            double xj, double yj, bool want_jac, bool lm_const, double *r, double *Ji,
                                            double *Jj, double *Jl, double *Jex = nullptr,
                                            const double *vel_i = nullptr, const double *vel_j = nullptr, double *Jtd = nullptr)
{   // ProjectionTdFactor (projection_td_factor.cpp:34-150): the caller passes the time-shifted points (:52-53);
    // Jtd (:139-144) additionally needs both velocities.
    const double sqrt_info = 460.0 / 1.5;
    const double ilam = 1.0 / lam;
    double pci[3] = {xi * ilam, yi * ilam, ilam};
    double t[3], pii[3], pw[3], d[3], pij[3], e[3], pcj[3];
    d_mv(ric, pci, t);
    pii[0] = t[0] + tic[0]; pii[1] = t[1] + tic[1]; pii[2] = t[2] + tic[2];
    d_mv(Ri, pii, t);
    pw[0] = t[0] + Pi[0]; pw[1] = t[1] + Pi[1]; pw[2] = t[2] + Pi[2];
    d[0] = pw[0] - Pj[0]; d[1] = pw[1] - Pj[1]; d[2] = pw[2] - Pj[2];
    d_mtv(Rj, d, pij);
    e[0] = pij[0] - tic[0]; e[1] = pij[1] - tic[1]; e[2] = pij[2] - tic[2];
    d_mtv(ric, e, pcj);
    const double dep = pcj[2];
    const double idep = 1.0 / dep;
    r[0] = sqrt_info * (pcj[0] * idep - xj);
    r[1] = sqrt_info * (pcj[1] * idep - yj);
    const double s = r[0] * r[0] + r[1] * r[1];
    // CauchyLoss(1): rho' = 1/(1+s), rho'' < 0 => Corrector scales residual and Jacobian by sqrt(rho')
    const double sum = 1.0 + s, inv = 1.0 / sum;
    const double rho0 = log(sum);
    const double sr = sqrt(inv > 2.2250738585072014e-308 ? inv : 2.2250738585072014e-308);
    if (want_jac) {
        double red[6] = {sqrt_info * idep, 0, -sqrt_info * pcj[0] * idep * idep, 0, sqrt_info * idep, -sqrt_info * pcj[1] * idep * idep};
        double A[9], B[9], C[9], S[9], Mx[9];
        // A = ric^T Rj^T
        double RjT[9] = {Rj[0], Rj[3], Rj[6], Rj[1], Rj[4], Rj[7], Rj[2], Rj[5], Rj[8]};
        d_mtm(ric, RjT, A);
        d_mm(A, Ri, B);                 // ric^T Rj^T Ri
        d_skew(pii, S); d_mm(B, S, C);
        // jaco_i = [A, -C]
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Ji[rr * 6 + c] = sr * (red[rr * 3] * A[c] + red[rr * 3 + 1] * A[3 + c] + red[rr * 3 + 2] * A[6 + c]);
                Ji[rr * 6 + 3 + c] = -sr * (red[rr * 3] * C[c] + red[rr * 3 + 1] * C[3 + c] + red[rr * 3 + 2] * C[6 + c]);
            }
        d_skew(pij, S); d_mtm(ric, S, Mx);     // ric^T skew(p_imu_j)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                Jj[rr * 6 + c] = -sr * (red[rr * 3] * A[c] + red[rr * 3 + 1] * A[3 + c] + red[rr * 3 + 2] * A[6 + c]);
                Jj[rr * 6 + 3 + c] = sr * (red[rr * 3] * Mx[c] + red[rr * 3 + 1] * Mx[3 + c] + red[rr * 3 + 2] * Mx[6 + c]);
            }
        if (!lm_const) {
            double tr[9], v[3], pts_i[3] = {xi, yi, 1.0};
            d_mm(B, ric, tr);
            d_mv(tr, pts_i, v);
            const double k = -ilam * ilam;
            Jl[0] = sr * (red[0] * v[0] + red[1] * v[1] + red[2] * v[2]) * k;
            Jl[1] = sr * (red[3] * v[0] + red[4] * v[1] + red[5] * v[2]) * k;
        } else { Jl[0] = 0; Jl[1] = 0; }
        if (Jex) {
            // jaco_ex = [ric^T (Rj^T Ri - I), -tmp_r skew(pc_i) + skew(tmp_r pc_i) + skew(ric^T (Rj^T (Ri tic + Pi - Pj) - tic))]
            double RjTRi[9], Mm[9], L3[9], tr[9], T1[9], v[3], S1[9], S2[9], u[3], w[3];
            d_mm(RjT, Ri, RjTRi);
#pragma unroll
            for (int k = 0; k < 9; ++k) Mm[k] = RjTRi[k] - ((k % 4 == 0) ? 1.0 : 0.0);
            d_mtm(ric, Mm, L3);
            d_mm(B, ric, tr);
            d_skew(pci, S); d_mm(tr, S, T1);
            d_mv(tr, pci, v); d_skew(v, S1);
            d_mv(Ri, tic, u);
            u[0] += Pi[0] - Pj[0]; u[1] += Pi[1] - Pj[1]; u[2] += Pi[2] - Pj[2];
            d_mtv(Rj, u, w);
            w[0] -= tic[0]; w[1] -= tic[1]; w[2] -= tic[2];
            d_mtv(ric, w, v); d_skew(v, S2);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    Jex[rr * 6 + c] = sr * (red[rr * 3] * L3[c] + red[rr * 3 + 1] * L3[3 + c] + red[rr * 3 + 2] * L3[6 + c]);
                    double e0 = -T1[c] + S1[c] + S2[c], e1 = -T1[3 + c] + S1[3 + c] + S2[3 + c], e2 = -T1[6 + c] + S1[6 + c] + S2[6 + c];
                    Jex[rr * 6 + 3 + c] = sr * (red[rr * 3] * e0 + red[rr * 3 + 1] * e1 + red[rr * 3 + 2] * e2);
                }
        }
        if (Jtd) {
            // reduce * ric^T Rj^T Ri ric * velocity_i / inv_dep_i * -1 + sqrt_info * velocity_j.head(2)
            double tr[9], v[3], vel[3] = {vel_i[0], vel_i[1], 0.0};
            d_mm(B, ric, tr);
            d_mv(tr, vel, v);
            Jtd[0] = sr * ((red[0] * v[0] + red[1] * v[1] + red[2] * v[2]) * ilam * -1.0 + sqrt_info * vel_j[0]);
            Jtd[1] = sr * ((red[3] * v[0] + red[4] * v[1] + red[5] * v[2]) * ilam * -1.0 + sqrt_info * vel_j[1]);
        }
    }
    r[0] *= sr; r[1] *= sr;
    return rho0;
}

// --------------------------------------------------------------------------
// IMU factor pieces.  sqrt_info = LLT(cov^-1).L^T is constant during a solve (the
// reference recomputes it in every Evaluate, imu_factor.h:66-69): computed once here.
// One warp per factor; S (15x15 upper) row-major in global scratch.
// --------------------------------------------------------------------------
static __device__ void imu_sqrt_info_warp(const double *cov, double *Sout, double *scr /* 2*225 doubles, per warp */)
{
    const int lane = threadIdx.x & 31;
    double *L = scr, *Inv = scr + 225;
    for (int i = lane; i < 225; i += 32) L[i] = cov[i];
    __syncwarp();
    // Cholesky (lower) of cov, column by column
    for (int j = 0; j < 15; ++j) {
        if (lane == 0) {
            double s = L[j * 15 + j];
            for (int k = 0; k < j; ++k) s -= L[j * 15 + k] * L[j * 15 + k];
            L[j * 15 + j] = sqrt(s);
        }
        __syncwarp();
        double ljj = L[j * 15 + j];
        int i = j + 1 + lane;
        if (i < 15) {
            double t = L[i * 15 + j];
            for (int k = 0; k < j; ++k) t -= L[i * 15 + k] * L[j * 15 + k];
            L[i * 15 + j] = t / ljj;
        }
        __syncwarp();
    }
    // inverse: lane c solves L L^T x = e_c
    if (lane < 15) {
        double x[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) x[i] = (i == lane) ? 1.0 : 0.0;
        for (int i = 0; i < 15; ++i) { double s = x[i]; for (int k = 0; k < i; ++k) s -= L[i * 15 + k] * x[k]; x[i] = s / L[i * 15 + i]; }
        for (int i = 14; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < 15; ++k) s -= L[k * 15 + i] * x[k]; x[i] = s / L[i * 15 + i]; }
        for (int i = 0; i < 15; ++i) Inv[i * 15 + lane] = x[i];
    }
    __syncwarp();
    // symmetrise, then Cholesky of the inverse
    for (int i = lane; i < 225; i += 32) { int r = i / 15, c = i - r * 15; L[i] = 0.5 * (Inv[r * 15 + c] + Inv[c * 15 + r]); }
    __syncwarp();
    for (int j = 0; j < 15; ++j) {
        if (lane == 0) {
            double s = L[j * 15 + j];
            for (int k = 0; k < j; ++k) s -= L[j * 15 + k] * L[j * 15 + k];
            L[j * 15 + j] = sqrt(s);
        }
        __syncwarp();
        double ljj = L[j * 15 + j];
        int i = j + 1 + lane;
        if (i < 15) {
            double t = L[i * 15 + j];
            for (int k = 0; k < j; ++k) t -= L[i * 15 + k] * L[j * 15 + k];
            L[i * 15 + j] = t / ljj;
        }
        __syncwarp();
    }
    for (int i = lane; i < 225; i += 32) { int r = i / 15, c = i - r * 15; Sout[i] = (c >= r) ? L[c * 15 + r] : 0.0; }
    __syncwarp();
}

// raw (un-whitened) residual, IntegrationBase::evaluate (integration_base.h:164-195)
struct ImuCtx {
    double Qi_inv[4], cdq[4], cdq_inv[4];
    double RiT[9];
    double sum_dt;
};
static __device__ void imu_residual_raw(const VrfImuPreint *pre, const double *pi, const double *sbi, const double *pj,
                                 const double *sbj, double g_norm, double *r, ImuCtx *cx)
{
    const double *J = pre->jacobian;
    const double dt = pre->sum_dt;
    double dba[3], dbg[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { dba[k] = sbi[3 + k] - pre->linearized_ba[k]; dbg[k] = sbi[6 + k] - pre->linearized_bg[k]; }
    double th[3], cdp[3], cdv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        th[i] = J[(3 + i) * 15 + 12] * dbg[0] + J[(3 + i) * 15 + 13] * dbg[1] + J[(3 + i) * 15 + 14] * dbg[2];
        cdp[i] = pre->delta_p[i] + J[i * 15 + 9] * dba[0] + J[i * 15 + 10] * dba[1] + J[i * 15 + 11] * dba[2] +
                 J[i * 15 + 12] * dbg[0] + J[i * 15 + 13] * dbg[1] + J[i * 15 + 14] * dbg[2];
        cdv[i] = pre->delta_v[i] + J[(6 + i) * 15 + 9] * dba[0] + J[(6 + i) * 15 + 10] * dba[1] + J[(6 + i) * 15 + 11] * dba[2] +
                 J[(6 + i) * 15 + 12] * dbg[0] + J[(6 + i) * 15 + 13] * dbg[1] + J[(6 + i) * 15 + 14] * dbg[2];
    }
    double dq[4] = {th[0] * 0.5, th[1] * 0.5, th[2] * 0.5, 1.0};
    d_qmul(pre->delta_q, dq, cx->cdq);
    d_qinv(pi + 3, cx->Qi_inv);
    d_q2R(cx->Qi_inv, cx->RiT);
    cx->sum_dt = dt;
    double v[3] = {pj[0] - pi[0] - sbi[0] * dt, pj[1] - pi[1] - sbi[1] * dt, 0.5 * g_norm * dt * dt + pj[2] - pi[2] - sbi[2] * dt}, w[3];
    d_qrot(cx->Qi_inv, v, w);
    r[0] = w[0] - cdp[0]; r[1] = w[1] - cdp[1]; r[2] = w[2] - cdp[2];
    double qij[4], qe[4];
    d_qinv(cx->cdq, cx->cdq_inv);
    d_qmul(cx->Qi_inv, pj + 3, qij);
    d_qmul(cx->cdq_inv, qij, qe);
    r[3] = 2 * qe[0]; r[4] = 2 * qe[1]; r[5] = 2 * qe[2];
    double v2[3] = {sbj[0] - sbi[0], sbj[1] - sbi[1], g_norm * dt + sbj[2] - sbi[2]};
    d_qrot(cx->Qi_inv, v2, w);
    r[6] = w[0] - cdv[0]; r[7] = w[1] - cdv[1]; r[8] = w[2] - cdv[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) { r[9 + k] = sbj[3 + k] - sbi[3 + k]; r[12 + k] = sbj[6 + k] - sbi[6 + k]; }
}

// column `c` (0..29: pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9) of the raw 15x30 local Jacobian
static __device__ void imu_jac_col(const VrfImuPreint *pre, const double *pi, const double *sbi, const double *pj,
                            const double *sbj, double g_norm, const ImuCtx *cx, int c, double *col)
{
#pragma unroll
    for (int k = 0; k < 15; ++k) col[k] = 0.0;
    const double dt = cx->sum_dt;
    const double *J = pre->jacobian;
    if (c < 3) {                     // d/dPi : O_P rows = -Ri^T
        for (int k = 0; k < 3; ++k) col[k] = -cx->RiT[k * 3 + c];
    } else if (c < 6) {              // d/dtheta_i
        const int cc = c - 3;
        double v[3] = {pj[0] - pi[0] - sbi[0] * dt, pj[1] - pi[1] - sbi[1] * dt, 0.5 * g_norm * dt * dt + pj[2] - pi[2] - sbi[2] * dt}, w[3], S[9];
        d_qrot(cx->Qi_inv, v, w); d_skew(w, S);
        for (int k = 0; k < 3; ++k) col[k] = S[k * 3 + cc];
        // -(Qleft(Qj^-1 Qi) * Qright(corrected_delta_q)).bottomRightCorner<3,3>()
        double qji[4], Qj_inv[4];
        d_qinv(pj + 3, Qj_inv); d_qmul(Qj_inv, pi + 3, qji);
        // Qleft(q) rows 1..3 (w,x,y,z order): [q_k | w I + skew(vec)] ; Qright(p) cols 1..3: [-vec^T ; w I - skew(vec)]
        const double *p = cx->cdq;
        double SL[9], SR[9];
        d_skew(qji, SL); d_skew(p, SR);
        for (int k = 0; k < 3; ++k) {
            // (QL row k+1) . (QR col cc+1)
            double a = qji[k] * (-p[cc]);
            for (int m = 0; m < 3; ++m) {
                double ql = (k == m ? qji[3] : 0.0) + SL[k * 3 + m];
                double qr = (m == cc ? p[3] : 0.0) - SR[m * 3 + cc];
                a += ql * qr;
            }
            col[3 + k] = -a;
        }
        double v2[3] = {sbj[0] - sbi[0], sbj[1] - sbi[1], g_norm * dt + sbj[2] - sbi[2]};
        d_qrot(cx->Qi_inv, v2, w); d_skew(w, S);
        for (int k = 0; k < 3; ++k) col[6 + k] = S[k * 3 + cc];
    } else if (c < 15) {             // speed-bias i
        const int cc = c - 6;
        if (cc < 3) {
            for (int k = 0; k < 3; ++k) { col[k] = -cx->RiT[k * 3 + cc] * dt; col[6 + k] = -cx->RiT[k * 3 + cc]; }
        } else if (cc < 6) {
            const int b = cc - 3;
            for (int k = 0; k < 3; ++k) { col[k] = -J[k * 15 + 9 + b]; col[6 + k] = -J[(6 + k) * 15 + 9 + b]; }
            col[9 + b] = -1.0;
        } else {
            const int b = cc - 6;
            for (int k = 0; k < 3; ++k) { col[k] = -J[k * 15 + 12 + b]; col[6 + k] = -J[(6 + k) * 15 + 12 + b]; }
            // -Qleft(Qj^-1 Qi delta_q).bottomRight * dq_dbg
            double Qj_inv[4], q1[4], q2[4], S[9];
            d_qinv(pj + 3, Qj_inv); d_qmul(Qj_inv, pi + 3, q1); d_qmul(q1, pre->delta_q, q2);
            d_skew(q2, S);
            for (int k = 0; k < 3; ++k) {
                double a = 0;
                for (int m = 0; m < 3; ++m) a += ((k == m ? q2[3] : 0.0) + S[k * 3 + m]) * J[(3 + m) * 15 + 12 + b];
                col[3 + k] = -a;
            }
            col[12 + b] = -1.0;
        }
    } else if (c < 18) {             // d/dPj
        const int cc = c - 15;
        for (int k = 0; k < 3; ++k) col[k] = cx->RiT[k * 3 + cc];
    } else if (c < 21) {             // d/dtheta_j: Qleft(cdq^-1 Qi^-1 Qj).bottomRight
        const int cc = c - 18;
        double q1[4], q2[4], S[9];
        d_qmul(cx->cdq_inv, cx->Qi_inv, q1); d_qmul(q1, pj + 3, q2);
        d_skew(q2, S);
        for (int k = 0; k < 3; ++k) col[3 + k] = (k == cc ? q2[3] : 0.0) + S[k * 3 + cc];
    } else {                         // speed-bias j
        const int cc = c - 21;
        if (cc < 3) { for (int k = 0; k < 3; ++k) col[6 + k] = cx->RiT[k * 3 + cc]; }
        else if (cc < 6) col[9 + cc - 3] = 1.0;
        else col[12 + cc - 6] = 1.0;
    }
}

// --------------------------------------------------------------------------
// prior residual: dx then r = r0 + J0 dx (marginalization_factor.cpp:353-400)
// --------------------------------------------------------------------------
static __device__ void prior_dx(const BaPriorStore *P, const double *pose, const double *sb, const double *ex, double td, double *dx)
{
    for (int b = threadIdx.x; b < P->n_blocks; b += blockDim.x) {
        const int kind = P->kind[b], index = P->index[b], size = P->size[b], idx = P->idx[b];
        const double *cur = kind == VRF_BLK_POSE ? pose + 7 * index : kind == VRF_BLK_SPEEDBIAS ? sb + 9 * index
                            : kind == VRF_BLK_TD ? &td : ex;
        const double *x0 = P->x0 + 9 * b;   // keep_block_data
        if (size != 7) {
            for (int k = 0; k < size; ++k) dx[idx + k] = cur[k] - x0[k];
        } else {
            for (int k = 0; k < 3; ++k) dx[idx + k] = cur[k] - x0[k];
            double q0i[4], qe[4];
            d_qinv(x0 + 3, q0i); d_qmul(q0i, cur + 3, qe);
            const double sg = (qe[3] >= 0) ? 1.0 : -1.0;
            for (int k = 0; k < 3; ++k) dx[idx + 3 + k] = sg * 2.0 * qe[k];
        }
    }
}

}  // namespace vrf
