// ba_kernels.cu -- sliding-window visual-inertial bundle adjustment on the GPU
// (Estimator::optimization, reference vins_estimator/src/estimator/estimator.cpp:1161-1578).
//
// One CTA (512 threads) per sequence runs the WHOLE solve on device -- linearisation,
// Jacobi scaling, per-landmark Schur reduction, dense Cholesky of the reduced camera
// system, traditional dogleg with trust-region control -- with no host round trip
// between iterations.  All arithmetic is FP64 (the reference/Ceres use double).
//
//   factor math            : ProjectionFactor::Evaluate (factor/projection_factor.cpp:22-130),
//                            IMUFactor::Evaluate (factor/imu_factor.h:20-205),
//                            MarginalizationFactor::Evaluate (factor/marginalization_factor.cpp:353-415),
//                            CauchyLoss(1.0) + Ceres Corrector (restated at marginalization_factor.cpp:39-72)
//   parameterisation       : PoseLocalParameterization (factor/pose_local_parameterization.cpp:3-28)
//   solver (third party)   : ceres::Solve with DENSE_SCHUR + DOGLEG, max_num_iterations = NUM_ITERATIONS
//                            (estimator.cpp:1348-1363); algorithm restated in oracle/ba_ref.c (parity
//                            UNPINNED against real Ceres)
//
// Data layout (HBM, per problem, SoA over landmarks so that lanes of a warp read
// consecutive words): lam[M], start[M], flag[M], obs_ptr[M+1], obs[O][2]; W[M][73]
// (landmark-to-pose / ex-pose / td coupling rows), hll/gl/diag/gd/gn/step per landmark.
// On chip: the 172x172 camera system lives in shared memory as a packed lower
// triangle (119 KB) and is Schur-reduced and Cholesky-factorised in place; the
// tangent layout is [pose f: 6f | speed-bias f: 66+9f | ex-pose: 165 | td: 171].
// The marginalization prior is consumed in information form (ba_dev.cuh, BaPriorStore).
//
// Roofline: ~100-160 KB touched and ~5 MFLOP FP64 per iteration per problem => FP64 dependency
// chains and shared-memory bandwidth bound; the only
// dense contraction is the 172^3/3 Cholesky.
#include "ba_math.cuh"
#include "ba_linesearch.cuh"

namespace vrf {

// --------------------------------------------------------------------------
// shared-memory frame of the solve kernel
// --------------------------------------------------------------------------
struct BaShared {
    double H[BA_NC * (BA_NC + 1) / 2];     // packed lower triangle: camera system, Schur-reduced & factorised in place
    double g[BA_NC], diag[BA_NC], gd[BA_NC], gn[BA_NC], jscale[BA_NC], tmp[BA_NC], y[BA_NC + 1], colv[BA_NC + 1];
    double Lp[8][BA_NC + 8];                 // current Cholesky panel, transposed: Lp[c][row]; row stride 180 = 4 mod 16 doubles, so that the
                                             // DMMA fragment loads of the trailing update (lane -> [lane & 3][row0 + (lane >> 2)]) are conflict-free
    double pose[BA_NF * 7], sb[BA_NF * 9], ex[7];
    double cpose[BA_NF * 7], csb[BA_NF * 9], cex[7];  // candidate
    double tdv[2];                                    // para_Td: current, candidate
    double R[BA_NF * 9], ric[9];
    double dx[VRF_PRIOR_MAX_DIM], pr[VRF_PRIOR_MAX_DIM];
    // prior block table / linearisation point / r0 / column map staged once per solve (L2 latency is ~800 cycles here)
    double px0[VRF_PRIOR_MAX_BLOCKS * 9], pr0[VRF_PRIOR_MAX_DIM];
    int pkind[VRF_PRIOR_MAX_BLOCKS], pindex[VRF_PRIOR_MAX_BLOCKS], psize[VRF_PRIOR_MAX_BLOCKS], pidx[VRF_PRIOR_MAX_BLOCKS];
    int pcol[VRF_PRIOR_MAX_DIM], pnb;
    double pc0;                                       // constant term of the prior (1/2 r0^T r0)
    double imuJ[BA_NF - 1][15 * 30];                 // whitened IMU Jacobians of the current linearisation
    double imur[BA_NF - 1][16];
    double red[BA_THREADS / 32];
    BaMeta meta;                                      // this CTA's problem descriptors (see k_ba_solve)
    BaProbDev prob;
    int pair_ptr[BA_NF * (BA_NF - 1) / 2 + 1];        // offsets of the frame pairs (host i, observer j) in BaProbDev::fac
    int pair_cnt[BA_NF * (BA_NF - 1) / 2 + 1];
    double d8[36];                                    // chol_diag8: the 8x8 diagonal block being factorised (packed lower triangle)
    int imu_ord[BA_NF], imu_neven;                    // IMU factors ordered [even j ... | odd j ...] for the two assembly passes
    double sc[16];
    int flag[8];
};

// Keep the frame inside the 196 KB shared-memory carve-out (1 KB is reserved per CTA): one step more (228 KB) leaves only
// 28 KB of L1 for the landmark arrays / coupling rows in HBM and slowed every phase that touches them by 20-60 %.
static_assert(sizeof(BaShared) <= 196 * 1024 - 1024, "BaShared must fit the 196 KB carve-out");

__device__ __forceinline__ bool col_active_dev(const BaMeta &m, int col)
{
    if (col < 66) { int f = col / 6; return f < m.nframes && !(f == 0 && !m.use_imu); }
    if (col < BA_COL_EX) { int f = (col - 66) / 9; return m.use_imu && f < m.nframes; }
    return col == BA_COL_TD ? m.td_active != 0 : m.ex_active != 0;
}

// Contributions of one batch of projection factors of a frame pair to the rows of the "global" block
// g = [ex-pose (6) | td (1)] (tangent columns 165..171), only when one of them is variable: Jg^T Ji, Jg^T Jj, Jg^T r
// (13 sums per row q) and the lower triangle of Jg^T Jg (28 sums), each summed over the warp's lanes with the
// reduce-scatter butterfly and accumulated over the pair's batches in accg[0..6] / accg[7..8] (lane 2e holds entry e).
// Jg: 2 x 7 row-major.
__device__ __noinline__ void accumulate_g(double *accg, int lane, const double *Ji, const double *Jj, const double *Jg, const double *r)
{
#pragma unroll 1
    for (int q = 0; q < 7; ++q) {
        double v[16];
        const double g0 = Jg[q], g1 = Jg[7 + q];
#pragma unroll
        for (int c = 0; c < 6; ++c) { v[c] = g0 * Ji[c] + g1 * Ji[6 + c]; v[6 + c] = g0 * Jj[c] + g1 * Jj[6 + c]; }
        v[12] = g0 * r[0] + g1 * r[1]; v[13] = 0; v[14] = 0; v[15] = 0;
        accg[q] += reduce_scatter16(v, lane);
    }
#pragma unroll 1
    for (int rd = 0; rd < 2; ++rd) {
        double v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int e = rd * 16 + t;
            int a = 0;
            while ((a + 1) * (a + 2) / 2 <= e) ++a;
            const int b = e - a * (a + 1) / 2;
            v[t] = (e < 28) ? Jg[a] * Jg[b] + Jg[7 + a] * Jg[7 + b] : 0.0;
        }
        accg[7 + rd] += reduce_scatter16(v, lane);
    }
}

#define BA_NPAIR (BA_NF * (BA_NF - 1) / 2)

// entry e of the 12x12 normal-equation block of one projection factor (host Jacobian Ji, observer Jacobian Jj,
// 2x6 row-major each): 0..35 Jj^T Ji, 36..56 lower triangle of Jj^T Jj, 57..77 of Ji^T Ji, 78..83 Jj^T r, 84..89 Ji^T r.
// Called with compile-time e (fully unrolled loops) so that the register arrays are indexed statically.
__device__ __forceinline__ double pair_entry(int e, const double *Ji, const double *Jj, const double *r)
{
    if (e < 36) { const int a = e / 6, b = e % 6; return Jj[a] * Ji[b] + Jj[6 + a] * Ji[6 + b]; }
    if (e < 78) {
        const double *J = e < 57 ? Jj : Ji;
        const int t = e < 57 ? e - 36 : e - 57;
        int a = 0;
        while ((a + 1) * (a + 2) / 2 <= t) ++a;
        const int b = t - a * (a + 1) / 2;
        return J[a] * J[b] + J[6 + a] * J[6 + b];
    }
    if (e < 84) { const int a = e - 78; return Jj[a] * r[0] + Jj[6 + a] * r[1]; }
    if (e < 90) { const int a = e - 84; return Ji[a] * r[0] + Ji[6 + a] * r[1]; }
    return 0.0;
}

// prior_dx (ba_math.cuh) on the block table staged in shared memory
__device__ __forceinline__ void prior_dx_smem(BaShared &sh, const double *pose, const double *sb, const double *ex, double td)
{
    for (int b = threadIdx.x; b < sh.pnb; b += BA_THREADS) {
        const int kind = sh.pkind[b], index = sh.pindex[b], size = sh.psize[b], idx = sh.pidx[b];
        const double *cur = kind == VRF_BLK_POSE ? pose + 7 * index : kind == VRF_BLK_SPEEDBIAS ? sb + 9 * index
                            : kind == VRF_BLK_TD ? &td : ex;
        const double *x0 = sh.px0 + 9 * b;   // keep_block_data
        if (size != 7) {
            for (int k = 0; k < size; ++k) sh.dx[idx + k] = cur[k] - x0[k];
        } else {
            for (int k = 0; k < 3; ++k) sh.dx[idx + k] = cur[k] - x0[k];
            double q0i[4], qe[4];
            d_qinv(x0 + 3, q0i); d_qmul(q0i, cur + 3, qe);
            const double sg = (qe[3] >= 0) ? 1.0 : -1.0;
            for (int k = 0; k < 3; ++k) sh.dx[idx + 3 + k] = sg * 2.0 * qe[k];
        }
    }
}

// index of the frame pair (host i, observer j > i) in the pair-major order
__device__ __forceinline__ int pair_index(int i, int j) { return i * (2 * BA_NF - i - 1) / 2 + (j - i - 1); }

// cost of all residual blocks at (pose, sb, lam); optionally the full linearisation into sh.H / sh.g / landmark arrays.
// __noinline__: the solve loop calls this from five sites; inlining produced a 51k-instruction kernel (800 KB of SASS)
// that thrashed the instruction cache (one resident CTA per SM, 16 warps in different code regions).
// GACT: ex-pose and/or td variable (two instantiations so that the common constant-extrinsic path keeps its registers).
//
// Bit-reproducible: the residual blocks are dealt to the warps by a dynamic queue, but no sum depends on which warp ran
// which task or when -- a task only produces partials that it alone owns (whitened IMU Jacobians in shared memory,
// per-pair and per-factor sums in the L2-resident scratch of ba_dev.cuh, its own cost), and the assembly pass adds them
// with one owner thread per destination in a fixed order.  No floating-point atomics.
template <bool GACT, bool DIRDER>
__device__ __noinline__ double ba_evaluate_t(const BaMeta &m, const BaProbDev &p, BaShared &sh, const double *pose, const double *sb,
                                const double *ex, const double td, const double *lam, const int mode, double *dphi_out = nullptr)
{
    // mode 0: cost only; 1: cost + full linearisation (H, g, W, hll, gl); 2: cost + directional derivative delta . gradient
    // for the trial points of the projected line search -- delta in sh.colv (camera columns) and p.y_l (landmarks), every
    // factor contributes r^T (J delta): no assembly, the linearisation at x is left alone
    // (DIRDER is a separate instantiation: the trial-point code stays out of the registers of the iteration's hot path)
    const bool lin = !DIRDER && mode == 1;
    constexpr bool dirder = DIRDER;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    if (lin) {
        for (int i = tid; i < BA_NC * (BA_NC + 1) / 2; i += BA_THREADS) sh.H[i] = 0.0;
        for (int i = tid; i < BA_NC; i += BA_THREADS) sh.g[i] = 0.0;
        // (the coupling rows W are zeroed once per solve: a linearisation rewrites the same entries every time)
    }
    for (int f = tid; f < BA_NF; f += BA_THREADS) d_q2R(pose + 7 * f + 3, sh.R + 9 * f);
    if (tid == 0) { d_q2R(ex + 3, sh.ric); sh.flag[2] = 0; }
    constexpr bool gact = GACT;
    const bool eprof = (m.debug & 16) != 0;
    long long et[5] = {0, 0, 0, 0, 0}, ec = eprof ? clock64() : 0;
#define EPROF(k) do { if (eprof) { long long t_ = clock64(); et[k] += t_ - ec; ec = t_; } } while (0)
    __syncthreads();
    EPROF(0);
    // ---- residual blocks as a dynamic task queue over the 16 warps: the IMU factors first (the longest tasks), then the
    // projection factors (linearisation: one task per frame pair; cost only: one task per 32 factors) ----
    const int nproj = lin ? BA_NPAIR : (sh.pair_ptr[BA_NPAIR] + 31) >> 5;
    const int ntask = m.nimu + nproj;
    for (;;) {
        int task = 0;
        if (lane == 0) task = atomicAdd(&sh.flag[2], 1);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntask) break;
        double tcost = 0.0, tdphi = 0.0;
        if (task < m.nimu) {
            // ---- IMU factor: one warp per factor; whitened residual and Jacobian go to shared memory, J^T J is formed by
            //      the assembly pass below ----
            const int f = task;
            const int j = m.imu_j[f], i = j - 1;
            const VrfImuPreint *pre = p.imu + (j - 1);
            const double *S = p.imuS + (size_t)(j - 1) * 225;
            double rr[15];
            ImuCtx cx;
            imu_residual_raw(pre, pose + 7 * i, sb + 9 * i, pose + 7 * j, sb + 9 * j, m.g_norm, rr, &cx);
            // whitened residual row `lane`
            double rw = 0;
            if (lane < 15) { for (int k = lane; k < 15; ++k) rw += S[lane * 15 + k] * rr[k]; tcost = 0.5 * rw * rw; }
            if (lin) {
                if (lane < 15) sh.imur[f][lane] = rw;
                if (lane < 30) {
                    double col[15];
                    imu_jac_col(pre, pose + 7 * i, sb + 9 * i, pose + 7 * j, sb + 9 * j, m.g_norm, &cx, lane, col);
                    for (int rI = 0; rI < 15; ++rI) { double a = 0; for (int k = rI; k < 15; ++k) a += S[rI * 15 + k] * col[k]; sh.imuJ[f][rI * 30 + lane] = a; }
                }
            } else if (dirder) {
                // delta . J^T r: lane c holds column c of the whitened Jacobian, the whitened residual rows come by shuffle
                double col[15];
#pragma unroll
                for (int k = 0; k < 15; ++k) col[k] = 0.0;
                if (lane < 30) imu_jac_col(pre, pose + 7 * i, sb + 9 * i, pose + 7 * j, sb + 9 * j, m.g_norm, &cx, lane, col);
                double gcol = 0;
                for (int rI = 0; rI < 15; ++rI) {
                    double a = 0;
                    for (int k = rI; k < 15; ++k) a += S[rI * 15 + k] * col[k];
                    gcol += a * __shfl_sync(0xffffffffu, rw, rI);
                }
                const int c = lane;
                const int tc = c < 6 ? 6 * i + c : c < 15 ? 66 + 9 * i + (c - 6) : c < 21 ? 6 * j + (c - 15) : 66 + 9 * j + (c - 21);
                if (lane < 30) tdphi = gcol * sh.colv[tc];
            }
        } else if (!lin) {
            // ---- projection factors, cost only: 32 consecutive factors of the pair-major list per warp, one lane per factor ----
            const int f = 32 * (task - m.nimu) + lane;
            if (f < sh.pair_ptr[BA_NPAIR]) {
                const int code = p.fac[f];
                const int l = code & 0xFFFF, j = code >> 16, i = p.start[l];
                const int o0 = p.obs_ptr[l];
                double r[2], Ji[12], Jj[12], Jl[2], xi, yi, xj, yj;
                obs_at(m, p, o0, td, xi, yi);
                obs_at(m, p, o0 + (j - i), td, xj, yj);
                if (!dirder)
                    tcost = 0.5 * proj_eval(pose + 7 * i, sh.R + 9 * i, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], xi, yi, xj, yj, false,
                                            p.lm_const[l] != 0, r, Ji, Jj, Jl);
                else {
                    // r^T (J delta) of this factor (corrected residual and Jacobians, like the linearisation)
                    double j0, j1;
                    if (!gact) {
                        tcost = 0.5 * proj_eval(pose + 7 * i, sh.R + 9 * i, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], xi, yi, xj, yj, true,
                                                p.lm_const[l] != 0, r, Ji, Jj, Jl);
                        j0 = 0; j1 = 0;
                    } else {
                        const int oj = o0 + (j - i);
                        double Je[12], Jt[2] = {0, 0};
                        tcost = 0.5 * proj_eval(pose + 7 * i, sh.R + 9 * i, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], xi, yi, xj, yj, true,
                                                p.lm_const[l] != 0, r, Ji, Jj, Jl, Je, m.td_active ? p.obs_vel + 2 * o0 : nullptr,
                                                m.td_active ? p.obs_vel + 2 * oj : nullptr, m.td_active ? Jt : nullptr);
                        j0 = Jt[0] * sh.colv[BA_COL_TD]; j1 = Jt[1] * sh.colv[BA_COL_TD];
                        if (m.ex_active) {
#pragma unroll
                            for (int c = 0; c < 6; ++c) { j0 += Je[c] * sh.colv[BA_COL_EX + c]; j1 += Je[6 + c] * sh.colv[BA_COL_EX + c]; }
                        }
                    }
                    const bool host_const = (i == 0 && !m.use_imu);
                    const double dl = p.y_l[l];
                    j0 += Jl[0] * dl; j1 += Jl[1] * dl;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double di = host_const ? 0.0 : sh.colv[6 * i + a], dj = sh.colv[6 * j + a];
                        j0 += Ji[a] * di + Jj[a] * dj; j1 += Ji[6 + a] * di + Jj[6 + a] * dj;
                    }
                    tdphi = j0 * r[0] + j1 * r[1];
                }
            }
        } else {
            // ---- projection factors, linearisation: one warp per frame pair (host i, observer j), one lane per factor ----
            // Every factor of the pair adds to the same 12x12 block of J^T J.  The 90 distinct sums (off-diagonal 6x6,
            // two diagonal lower triangles, two gradient 6-vectors) are reduced over the warp's lanes with a
            // reduce-scatter butterfly (one shuffle per value instead of five), accumulated in registers over the
            // pair's batches and written to the pair's slot of p.pair_part.  The landmark sums (host-frame part of the W
            // row, hll, gl) of a factor go to the factor's slot of p.fpart; the observer part of the W row has this factor
            // as its only contributor and is stored directly.
            const int pi = task - m.nimu;
            int i = 0, rem = pi;
            while (rem >= BA_NF - 1 - i) { rem -= BA_NF - 1 - i; ++i; }
            const int j = i + 1 + rem;
            double acc[6] = {0, 0, 0, 0, 0, 0};
            double accg[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            // the pair's factors are contiguous in the pair-major list built once per solve: full batches of 32 factors
            const int f_end = sh.pair_ptr[pi + 1];
            for (int fb = sh.pair_ptr[pi]; fb < f_end; fb += 32) {
                const bool act = fb + lane < f_end;
                const int l = act ? (p.fac[fb + lane] & 0xFFFF) : 0;
                const int o0 = act ? p.obs_ptr[l] : 0;
                double r[2] = {0, 0}, Ji[12], Jj[12], Jl[2] = {0, 0};
                double Jg[14];          // [ex-pose | td] block, only touched when one of them is variable
#pragma unroll
                for (int k = 0; k < 12; ++k) { Ji[k] = 0; Jj[k] = 0; }
                if (gact) {
#pragma unroll
                    for (int k = 0; k < 14; ++k) Jg[k] = 0;
                }
                if (act) {
                    const int oj = o0 + (j - i);
                    double xi, yi, xj, yj, rho0;
                    obs_at(m, p, o0, td, xi, yi);
                    obs_at(m, p, oj, td, xj, yj);
                    double *Wl = p.W + (size_t)l * BA_WS;
                    double *fp = p.fpart + (size_t)oj * BA_FP_STRIDE;
                    if (!gact)
                        rho0 = proj_eval(pose + 7 * i, sh.R + 9 * i, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], xi, yi, xj, yj, true,
                                         p.lm_const[l] != 0, r, Ji, Jj, Jl);
                    else {
                        double Je[12], Jt[2] = {0, 0};
                        rho0 = proj_eval(pose + 7 * i, sh.R + 9 * i, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], xi, yi, xj, yj, true,
                                         p.lm_const[l] != 0, r, Ji, Jj, Jl, Je, m.td_active ? p.obs_vel + 2 * o0 : nullptr,
                                         m.td_active ? p.obs_vel + 2 * oj : nullptr, m.td_active ? Jt : nullptr);
#pragma unroll
                        for (int c = 0; c < 6; ++c) { Jg[c] = m.ex_active ? Je[c] : 0.0; Jg[7 + c] = m.ex_active ? Je[6 + c] : 0.0; }
                        Jg[6] = Jt[0]; Jg[13] = Jt[1];
#pragma unroll
                        for (int q = 0; q < 7; ++q) fp[8 + q] = Jg[q] * Jl[0] + Jg[7 + q] * Jl[1];
                    }
                    tcost += 0.5 * rho0;
                    // a constant parameter block is not part of the Ceres program: no Jacobian columns (para_Pose[0] in VO mode,
                    // estimator.cpp:1182-1185)
                    if (i == 0 && !m.use_imu) {
#pragma unroll
                        for (int k = 0; k < 12; ++k) Ji[k] = 0.0;
                    }
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        Wl[6 * j + a] = Jj[a] * Jl[0] + Jj[6 + a] * Jl[1];
                        fp[a] = Ji[a] * Jl[0] + Ji[6 + a] * Jl[1];
                    }
                    fp[6] = Jl[0] * Jl[0] + Jl[1] * Jl[1];
                    fp[7] = Jl[0] * r[0] + Jl[1] * r[1];
                }
                if (gact) accumulate_g(accg, lane, Ji, Jj, Jg, r);
#pragma unroll
                for (int ps = 0; ps < 6; ++ps) {
                    double v[16];
#pragma unroll
                    for (int t = 0; t < 16; ++t) v[t] = pair_entry(ps * 16 + t, Ji, Jj, r);
                    acc[ps] += reduce_scatter16(v, lane);
                }
            }
            if (!(lane & 1)) {
                double *pp = p.pair_part + (size_t)pi * BA_PP_STRIDE + (lane >> 1);
#pragma unroll
                for (int ps = 0; ps < 6; ++ps) pp[16 * ps] = acc[ps];
                if (gact) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) pp[96 + 16 * q] = accg[q];
                }
            }
        }
        tcost = warp_sum_d(tcost);
        if (dirder) tdphi = warp_sum_d(tdphi);
        if (lane == 0) { p.task_cost[task] = tcost; if (dirder) p.task_cost[BA_MAX_TASKS + task] = tdphi; }
    }
    EPROF(1);
    __syncthreads();
    double cost = 0.0, dphi = 0.0;
    for (int t = tid; t < ntask; t += BA_THREADS) { cost += p.task_cost[t]; if (dirder) dphi += p.task_cost[BA_MAX_TASKS + t]; }
    if (lin) {
        // ---- assembly, one owner thread per destination, contributions added in a fixed order ----
        // (a) projection factors: pose blocks and gradient from the per-pair sums, landmark sums from the per-factor slots
        const int nfp = gact ? 15 : 8;
        const int nA = BA_NPAIR * 36, nB = BA_NF * 21, nC = BA_NF * 6, nD = gact ? 7 * BA_NF * 6 + 7 + 28 : 0, nG = m.M * nfp;
        for (int it = tid; it < nA + nB + nC + nD + nG; it += BA_THREADS) {
            int e = it;
            if (e < nA) {
                // off-diagonal 6 x 6 block Jj^T Ji of pair (i, j): this pair is its only projection contributor
                const int pi = e / 36, t = e - pi * 36;
                int i = 0, rem = pi;
                while (rem >= BA_NF - 1 - i) { rem -= BA_NF - 1 - i; ++i; }
                const int j = i + 1 + rem;
                sh.H[pk(6 * j + t / 6, 6 * i + t % 6)] = p.pair_part[(size_t)pi * BA_PP_STRIDE + t];
                continue;
            }
            e -= nA;
            if (e < nB + nC) {
                // diagonal 6 x 6 block (lower triangle, 21 entries) / gradient (6 entries) of frame f: frame f is the observer of
                // the pairs (i < f, f) and the host of the pairs (f, j > f)
                const bool grad = e >= nB;
                if (grad) e -= nB;
                const int per = grad ? 6 : 21;
                const int f = e / per, t = e - f * per;
                const int so = grad ? 78 + t : 36 + t, sh_ = grad ? 84 + t : 57 + t;     // slot as observer / as host
                double sum = 0.0;
                for (int i = 0; i < f; ++i) sum += p.pair_part[(size_t)pair_index(i, f) * BA_PP_STRIDE + so];
                for (int j = f + 1; j < BA_NF; ++j) sum += p.pair_part[(size_t)pair_index(f, j) * BA_PP_STRIDE + sh_];
                if (grad) sh.g[6 * f + t] = sum;
                else {
                    int a = 0;
                    while ((a + 1) * (a + 2) / 2 <= t) ++a;
                    sh.H[pk(6 * f + a, 6 * f + (t - a * (a + 1) / 2))] = sum;
                }
                continue;
            }
            e -= nB + nC;
            if (e < nD) {
                double sum = 0.0;
                if (e < 7 * BA_NF * 6) {
                    // H[ex/td row q, pose column 6 f + c]
                    const int q = e / (BA_NF * 6), fc = e - q * (BA_NF * 6), f = fc / 6, c = fc - 6 * f;
                    for (int i = 0; i < f; ++i) sum += p.pair_part[(size_t)pair_index(i, f) * BA_PP_STRIDE + 96 + 16 * q + 6 + c];
                    for (int j = f + 1; j < BA_NF; ++j) sum += p.pair_part[(size_t)pair_index(f, j) * BA_PP_STRIDE + 96 + 16 * q + c];
                    sh.H[pk(BA_COL_EX + q, 6 * f + c)] = sum;
                } else if (e < 7 * BA_NF * 6 + 7) {
                    const int q = e - 7 * BA_NF * 6;
                    for (int pi = 0; pi < BA_NPAIR; ++pi) sum += p.pair_part[(size_t)pi * BA_PP_STRIDE + 96 + 16 * q + 12];
                    sh.g[BA_COL_EX + q] = sum;
                } else {
                    const int t = e - (7 * BA_NF * 6 + 7);
                    for (int pi = 0; pi < BA_NPAIR; ++pi) sum += p.pair_part[(size_t)pi * BA_PP_STRIDE + 208 + t];
                    int a = 0;
                    while ((a + 1) * (a + 2) / 2 <= t) ++a;
                    sh.H[pk(BA_COL_EX + a, BA_COL_EX + t - a * (a + 1) / 2)] = sum;
                }
                continue;
            }
            e -= nD;
            {
                // landmark l, value a: sum over the landmark's factors in observation order
                const int l = e / nfp, a = e - l * nfp;
                const int o0 = p.obs_ptr[l], o1 = p.obs_ptr[l + 1];
                double sum = 0.0;
                for (int o = o0 + 1; o < o1; ++o) sum += p.fpart[(size_t)o * BA_FP_STRIDE + a];
                if (a < 6) p.W[(size_t)l * BA_WS + 6 * p.start[l] + a] = sum;
                else if (a == 6) p.hll[l] = sum;
                else if (a == 7) p.gl[l] = sum;
                else p.W[(size_t)l * BA_WS + 66 + (a - 8)] = sum;
            }
        }
        __syncthreads();
        // (b) IMU factors: J^T J (465 entries of the 30 x 30 lower triangle) and J^T r (30) per factor from the whitened
        //     Jacobians in shared memory.  Factors with equal parity of j touch disjoint frames: two passes, plain adds.
        for (int pass = 0; pass < 2; ++pass) {
            const int f_lo = pass ? sh.imu_neven : 0, f_hi = pass ? m.nimu : sh.imu_neven;
            for (int it = tid + f_lo * 495; it < f_hi * 495; it += BA_THREADS) {
                const int fo = it / 495, e = it - fo * 495;
                const int f = sh.imu_ord[fo];
                const int j = m.imu_j[f], i = j - 1;
                const double *J = sh.imuJ[f];
                auto tcol = [&](int c) { return c < 6 ? 6 * i + c : c < 15 ? 66 + 9 * i + (c - 6) : c < 21 ? 6 * j + (c - 15) : 66 + 9 * j + (c - 21); };
                if (e < 465) {
                    int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
                    while (a * (a + 1) / 2 > e) --a;
                    while ((a + 1) * (a + 2) / 2 <= e) ++a;
                    const int b = e - a * (a + 1) / 2;
                    double h = 0;
#pragma unroll
                    for (int k = 0; k < 15; ++k) h += J[k * 30 + a] * J[k * 30 + b];
                    sh.H[pk(tcol(a), tcol(b))] += h;
                } else {
                    const int c = e - 465;
                    double gsum = 0;
#pragma unroll
                    for (int k = 0; k < 15; ++k) gsum += J[k * 30 + c] * sh.imur[f][k];
                    sh.g[tcol(c)] += gsum;
                }
            }
            __syncthreads();
        }
    }
    EPROF(3);
    EPROF(2);
    // ---- prior (MarginalizationFactor::Evaluate, marginalization_factor.cpp:353-415), in information form:
    //      1/2 |r0 + J0 dx|^2 = c0 + dx.(gp + HP dx / 2),  J0^T (r0 + J0 dx) = gp + HP dx,  HP = J0^T J0, gp = J0^T r0 ----
    const BaPriorStore *P = p.prior;
    const int np_ = (P && P->valid) ? P->n : 0;
    if (np_ > 0) {
        const int n = np_;
        const double *HPm = P->J0;              // information form (see the kernel's set-up)
        // w = HP dx: one warp per row, lanes along the row (coalesced).  For the usual sizes (n <= 96) a warp's row
        // fragments are requested before dx exists, so that the L2 latency overlaps prior_dx.
        const bool small = n <= 96;
        double jv[6][3];
        if (small) {
#pragma unroll
            for (int q = 0; q < 6; ++q)
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const int rI = warp + nwarp * q, k = lane + 32 * kk;
                    jv[q][kk] = (rI < n && k < n) ? __ldg(&HPm[(size_t)rI * n + k]) : 0.0;
                }
        }
        prior_dx_smem(sh, pose, sb, ex, td);
        __syncthreads();
        if (tid == 0) cost += sh.pc0;
        if (small) {
            double dxv[3], a[6];
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) dxv[kk] = (lane + 32 * kk < n) ? sh.dx[lane + 32 * kk] : 0.0;
#pragma unroll
            for (int q = 0; q < 6; ++q) a[q] = jv[q][0] * dxv[0] + jv[q][1] * dxv[1] + jv[q][2] * dxv[2];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < 6; ++q) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int rI = warp + nwarp * q;
                if (lane == 0 && rI < n) { sh.pr[rI] = a[q]; cost += sh.dx[rI] * (sh.pr0[rI] + 0.5 * a[q]); }
            }
        } else {
            for (int rI = warp; rI < n; rI += nwarp) {
                const double *row = HPm + (size_t)rI * n;
                double a = 0;
                for (int k = lane; k < n; k += 32) a += row[k] * sh.dx[k];
                a = warp_sum_d(a);
                if (lane == 0) { sh.pr[rI] = a; cost += sh.dx[rI] * (sh.pr0[rI] + 0.5 * a); }
            }
        }
        __syncthreads();
        if (dirder) {
            for (int a = tid; a < n; a += BA_THREADS) {
                const int ca = sh.pcol[a];
                if (ca >= 0) dphi += (sh.pr0[a] + sh.pr[a]) * sh.colv[ca];
            }
        }
        if (lin) {
            // g += gp + HP dx ; H += HP, both through the column map (constant blocks drop out; distinct prior columns map to
            // distinct tangent columns, so every destination has one writer)
            for (int a = tid; a < n; a += BA_THREADS) {
                const int ca = sh.pcol[a];
                if (ca >= 0) sh.g[ca] += sh.pr0[a] + sh.pr[a];
            }
#pragma unroll 4
            for (int e = tid; e < n * n; e += BA_THREADS) {
                int a2 = e / n, b = e - a2 * n;
                if (b > a2) continue;
                int ca = sh.pcol[a2], cb = sh.pcol[b];
                if (ca < 0 || cb < 0) continue;
                sh.H[pk(ca, cb)] += __ldg(&HPm[(size_t)a2 * n + b]);
            }
        }
    }
    __syncthreads();
    EPROF(4);
    if (eprof && blockIdx.x == 0 && (tid == 0 || tid == 320 || tid == 500))
        printf("evaluate mode=%d tid=%d zero=%lld tasks=%lld diag=%lld wait=%lld prior=%lld\n", mode, tid, et[0], et[1], et[2], et[3], et[4]);
#undef EPROF
    if (dirder) { const double d_ = block_sum(dphi, sh.red); __syncthreads(); *dphi_out = d_; }
    return block_sum(cost, sh.red);
}

__device__ __forceinline__ double ba_evaluate(const BaMeta &m, const BaProbDev &p, BaShared &sh, const double *pose, const double *sb,
                                              const double *ex, const double td, const double *lam, int mode, double *dphi_out = nullptr)
{
    if (mode == 2)
        return (m.ex_active || m.td_active) ? ba_evaluate_t<true, true>(m, p, sh, pose, sb, ex, td, lam, mode, dphi_out)
                                            : ba_evaluate_t<false, true>(m, p, sh, pose, sb, ex, td, lam, mode, dphi_out);
    return (m.ex_active || m.td_active) ? ba_evaluate_t<true, false>(m, p, sh, pose, sb, ex, td, lam, mode, dphi_out)
                                        : ba_evaluate_t<false, false>(m, p, sh, pose, sb, ex, td, lam, mode, dphi_out);
}

// D(8x8) = A(8x4) B(4x8) + D on the FP64 tensor cores (SASS: DMMA.8x8x4); see the trailing update of the Cholesky below
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 8x8 diagonal block (rows/cols j0 .. j0+nbp-1) of the blocked Cholesky, factorised by ONE warp.  This block is the serial
// spine of the factorisation (22 panels x 8 dependent pivots per solve): every lane keeps the WHOLE lower triangle (36 doubles)
// in registers and runs the 8 pivots redundantly, so that a pivot costs one rsqrt + one multiply + one FMA of latency and no
// shuffles (the previous version held one row per lane and paid 8 shuffle round trips per pivot: ~5.5 k cycles per block,
// the critical path of the panel loop; this one ~1.2 k).
// `update`: first subtract the rank-8 contribution of the panel staged in sh.Lp (look-ahead), two entries per lane, gathered
// through sh.d8.  Writes the factor back to sh.H, the reciprocal pivots to sh.colv[0..7], and raises sh.flag[0] on a
// non-positive pivot.  Entries outside the nbp x nbp block are treated as the identity.
__device__ __noinline__ void chol_diag8(BaShared &sh, int j0, int nbp, int lane, bool update, bool dprof = false)
{
    long long dt0 = dprof ? clock64() : 0, dt1 = 0, dt2 = 0, dt3 = 0;
    {
        // the block as one DMMA tile: lane -> row m8, columns 2 kq, 2 kq + 1 (C fragment of m8n8k4); with `update` the rank-8
        // contribution of the staged panel is subtracted by two DMMAs (A = -B^T = -Lp[.][j0 + .])
        const int m8 = lane >> 2, kq = lane & 3, row = j0 + m8, c = 2 * kq;
        const bool v0 = m8 < nbp && c <= m8, v1 = m8 < nbp && c + 1 <= m8;
        const double *rp = sh.H + row * (row + 1) / 2 + j0;
        double c0 = v0 ? rp[c] : 0.0, c1 = v1 ? rp[c + 1] : 0.0;
        if (update) {
            const double b0 = sh.Lp[kq][row], b1 = sh.Lp[4 + kq][row];
            dmma884(c0, c1, -b0, b0);
            dmma884(c0, c1, -b1, b1);
        }
        if (c <= m8) sh.d8[m8 * (m8 + 1) / 2 + c] = v0 ? c0 : (m8 == c ? 1.0 : 0.0);
        if (c + 1 <= m8) sh.d8[m8 * (m8 + 1) / 2 + c + 1] = v1 ? c1 : (m8 == c + 1 ? 1.0 : 0.0);
    }
    __syncwarp();
    if (dprof) dt1 = clock64();
    double a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) a[r][c] = sh.d8[r * (r + 1) / 2 + c];
    if (dprof) dt2 = clock64();
    bool okp = true;
    double rinv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double dj = a[j][j];
        if (!(dj > 0.0)) okp = false;
        // FP64 sqrt/div have ~500-cycle latencies on this part: one rsqrt per pivot instead
        rinv[j] = rsqrt(dj);
        a[j][j] = dj * rinv[j];
#pragma unroll
        for (int r = j + 1; r < 8; ++r) a[r][j] *= rinv[j];
#pragma unroll
        for (int c = j + 1; c < 8; ++c)
#pragma unroll
            for (int r = c; r < 8; ++r) a[r][c] -= a[r][j] * a[c][j];
    }
    if (dprof) dt3 = clock64();
    // every lane holds the same factor; lane 0 writes it back (32 lanes storing to one address serialise: measured 1 k cycles)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (r < nbp && lane == 0) {
            double *wp = sh.H + (j0 + r) * (j0 + r + 1) / 2 + j0;
#pragma unroll
            for (int c = 0; c <= r; ++c) wp[c] = a[r][c];
            sh.colv[r] = rinv[r];                    // 1 / L[j0+r, j0+r]
        }
    }
    if (!okp && lane == 0) sh.flag[0] = 1;
    if (dprof && lane == 0)
        printf("diag8 j0=%d: update+gather %lld  load %lld  factor %lld  store %lld cycles\n", j0, dt1 - dt0, dt2 - dt1, dt3 - dt2, clock64() - dt3);
}

// scale a fresh linearisation: H_s = D H D, g_s = D g, W_s, hll_s, gl_s (Jacobi scaling of the Jacobian columns)
__device__ __noinline__ void ba_scale(const BaMeta &m, const BaProbDev &p, BaShared &sh)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    for (int a = warp; a < BA_NC; a += nwarp) {              // one packed row per warp, lanes along the row
        double *row = sh.H + a * (a + 1) / 2;
        const double sa = sh.jscale[a];
        for (int b = lane; b <= a; b += 32) row[b] *= sa * sh.jscale[b];
    }
    for (int c = tid; c < BA_NC; c += BA_THREADS) sh.g[c] *= sh.jscale[c];
    // The coupling rows W stay UNSCALED in HBM/L2 (one read-modify-write pass over M x 73 doubles less per linearisation): their
    // four readers apply the factor jscale_l[l] * jscale[col] on the fly (bit-identical products, see wsc()).
    for (int l = tid; l < m.M; l += BA_THREADS) { const double sl = p.jscale_l[l]; p.hll[l] *= sl * sl; p.gl[l] *= sl; }
    __syncthreads();
}

// scaled coupling entry: W_s[l][k] = W[l][k] * (jscale_l[l] * jscale[col(k)])  (what ba_scale used to store in place)
__device__ __forceinline__ double wsc(const BaShared &sh, const double *Wl, int k, double sl)
{
    return Wl[k] * (sl * sh.jscale[wcol(k)]);
}

// (sum over camera + landmark entries of a_c*b_c) helper: camera part from smem arrays, landmark part from global
__device__ double dot_full(const BaMeta &m, const double *ac, const double *bc, const double *al, const double *bl, double *s_red)
{
    double v = 0;
    for (int i = threadIdx.x; i < BA_NC; i += BA_THREADS) v += ac[i] * bc[i];
    for (int l = threadIdx.x; l < m.M; l += BA_THREADS) v += al[l] * bl[l];
    return block_sum(v, s_red);
}

// the trust-region step in the parameters' tangent space: delta = (c1 gd + c2 gn) / diag * jscale (dogleg coefficients c1, c2)
__device__ __forceinline__ double ba_dcol(const BaShared &sh, double c1, double c2, int c)
{
    return (c1 * sh.gd[c] + c2 * sh.gn[c]) / sh.diag[c] * sh.jscale[c];
}
__device__ __forceinline__ double ba_dlm(const BaProbDev &p, double c1, double c2, int l)
{
    return (c1 * p.gd_l[l] + c2 * p.gn_l[l]) / p.diag_l[l] * p.jscale_l[l];
}

// candidate = x (+) t delta into sh.cpose / sh.csb / sh.cex / sh.tdv[1] / p.clam; t = 1 except for the trial points of the line search
__device__ __noinline__ void ba_make_candidate(const BaMeta &m, const BaProbDev &p, BaShared &sh, double c1, double c2, const double t)
{
    const int tid = threadIdx.x;
    for (int f = tid; f < BA_NF; f += BA_THREADS) {
        if (col_active_dev(m, 6 * f)) {
            double dl[6];
            for (int k = 0; k < 6; ++k) dl[k] = ba_dcol(sh, c1, c2, 6 * f + k) * t;
            d_pose_plus(sh.pose + 7 * f, dl, sh.cpose + 7 * f);
        } else for (int k = 0; k < 7; ++k) sh.cpose[7 * f + k] = sh.pose[7 * f + k];
        for (int k = 0; k < 9; ++k) {
            const int c = 66 + 9 * f + k;
            sh.csb[9 * f + k] = sh.sb[9 * f + k] + (col_active_dev(m, c) ? ba_dcol(sh, c1, c2, c) * t : 0.0);
        }
    }
    if (tid == 64) {
        if (m.ex_active) {
            double dl[6];
            for (int k = 0; k < 6; ++k) dl[k] = ba_dcol(sh, c1, c2, BA_COL_EX + k) * t;
            d_pose_plus(sh.ex, dl, sh.cex);
        } else for (int k = 0; k < 7; ++k) sh.cex[k] = sh.ex[k];
        sh.tdv[1] = sh.tdv[0] + (m.td_active ? ba_dcol(sh, c1, c2, BA_COL_TD) * t : 0.0);
    }
    for (int l = tid; l < m.M; l += BA_THREADS) {
        double v = p.lam[l];
        if (!p.lm_const[l]) {
            v += ba_dlm(p, c1, c2, l) * t;
            v = fmin(v, p.lm_ub[l]);           // ParameterBlock::Plus projects onto the bounds
        }
        p.clam[l] = v;
    }
    __syncthreads();
}

// Ceres' projected Armijo line search of bound-constrained problems (TrustRegionMinimizer::DoLineSearch; statement and
// defaults: ba_linesearch.cuh, oracle/ba_ref.c), entered when the full step fails f(x [+] delta) <= f(x) + 1e-4 g.delta.
// phi(t) = f(x [+] t delta) with the bounds projection inside Plus, phi'(t) = delta . gradient at the trial point.  The step
// is contracted by polynomial interpolation until the test holds (<= 20 iterations, step size >= 1e-9 / |delta|_inf).  A trial
// point costs one evaluation of all factors with Jacobians, each contributing r^T (J delta): the linearisation at x stays
// untouched.  On success the candidate arrays hold x [+] t delta and *cand_cost its cost; a failed search leaves the full step.
__device__ __noinline__ bool ba_line_search(const BaMeta &m, const BaProbDev &p, BaShared &sh, double c1, double c2, double x_cost,
                                            double sTg, double *cand_cost)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    const int M = m.M;
    double dmax = 0;
    // the direction for the trial evaluations (both arrays are idle between two factorisations)
    for (int c = tid; c < BA_NC; c += BA_THREADS) { const double d = ba_dcol(sh, c1, c2, c); sh.colv[c] = d; dmax = fmax(dmax, fabs(d)); }
    for (int l = tid; l < M; l += BA_THREADS) { const double d = p.lm_const[l] ? 0.0 : ba_dlm(p, c1, c2, l); p.y_l[l] = d; dmax = fmax(dmax, fabs(d)); }
    for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    __syncthreads();
    if (lane == 0) sh.red[warp] = dmax;
    __syncthreads();
    dmax = 0;
    for (int i = 0; i < nwarp; ++i) dmax = fmax(dmax, sh.red[i]);
    __syncthreads();
    const LsSample lower = {0.0, x_cost, sTg, true, true};
    LsSample prev = {0.0, 0.0, 0.0, false, false}, cur = {1.0, *cand_cost, 0.0, false, true};
    bool ok = false;
    for (int it = 0;; ++it) {
        if (it > 0) {
            if (it >= 20) break;                                   // max_num_line_search_step_size_iterations
            // (one warp evaluates the polynomial step, the others wait: the samples are identical in every thread)
            if (warp == 0) { const double t_ = ls_interpolating_step(lower, prev, cur, 1e-3 * cur.x, 0.6 * cur.x); if (lane == 0) sh.sc[0] = t_; }
            __syncthreads();
            const double t = sh.sc[0];
            __syncthreads();
            if (t * dmax < 1e-9) break;                            // min_line_search_step_size
            prev = cur;
            cur.x = t;
            ba_make_candidate(m, p, sh, c1, c2, t);
        }
        double dg = 0;
        cur.value = ba_evaluate(m, p, sh, sh.cpose, sh.csb, sh.cex, sh.tdv[1], p.clam, 2, &dg);
        cur.gradient = dg;
        cur.value_ok = isfinite(cur.value) && isfinite(cur.gradient);
        cur.grad_ok = cur.value_ok;
        if (cur.value_ok && cur.value <= x_cost + 1e-4 * sTg * cur.x) { ok = it > 0; break; }
    }
    if (ok) *cand_cost = cur.value;                    // the candidate arrays already hold x [+] t delta
    else ba_make_candidate(m, p, sh, c1, c2, 1.0);    // failed search: the step stays as it was
    return ok;
}


__global__ void __launch_bounds__(BA_THREADS, 1)
k_ba_solve(const BaMeta *metas, const BaProbDev *probs, BaOutDev *outs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BaShared &sh = *reinterpret_cast<BaShared *>(smem_raw);
    // The problem descriptors live in SHARED memory, one copy per CTA.  As per-thread copies they were forced into local
    // memory (they are passed by reference to the __noinline__ phases): 512 threads x 2.8 KB of stack = 1.4 MB per CTA behind
    // a 36 KB L1, so every `p.field` pointer load on the way to a global access missed to L2 (and 148 CTAs x 1.4 MB spilled
    // out of L2: 1.1 GB of DRAM traffic per launch, profiles/r02_kernels.md).
    if (threadIdx.x < (int)(sizeof(BaMeta) / sizeof(int))) reinterpret_cast<int *>(&sh.meta)[threadIdx.x] = reinterpret_cast<const int *>(metas + blockIdx.x)[threadIdx.x];
    if (threadIdx.x < (int)(sizeof(BaProbDev) / sizeof(int))) reinterpret_cast<int *>(&sh.prob)[threadIdx.x] = reinterpret_cast<const int *>(probs + blockIdx.x)[threadIdx.x];
    __syncthreads();
    const BaMeta &m = sh.meta;
    const BaProbDev &p = sh.prob;
    BaOutDev &out = outs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    const int M = m.M;
    const long long t_kernel00 = clock64();

    for (int i = tid; i < BA_NF * 7; i += BA_THREADS) sh.pose[i] = p.pose0[i];
    for (int i = tid; i < BA_NF * 9; i += BA_THREADS) sh.sb[i] = p.sb0[i];
    if (tid < 7) sh.ex[tid] = p.ex0[tid];
    if (tid == 0) { sh.tdv[0] = *p.td0; sh.tdv[1] = *p.td0; }
    // TrustRegionMinimizer::Init of a bound-constrained problem projects the start point onto the bounds (Plus(x, 0)) before
    // the first evaluation: only the estimate_flag == 2 landmarks carry one (estimator.cpp:1293-1298)
    for (int l = tid; l < M; l += BA_THREADS) p.lam[l] = p.lm_const[l] ? p.lam0[l] : fmin(p.lam0[l], p.lm_ub[l]);
    // coupling rows: zeroed once, every linearisation then rewrites the same entries (host frame, observer frames, ex-pose / td)
    for (int i = tid; i < M * BA_WS; i += BA_THREADS) p.W[i] = 0.0;
    if (tid == 0) {
        int k = 0;
        for (int f = 0; f < m.nimu; ++f) if (!(m.imu_j[f] & 1)) sh.imu_ord[k++] = f;
        sh.imu_neven = k;
        for (int f = 0; f < m.nimu; ++f) if (m.imu_j[f] & 1) sh.imu_ord[k++] = f;
    }
    const int ws = (m.ex_active || m.td_active) ? BA_WS : 66;      // used width of the landmark coupling rows
    __syncthreads();
    // ---- once per solve: IMU information square roots, prior normal matrix ----
    {
        double *scr = reinterpret_cast<double *>(sh.imuJ) + warp * 450;   // 10 warps x 450 doubles fit in imuJ
        if (warp < m.nimu && warp < BA_NF - 1) imu_sqrt_info_warp(p.imu[m.imu_j[warp] - 1].covariance, p.imuS + (size_t)(m.imu_j[warp] - 1) * 225, scr);
        for (int f = warp + nwarp; f < m.nimu; f += nwarp) imu_sqrt_info_warp(p.imu[m.imu_j[f] - 1].covariance, p.imuS + (size_t)(m.imu_j[f] - 1) * 225, scr);
        const BaPriorStore *P = p.prior;
        const int n = (P && P->valid) ? P->n : 0;
        if (n > 0 && P->form == 0) {
            // an uploaded (linearized_jacobians, linearized_residuals) prior: J0 <- J0^T J0, r0 <- J0^T r0, c0 = r0^T r0 / 2,
            // in place (through the HP scratch); from here on every kernel sees the information form
            BaPriorStore *Pw = const_cast<BaPriorStore *>(P);
            for (int e = tid; e < n * n; e += BA_THREADS) {
                int a = e / n, b = e - a * n;
                if (b > a) continue;
                double h = 0;
                for (int rI = 0; rI < n; ++rI) h += P->J0[(size_t)rI * n + a] * P->J0[(size_t)rI * n + b];
                p.HP[(size_t)a * n + b] = h; p.HP[(size_t)b * n + a] = h;
            }
            double cc = 0;
            for (int a = tid; a < n; a += BA_THREADS) {
                double gsum = 0;
                for (int rI = 0; rI < n; ++rI) gsum += P->J0[(size_t)rI * n + a] * P->r0[rI];
                sh.pr[a] = gsum;
                cc += 0.5 * P->r0[a] * P->r0[a];
            }
            cc = block_sum(cc, sh.red);              // (barriers inside: every read of J0 / r0 above is complete)
            for (int e = tid; e < n * n; e += BA_THREADS) Pw->J0[e] = p.HP[e];
            for (int a = tid; a < n; a += BA_THREADS) Pw->r0[a] = sh.pr[a];
            if (tid == 0) { Pw->c0 = cc; Pw->form = 1; }
            __threadfence();
            __syncthreads();
        }
        if (tid == 0) sh.pc0 = n > 0 ? P->c0 : 0.0;
        // prior column -> tangent column (constant blocks drop out); block table, x0 and r0 staged in shared memory
        if (tid == 0) sh.pnb = n > 0 ? P->n_blocks : 0;
        if (n > 0) {
            for (int b = tid; b < P->n_blocks; b += BA_THREADS) {
                const int kind = P->kind[b], ls = P->size[b] == 7 ? 6 : P->size[b];
                int col = kind == VRF_BLK_POSE ? 6 * P->index[b] : kind == VRF_BLK_SPEEDBIAS ? 66 + 9 * P->index[b]
                          : kind == VRF_BLK_EXPOSE ? BA_COL_EX : kind == VRF_BLK_TD ? BA_COL_TD : -1;
                if (col >= 0 && !col_active_dev(m, col)) col = -1;
                for (int c = 0; c < ls; ++c) { p.colmap[P->idx[b] + c] = col < 0 ? -1 : col + c; sh.pcol[P->idx[b] + c] = col < 0 ? -1 : col + c; }
                sh.pkind[b] = kind; sh.pindex[b] = P->index[b]; sh.psize[b] = P->size[b]; sh.pidx[b] = P->idx[b];
            }
            for (int i = tid; i < P->n_blocks * 9; i += BA_THREADS) sh.px0[i] = P->x0[i];
            for (int i = tid; i < n; i += BA_THREADS) sh.pr0[i] = P->r0[i];
        }
    }
    __syncthreads();
    __threadfence_block();

    // ---- once per solve: the projection factors in frame-pair-major order (pair (i, j) = host frame i, observer frame j > i;
    //      within a pair in landmark order, so the list -- and every sum formed over it -- is reproducible) ----
    {
        for (int pass = 0; pass < 2; ++pass) {
            for (int pi = warp; pi < BA_NPAIR; pi += nwarp) {
                int i = 0, rem = pi;
                while (rem >= BA_NF - 1 - i) { rem -= BA_NF - 1 - i; ++i; }
                const int j = i + 1 + rem;
                int off = pass ? sh.pair_ptr[pi] : 0;
                for (int b0 = 0; b0 < M; b0 += 32) {
                    const int l = b0 + lane;
                    const bool a = l < M && p.start[l] == i && (p.obs_ptr[l + 1] - p.obs_ptr[l] - 1) >= (j - i);
                    const unsigned mask = __ballot_sync(0xffffffffu, a);
                    if (pass && a) p.fac[off + __popc(mask & ((1u << lane) - 1u))] = l | (j << 16);
                    off += __popc(mask);
                }
                if (!pass && lane == 0) sh.pair_cnt[pi] = off;
            }
            __syncthreads();
            if (!pass) {
                if (tid == 0) {
                    int acc = 0;
                    for (int pi = 0; pi < BA_NPAIR; ++pi) { sh.pair_ptr[pi] = acc; acc += sh.pair_cnt[pi]; }
                    sh.pair_ptr[BA_NPAIR] = acc;
                }
                __syncthreads();
            }
        }
        __threadfence_block();
    }

    const long long t_kernel0 = t_kernel00;
    long long tprof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tmark = clock64();
#define TPROF(k) do { long long t_ = clock64(); tprof[k] += t_ - tmark; tmark = t_; } while (0)
    double x_cost = ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, 1);
    TPROF(0);
    const double initial_cost = x_cost;
    // Jacobi scaling (once): 1 / (1 + ||column||)
    for (int c = tid; c < BA_NC; c += BA_THREADS) sh.jscale[c] = 1.0 / (1.0 + sqrt(sh.H[pk(c, c)]));
    for (int l = tid; l < M; l += BA_THREADS) p.jscale_l[l] = 1.0 / (1.0 + sqrt(p.hll[l]));
    __syncthreads();

    double radius = 1e4, mu = 1e-8, alpha = 0, dogleg_norm = 0;
    double Quu = 0, gdn2 = 0;           // u^T H u, |gd|^2
    double ytg = 0, yDy = 0, utg = 0, uDy = 0, gnn2 = 0, gdgn = 0;
    const double min_mu = 1e-8, max_mu = 1.0, mu_inc = 10.0;
    int reuse = 0, invalid = 0, need_scale = 1, iterations = 0, successful = 0, termination = 0, status = VRF_OK;
    double x_norm = 0, gradient_max = 0;
    const int max_iter = m.max_iter;

    auto xnorm2 = [&](const double *pose, const double *sb, const double *lam) {
        double v = 0;
        if (m.ex_active && tid < 7) v += sh.ex[tid] * sh.ex[tid];
        if (m.td_active && tid == 7) v += sh.tdv[0] * sh.tdv[0];
        for (int i = tid; i < BA_NF * 7; i += BA_THREADS) if (col_active_dev(m, 6 * (i / 7))) v += pose[i] * pose[i];
        for (int i = tid; i < BA_NF * 9; i += BA_THREADS) if (col_active_dev(m, 66 + 9 * (i / 9))) v += sb[i] * sb[i];
        for (int l = tid; l < M; l += BA_THREADS) if (!p.lm_const[l]) v += lam[l] * lam[l];
        return block_sum(v, sh.red);
    };
    x_norm = sqrt(xnorm2(sh.pose, sh.sb, p.lam));
    // Solver::Options::is_constrained: some non-constant parameter block carries a bound
    int armijo_failures = 0;
    bool constrained;
    {
        double nb_ = 0;
        for (int l = tid; l < M; l += BA_THREADS) if (!p.lm_const[l] && isfinite(p.lm_ub[l])) nb_ += 1;
        constrained = block_sum(nb_, sh.red) > 0;
    }

    while (true) {
        if (need_scale) {
            ba_scale(m, p, sh);
            // projected gradient max-norm (bounds): x - Plus(x, -g)
            {
                double mx = 0;
                for (int f = tid; f < BA_NF; f += BA_THREADS) {
                    if (col_active_dev(m, 6 * f)) {
                        double dl[6], o[7];
                        for (int k = 0; k < 6; ++k) dl[k] = -sh.g[6 * f + k] / sh.jscale[6 * f + k];
                        d_pose_plus(sh.pose + 7 * f, dl, o);
                        for (int k = 0; k < 7; ++k) mx = fmax(mx, fabs(sh.pose[7 * f + k] - o[k]));
                    }
                    if (col_active_dev(m, 66 + 9 * f))
                        for (int k = 0; k < 9; ++k) mx = fmax(mx, fabs(sh.g[66 + 9 * f + k] / sh.jscale[66 + 9 * f + k]));
                }
                if (m.ex_active && tid == 32) {
                    double dl[6], o[7];
                    for (int k = 0; k < 6; ++k) dl[k] = -sh.g[BA_COL_EX + k] / sh.jscale[BA_COL_EX + k];
                    d_pose_plus(sh.ex, dl, o);
                    for (int k = 0; k < 7; ++k) mx = fmax(mx, fabs(sh.ex[k] - o[k]));
                }
                if (m.td_active && tid == 33) mx = fmax(mx, fabs(sh.g[BA_COL_TD] / sh.jscale[BA_COL_TD]));
                for (int l = tid; l < M; l += BA_THREADS) {
                    if (p.lm_const[l]) continue;
                    double v = p.lam[l] + (-p.gl[l] / p.jscale_l[l]);
                    v = fmin(v, p.lm_ub[l]);
                    mx = fmax(mx, fabs(p.lam[l] - v));
                }
                // block max via sum-free reduction
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                __syncthreads();
                if (lane == 0) sh.red[warp] = mx;
                __syncthreads();
                mx = 0;
                for (int i = 0; i < nwarp; ++i) mx = fmax(mx, sh.red[i]);
                gradient_max = mx;
            }
            need_scale = 0;
            TPROF(1);
        }
        if (iterations >= max_iter) { termination = 0; break; }
        if (gradient_max <= 1e-10) { termination = 2; break; }
        if (radius <= 1e-32) { termination = 4; break; }
        ++iterations;

        bool step_ok = true;
        if (!reuse) {
            reuse = 1;
            // diagonal_ = sqrt(clamp(diag(J^T J))), gradient_ /= diagonal_
            for (int c = tid; c < BA_NC; c += BA_THREADS) {
                double dg = sqrt(fmin(fmax(sh.H[pk(c, c)], 1e-6), 1e32));
                sh.diag[c] = dg; sh.gd[c] = sh.g[c] / dg; sh.tmp[c] = sh.g[c] / (dg * dg);      // tmp = u_c = D^-2 g
            }
            for (int l = tid; l < M; l += BA_THREADS) {
                double dg = sqrt(fmin(fmax(p.hll[l], 1e-6), 1e32));
                p.diag_l[l] = dg; p.gd_l[l] = p.gl[l] / dg; p.u_l[l] = p.lm_const[l] ? 0.0 : p.gl[l] / (dg * dg);
            }
            __syncthreads();
            // Cauchy point: alpha = |gd|^2 / (u^T H u), H = J^T J of the scaled problem (before it is reduced in place)
            {
                // u_c^T H_cc u_c over the packed lower triangle: one row per warp pass, lanes along the row (coalesced, no index
                // arithmetic per element); off-diagonal entries count twice
                double v = 0;
                for (int a = warp; a < BA_NC; a += nwarp) {
                    const double *row = sh.H + a * (a + 1) / 2;
                    double hv = 0;
                    for (int b = lane; b <= a; b += 32) hv += row[b] * sh.tmp[b] * (b == a ? 1.0 : 2.0);
                    v += sh.tmp[a] * hv;
                }
                // (four landmarks per pass: their coupling rows live in L2, ~700 cycles away -- the loads of a pass are issued together
                // and the four shuffle reductions interleave; the sums are added in landmark order as before)
                for (int l0 = warp; l0 < M; l0 += 4 * nwarp) {
                    double wv[4];
                    bool on[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int l = l0 + u * nwarp;
                        on[u] = l < M && !p.lm_const[l];
                        wv[u] = 0;
                        if (on[u]) {
                            const double *Wl = p.W + (size_t)l * BA_WS;
                            const double sl = p.jscale_l[l];
                            for (int k = lane; k < ws; k += 32) wv[u] += wsc(sh, Wl, k, sl) * sh.tmp[wcol(k)];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) wv[u] = warp_sum_d(wv[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int l = l0 + u * nwarp;
                        if (on[u] && lane == 0) v += 2.0 * p.u_l[l] * wv[u] + p.hll[l] * p.u_l[l] * p.u_l[l];
                    }
                }
                Quu = block_sum(v, sh.red);
                double gsq = 0;
                for (int c = tid; c < BA_NC; c += BA_THREADS) gsq += sh.gd[c] * sh.gd[c];
                for (int l = tid; l < M; l += BA_THREADS) if (!p.lm_const[l]) gsq += p.gd_l[l] * p.gd_l[l];
                gdn2 = block_sum(gsq, sh.red);
                alpha = gdn2 / Quu;
                utg = 0;
                {
                    double t = 0;
                    for (int c = tid; c < BA_NC; c += BA_THREADS) t += sh.tmp[c] * sh.g[c];
                    for (int l = tid; l < M; l += BA_THREADS) if (!p.lm_const[l]) t += p.u_l[l] * p.gl[l];
                    utg = block_sum(t, sh.red);
                }
            }
            TPROF(2);
            // Gauss-Newton step: (H + mu D^2) y = g by Schur complement + Cholesky, retried with larger mu
            bool solved = false;
            while (mu < max_mu) {
                // rhs
                for (int c = tid; c < BA_NC; c += BA_THREADS) { sh.y[c] = sh.g[c]; }
                __syncthreads();
                // per-landmark pivots; reduce rhs
                if (tid == 0) sh.flag[0] = 0;
                __syncthreads();
                // pivots of all landmarks, one landmark per thread (division / sqrt once per landmark, not once per warp pass);
                // y_l is free until the back-substitution and carries g_l / h_l to the reduction of the right-hand side
                for (int l = tid; l < M; l += BA_THREADS) {
                    if (p.lm_const[l]) continue;
                    const double hl = p.hll[l] + mu * p.diag_l[l] * p.diag_l[l];
                    if (!(hl > 0)) { sh.flag[0] = 1; continue; }
                    const double hi_ = 1.0 / hl;
                    p.hinv_l[l] = hi_; p.shinv_l[l] = sqrt(hi_); p.y_l[l] = p.gl[l] / hl;
                }
                __syncthreads();
                {
                    // y_c -= sum_l w_l (g_l / h_l): per-warp partial rows in registers (lane <-> columns lane, lane + 32, lane + 64),
                    // parked in the (idle) tile buffer and added by one thread per column in warp order
                    double *part = reinterpret_cast<double *>(sh.imuJ);      // [nwarp][96]
                    double a0 = 0, a1 = 0, a2 = 0;
                    for (int l0 = warp; l0 < M; l0 += 4 * nwarp) {
                        double w0[4], w1[4], w2[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {                  // the four rows' loads in flight together
                            const int l = l0 + u * nwarp;
                            w0[u] = 0; w1[u] = 0; w2[u] = 0;
                            if (l < M && !p.lm_const[l] && (p.hll[l] + mu * p.diag_l[l] * p.diag_l[l] > 0)) {
                                const double *Wl = p.W + (size_t)l * BA_WS;
                                const double gl_h = p.y_l[l], sl = p.jscale_l[l];
                                w0[u] = wsc(sh, Wl, lane, sl) * gl_h;
                                w1[u] = wsc(sh, Wl, lane + 32, sl) * gl_h;
                                if (lane + 64 < ws) w2[u] = wsc(sh, Wl, lane + 64, sl) * gl_h;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) { a0 += w0[u]; a1 += w1[u]; a2 += w2[u]; }
                    }
                    part[warp * 96 + lane] = a0; part[warp * 96 + 32 + lane] = a1; part[warp * 96 + 64 + lane] = a2;
                    __syncthreads();
                    if (tid < ws) {
                        double t = 0;
#pragma unroll
                        for (int w = 0; w < BA_THREADS / 32; ++w) t += part[w * 96 + tid];
                        sh.y[wcol(tid)] -= t;
                    }
                }
                __syncthreads();
                // S = H + mu D^2 - W^T diag(1/h) W on the pose block (66 columns; + ex-pose and td when variable: ws = 73).
                // W is streamed through a shared-memory tile (landmarks x ws, pre-multiplied by 1/sqrt(h), rows padded to a
                // multiple of 3).  A thread owns one 3x3 block of the lower block-triangle and, when the thread count allows,
                // one of two interleaved landmark subsets: 6 shared-memory loads per 9 FMAs instead of 18 (the phase is
                // bound by shared-memory bandwidth).
                {
                    double *tile = reinterpret_cast<double *>(sh.imuJ);      // 4500 doubles available
                    const int nb3 = (ws + 2) / 3, wsp = 3 * nb3;              // 22 blocks / stride 66, or 25 / 75
                    const int tl = 4500 / wsp;                                // 68 or 60 landmarks per tile
                    const int nblk = nb3 * (nb3 + 1) / 2;                     // 253 or 325
                    const int ngrp = (2 * nblk <= BA_THREADS) ? 2 : 1;
                    const bool actv = tid < ngrp * nblk;
                    const int grp = tid / nblk, blk = tid - grp * nblk;
                    int bi = 0;
                    while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
                    const int bj = blk - bi * (bi + 1) / 2;
                    double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                    for (int t0 = 0; t0 < M; t0 += tl) {
                        const int nt = min(tl, M - t0);
                        for (int lr0 = warp; lr0 < nt; lr0 += 4 * nwarp) {
                            double wq[4][3];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {              // four rows' loads in flight together (wsp <= 75: three entries per lane)
                                const int lr = lr0 + u * nwarp;
#pragma unroll
                                for (int q = 0; q < 3; ++q) wq[u][q] = 0.0;
                                if (lr < nt) {
                                    const int l = t0 + lr;
                                    const double sc_ = p.lm_const[l] ? 0.0 : p.shinv_l[l], sl = p.jscale_l[l];
                                    const double *Wl = p.W + (size_t)l * BA_WS;
#pragma unroll
                                    for (int q = 0; q < 3; ++q) { const int k = lane + 32 * q; if (k < ws) wq[u][q] = wsc(sh, Wl, k, sl) * sc_; }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int lr = lr0 + u * nwarp;
                                if (lr < nt) {
#pragma unroll
                                    for (int q = 0; q < 3; ++q) { const int k = lane + 32 * q; if (k < wsp) tile[lr * wsp + k] = wq[u][q]; }
                                }
                            }
                        }
                        __syncthreads();
                        if (actv) {
                            const double *ta = tile + 3 * bi, *tb = tile + 3 * bj;
#pragma unroll 2
                            for (int l = grp; l < nt; l += ngrp) {
                                const double a0 = ta[l * wsp], a1 = ta[l * wsp + 1], a2 = ta[l * wsp + 2];
                                const double b0 = tb[l * wsp], b1 = tb[l * wsp + 1], b2 = tb[l * wsp + 2];
                                acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[0][2] += a0 * b2;
                                acc[1][0] += a1 * b0; acc[1][1] += a1 * b1; acc[1][2] += a1 * b2;
                                acc[2][0] += a2 * b0; acc[2][1] += a2 * b1; acc[2][2] += a2 * b2;
                            }
                        }
                        __syncthreads();
                    }
                    // the second landmark subset hands its block to the first through the (now idle) tile: one writer per entry
                    if (actv && grp == 1) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) tile[blk * 9 + 3 * i + j] = acc[i][j];
                    }
                    __syncthreads();
                    if (actv && grp == 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i)
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const int r = 3 * bi + i, c = 3 * bj + j;
                                const double a = ngrp == 2 ? acc[i][j] + tile[blk * 9 + 3 * i + j] : acc[i][j];
                                if (r < ws && c <= r) sh.H[pk(wcol(r), wcol(c))] -= a;
                            }
                    }
                }
                __syncthreads();      // the diagonal entries are touched again just below by other threads
                for (int c = tid; c < BA_NC; c += BA_THREADS) sh.H[pk(c, c)] += mu * sh.diag[c] * sh.diag[c];
                __syncthreads();
                TPROF(3);
                // in-place blocked Cholesky (lower, packed, panel width 8) of the 172x172 reduced camera system.
                // The right-hand side rides along as an extra row (index BA_NC): forward substitution for free.
                // Per panel: (2) one thread per row solves the panel's triangular system, (3) rank-8 update of the
                // trailing matrix by warps 1..15 (4 rows x 1 column register tiles) WHILE warp 0 updates and factorises
                // the next panel's 8x8 diagonal block (look-ahead: the serial pivot chain is off the critical path)
                // => 2 barriers / panel.
                bool bad = sh.flag[0] != 0;
                if (!bad) {
                    if (warp == 0) chol_diag8(sh, 0, min(8, BA_NC), lane, false);
                    __syncthreads();
                    bad = sh.flag[0] != 0;
                }
                const bool cprof = (m.debug & 32) != 0 && blockIdx.x == 0 && iterations == 1;
                long long cp_solve = 0, cp_work = 0, cp_wait1 = 0, cp_wait2 = 0, cp_t = cprof ? clock64() : 0;
#define CPROF(acc) do { if (cprof) { long long t_ = clock64(); acc += t_ - cp_t; cp_t = t_; } } while (0)
                for (int j0 = 0; j0 < BA_NC && !bad; j0 += 8) {
                    const int nbp = min(8, BA_NC - j0);
                    // (2) rows below the panel (incl. the rhs row): x * Ld^T = a ; the solved panel is also
                    //     staged transposed in Lp so that step (3) reads it without bank conflicts
                    for (int i2 = j0 + nbp + tid; i2 <= BA_NC; i2 += BA_THREADS) {
                        double *ri = (i2 < BA_NC) ? sh.H + i2 * (i2 + 1) / 2 : sh.y;
                        double x[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            x[c] = 0.0;
                            if (c < nbp) {
                                const double *rc = sh.H + (j0 + c) * (j0 + c + 1) / 2;
                                double t = ri[j0 + c];
#pragma unroll
                                for (int k = 0; k < 8; ++k) if (k < c) t -= x[k] * rc[j0 + k];
                                x[c] = t * sh.colv[c];
                                ri[j0 + c] = x[c];
                            }
                            sh.Lp[c][i2] = x[c];
                        }
                    }
                    CPROF(cp_solve);
                    __syncthreads();
                    CPROF(cp_wait1);
                    // (3) trailing update H[ii,kk] -= sum_c L[ii,j0+c] L[kk,j0+c]
                    const int t0 = j0 + nbp;
                    const int nb2 = min(8, BA_NC - t0);            // width of the next panel (<= 0: none)
                    if (warp == 0) {
                        if (nb2 > 0) chol_diag8(sh, t0, nb2, lane, true, cprof && (j0 == 16 || j0 == 80));
                    } else {
                        // FP64 tensor cores: the rank-8 update is cut into 8 x 8 tiles C(I, K) -= A(I) B(K)^T of the global
                        // 8-aligned tile grid, A(I)[m][p] = L[8 I + m][j0 + p] = Lp[p][8 I + m]; two mma.sync.m8n8k4.f64 (DMMA) per
                        // tile instead of 64 DFMA + ~300 address / predicate / shared-memory instructions.  Fragments (PTX ISA,
                        // m8n8k4 .f64): A, B: lane -> [lane >> 2][lane & 3]; C: lane -> row lane >> 2, columns 2 (lane & 3), +1.
                        // Tiles are dealt to the 15 warps cyclically over (7 I + K); entries outside the lower triangle / above
                        // the next diagonal block / beyond the rhs row (index BA_NC) are computed and discarded.
                        const int w1 = warp - 1, nw1 = BA_THREADS / 32 - 1;
                        const int R0 = t0 + max(nb2, 0);                   // first row of the trailing rows below the next diagonal block
                        const int m8 = lane >> 2, kq = lane & 3;
                        const int K0 = t0 >> 3, I0 = R0 >> 3, IL = BA_NC >> 3;       // tile rows I0 .. IL, tile row I holds tiles K0 .. I
                        // tiles in row-major order of the lower block triangle, tile t -> warp t % 15; (I, K) advance without divisions.
                        // (Measured: two tiles in flight per pass with split accumulators is not faster -- the phase is bound by the
                        // shared-memory pipe, not by the 26-cycle DMMA accumulator latency.)
                        int I = I0, K = K0 + w1;
                        while (I <= IL && K > I) { K -= I - K0 + 1; ++I; }
                        int Icur = -1, row = 0, cmax = 0;
                        bool rowok = false;
                        double a0 = 0, a1 = 0;
                        double *rp = sh.y;
                        while (I <= IL) {
                            if (I != Icur) {
                                Icur = I;
                                row = 8 * I + m8;
                                a0 = -sh.Lp[kq][row]; a1 = -sh.Lp[4 + kq][row];
                                rowok = row >= R0 && row <= BA_NC;
                                rp = row < BA_NC ? sh.H + row * (row + 1) / 2 : sh.y;
                                cmax = min(row, BA_NC - 1);
                            }
                            const int col = 8 * K + 2 * kq;
                            const double b0 = sh.Lp[kq][8 * K + m8], b1 = sh.Lp[4 + kq][8 * K + m8];
                            const bool v0 = rowok && col >= t0 && col <= cmax, v1 = rowok && col + 1 >= t0 && col + 1 <= cmax;
                            double c0 = v0 ? rp[col] : 0.0, c1 = v1 ? rp[col + 1] : 0.0;
                            dmma884(c0, c1, a0, b0);
                            dmma884(c0, c1, a1, b1);
                            if (v0) rp[col] = c0;
                            if (v1) rp[col + 1] = c1;
                            K += nw1;
                            while (I <= IL && K > I) { K -= I - K0 + 1; ++I; }
                        }
                    }
                    CPROF(cp_work);
                    __syncthreads();
                    CPROF(cp_wait2);
                    if (sh.flag[0]) { bad = true; break; }
                }
                if (cprof && lane == 0 && (warp == 0 || warp == 1 || warp == 8 || warp == 15))
                    printf("chol warp %2d: panel solve %lld  wait %lld  %s %lld  wait %lld cycles (22 panels)\n", warp, cp_solve, cp_wait1,
                           warp == 0 ? "diag8" : "trailing", cp_work, cp_wait2);
#undef CPROF
                TPROF(4);
                if (!bad) {
                    // back substitution L^T x = z by one warp with the solution vector in registers (lane holds entries
                    // lane, lane + 32, ...): per step one multiply by the stored reciprocal pivot, one broadcast and one
                    // FMA per register; the rows of L are prefetched one step ahead.
                    for (int c = tid; c < BA_NC; c += BA_THREADS) sh.colv[c] = 1.0 / sh.H[pk(c, c)];
                    __syncthreads();
                    if (warp == 0) {
                        constexpr int NS = (BA_NC + 31) / 32;
                        double yv[NS], cur[NS], nxt[NS];
#pragma unroll
                        for (int q = 0; q < NS; ++q) { const int k = lane + 32 * q; yv[q] = k < BA_NC ? sh.y[k] : 0.0; }
                        {
                            const double *row = sh.H + (BA_NC - 1) * BA_NC / 2;
#pragma unroll
                            for (int q = 0; q < NS; ++q) { const int k = lane + 32 * q; cur[q] = k < BA_NC - 1 ? row[k] : 0.0; }
                        }
                        double rcur = sh.colv[BA_NC - 1];
#pragma unroll
                        for (int mq = NS - 1; mq >= 0; --mq) {
                            const int ihi = min(32 * mq + 31, BA_NC - 1);
                            for (int i = ihi; i >= 32 * mq; --i) {
                                // prefetch row i-1 (entries k < i-1)
                                if (i > 0) {
                                    const double *row = sh.H + (i - 1) * i / 2;
#pragma unroll
                                    for (int q = 0; q <= mq; ++q) { const int k = lane + 32 * q; nxt[q] = k < i - 1 ? row[k] : 0.0; }
                                }
                                const double rnx = sh.colv[i > 0 ? i - 1 : 0];
                                const double xi = __shfl_sync(0xffffffffu, yv[mq] * rcur, i & 31);
                                if (lane == (i & 31)) yv[mq] = xi;
#pragma unroll
                                for (int q = 0; q <= mq; ++q) {
                                    const int k = lane + 32 * q;
                                    if (k < i) yv[q] -= cur[q] * xi;
                                }
#pragma unroll
                                for (int q = 0; q <= mq; ++q) cur[q] = nxt[q];
                                rcur = rnx;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < NS; ++q) { const int k = lane + 32 * q; if (k < BA_NC) sh.y[k] = yv[q]; }
                    }
                    __syncthreads();
                    double fin = 0;
                    for (int c = tid; c < BA_NC; c += BA_THREADS) if (!isfinite(sh.y[c])) fin += 1;
                    fin = block_sum(fin, sh.red);
                    if (fin == 0) {
                        // back-substitute the landmarks: y_l = (g_l - w_l . y_c) / h_l
                        for (int l0 = warp; l0 < M; l0 += 4 * nwarp) {
                            double v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {              // the four rows' loads in flight together
                                const int l = l0 + u * nwarp;
                                v[u] = 0;
                                if (l < M && !p.lm_const[l]) {
                                    const double *Wl = p.W + (size_t)l * BA_WS;
                                    const double sl = p.jscale_l[l];
                                    for (int k = lane; k < ws; k += 32) v[u] += wsc(sh, Wl, k, sl) * sh.y[wcol(k)];
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = warp_sum_d(v[u]);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int l = l0 + u * nwarp;
                                if (l < M && lane == 0) p.y_l[l] = p.lm_const[l] ? 0.0 : (p.gl[l] - v[u]) * p.hinv_l[l];
                            }
                        }
                        solved = true;
                    }
                }
                __syncthreads();
                if (solved) break;
                // failed: the in-place factorisation destroyed H => re-linearise and retry with a larger mu
                mu *= mu_inc;
                ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, 1);
                ba_scale(m, p, sh);
            }
            TPROF(5);
            if (!solved) { status = VRF_SOFT_NOT_SPD; step_ok = false; }
            else {
                // gauss_newton_step_ = -diag * y ; scalar products needed by the dogleg model
                double a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
                for (int c = tid; c < BA_NC; c += BA_THREADS) {
                    const double yc = sh.y[c], dg = sh.diag[c];
                    sh.gn[c] = -dg * yc;
                    a1 += yc * sh.g[c]; a2 += dg * dg * yc * yc; a3 += dg * dg * sh.tmp[c] * yc;
                    a4 += sh.gn[c] * sh.gn[c]; a5 += sh.gd[c] * sh.gn[c];
                }
                for (int l = tid; l < M; l += BA_THREADS) {
                    if (p.lm_const[l]) { p.gn_l[l] = 0; continue; }
                    const double yc = p.y_l[l], dg = p.diag_l[l];
                    const double gnv = -dg * yc;
                    p.gn_l[l] = gnv;
                    a1 += yc * p.gl[l]; a2 += dg * dg * yc * yc; a3 += dg * dg * p.u_l[l] * yc;
                    a4 += gnv * gnv; a5 += p.gd_l[l] * gnv;
                }
                ytg = block_sum(a1, sh.red); yDy = block_sum(a2, sh.red); uDy = block_sum(a3, sh.red);
                gnn2 = block_sum(a4, sh.red); gdgn = block_sum(a5, sh.red);
            }
        }
        if (step_ok) {
            // ComputeTraditionalDoglegStep: step_d = c1 * gd + c2 * gn (dogleg space), step = step_d / diag
            const double gnorm = sqrt(gdn2), gnn = sqrt(gnn2);
            double c1, c2;
            if (gnn <= radius) { c1 = 0; c2 = 1; dogleg_norm = gnn; }
            else if (gnorm * alpha >= radius) { c1 = -(radius / gnorm); c2 = 0; dogleg_norm = radius; }
            else {
                const double b_dot_a = -alpha * gdgn;
                const double a_sq = (alpha * gnorm) * (alpha * gnorm);
                const double bma = a_sq - 2 * b_dot_a + gnn * gnn;
                const double cc = b_dot_a - a_sq;
                const double dd = sqrt(cc * cc + bma * (radius * radius - a_sq));
                const double beta = (cc <= 0) ? (dd - cc) / bma : (radius * radius - a_sq) / (dd + cc);
                c1 = -alpha * (1.0 - beta); c2 = beta;
                dogleg_norm = sqrt(c1 * c1 * gdn2 + 2 * c1 * c2 * gdgn + c2 * c2 * gnn2);
            }
            // model_cost_change = -s^T g - 0.5 s^T H s with s = c1 u - c2 y (u = D^-2 g), all from stored scalars:
            //   u^T H u = Quu ; y^T H y = y^T g - mu y^T D^2 y ; u^T H y = u^T g - mu u^T D^2 y
            const double sTg = c1 * utg - c2 * ytg;
            const double sHs = c1 * c1 * Quu - 2 * c1 * c2 * (utg - mu * uDy) + c2 * c2 * (ytg - mu * yDy);
            const double mcc = -sTg - 0.5 * sHs;
            if (!(mcc > 0.0)) step_ok = false;
            else {
                invalid = 0;
                // candidate = x (+) (step * jscale)
                ba_make_candidate(m, p, sh, c1, c2, 1.0);
                TPROF(6);
                double cand_cost = ba_evaluate(m, p, sh, sh.cpose, sh.csb, sh.cex, sh.tdv[1], p.clam, 0);
                TPROF(7);
                // Ceres' projected Armijo line search of bound-constrained problems: the full step passing
                // f(x [+] delta) <= f(x) + 1e-4 g.delta is the common case and costs nothing (ba_line_search above)
                if (constrained && (!isfinite(cand_cost) || cand_cost > x_cost + 1e-4 * sTg)) {
                    ++armijo_failures;
                    ba_line_search(m, p, sh, c1, c2, x_cost, sTg, &cand_cost);
                }
                // step norm over the non-constant blocks (ambient space)
                double sn = 0;
                for (int i = tid; i < BA_NF * 7; i += BA_THREADS) if (col_active_dev(m, 6 * (i / 7))) { double dd = sh.pose[i] - sh.cpose[i]; sn += dd * dd; }
                for (int i = tid; i < BA_NF * 9; i += BA_THREADS) if (col_active_dev(m, 66 + 9 * (i / 9))) { double dd = sh.sb[i] - sh.csb[i]; sn += dd * dd; }
                for (int l = tid; l < M; l += BA_THREADS) if (!p.lm_const[l]) { double dd = p.lam[l] - p.clam[l]; sn += dd * dd; }
                if (m.ex_active && tid < 7) { double dd = sh.ex[tid] - sh.cex[tid]; sn += dd * dd; }
                if (m.td_active && tid == 7) { double dd = sh.tdv[0] - sh.tdv[1]; sn += dd * dd; }
                const double step_norm = sqrt(block_sum(sn, sh.red));
                if (step_norm <= 1e-8 * (x_norm + 1e-8)) { termination = 3; break; }
                const double cost_change = x_cost - cand_cost;
                if (fabs(cost_change) <= 1e-6 * x_cost) { termination = 1; break; }
                const double rho = cost_change / mcc;
                if ((m.debug & 1) && tid == 0 && blockIdx.x == 0)
                    printf("gpu it %d cost %.6f cand %.6f mcc %.6g rho %.4f radius %.3g |step| %.3g mu %.1e gradmax %.3g c1 %.4g c2 %.4g\n",
                           iterations, x_cost, cand_cost, mcc, rho, radius, dogleg_norm, mu, gradient_max, c1, c2);
                if (rho > 1e-3) {
                    for (int i = tid; i < BA_NF * 7; i += BA_THREADS) sh.pose[i] = sh.cpose[i];
                    for (int i = tid; i < BA_NF * 9; i += BA_THREADS) sh.sb[i] = sh.csb[i];
                    for (int l = tid; l < M; l += BA_THREADS) p.lam[l] = p.clam[l];
                    if (tid < 7) sh.ex[tid] = sh.cex[tid];
                    if (tid == 7) sh.tdv[0] = sh.tdv[1];
                    __syncthreads();
                    x_cost = cand_cost;
                    x_norm = sqrt(xnorm2(sh.pose, sh.sb, p.lam));
                    TPROF(6);
                    ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, 1);
                    TPROF(0);
                    need_scale = 1;
                    ++successful;
                    if (rho < 0.25) radius *= 0.5;
                    if (rho > 0.75) radius = fmax(radius, 3.0 * dogleg_norm);
                    mu = fmax(min_mu, 2.0 * mu / mu_inc);
                    reuse = 0;
                } else {
                    radius *= 0.5;
                    reuse = 1;
                }
                continue;
            }
        }
        if (++invalid >= 5) { termination = 5; break; }
        mu *= mu_inc;
        reuse = 0;
        if (status == VRF_SOFT_NOT_SPD && mu >= max_mu) { termination = 5; break; }
        // H was consumed by the failed factorisation attempts: rebuild it
        ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, 1);
        need_scale = 1;
    }
    __syncthreads();
    // ---- double2vector gauge fix (estimator.cpp:985-1111) + vector2double re-pack for marginalization ----
    if (tid == 0) {
        double R0[9], R00[9], ypr0[3], ypr00[3], rot[9];
        d_q2R(p.pose0 + 3, R0);
        d_R2ypr(R0, ypr0);
        if (m.use_imu) {
            d_q2R(sh.pose + 3, R00);
            d_R2ypr(R00, ypr00);
            const double yd = (ypr0[0] - ypr00[0]) / 180.0 * 3.14159265358979323846;
            rot[0] = cos(yd); rot[1] = -sin(yd); rot[2] = 0; rot[3] = sin(yd); rot[4] = cos(yd); rot[5] = 0; rot[6] = 0; rot[7] = 0; rot[8] = 1;
            if (fabs(fabs(ypr0[1]) - 90) < 1.0 || fabs(fabs(ypr00[1]) - 90) < 1.0) {
                double R00T[9] = {R00[0], R00[3], R00[6], R00[1], R00[4], R00[7], R00[2], R00[5], R00[8]};
                d_mm(R0, R00T, rot);
            }
        } else { for (int k = 0; k < 9; ++k) rot[k] = (k % 4 == 0) ? 1.0 : 0.0; }
        for (int k = 0; k < 9; ++k) sh.tmp[k] = rot[k];
    }
    __syncthreads();
    for (int f = tid; f < BA_NF; f += BA_THREADS) {
        const double *rot = sh.tmp;
        double q[4] = {sh.pose[7 * f + 3], sh.pose[7 * f + 4], sh.pose[7 * f + 5], sh.pose[7 * f + 6]}, R[9], Rf[9];
        double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        q[0] /= nq; q[1] /= nq; q[2] /= nq; q[3] /= nq;
        d_q2R(q, R);
        if (m.use_imu) {
            d_mm(rot, R, Rf);
            double dd[3] = {sh.pose[7 * f] - sh.pose[0], sh.pose[7 * f + 1] - sh.pose[1], sh.pose[7 * f + 2] - sh.pose[2]}, t[3], v[3];
            d_mv(rot, dd, t);
            d_mv(rot, sh.sb + 9 * f, v);
            for (int k = 0; k < 3; ++k) {
                out.Ps[3 * f + k] = t[k] + p.pose0[k]; out.Vs[3 * f + k] = v[k];
                out.Bas[3 * f + k] = sh.sb[9 * f + 3 + k]; out.Bgs[3 * f + k] = sh.sb[9 * f + 6 + k];
            }
        } else {
            for (int k = 0; k < 9; ++k) Rf[k] = R[k];
            for (int k = 0; k < 3; ++k) { out.Ps[3 * f + k] = sh.pose[7 * f + k]; out.Vs[3 * f + k] = 0; out.Bas[3 * f + k] = 0; out.Bgs[3 * f + k] = 0; }
        }
        for (int k = 0; k < 9; ++k) out.Rs[9 * f + k] = Rf[k];
        // vector2double (estimator.cpp:936-981)
        for (int k = 0; k < 3; ++k) out.mpose[7 * f + k] = out.Ps[3 * f + k];
        d_R2q(Rf, out.mpose + 7 * f + 3);
        for (int k = 0; k < 3; ++k) {
            out.msb[9 * f + k] = m.use_imu ? out.Vs[3 * f + k] : sh.sb[9 * f + k];
            out.msb[9 * f + 3 + k] = sh.sb[9 * f + 3 + k]; out.msb[9 * f + 6 + k] = sh.sb[9 * f + 6 + k];
        }
    }
    if (tid == 0) {
        for (int k = 0; k < 3; ++k) out.mex[k] = sh.ex[k];
        if (m.use_imu) {
            double q[4] = {sh.ex[3], sh.ex[4], sh.ex[5], sh.ex[6]}, R[9];
            double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            q[0] /= nq; q[1] /= nq; q[2] /= nq; q[3] /= nq;
            d_q2R(q, R); d_R2q(R, out.mex + 3);
        } else for (int k = 3; k < 7; ++k) out.mex[k] = sh.ex[k];
    }
    // setDepth / getDepthVector round trip of the inverse depths (feature_manager.cpp:197-223,302-324)
    for (int l = tid; l < M; l += BA_THREADS) { double depth = 1.0 / p.lam[l]; p.clam[l] = 1.0 / depth; p.lam_out[l] = p.lam[l]; }
    __syncthreads();
    // ---- outputs ----
    double nf_ = 0;
    for (int i = tid; i < BA_NF * 7; i += BA_THREADS) { out.pose[i] = sh.pose[i]; if (!isfinite(sh.pose[i])) nf_ += 1; }
    nf_ = block_sum(nf_, sh.red);
    if (nf_ > 0) status = VRF_SOFT_NONFINITE;
    for (int i = tid; i < BA_NF * 9; i += BA_THREADS) out.sb[i] = sh.sb[i];
    if (tid < 7) out.ex[tid] = sh.ex[tid];
    if (tid == 7) { out.td = sh.tdv[0]; out.mtd = sh.tdv[0]; }
    __syncthreads();
    if (tid == 0) {
        out.status = status; out.iterations = iterations; out.successful = successful; out.termination = termination;
        out.armijo_failures = armijo_failures;
        out.initial_cost = initial_cost; out.final_cost = x_cost;
        for (int k = 0; k < 8; ++k) out.prof[k] = tprof[k];
        out.prof2[7] = clock64() - t_kernel0;
    }
}

size_t ba_solve_smem_bytes() { return sizeof(BaShared); }

// The opt-in for > 48 KB of dynamic shared memory is a per-device attribute: set by ba_create() for the handle's device
// (a process may open handles on several GPUs).
int ba_solve_configure()
{
    return cudaFuncSetAttribute(k_ba_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BaShared)) == cudaSuccess ? 0 : -1;
}

int ba_solve_launch(const BaMeta *d_meta, const BaProbDev *d_prob, BaOutDev *d_out, int n, LaunchCtx &lc)
{
    lc.begin(K_BA_SOLVE);
    k_ba_solve<<<n, BA_THREADS, sizeof(BaShared), lc.st>>>(d_meta, d_prob, d_out);
    lc.end();
    return 0;
}

}  // namespace vrf
