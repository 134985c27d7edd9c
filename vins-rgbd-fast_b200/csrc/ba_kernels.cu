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
// consecutive words): lam[M], start[M], flag[M], obs_ptr[M+1], obs[O][2]; W[M][66]
// (landmark-to-pose coupling rows), hll/gl/diag/gd/gn/step per landmark.
// On chip: the 171x171 camera system lives in shared memory as a packed lower
// triangle (117 KB) and is Schur-reduced and Cholesky-factorised in place; the
// tangent layout is [pose f: 6f | speed-bias f: 66+9f | ex-pose: 165].
//
// Roofline: ~100-160 KB touched and ~5 MFLOP FP64 per iteration per problem => compute /
// latency bound; the only dense contraction is the 171^3/3 Cholesky.
#include "ba_math.cuh"

namespace vrf {

// --------------------------------------------------------------------------
// shared-memory frame of the solve kernel
// --------------------------------------------------------------------------
struct BaShared {
    double H[BA_NC * (BA_NC + 1) / 2];     // packed lower triangle: camera system, Schur-reduced & factorised in place
    double g[BA_NC], diag[BA_NC], gd[BA_NC], gn[BA_NC], jscale[BA_NC], tmp[BA_NC], y[BA_NC + 1], colv[BA_NC + 1];
    double Lp[8][BA_NC + 5];                 // current Cholesky panel, transposed: Lp[c][row] (bank-conflict-free trailing update)
    double pose[BA_NF * 7], sb[BA_NF * 9], ex[7];
    double cpose[BA_NF * 7], csb[BA_NF * 9], cex[7];  // candidate
    double tdv[2];                                    // para_Td: current, candidate
    double R[BA_NF * 9], ric[9];
    double dx[VRF_PRIOR_MAX_DIM], pr[VRF_PRIOR_MAX_DIM];
    double imuJ[BA_NF - 1][15 * 30];                 // whitened IMU Jacobians of the current linearisation
    double imur[BA_NF - 1][16];
    double red[BA_THREADS / 32];
    double sc[16];
    int flag[8];
};

__device__ __forceinline__ bool col_active_dev(const BaMeta &m, int col)
{
    if (col < 66) { int f = col / 6; return f < m.nframes && !(f == 0 && !m.use_imu); }
    if (col < BA_COL_EX) { int f = (col - 66) / 9; return m.use_imu && f < m.nframes; }
    return col == BA_COL_TD ? m.td_active != 0 : m.ex_active != 0;
}

// Contributions of one batch of projection factors of the pair (host i, observer j) to the rows of the "global"
// block g = [ex-pose (6) | td (1)] (tangent columns 165..171), only when one of them is variable:
// Jg^T Ji, Jg^T Jj, Jg^T r and the lower triangle of Jg^T Jg, each summed over the warp's lanes with the
// reduce-scatter butterfly and added to the shared system.  Jg: 2 x 7 row-major.
__device__ __forceinline__ double reduce_scatter16(double (&v)[16], int lane);
__device__ __noinline__ void accumulate_g(double *H, double *g, int i, int j, int lane, const double *Ji, const double *Jj,
                                          const double *Jg, const double *r)
{
#pragma unroll 1
    for (int q = 0; q < 7; ++q) {
        double v[16];
        const double g0 = Jg[q], g1 = Jg[7 + q];
#pragma unroll
        for (int c = 0; c < 6; ++c) { v[c] = g0 * Ji[c] + g1 * Ji[6 + c]; v[6 + c] = g0 * Jj[c] + g1 * Jj[6 + c]; }
        v[12] = g0 * r[0] + g1 * r[1]; v[13] = 0; v[14] = 0; v[15] = 0;
        const double s_ = reduce_scatter16(v, lane);
        if (!(lane & 1) && s_ != 0.0) {
            const int e = lane >> 1;
            if (e < 6) atomicAdd(&H[pk(BA_COL_EX + q, 6 * i + e)], s_);
            else if (e < 12) atomicAdd(&H[pk(BA_COL_EX + q, 6 * j + e - 6)], s_);
            else if (e == 12) atomicAdd(&g[BA_COL_EX + q], s_);
        }
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
        const double s_ = reduce_scatter16(v, lane);
        const int e = rd * 16 + (lane >> 1);
        if (!(lane & 1) && e < 28 && s_ != 0.0) {
            int a = 0;
            while ((a + 1) * (a + 2) / 2 <= e) ++a;
            atomicAdd(&H[pk(BA_COL_EX + a, BA_COL_EX + e - a * (a + 1) / 2)], s_);
        }
    }
}

#define BA_NPAIR (BA_NF * (BA_NF - 1) / 2)
static_assert(BA_NPAIR * 54 <= (BA_NF - 1) * 15 * 30, "pair partials must fit in imuJ");

__device__ __forceinline__ int pair_index(int i, int j) { return i * (2 * BA_NF - i - 1) / 2 + (j - i - 1); }

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

// cost of all residual blocks at (pose, sb, lam); optionally the full linearisation into sh.H / sh.g / landmark arrays.
// __noinline__: the solve loop calls this from five sites; inlining produced a 51k-instruction kernel (800 KB of SASS)
// that thrashed the instruction cache (one resident CTA per SM, 16 warps in different code regions).
// GACT: ex-pose and/or td variable (two instantiations so that the common constant-extrinsic path keeps its registers).
template <bool GACT>
__device__ __noinline__ double ba_evaluate_t(const BaMeta &m, const BaProbDev &p, BaShared &sh, const double *pose, const double *sb,
                                const double *ex, const double td, const double *lam, bool lin)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    if (lin) {
        for (int i = tid; i < BA_NC * (BA_NC + 1) / 2; i += BA_THREADS) sh.H[i] = 0.0;
        for (int i = tid; i < BA_NC; i += BA_THREADS) sh.g[i] = 0.0;
        for (int i = tid; i < BA_NPAIR * 54; i += BA_THREADS) (&sh.imuJ[0][0])[i] = 0.0;      // per-pair partials (see below)
        for (int i = tid; i < m.M * BA_WS; i += BA_THREADS) p.W[i] = 0.0;
        for (int l = tid; l < m.M; l += BA_THREADS) { p.hll[l] = 0.0; p.gl[l] = 0.0; }
    }
    for (int f = tid; f < BA_NF; f += BA_THREADS) d_q2R(pose + 7 * f + 3, sh.R + 9 * f);
    if (tid == 0) d_q2R(ex + 3, sh.ric);
    constexpr bool gact = GACT;
    const bool eprof = (m.debug & 16) != 0;
    long long et[5] = {0, 0, 0, 0, 0}, ec = eprof ? clock64() : 0;
#define EPROF(k) do { if (eprof) { long long t_ = clock64(); et[k] += t_ - ec; ec = t_; } } while (0)
    __syncthreads();
    EPROF(0);
    double cost = 0.0;
    if (!lin) {
        // ---- projection factors, cost only: two landmarks per warp (16 lanes each), one lane per factor ----
        const int half = lane >> 4, hl = lane & 15;
        for (int lp = warp; 2 * lp < m.M; lp += nwarp) {
            const int l = 2 * lp + half;
            if (l >= m.M) continue;
            const int o0 = p.obs_ptr[l], nf = p.obs_ptr[l + 1] - o0 - 1;
            if (hl >= nf) continue;
            const int i = p.start[l], j = i + 1 + hl;
            double r[2], Ji[12], Jj[12], Jl[2], xi, yi, xj, yj;
            obs_at(m, p, o0, td, xi, yi);
            obs_at(m, p, o0 + 1 + hl, td, xj, yj);
            cost += 0.5 * proj_eval(pose + 7 * i, sh.R + 9 * i, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], xi, yi, xj, yj, false,
                                    p.lm_const[l] != 0, r, Ji, Jj, Jl);
        }
    } else {
        // ---- projection factors, linearisation: one warp per frame pair (host i, observer j), one lane per factor ----
        // Every factor of the pair adds to the same 12x12 block of J^T J.  The 90 distinct sums (off-diagonal 6x6,
        // two diagonal lower triangles, two gradient 6-vectors) are reduced over the warp's lanes with a
        // reduce-scatter butterfly (one shuffle per value instead of five) and accumulated in registers over the
        // pair's batches: no shared-memory atomics.  The off-diagonal block has a single owner and is stored
        // directly; the diagonal parts go to per-pair partials that are summed per frame afterwards.
        double *part = &sh.imuJ[0][0];              // [55][54]; imuJ is rewritten by the IMU phase below
        const int nbatch = (m.M + 31) >> 5;
        for (int pi = warp; pi < BA_NPAIR; pi += nwarp) {
            int i = 0, rem = pi;
            while (rem >= BA_NF - 1 - i) { rem -= BA_NF - 1 - i; ++i; }
            const int j = i + 1 + rem;
            double acc[6] = {0, 0, 0, 0, 0, 0};
            bool any_pair = false;
            for (int bt = 0; bt < nbatch; ++bt) {
                const int l = 32 * bt + lane;
                bool act = false;
                int o0 = 0;
                if (l < m.M && p.start[l] == i) {
                    o0 = p.obs_ptr[l];
                    act = (p.obs_ptr[l + 1] - o0 - 1) >= (j - i);
                }
                if (!__any_sync(0xffffffffu, act)) continue;
                any_pair = true;
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
                        for (int q = 0; q < 7; ++q) atomicAdd(&Wl[66 + q], Jg[q] * Jl[0] + Jg[7 + q] * Jl[1]);
                    }
                    cost += 0.5 * rho0;
                    // landmark rows: the observer part has one contributor, the rest sums over the landmark's factors
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        Wl[6 * j + a] = Jj[a] * Jl[0] + Jj[6 + a] * Jl[1];
                        atomicAdd(&Wl[6 * i + a], Ji[a] * Jl[0] + Ji[6 + a] * Jl[1]);
                    }
                    atomicAdd(&p.hll[l], Jl[0] * Jl[0] + Jl[1] * Jl[1]);
                    atomicAdd(&p.gl[l], Jl[0] * r[0] + Jl[1] * r[1]);
                }
                if (gact) accumulate_g(sh.H, sh.g, i, j, lane, Ji, Jj, Jg, r);
#pragma unroll
                for (int ps = 0; ps < 6; ++ps) {
                    double v[16];
#pragma unroll
                    for (int t = 0; t < 16; ++t) v[t] = pair_entry(ps * 16 + t, Ji, Jj, r);
                    acc[ps] += reduce_scatter16(v, lane);
                }
            }
            if (any_pair && !(lane & 1)) {
#pragma unroll
                for (int ps = 0; ps < 6; ++ps) {
                    const int e = ps * 16 + (lane >> 1);
                    if (e < 36) sh.H[pk(6 * j + e / 6, 6 * i + e % 6)] = acc[ps];
                    else if (e < 90) part[pi * 54 + (e - 36)] = acc[ps];
                }
            }
        }
        __syncthreads();
        // diagonal blocks and gradient of frame f: sum of the partials of every pair f takes part in
        for (int t = tid; t < BA_NF * 27; t += BA_THREADS) {
            const int f = t / 27, q = t - f * 27;
            double sum = 0;
            for (int i = 0; i < f; ++i) sum += part[pair_index(i, f) * 54 + (q < 21 ? q : 42 + (q - 21))];            // f observes
            for (int j = f + 1; j < BA_NF; ++j) sum += part[pair_index(f, j) * 54 + (q < 21 ? 21 + q : 48 + (q - 21))];   // f hosts
            if (q < 21) {
                int a = 0;
                while ((a + 1) * (a + 2) / 2 <= q) ++a;
                sh.H[pk(6 * f + a, 6 * f + (q - a * (a + 1) / 2))] = sum;
            } else
                sh.g[6 * f + (q - 21)] = sum;
        }
        __syncthreads();
    }
    EPROF(1);
    // ---- IMU factors: one warp per factor ----
    for (int f = warp; f < m.nimu; f += nwarp) {
        const int j = m.imu_j[f], i = j - 1;
        const VrfImuPreint *pre = p.imu + (j - 1);
        const double *S = p.imuS + (size_t)(j - 1) * 225;
        double rr[15];
        ImuCtx cx;
        imu_residual_raw(pre, pose + 7 * i, sb + 9 * i, pose + 7 * j, sb + 9 * j, m.g_norm, rr, &cx);
        // whitened residual row `lane`
        double rw = 0;
        if (lane < 15) { for (int k = lane; k < 15; ++k) rw += S[lane * 15 + k] * rr[k]; cost += 0.5 * rw * rw; }
        if (!lin) continue;
        if (lane < 15) sh.imur[f][lane] = rw;
        if (lane < 30) {
            double col[15];
            imu_jac_col(pre, pose + 7 * i, sb + 9 * i, pose + 7 * j, sb + 9 * j, m.g_norm, &cx, lane, col);
            for (int rI = 0; rI < 15; ++rI) { double a = 0; for (int k = rI; k < 15; ++k) a += S[rI * 15 + k] * col[k]; sh.imuJ[f][rI * 30 + lane] = a; }
        }
        __syncwarp();
        // accumulate J^T J and J^T r ; tangent column of local column c
        auto tcol = [&](int c) { return c < 6 ? 6 * i + c : c < 15 ? 66 + 9 * i + (c - 6) : c < 21 ? 6 * j + (c - 15) : 66 + 9 * j + (c - 21); };
        for (int e = lane; e < 30 * 31 / 2; e += 32) {
            int a = 0, rem = e;
            while (rem > a) { rem -= (a + 1); ++a; }       // e -> (a, b<=a) in the 30x30 lower triangle
            int b = rem;
            double h = 0;
            for (int k = 0; k < 15; ++k) h += sh.imuJ[f][k * 30 + a] * sh.imuJ[f][k * 30 + b];
            atomicAdd(&sh.H[pk(tcol(a), tcol(b))], h);
        }
        if (lane < 30) {
            double gsum = 0;
            for (int k = 0; k < 15; ++k) gsum += sh.imuJ[f][k * 30 + lane] * sh.imur[f][k];
            atomicAdd(&sh.g[tcol(lane)], gsum);
        }
    }
    EPROF(2);
    __syncthreads();
    EPROF(3);
    // ---- prior ----
    const BaPriorStore *P = p.prior;
    const int np_ = (P && P->valid) ? P->n : 0;
    if (np_ > 0) {
        prior_dx(P, pose, sb, ex, td, sh.dx);
        __syncthreads();
        const int n = np_;
        for (int rI = tid; rI < n; rI += BA_THREADS) {
            double a = P->r0[rI];
            const double *row = P->J0 + (size_t)rI * n;
            for (int k = 0; k < n; ++k) a += row[k] * sh.dx[k];
            sh.pr[rI] = a;
            cost += 0.5 * a * a;
        }
        __syncthreads();
        if (lin) {
            // g += J0^T r ; H += J0^T J0 (precomputed HP), both through the column map
            for (int a = tid; a < n; a += BA_THREADS) {
                int ca = p.colmap[a];
                if (ca < 0) continue;
                double gsum = 0;
                for (int rI = 0; rI < n; ++rI) gsum += P->J0[(size_t)rI * n + a] * sh.pr[rI];
                sh.g[ca] += gsum;
            }
            for (int e = tid; e < n * n; e += BA_THREADS) {
                int a = e / n, b = e - a * n;
                if (b > a) continue;
                int ca = p.colmap[a], cb = p.colmap[b];
                if (ca < 0 || cb < 0) continue;
                sh.H[pk(ca, cb)] += p.HP[(size_t)a * n + b];
            }
        }
    }
    __syncthreads();
    EPROF(4);
    if (eprof && blockIdx.x == 0 && (tid == 0 || tid == 320 || tid == 500))
        printf("evaluate lin=%d tid=%d zero=%lld proj=%lld imu=%lld wait=%lld prior=%lld\n", (int)lin, tid, et[0], et[1], et[2], et[3], et[4]);
#undef EPROF
    return block_sum(cost, sh.red);
}

__device__ __forceinline__ double ba_evaluate(const BaMeta &m, const BaProbDev &p, BaShared &sh, const double *pose, const double *sb,
                                              const double *ex, const double td, const double *lam, bool lin)
{
    return (m.ex_active || m.td_active) ? ba_evaluate_t<true>(m, p, sh, pose, sb, ex, td, lam, lin)
                                        : ba_evaluate_t<false>(m, p, sh, pose, sb, ex, td, lam, lin);
}

// scale a fresh linearisation: H_s = D H D, g_s = D g, W_s, hll_s, gl_s (Jacobi scaling of the Jacobian columns)
__device__ __noinline__ void ba_scale(const BaMeta &m, const BaProbDev &p, BaShared &sh)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    for (int a = tid; a < BA_NC; a += BA_THREADS) {          // one packed row per thread
        double *row = sh.H + a * (a + 1) / 2;
        const double sa = sh.jscale[a];
        for (int b = 0; b <= a; ++b) row[b] *= sa * sh.jscale[b];
    }
    for (int c = tid; c < BA_NC; c += BA_THREADS) sh.g[c] *= sh.jscale[c];
    for (int l = warp; l < m.M; l += nwarp) {
        const double sl = p.jscale_l[l];
        double *Wl = p.W + (size_t)l * BA_WS;
        for (int k = lane; k < BA_WS; k += 32) Wl[k] *= sl * sh.jscale[wcol(k)];
        if (lane == 0) { p.hll[l] *= sl * sl; p.gl[l] *= sl; }
    }
    __syncthreads();
}

// (sum over camera + landmark entries of a_c*b_c) helper: camera part from smem arrays, landmark part from global
__device__ double dot_full(const BaMeta &m, const double *ac, const double *bc, const double *al, const double *bl, double *s_red)
{
    double v = 0;
    for (int i = threadIdx.x; i < BA_NC; i += BA_THREADS) v += ac[i] * bc[i];
    for (int l = threadIdx.x; l < m.M; l += BA_THREADS) v += al[l] * bl[l];
    return block_sum(v, s_red);
}

__global__ void __launch_bounds__(BA_THREADS, 1)
k_ba_solve(const BaMeta *metas, const BaProbDev *probs, BaOutDev *outs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BaShared &sh = *reinterpret_cast<BaShared *>(smem_raw);
    const BaMeta m = metas[blockIdx.x];
    const BaProbDev p = probs[blockIdx.x];
    BaOutDev &out = outs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    const int M = m.M;
    const long long t_kernel00 = clock64();

    for (int i = tid; i < BA_NF * 7; i += BA_THREADS) sh.pose[i] = p.pose0[i];
    for (int i = tid; i < BA_NF * 9; i += BA_THREADS) sh.sb[i] = p.sb0[i];
    if (tid < 7) sh.ex[tid] = p.ex0[tid];
    if (tid == 0) { sh.tdv[0] = *p.td0; sh.tdv[1] = *p.td0; }
    for (int l = tid; l < M; l += BA_THREADS) p.lam[l] = p.lam0[l];
    const int ws = (m.ex_active || m.td_active) ? BA_WS : 66;      // used width of the landmark coupling rows
    __syncthreads();
    // ---- once per solve: IMU information square roots, prior normal matrix ----
    {
        double *scr = reinterpret_cast<double *>(sh.imuJ) + warp * 450;   // 10 warps x 450 doubles fit in imuJ
        if (warp < m.nimu && warp < BA_NF - 1) imu_sqrt_info_warp(p.imu[m.imu_j[warp] - 1].covariance, p.imuS + (size_t)(m.imu_j[warp] - 1) * 225, scr);
        for (int f = warp + nwarp; f < m.nimu; f += nwarp) imu_sqrt_info_warp(p.imu[m.imu_j[f] - 1].covariance, p.imuS + (size_t)(m.imu_j[f] - 1) * 225, scr);
        const BaPriorStore *P = p.prior;
        const int n = (P && P->valid) ? P->n : 0;
        for (int e = tid; e < n * n; e += BA_THREADS) {
            int a = e / n, b = e - a * n;
            if (b > a) continue;
            double h = 0;
            for (int rI = 0; rI < n; ++rI) h += P->J0[(size_t)rI * n + a] * P->J0[(size_t)rI * n + b];
            p.HP[(size_t)a * n + b] = h;
        }
        // prior column -> tangent column (constant blocks drop out)
        if (n > 0)
            for (int b = tid; b < P->n_blocks; b += BA_THREADS) {
                const int kind = P->kind[b], ls = P->size[b] == 7 ? 6 : P->size[b];
                int col = kind == VRF_BLK_POSE ? 6 * P->index[b] : kind == VRF_BLK_SPEEDBIAS ? 66 + 9 * P->index[b]
                          : kind == VRF_BLK_EXPOSE ? BA_COL_EX : kind == VRF_BLK_TD ? BA_COL_TD : -1;
                if (col >= 0 && !col_active_dev(m, col)) col = -1;
                for (int c = 0; c < ls; ++c) p.colmap[P->idx[b] + c] = col < 0 ? -1 : col + c;
            }
    }
    __syncthreads();
    __threadfence_block();

    const long long t_kernel0 = t_kernel00;
    long long tprof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tmark = clock64();
#define TPROF(k) do { long long t_ = clock64(); tprof[k] += t_ - tmark; tmark = t_; } while (0)
    double x_cost = ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, true);
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
                double v = 0;
                for (int a = tid; a < BA_NC; a += BA_THREADS) {
                    double hv = 0;
                    for (int b = 0; b < BA_NC; ++b) hv += sh.H[pk(a, b)] * sh.tmp[b];
                    v += sh.tmp[a] * hv;
                }
                for (int l = warp; l < M; l += nwarp) {
                    if (p.lm_const[l]) continue;
                    const double *Wl = p.W + (size_t)l * BA_WS;
                    double wv = 0;
                    for (int k = lane; k < ws; k += 32) wv += Wl[k] * sh.tmp[wcol(k)];
                    wv = warp_sum_d(wv);
                    if (lane == 0) v += 2.0 * p.u_l[l] * wv + p.hll[l] * p.u_l[l] * p.u_l[l];
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
                for (int l = warp; l < M; l += nwarp) {
                    if (p.lm_const[l]) continue;
                    const double hl = p.hll[l] + mu * p.diag_l[l] * p.diag_l[l];
                    if (!(hl > 0)) { if (lane == 0) sh.flag[0] = 1; continue; }
                    const double *Wl = p.W + (size_t)l * BA_WS;
                    const double gl_h = p.gl[l] / hl;
                    for (int k = lane; k < ws; k += 32) { double w = Wl[k]; if (w != 0.0) atomicAdd(&sh.y[wcol(k)], -w * gl_h); }
                    if (lane == 0) p.hinv_l[l] = 1.0 / hl;
                }
                __syncthreads();
                // S = H + mu D^2 - W^T diag(1/h) W on the pose block (66 columns; + ex-pose and td when variable: ws = 73).
                // W is streamed through a shared-memory tile (up to 64 landmarks x ws, pre-multiplied by 1/sqrt(h));
                // every thread owns a fixed set of the ws(ws+1)/2 output entries, so the reduction is deterministic
                // and atomic-free.
                {
                    double *tile = reinterpret_cast<double *>(sh.imuJ);      // 4500 doubles available
                    const int tl = ws == 66 ? 64 : 61;                        // 64*66 = 4224, 61*73 = 4453
                    int ea[6], eb[6], ne = 0;
                    double acc[6] = {0, 0, 0, 0, 0, 0};
                    for (int e = tid; e < ws * (ws + 1) / 2 && ne < 6; e += BA_THREADS) {
                        int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                        while (a * (a + 1) / 2 > e) --a;
                        while ((a + 1) * (a + 2) / 2 <= e) ++a;
                        ea[ne] = a; eb[ne] = e - a * (a + 1) / 2; ++ne;
                    }
                    for (int t0 = 0; t0 < M; t0 += tl) {
                        const int nt = min(tl, M - t0);
                        for (int q = tid; q < nt * ws; q += BA_THREADS) {
                            int l = t0 + q / ws, k = q - (q / ws) * ws;
                            tile[q] = p.lm_const[l] ? 0.0 : p.W[(size_t)l * BA_WS + k] * sqrt(p.hinv_l[l]);
                        }
                        __syncthreads();
                        for (int q = 0; q < ne; ++q) {
                            double s_ = 0;
                            const double *ta = tile + ea[q], *tb = tile + eb[q];
                            for (int l = 0; l < nt; ++l) s_ += ta[l * ws] * tb[l * ws];
                            acc[q] += s_;
                        }
                        __syncthreads();
                    }
                    for (int q = 0; q < ne; ++q) sh.H[pk(wcol(ea[q]), wcol(eb[q]))] -= acc[q];
                }
                __syncthreads();      // the diagonal entries are touched again just below by other threads
                for (int c = tid; c < BA_NC; c += BA_THREADS) sh.H[pk(c, c)] += mu * sh.diag[c] * sh.diag[c];
                __syncthreads();
                TPROF(3);
                // in-place blocked Cholesky (lower, packed, panel width 8) of the 171x171 reduced camera system.
                // The right-hand side rides along as an extra row (index BA_NC): forward substitution for free.
                // Per panel: (1) warp 0 factors the 8x8 diagonal block, (2) one thread per row solves the
                // panel's triangular system, (3) rank-8 update of the trailing matrix  => 3 barriers / panel.
                bool bad = sh.flag[0] != 0;
                for (int j0 = 0; j0 < BA_NC && !bad; j0 += 8) {
                    const int nbp = min(8, BA_NC - j0);
                    if (warp == 0) {
                        // (1) the 8x8 diagonal block lives in registers of warp 0 (lane r = row r), factorised with shuffles
                        const int rr = lane;
                        double a8[8];
                        const int rrow = j0 + min(rr, nbp - 1);
                        const double *rp = sh.H + rrow * (rrow + 1) / 2 + j0;
#pragma unroll
                        for (int c = 0; c < 8; ++c) a8[c] = (rr < nbp && c <= rr && c < nbp) ? rp[c] : 0.0;
                        bool okp = true;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (j < nbp) {
                                const double dj = __shfl_sync(0xffffffffu, a8[j], j);
                                if (!(dj > 0.0)) okp = false;
                                // FP64 sqrt/div have ~500-cycle latencies on this part: one rsqrt per pivot instead
                                const double rinv = rsqrt(dj);
                                const double ljj = dj * rinv;
                                const double lrj = (rr == j) ? ljj : a8[j] * rinv;
                                if (rr >= j) a8[j] = lrj;
                                if (lane == 0) sh.colv[j] = rinv;            // 1 / L[j0+j, j0+j]
#pragma unroll
                                for (int c = j + 1; c < 8; ++c) {
                                    const double lcj = __shfl_sync(0xffffffffu, a8[j], c);
                                    if (c < nbp && rr >= c) a8[c] -= lrj * lcj;
                                }
                            }
                        }
                        if (rr < nbp) {
                            double *wp = sh.H + (j0 + rr) * (j0 + rr + 1) / 2 + j0;
#pragma unroll
                            for (int c = 0; c < 8; ++c) if (c <= rr && c < nbp) wp[c] = a8[c];
                        }
                        if (!okp && lane == 0) sh.flag[0] = 1;
                    }
                    __syncthreads();
                    if (sh.flag[0]) { bad = true; break; }
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
                    __syncthreads();
                    // (3) trailing update H[ii,kk] -= sum_c L[ii,j0+c] L[kk,j0+c]; 16 x 32 thread tile over (ii, kk)
                    {
                        const int tx = tid & 31, ty = tid >> 5;
                        const int t0 = j0 + nbp;
                        for (int ii = t0 + ty; ii <= BA_NC; ii += BA_THREADS / 32) {
                            double lv[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) lv[c] = sh.Lp[c][ii];
                            double *row = (ii < BA_NC) ? sh.H + ii * (ii + 1) / 2 : sh.y;
                            const int kend = (ii < BA_NC) ? ii : BA_NC - 1;
                            for (int kk = t0 + tx; kk <= kend; kk += 32) {
                                double acc = 0;
#pragma unroll
                                for (int c = 0; c < 8; ++c) acc += lv[c] * sh.Lp[c][kk];
                                row[kk] -= acc;
                            }
                        }
                    }
                    __syncthreads();
                }
                TPROF(4);
                if (!bad) {
                    // back substitution L^T x = z by one warp (warp-level sync only)
                    if (warp == 0) {
                        for (int i = BA_NC - 1; i >= 0; --i) {
                            if (lane == 0) sh.y[i] /= sh.H[pk(i, i)];
                            __syncwarp();
                            const double yi = sh.y[i];
                            const double *row = sh.H + i * (i + 1) / 2;
                            for (int k = lane; k < i; k += 32) sh.y[k] -= row[k] * yi;
                            __syncwarp();
                        }
                    }
                    __syncthreads();
                    double fin = 0;
                    for (int c = tid; c < BA_NC; c += BA_THREADS) if (!isfinite(sh.y[c])) fin += 1;
                    fin = block_sum(fin, sh.red);
                    if (fin == 0) {
                        // back-substitute the landmarks: y_l = (g_l - w_l . y_c) / h_l
                        for (int l = warp; l < M; l += nwarp) {
                            double v = 0;
                            if (!p.lm_const[l]) {
                                const double *Wl = p.W + (size_t)l * BA_WS;
                                for (int k = lane; k < ws; k += 32) v += Wl[k] * sh.y[wcol(k)];
                                v = warp_sum_d(v);
                                v = (p.gl[l] - v) * p.hinv_l[l];
                            }
                            if (lane == 0) p.y_l[l] = v;
                        }
                        solved = true;
                    }
                }
                __syncthreads();
                if (solved) break;
                // failed: the in-place factorisation destroyed H => re-linearise and retry with a larger mu
                mu *= mu_inc;
                ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, true);
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
                for (int f = tid; f < BA_NF; f += BA_THREADS) {
                    if (col_active_dev(m, 6 * f)) {
                        double dl[6];
                        for (int k = 0; k < 6; ++k) { int c = 6 * f + k; dl[k] = (c1 * sh.gd[c] + c2 * sh.gn[c]) / sh.diag[c] * sh.jscale[c]; }
                        d_pose_plus(sh.pose + 7 * f, dl, sh.cpose + 7 * f);
                    } else for (int k = 0; k < 7; ++k) sh.cpose[7 * f + k] = sh.pose[7 * f + k];
                    for (int k = 0; k < 9; ++k) {
                        int c = 66 + 9 * f + k;
                        sh.csb[9 * f + k] = sh.sb[9 * f + k] + (col_active_dev(m, c) ? (c1 * sh.gd[c] + c2 * sh.gn[c]) / sh.diag[c] * sh.jscale[c] : 0.0);
                    }
                }
                if (tid == 64) {
                    if (m.ex_active) {
                        double dl[6];
                        for (int k = 0; k < 6; ++k) { int c = BA_COL_EX + k; dl[k] = (c1 * sh.gd[c] + c2 * sh.gn[c]) / sh.diag[c] * sh.jscale[c]; }
                        d_pose_plus(sh.ex, dl, sh.cex);
                    } else for (int k = 0; k < 7; ++k) sh.cex[k] = sh.ex[k];
                    const int c = BA_COL_TD;
                    sh.tdv[1] = sh.tdv[0] + (m.td_active ? (c1 * sh.gd[c] + c2 * sh.gn[c]) / sh.diag[c] * sh.jscale[c] : 0.0);
                }
                for (int l = tid; l < M; l += BA_THREADS) {
                    double v = p.lam[l];
                    if (!p.lm_const[l]) {
                        v += (c1 * p.gd_l[l] + c2 * p.gn_l[l]) / p.diag_l[l] * p.jscale_l[l];
                        v = fmin(v, p.lm_ub[l]);           // ParameterBlock::Plus projects onto the bounds
                    }
                    p.clam[l] = v;
                }
                __syncthreads();
                TPROF(6);
                const double cand_cost = ba_evaluate(m, p, sh, sh.cpose, sh.csb, sh.cex, sh.tdv[1], p.clam, false);
                TPROF(7);
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
                    ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, true);
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
        ba_evaluate(m, p, sh, sh.pose, sh.sb, sh.ex, sh.tdv[0], p.lam, true);
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
        out.initial_cost = initial_cost; out.final_cost = x_cost;
        for (int k = 0; k < 8; ++k) out.prof[k] = tprof[k];
        out.prof2[7] = clock64() - t_kernel0;
    }
}

size_t ba_solve_smem_bytes() { return sizeof(BaShared); }

int ba_solve_launch(const BaMeta *d_meta, const BaProbDev *d_prob, BaOutDev *d_out, int n, LaunchCtx &lc)
{
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_ba_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BaShared)) != cudaSuccess) return -1;
        configured = true;
    }
    lc.begin(K_BA_SOLVE);
    k_ba_solve<<<n, BA_THREADS, sizeof(BaShared), lc.st>>>(d_meta, d_prob, d_out);
    lc.end();
    return 0;
}

}  // namespace vrf
