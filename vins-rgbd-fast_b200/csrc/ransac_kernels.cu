// ransac_kernels.cu -- FeatureTracker::rejectWithF (reference
// vins_estimator/src/feature_tracker/feature_tracker.cpp:441-473):
//   lift both point sets, re-project to the virtual pinhole camera (f = FOCAL_LENGTH,
//   c = (COL/2, ROW/2)) as float, cv::findFundamentalMat(FM_RANSAC, F_THRESHOLD, 0.99),
//   keep the inliers.
//
// cv::findFundamentalMat is third-party (OpenCV, not vendored in the reference).  Its
// RANSAC (modules/calib3d: ptsetreg.cpp RANSACPointSetRegistrator + fundam.cpp
// FMEstimatorCallback, OpenCV 4.13) is restated here and pinned empirically against
// cv2 4.13 (tests/test_oracle_ransac.py pins the numpy restatement, the GPU tests
// pin this kernel): identical inlier masks.
//   - cv::RNG (multiply-with-carry, seed (uint64)-1), uniform(0,count) sampling with
//     duplicate rejection, collinearity check of the last sample point,
//   - 7-point solver: null space of the 7x9 epipolar system, det(l*f1+(1-l)*f2)=0 cubic
//     (cv::solveCubic formulae), up to 3 models per sample,
//   - symmetric epipolar error max(d1^2/|l1|^2, d2^2/|l2|^2) in double on the float
//     points, rounded to float and compared with (float)(thr*thr),
//   - model accepted iff goodCount > max(maxGoodCount, 6); niters updated with
//     RANSACUpdateNumIters(0.99, outlier ratio, 7, niters) (initially 1000).
// B200 mapping: one CTA per sequence; hypotheses are generated sequentially by one
// thread (the RNG stream is sequential by definition) in batches of 32, solved one per
// lane, scored by all warps (ballot/popc reductions), and committed in iteration order.
// Only the null-space basis differs from OpenCV (pivoted Gauss-Jordan on Hartley-
// normalised points instead of LAPACK SVD): the rank-2 members of the pencil are basis
// independent, so the candidate F set is the same up to rounding (~1e-12).
//
// For 8 <= n < 15 cv::findFundamentalMat(FM_RANSAC) switches to LMedS (fundam.cpp: `npoints >= 15` selects RANSAC;
// ptsetreg.cpp LMeDSPointSetRegistrator::run): the same RNG / getSubset / 7-point machinery, a fixed
// RANSACUpdateNumIters(0.99, 0.45, 7, 1000) = 300 iterations, per model the median (nth_element at count / 2) of the float
// residuals, strict-minimum selection in iteration order, sigma = max(2.5 * 1.4826 * (1 + 5 / (n - 7)) * sqrt(median), 0.001)
// and the mask err <= (float)(sigma^2).  Restated here with the same structure.  Pinned against cv2 at n = 14 (identical
// masks, tests/test_oracle_ransac.py and the GPU tests).  For n <= 13 the median index falls inside the 7 residuals of the
// minimal sample, i.e. on rounding noise of OpenCV's LAPACK SVD (~1e-25): cv2 itself returns a different 7-survivor mask
// when one input coordinate moves by 1 ulp (shown in tests/test_oracle_ransac.py), so that regime has the reference's
// semantics (LMedS runs, typically exactly the 7 points of one minimal sample survive) but cannot be bit-pinned.
#include "common.cuh"
#include "handle.h"

namespace vrf {

#define RS_THREADS 128
#define RS_BATCH 32

struct CvRng {
    unsigned long long state;
    __device__ unsigned next()
    {
        state = (unsigned long long)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// modules/calib3d/src/precomp.hpp haveCollinearPoints: only the last point is tested.
__device__ bool rs_collinear(const float2 *m, const int *idx, int count)
{
    const int i = count - 1;
    const float2 pi = m[idx[i]];
    for (int j = 0; j < i; ++j) {
        float2 pj = m[idx[j]];
        double dx1 = (double)pj.x - (double)pi.x, dy1 = (double)pj.y - (double)pi.y;
        for (int k = 0; k < j; ++k) {
            float2 pk = m[idx[k]];
            double dx2 = (double)pk.x - (double)pi.x, dy2 = (double)pk.y - (double)pi.y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= 1.1920929e-07 * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
                return true;
        }
    }
    return false;
}

// cv::solveCubic for a0 != 0 plus the degenerate branches; returns number of roots.
__device__ int rs_solve_cubic(double a0, double a1, double a2, double a3, double *x)
{
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) return 0;
            x[0] = -a3 / a2;
            return 1;
        }
        double dd = a2 * a2 - 4 * a1 * a3;
        if (dd < 0) return 0;
        dd = sqrt(dd);
        double q1 = (-a2 + dd) * 0.5, q2 = (a2 + dd) * -0.5;
        if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
        else { x[0] = q2 / a1; x[1] = a3 / q2; }
        return dd > 0 ? 2 : 1;
    }
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    double Qc = Q * Q * Q;
    double dd = Qc - R * R;
    if (dd > 0) {
        double theta = acos(R / sqrt(Qc));
        double sq = sqrt(Q);
        double t0 = -2 * sq, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * cos(t1) - t2;
        x[1] = t0 * cos(t1 + (2. * 3.1415926535897932384626433832795 / 3)) - t2;
        x[2] = t0 * cos(t1 + (4. * 3.1415926535897932384626433832795 / 3)) - t2;
        return 3;
    }
    if (dd == 0) {
        if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
        else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
        return 2;
    }
    dd = sqrt(-dd);
    double e = pow(dd + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}

// 7-point solver (fundam.cpp run7Point).  A: caller-provided 7x9 scratch.  Writes up to
// 3 row-major 3x3 models to F and returns their count.
__device__ int rs_run7point(const float2 *m1, const float2 *m2, const int *idx, double *A, double *F)
{
    // Hartley normalisation (numerical conditioning only; the solution set is unchanged)
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
    for (int i = 0; i < 7; ++i) {
        c1x += m1[idx[i]].x; c1y += m1[idx[i]].y; c2x += m2[idx[i]].x; c2y += m2[idx[i]].y;
    }
    c1x /= 7; c1y /= 7; c2x /= 7; c2y /= 7;
    double s1 = 0, s2 = 0;
    for (int i = 0; i < 7; ++i) {
        double ax = m1[idx[i]].x - c1x, ay = m1[idx[i]].y - c1y, bx = m2[idx[i]].x - c2x, by = m2[idx[i]].y - c2y;
        s1 += sqrt(ax * ax + ay * ay); s2 += sqrt(bx * bx + by * by);
    }
    s1 /= 7; s2 /= 7;
    if (s1 < 1.1920929e-07 || s2 < 1.1920929e-07) return 0;
    s1 = 1.4142135623730951 / s1; s2 = 1.4142135623730951 / s2;
    for (int i = 0; i < 7; ++i) {
        double x0 = (m1[idx[i]].x - c1x) * s1, y0 = (m1[idx[i]].y - c1y) * s1;
        double x1 = (m2[idx[i]].x - c2x) * s2, y1 = (m2[idx[i]].y - c2y) * s2;
        double *r = A + i * 9;
        r[0] = x1 * x0; r[1] = x1 * y0; r[2] = x1; r[3] = y1 * x0; r[4] = y1 * y0; r[5] = y1;
        r[6] = x0; r[7] = y0; r[8] = 1.0;
    }
    // Gauss-Jordan with complete pivoting -> [I7 | B] in permuted column order
    int perm[9];
    for (int j = 0; j < 9; ++j) perm[j] = j;
    for (int k = 0; k < 7; ++k) {
        int pr = k, pc = k;
        double best = -1;
        for (int i = k; i < 7; ++i)
            for (int j = k; j < 9; ++j) {
                double v = fabs(A[i * 9 + j]);
                if (v > best) { best = v; pr = i; pc = j; }
            }
        if (best < 1e-13) return 0;
        if (pr != k) for (int j = 0; j < 9; ++j) { double t = A[k * 9 + j]; A[k * 9 + j] = A[pr * 9 + j]; A[pr * 9 + j] = t; }
        if (pc != k) {
            for (int i = 0; i < 7; ++i) { double t = A[i * 9 + k]; A[i * 9 + k] = A[i * 9 + pc]; A[i * 9 + pc] = t; }
            int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t;
        }
        double inv = 1.0 / A[k * 9 + k];
        for (int j = k; j < 9; ++j) A[k * 9 + j] *= inv;
        for (int i = 0; i < 7; ++i) {
            if (i == k) continue;
            double f = A[i * 9 + k];
            if (f != 0.0) for (int j = k; j < 9; ++j) A[i * 9 + j] -= f * A[k * 9 + j];
        }
    }
    double f1[9], f2[9];
    for (int j = 0; j < 9; ++j) { f1[j] = 0; f2[j] = 0; }
    f1[perm[7]] = 1.0; f2[perm[8]] = 1.0;
    for (int i = 0; i < 7; ++i) { f1[perm[i]] = -A[i * 9 + 7]; f2[perm[i]] = -A[i * 9 + 8]; }
    // normalise the basis vectors (scale only)
    double n1 = 0, n2 = 0;
    for (int j = 0; j < 9; ++j) { n1 += f1[j] * f1[j]; n2 += f2[j] * f2[j]; }
    n1 = 1.0 / sqrt(n1); n2 = 1.0 / sqrt(n2);
    for (int j = 0; j < 9; ++j) { f1[j] *= n1; f2[j] *= n2; }
    // fundam.cpp: f ~ lambda*f1 + (1-lambda)*f2, det(f) = 0
    for (int j = 0; j < 9; ++j) f1[j] -= f2[j];
    double c[4];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
           f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
           f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7]; t1 = f1[3] * f1[8] - f1[5] * f1[6]; t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
           f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
           f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    double r[3];
    int n = rs_solve_cubic(c[0], c[1], c[2], c[3], r);
    if (n < 1 || n > 3) return 0;
    for (int k = 0; k < n; ++k) {
        double lambda = r[k], mu = 1.0;
        double s = f1[8] * r[k] + f2[8];
        double Fn[9];
        if (fabs(s) > 2.220446049250313e-16) { mu = 1. / s; lambda *= mu; Fn[8] = 1.0; }
        else Fn[8] = 0.0;
        for (int i = 0; i < 8; ++i) Fn[i] = f1[i] * lambda + f2[i] * mu;
        // de-normalise: F = T2^T * Fn * T1, T = [s 0 -c.x*s; 0 s -c.y*s; 0 0 1]
        double G[9];   // Fn * T1
        for (int i = 0; i < 3; ++i) {
            G[i * 3 + 0] = Fn[i * 3 + 0] * s1;
            G[i * 3 + 1] = Fn[i * 3 + 1] * s1;
            G[i * 3 + 2] = -Fn[i * 3 + 0] * c1x * s1 - Fn[i * 3 + 1] * c1y * s1 + Fn[i * 3 + 2];
        }
        double *o = F + k * 9;
        for (int j = 0; j < 3; ++j) {
            o[0 * 3 + j] = s2 * G[0 * 3 + j];
            o[1 * 3 + j] = s2 * G[1 * 3 + j];
            o[2 * 3 + j] = -c2x * s2 * G[0 * 3 + j] - c2y * s2 * G[1 * 3 + j] + G[2 * 3 + j];
        }
    }
    return n;
}

// FMEstimatorCallback::computeError for one point; returns the float error.
__device__ __forceinline__ float rs_error(const double *F, float2 p1, float2 p2)
{
    double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    double a = F[0] * x1 + F[1] * y1 + F[2];
    double b = F[3] * x1 + F[4] * y1 + F[5];
    double c = F[6] * x1 + F[7] * y1 + F[8];
    double s2 = 1. / (a * a + b * b);
    double d2 = x2 * a + y2 * b + c;
    a = F[0] * x2 + F[3] * y2 + F[6];
    b = F[1] * x2 + F[4] * y2 + F[7];
    c = F[2] * x2 + F[5] * y2 + F[8];
    double s1 = 1. / (a * a + b * b);
    double d1 = x1 * a + y1 * b + c;
    return (float)fmax(d1 * d1 * s1, d2 * d2 * s2);
}

// RANSACUpdateNumIters (ptsetreg.cpp)
__device__ int rs_update_niters(double p, double ep, int modelPoints, int maxIters)
{
    p = fmax(p, 0.); p = fmin(p, 1.);
    ep = fmax(ep, 0.); ep = fmin(ep, 1.);
    double num = fmax(1. - p, 2.2250738585072014e-308);
    double denom = 1. - pow(1. - ep, (double)modelPoints);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= maxIters * (-denom)) ? maxIters : __double2int_rn(num / denom);
}

__global__ void __launch_bounds__(RS_THREADS)
k_ransac(FrontCfg c, const SeqCall *calls, FrontDev d)
{
    __shared__ float2 s_m1[VRF_CAP], s_m2[VRF_CAP];
    __shared__ double s_A[RS_BATCH][63];
    __shared__ double s_F[RS_BATCH * 3][9];
    __shared__ double s_best[9];
    __shared__ int s_idx[RS_BATCH][7];
    __shared__ int s_nmod[RS_BATCH];
    __shared__ int s_cnt[RS_BATCH * 3];
    __shared__ float s_med[RS_BATCH * 3];      // LMedS: median residual of every model of the batch
    __shared__ int s_ctl[4];        // 0: batch size, 1: continue flag, 2: have best, 3: first subset that failed the parallel collinearity test
    __shared__ int s_want;
    __shared__ unsigned long long s_rng[RS_BATCH];      // RNG state after the draws of each subset of the batch
    const SeqCall call = calls[blockIdx.x];
    if (!call.pub) return;
    const int seq = call.seq;
    const size_t base = (size_t)seq * VRF_CAP;
    const int n = d.t_n[seq];
    if (n < 8) return;              // rejectWithF: if (forw_pts.size() >= 8)
    const bool lmeds = n < 15;      // fundam.cpp: RANSAC needs npoints >= 15, else LMedS
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // un_cur_pts / un_forw_pts (feature_tracker.cpp:446-459)
    for (int i = tid; i < n; i += RS_THREADS) {
        float2 p = d.t_prev[base + i], q = d.t_forw[base + i];
        double mx, my;
        cam_lift(c, (double)p.x, (double)p.y, mx, my);
        s_m1[i] = make_float2((float)(c.focal * mx / 1.0 + c.cols / 2.0), (float)(c.focal * my / 1.0 + c.rows / 2.0));
        cam_lift(c, (double)q.x, (double)q.y, mx, my);
        s_m2[i] = make_float2((float)(c.focal * mx / 1.0 + c.cols / 2.0), (float)(c.focal * my / 1.0 + c.rows / 2.0));
    }
    __syncthreads();
    const float thr = (float)(c.f_thr * c.f_thr);
    // sequential control state lives in thread 0's registers
    CvRng rng; rng.state = 0xFFFFFFFFFFFFFFFFULL;
    int niters = 1000, iter = 0, maxGood = 0;
    double minMedian = 1.7976931348623157e308;
    if (lmeds) { niters = rs_update_niters(0.99, 0.45, 7, 1000); niters = max(niters, 3); }
    const int max_attempts = lmeds ? 1000 : 10000;      // getSubset(): RANSAC passes 10000, LMedS the default
    if (tid == 0) s_ctl[2] = 0;
    while (true) {
        // getSubset() for up to 32 hypotheses.  The RNG stream is sequential by definition, but the collinearity test of a
        // drawn subset (the expensive part: 2 x 15 triple products in double) is not: thread 0 draws all index sets assuming
        // every subset passes (recording the RNG state after each), 32 lanes test them in parallel, and only if a subset
        // fails -- rare -- does thread 0 redo the stream from that subset on with the reference's sequential retry rule.
        if (tid == 0) {
            const int want0 = min(RS_BATCH, niters - iter);        // niters / iter live in thread 0's registers
            s_want = want0;
            for (int nb = 0; nb < want0; ++nb) {
                int *idx = s_idx[nb];
                for (int i = 0; i < 7; ++i) {
                    int v;
                    bool dup;
                    do {
                        v = rng.uniform(0, n);
                        dup = false;
                        for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
                    } while (dup);
                    idx[i] = v;
                }
                s_rng[nb] = rng.state;
            }
        }
        __syncthreads();
        const int want = s_want;
        if (warp == 0) {
            const bool bad = lane < want && (rs_collinear(s_m1, s_idx[lane], 7) || rs_collinear(s_m2, s_idx[lane], 7));
            const unsigned badmask = __ballot_sync(0xffffffffu, bad);
            if (lane == 0) s_ctl[3] = badmask ? __ffs(badmask) - 1 : want;
        }
        __syncthreads();
        if (tid == 0) {
            int nb = s_ctl[3];
            int cont = 1;
            if (nb < want) {
                // subset nb failed its first attempt: continue the stream after its draws, sequentially (attempt 2, 3, ...)
                rng.state = s_rng[nb];
                int first_att = 1;
                for (; nb < want; ++nb) {
                    bool found = false;
                    for (int att = first_att; att < max_attempts && !found; ++att) {
                        int *idx = s_idx[nb];
                        for (int i = 0; i < 7; ++i) {
                            int v;
                            bool dup;
                            do {
                                v = rng.uniform(0, n);
                                dup = false;
                                for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
                            } while (dup);
                            idx[i] = v;
                        }
                        found = !rs_collinear(s_m1, idx, 7) && !rs_collinear(s_m2, idx, 7);
                    }
                    first_att = 0;
                    if (!found) { cont = 0; break; }
                }
            }
            s_ctl[0] = nb;
            s_ctl[1] = cont;
        }
        __syncthreads();
        const int nb = s_ctl[0];
        if (tid < nb) s_nmod[tid] = rs_run7point(s_m1, s_m2, s_idx[tid], s_A[tid], s_F[tid * 3]);
        __syncthreads();
        // score every model of the batch: warps over models, lanes over points
        for (int m = warp; m < nb * 3; m += RS_THREADS / 32) {
            const int hb = m / 3, k = m - hb * 3;
            if (k >= s_nmod[hb]) continue;
            const double *F = s_F[m];
            if (lmeds) {
                // median = element count/2 of the sorted residuals (std::nth_element on the float bit patterns, n <= 14):
                // lane i ranks its residual by counting (ties broken by index), the lane of rank n/2 holds the median
                const float e = lane < n ? rs_error(F, s_m1[lane], s_m2[lane]) : 0.f;
                int rank = 0;
                for (int j = 0; j < n; ++j) {
                    const float ej = __shfl_sync(0xffffffffu, e, j);
                    rank += (ej < e || (ej == e && j < lane)) ? 1 : 0;
                }
                if (lane < n && rank == n / 2) s_med[m] = e;
                continue;
            }
            int cnt = 0;
            for (int i = lane; i < n; i += 32) cnt += (rs_error(F, s_m1[i], s_m2[i]) <= thr) ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0) s_cnt[m] = cnt;
        }
        __syncthreads();
        if (tid == 0) {
            bool stop = false;
            for (int hb = 0; hb < nb && !stop; ++hb) {
                for (int k = 0; k < s_nmod[hb]; ++k) {
                    if (lmeds) {
                        const double median = (double)s_med[hb * 3 + k];
                        if (median < minMedian) {
                            minMedian = median;
                            for (int j = 0; j < 9; ++j) s_best[j] = s_F[hb * 3 + k][j];
                            s_ctl[2] = 1;
                        }
                        continue;
                    }
                    int good = s_cnt[hb * 3 + k];
                    if (good > max(maxGood, 6)) {
                        maxGood = good;
                        for (int j = 0; j < 9; ++j) s_best[j] = s_F[hb * 3 + k][j];
                        s_ctl[2] = 1;
                        niters = rs_update_niters(0.99, (double)(n - good) / n, 7, niters);
                    }
                }
                ++iter;
                if (iter >= niters) stop = true;
            }
            if (!s_ctl[1] || iter >= niters) s_ctl[1] = 0; else s_ctl[1] = 1;
        }
        __syncthreads();
        if (!s_ctl[1]) break;
    }
    // inlier mask of the best model (== the mask OpenCV kept); none => all rejected
    const int have = s_ctl[2];
    float thr_final = thr;
    if (lmeds) {
        // thread 0 owns minMedian: broadcast the final threshold (float)(sigma^2) through shared memory
        if (tid == 0) {
            double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * sqrt(minMedian);
            sigma = fmax(sigma, 0.001);
            s_med[0] = (float)(sigma * sigma);
        }
        __syncthreads();
        thr_final = s_med[0];
    }
    for (int i = tid; i < n; i += RS_THREADS) {
        uint8_t keep = 0;
        if (have) keep = (rs_error(s_best, s_m1[i], s_m2[i]) <= thr_final) ? 1 : 0;
        d.t_keep[base + i] = keep;
    }
}

int ransac_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, LaunchCtx &lc)
{
    lc.begin(K_RANSAC);
    k_ransac<<<ncalls, RS_THREADS, 0, lc.st>>>(c, d_calls, d);
    lc.end();
    return 0;
}

}  // namespace vrf
