// fm_kernels.cu -- the steps either side of Estimator::optimization() (C ABI in
// include/vrf_fm.h), batched over independent sequences:
//   k_fm_triangulate  FeatureManager::triangulateWithDepth   (feature_manager/feature_manager.cpp:386-543)
//   k_fm_check        Estimator::movingConsistencyCheck      (estimator/estimator.cpp:1944-2009)
//   k_imu_preint      IntegrationBase::propagate / midPointIntegration (factor/integration_base.h:13-162)
// One warp per landmark (per IMU segment); all arithmetic FP64 like the reference.  The calls are stateless: the host
// side packs the caller's arrays into one pinned staging buffer, one H2D copy, one launch, one D2H copy of the
// result section.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>

#include "ba_math.cuh"

namespace vrf {

struct FmHdr {
    double Ps[VRF_NUM_FRAMES * 3], Rs[VRF_NUM_FRAMES * 9], tic[3], ric[9];
};

struct FmState {
    unsigned char *h_buf = nullptr, *d_buf = nullptr;
    size_t cap = 0;
};

#define FM_WPB 8

// camera pose of window frame f: Rc = Rs[f] ric, tc = Ps[f] + Rs[f] tic   (12 doubles: Rc row-major | tc)
__device__ __forceinline__ void fm_cam_pose(const FmHdr &H, int f, double *o)
{
    d_mm(H.Rs + 9 * f, H.ric, o);
    double t[3];
    d_mv(H.Rs + 9 * f, H.tic, t);
    o[9] = H.Ps[3 * f] + t[0]; o[10] = H.Ps[3 * f + 1] + t[1]; o[11] = H.Ps[3 * f + 2] + t[2];
}

// right singular vector of the smallest singular value of A (rows x 4) by one-sided Jacobi (Hestenes); returns v[2] / v[3]
// (Eigen::JacobiSVD(...).matrixV().rightCols<1>() in the reference, feature_manager.cpp:497-501)
__device__ double fm_svd_depth(double *A, int rows)
{
    double V[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int r = 0; r < rows; ++r) { const double a = A[4 * r + p], b = A[4 * r + q]; al += a * a; be += b * b; ga += a * b; }
                if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < rows; ++r) {
                    const double a = A[4 * r + p], b = A[4 * r + q];
                    A[4 * r + p] = c * a - s * b; A[4 * r + q] = s * a + c * b;
                }
                for (int r = 0; r < 4; ++r) {
                    const double a = V[4 * r + p], b = V[4 * r + q];
                    V[4 * r + p] = c * a - s * b; V[4 * r + q] = s * a + c * b;
                }
            }
        if (!rotated) break;
    }
    int best = 0;
    double bn = 0;
    for (int c = 0; c < 4; ++c) {
        double nn = 0;
        for (int r = 0; r < rows; ++r) nn += A[4 * r + c] * A[4 * r + c];
        if (c == 0 || nn < bn) { bn = nn; best = c; }
    }
    return V[4 * 2 + best] / V[4 * 3 + best];
}

__global__ void __launch_bounds__(FM_WPB * 32)
k_fm_triangulate(const FmHdr *hdr, const int *lm_prob, const int *lm_start, const int *obs_beg, const int *obs_cnt,
                 const double *pts, const double *dep, double *est, int *flag, const uint8_t *dyn, int totM,
                 double depth_min, double depth_max)
{
    __shared__ double s_cam[FM_WPB][VRF_NUM_FRAMES][12];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int g = blockIdx.x * FM_WPB + wib; g < totM; g += gridDim.x * FM_WPB) {
        if (est[g] > 0 || dyn[g]) continue;                               // :390-395
        const int nk = obs_cnt[g], i0 = lm_start[g];
        if (!(nk >= 2 && i0 < VRF_WINDOW_SIZE - 2)) continue;            // :396-398
        const FmHdr &H = hdr[lm_prob[g]];
        const int ob = obs_beg[g];
        __syncwarp();
        if (lane < nk) fm_cam_pose(H, i0 + lane, s_cam[wib][lane]);
        __syncwarp();
        const double *Rr = s_cam[wib][0], *tr = s_cam[wib][0] + 9;       // host frame (:402-403)
        int nver = 0, nrough = 0;
        double sver = 0, srough = 0;
        for (int p = lane; p < nk * nk; p += 32) {
            const int k = p / nk, j = p - k * nk;
            if (k == j) continue;
            const double dk = dep[ob + k];
            if (dk == 0) continue;                                        // :415-419
            const double *R0 = s_cam[wib][k], *t0 = R0 + 9, *R1 = s_cam[wib][j], *t1 = R1 + 9;
            const double point0[3] = {pts[2 * (ob + k)] * dk, pts[2 * (ob + k) + 1] * dk, dk};
            double d10[3] = {t1[0] - t0[0], t1[1] - t0[1], t1[2] - t0[2]}, t20[3], R20[9], a[3], b[3];
            d_mtv(R0, d10, t20);                                          // R0^T (t1 - t0)
            d_mtm(R0, R1, R20);                                           // R0^T R1
            d_mtv(R20, point0, a);
            d_mtv(R20, t20, b);
            const double px = a[0] - b[0], py = a[1] - b[1], pz = a[2] - b[2];
            const double rx = pts[2 * (ob + j)] - px / pz, ry = pts[2 * (ob + j) + 1] - py / pz;
            if (sqrt(rx * rx + ry * ry) < 10.0 / 460) {                  // :444
                double d0r[3] = {t0[0] - tr[0], t0[1] - tr[1], t0[2] - tr[2]}, t2r[3], R2r[9], pr[3];
                d_mtv(Rr, d0r, t2r);
                d_mtm(Rr, R0, R2r);
                d_mv(R2r, point0, pr);
                const double z = pr[2] + t2r[2];
                if (dk > depth_max) { ++nrough; srough += z; } else { ++nver; sver += z; }
            }
        }
        const unsigned nodep = __ballot_sync(0xffffffffu, lane < nk && dep[ob + (lane < nk ? lane : 0)] == 0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nver += __shfl_xor_sync(0xffffffffu, nver, o); nrough += __shfl_xor_sync(0xffffffffu, nrough, o);
            sver += __shfl_xor_sync(0xffffffffu, sver, o); srough += __shfl_xor_sync(0xffffffffu, srough, o);
        }
        if (lane != 0) continue;
        double e;
        int fl;
        if (nver > 0) { e = sver / nver; fl = 1; }                        // :528-535
        else if (nrough > 0) { e = srough / nrough; fl = 0; }             // :519-526
        else if (__popc(nodep) == nk) {                                   // :464-513 no depth anywhere: linear triangulation
            double A[2 * VRF_NUM_FRAMES * 4];
            const double *R0 = s_cam[wib][0], *t0 = R0 + 9;
            for (int k = 0; k < nk; ++k) {
                const double *R1 = s_cam[wib][k], *t1 = R1 + 9;
                double d10[3] = {t1[0] - t0[0], t1[1] - t0[1], t1[2] - t0[2]}, t[3], R[9], Rtt[3];
                d_mtv(R0, d10, t);
                d_mtm(R0, R1, R);
                d_mtv(R, t, Rtt);
                // P = [R^T | -R^T t]
                double P[12] = {R[0], R[3], R[6], -Rtt[0], R[1], R[4], R[7], -Rtt[1], R[2], R[5], R[8], -Rtt[2]};
                const double x = pts[2 * (ob + k)], y = pts[2 * (ob + k) + 1];
                const double nrm = sqrt(x * x + y * y + 1.0);
                const double f0 = x / nrm, f1 = y / nrm, f2 = 1.0 / nrm;
                for (int c = 0; c < 4; ++c) {
                    A[4 * (2 * k) + c] = f0 * P[8 + c] - f2 * P[c];
                    A[4 * (2 * k + 1) + c] = f1 * P[8 + c] - f2 * P[4 + c];
                }
            }
            const double svd_method = fm_svd_depth(A, 2 * nk);
            e = (svd_method < depth_min) ? depth_max : svd_method;
            fl = 2;
        } else continue;                                                  // :514-517
        if (e < 0.1) { e = VRF_INIT_DEPTH; fl = 0; }                      // :537-541
        est[g] = e; flag[g] = fl;
    }
}

__global__ void __launch_bounds__(FM_WPB * 32)
k_fm_check(const FmHdr *hdr, const int *lm_prob, const int *lm_start, const int *obs_beg, const int *obs_cnt,
           const double *pts, const double *est, uint8_t *dyn, uint8_t *rem, int totM, double focal)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int g = blockIdx.x * FM_WPB + wib; g < totM; g += gridDim.x * FM_WPB) {
        const int nk = obs_cnt[g], i0 = lm_start[g];
        if (!(nk >= 2 && i0 < VRF_WINDOW_SIZE - 2)) continue;            // estimator.cpp:1970-1972
        const double depth = est[g];
        if (depth < 0) continue;                                          // :1974-1976
        const FmHdr &H = hdr[lm_prob[g]];
        const int ob = obs_beg[g];
        double err = 0, err3 = 0;
        if (lane >= 1 && lane < nk) {
            const int j = i0 + lane;
            const double uvi[3] = {pts[2 * ob] * depth, pts[2 * ob + 1] * depth, depth};
            double a[3], b[3], w[3], c[3], d[3];
            d_mv(H.ric, uvi, a);
            a[0] += H.tic[0]; a[1] += H.tic[1]; a[2] += H.tic[2];
            d_mv(H.Rs + 9 * i0, a, b);
            w[0] = b[0] + H.Ps[3 * i0] - H.Ps[3 * j]; w[1] = b[1] + H.Ps[3 * i0 + 1] - H.Ps[3 * j + 1]; w[2] = b[2] + H.Ps[3 * i0 + 2] - H.Ps[3 * j + 2];
            d_mtv(H.Rs + 9 * j, w, c);
            c[0] -= H.tic[0]; c[1] -= H.tic[1]; c[2] -= H.tic[2];
            d_mtv(H.ric, c, d);                                           // pts_cj
            const double xj = pts[2 * (ob + lane)], yj = pts[2 * (ob + lane) + 1];
            const double rx = d[0] / d[2] - xj, ry = d[1] / d[2] - yj;
            err = sqrt(rx * rx + ry * ry);                                // reprojectionError :1944-1954
            const double ex = d[0] - xj, ey = d[1] - yj, ez = d[2] - 1.0;
            err3 = sqrt(ex * ex + ey * ey + ez * ez) / depth;             // reprojectionError3D :1956-1963
        }
        err = warp_sum_d(err); err3 = warp_sum_d(err3);
        if (lane == 0) {
            const int cnt = nk - 1;
            if (focal * err / cnt > 10 || err3 / cnt > 2.0) { dyn[g] = 1; rem[g] = 1; }     // :1998-2006
            else dyn[g] = 0;
        }
    }
}

// ---------------------------------------------------------------------------
// IMU pre-integration: one warp per segment, samples in sequence.
// ---------------------------------------------------------------------------
#define PI_WPB 4
struct PreintSeg {
    double acc_0[3], gyr_0[3], ba[3], bg[3];
    int n, off;           // samples, offset into the sample arrays
};

__global__ void __launch_bounds__(PI_WPB * 32)
k_imu_preint(const PreintSeg *segs, const double *dts, const double *accs, const double *gyrs, VrfImuPreint *out, int nseg,
             double acc_n, double gyr_n, double acc_w, double gyr_w)
{
    __shared__ double s_J[PI_WPB][225], s_C[PI_WPB][225], s_F[PI_WPB][225], s_T[PI_WPB][225], s_V[PI_WPB][15 * 18];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int sgi = blockIdx.x * PI_WPB + wib;
    if (sgi >= nseg) return;
    const PreintSeg sg = segs[sgi];
    double *J = s_J[wib], *C = s_C[wib], *F = s_F[wib], *T = s_T[wib], *V = s_V[wib];
    for (int e = lane; e < 225; e += 32) { J[e] = (e % 16 == 0) ? 1.0 : 0.0; C[e] = 0.0; }
    double dp[3] = {0, 0, 0}, dq[4] = {0, 0, 0, 1}, dv[3] = {0, 0, 0}, sum_dt = 0;
    double a0[3] = {sg.acc_0[0], sg.acc_0[1], sg.acc_0[2]}, g0[3] = {sg.gyr_0[0], sg.gyr_0[1], sg.gyr_0[2]};
    const double nz[6] = {acc_n * acc_n, gyr_n * gyr_n, acc_n * acc_n, gyr_n * gyr_n, acc_w * acc_w, gyr_w * gyr_w};
    __syncwarp();
    for (int s = 0; s < sg.n; ++s) {
        const double dt = dts[sg.off + s];
        const double a1[3] = {accs[3 * (sg.off + s)], accs[3 * (sg.off + s) + 1], accs[3 * (sg.off + s) + 2]};
        const double g1[3] = {gyrs[3 * (sg.off + s)], gyrs[3 * (sg.off + s) + 1], gyrs[3 * (sg.off + s) + 2]};
        // midPointIntegration (integration_base.h:56-72), every lane redundantly
        const double a0x[3] = {a0[0] - sg.ba[0], a0[1] - sg.ba[1], a0[2] - sg.ba[2]};
        const double a1x[3] = {a1[0] - sg.ba[0], a1[1] - sg.ba[1], a1[2] - sg.ba[2]};
        const double wx[3] = {0.5 * (g0[0] + g1[0]) - sg.bg[0], 0.5 * (g0[1] + g1[1]) - sg.bg[1], 0.5 * (g0[2] + g1[2]) - sg.bg[2]};
        double un0[3], un1[3], rq[4];
        d_qrot(dq, a0x, un0);
        const double hq[4] = {wx[0] * dt / 2, wx[1] * dt / 2, wx[2] * dt / 2, 1.0};
        d_qmul(dq, hq, rq);
        d_qrot(rq, a1x, un1);
        const double un[3] = {0.5 * (un0[0] + un1[0]), 0.5 * (un0[1] + un1[1]), 0.5 * (un0[2] + un1[2])};
        double rp[3], rv[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { rp[k] = dp[k] + dv[k] * dt + 0.5 * un[k] * dt * dt; rv[k] = dv[k] + un[k] * dt; }
        // Jacobian blocks (:74-131)
        double R0[9], R1[9], Sw[9], Sa0[9], Sa1[9], IW[9], R0a0[9], R1a1[9], R1a1IW[9];
        d_q2R(dq, R0); d_q2R(rq, R1);
        d_skew(wx, Sw); d_skew(a0x, Sa0); d_skew(a1x, Sa1);
#pragma unroll
        for (int k = 0; k < 9; ++k) IW[k] = ((k % 4 == 0) ? 1.0 : 0.0) - Sw[k] * dt;
        d_mm(R0, Sa0, R0a0); d_mm(R1, Sa1, R1a1); d_mm(R1a1, IW, R1a1IW);
        for (int e = lane; e < 225; e += 32) F[e] = 0.0;
        for (int e = lane; e < 270; e += 32) V[e] = 0.0;
        __syncwarp();
        if (lane < 9) {
            const int r = lane / 3, c = lane % 3, k = lane;
            const double id = (r == c) ? 1.0 : 0.0;
            F[(0 + r) * 15 + 0 + c] = id;
            F[(0 + r) * 15 + 3 + c] = -0.25 * R0a0[k] * dt * dt + -0.25 * R1a1IW[k] * dt * dt;
            F[(0 + r) * 15 + 6 + c] = id * dt;
            F[(0 + r) * 15 + 9 + c] = -0.25 * (R0[k] + R1[k]) * dt * dt;
            F[(0 + r) * 15 + 12 + c] = -0.25 * R1a1[k] * dt * dt * -dt;
            F[(3 + r) * 15 + 3 + c] = IW[k];
            F[(3 + r) * 15 + 12 + c] = -1.0 * id * dt;
            F[(6 + r) * 15 + 3 + c] = -0.5 * R0a0[k] * dt + -0.5 * R1a1IW[k] * dt;
            F[(6 + r) * 15 + 6 + c] = id;
            F[(6 + r) * 15 + 9 + c] = -0.5 * (R0[k] + R1[k]) * dt;
            F[(6 + r) * 15 + 12 + c] = -0.5 * R1a1[k] * dt * -dt;
            F[(9 + r) * 15 + 9 + c] = id;
            F[(12 + r) * 15 + 12 + c] = id;
            const double v03 = 0.25 * -R1a1[k] * dt * dt * 0.5 * dt, v63 = 0.5 * -R1a1[k] * dt * 0.5 * dt;
            V[(0 + r) * 18 + 0 + c] = 0.25 * R0[k] * dt * dt;
            V[(0 + r) * 18 + 3 + c] = v03;
            V[(0 + r) * 18 + 6 + c] = 0.25 * R1[k] * dt * dt;
            V[(0 + r) * 18 + 9 + c] = v03;
            V[(3 + r) * 18 + 3 + c] = 0.5 * id * dt;
            V[(3 + r) * 18 + 9 + c] = 0.5 * id * dt;
            V[(6 + r) * 18 + 0 + c] = 0.5 * R0[k] * dt;
            V[(6 + r) * 18 + 3 + c] = v63;
            V[(6 + r) * 18 + 6 + c] = 0.5 * R1[k] * dt;
            V[(6 + r) * 18 + 9 + c] = v63;
            V[(9 + r) * 18 + 12 + c] = id * dt;
            V[(12 + r) * 18 + 15 + c] = id * dt;
        }
        __syncwarp();
        // jacobian = F * jacobian
        for (int e = lane; e < 225; e += 32) {
            const int r = e / 15, c = e - r * 15;
            double a = 0;
            for (int k = 0; k < 15; ++k) a += F[r * 15 + k] * J[k * 15 + c];
            T[e] = a;
        }
        __syncwarp();
        for (int e = lane; e < 225; e += 32) J[e] = T[e];
        // covariance = F * covariance * F^T + V * noise * V^T
        for (int e = lane; e < 225; e += 32) {
            const int r = e / 15, c = e - r * 15;
            double a = 0;
            for (int k = 0; k < 15; ++k) a += F[r * 15 + k] * C[k * 15 + c];
            T[e] = a;
        }
        __syncwarp();
        for (int e = lane; e < 225; e += 32) {
            const int r = e / 15, c = e - r * 15;
            double a = 0, b = 0;
            for (int k = 0; k < 15; ++k) a += T[r * 15 + k] * F[c * 15 + k];
            for (int k = 0; k < 18; ++k) b += V[r * 18 + k] * nz[k / 3] * V[c * 18 + k];
            C[e] = a + b;
        }
        __syncwarp();
        // propagate (:133-162)
        const double qn = sqrt(rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3]);
#pragma unroll
        for (int k = 0; k < 3; ++k) { dp[k] = rp[k]; dv[k] = rv[k]; a0[k] = a1[k]; g0[k] = g1[k]; }
#pragma unroll
        for (int k = 0; k < 4; ++k) dq[k] = rq[k] / qn;
        sum_dt += dt;
    }
    VrfImuPreint &o = out[sgi];
    if (lane == 0) {
        o.sum_dt = sum_dt;
        for (int k = 0; k < 3; ++k) { o.delta_p[k] = dp[k]; o.delta_v[k] = dv[k]; o.linearized_ba[k] = sg.ba[k]; o.linearized_bg[k] = sg.bg[k]; }
        for (int k = 0; k < 4; ++k) o.delta_q[k] = dq[k];
    }
    for (int e = lane; e < 225; e += 32) { o.jacobian[e] = J[e]; o.covariance[e] = C[e]; }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#define FCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { snprintf(h->errbuf, sizeof(h->errbuf), "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e__)); return VRF_ERR_CUDA; } } while (0)

static int fm_reserve(vrf_handle *h, size_t bytes)
{
    if (!h->fm) h->fm = new (std::nothrow) FmState();
    if (!h->fm) return VRF_ERR_ARG;
    FmState *f = h->fm;
    if (bytes <= f->cap) return VRF_OK;
    if (f->h_buf) cudaFreeHost(f->h_buf);
    if (f->d_buf) cudaFree(f->d_buf);
    f->h_buf = nullptr; f->d_buf = nullptr; f->cap = 0;
    const size_t cap = bytes + bytes / 2 + 4096;
    FCK(cudaMallocHost((void **)&f->h_buf, cap));
    FCK(cudaMalloc((void **)&f->d_buf, cap));
    f->cap = cap;
    return VRF_OK;
}

void fm_destroy(vrf_handle *h)
{
    if (!h->fm) return;
    if (h->fm->h_buf) cudaFreeHost(h->fm->h_buf);
    if (h->fm->d_buf) cudaFree(h->fm->d_buf);
    delete h->fm;
    h->fm = nullptr;
}

static size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

// layout of one packed call; all offsets are 16-byte aligned
struct FmLayout {
    size_t hdr, prob, start, obeg, ocnt, pts, dep, est, flag, dyn, rem, total;
    size_t out_begin, out_end;
};
static FmLayout fm_layout(int n, int totM, int totO)
{
    FmLayout L;
    size_t o = 0;
    L.hdr = o; o = al16(o + sizeof(FmHdr) * n);
    L.prob = o; o = al16(o + sizeof(int) * totM);
    L.start = o; o = al16(o + sizeof(int) * totM);
    L.obeg = o; o = al16(o + sizeof(int) * totM);
    L.ocnt = o; o = al16(o + sizeof(int) * totM);
    L.pts = o; o = al16(o + sizeof(double) * 2 * totO);
    L.dep = o; o = al16(o + sizeof(double) * totO);
    L.out_begin = o;
    L.est = o; o = al16(o + sizeof(double) * totM);
    L.flag = o; o = al16(o + sizeof(int) * totM);
    L.dyn = o; o = al16(o + totM);
    L.rem = o; o = al16(o + totM);
    L.out_end = o;
    L.total = o;
    return L;
}

static int fm_run(vrf_handle *h, int n, VrfFmProblem *probs, bool triangulate)
{
    if (!h || n < 0 || (n > 0 && !probs)) return VRF_ERR_ARG;
    if (n == 0) return VRF_OK;
    long long totM = 0, totO = 0;
    for (int i = 0; i < n; ++i) {
        const VrfFmProblem &p = probs[i];
        if (p.n_landmarks < 0 || p.n_obs < 0) return VRF_ERR_ARG;
        if (p.n_landmarks > 0 && (!p.lm_start_frame || !p.lm_obs_ptr || !p.obs_pts || !p.estimated_depth || !p.is_dynamic)) return VRF_ERR_ARG;
        if (triangulate && p.n_landmarks > 0 && (!p.obs_depth || !p.estimate_flag)) return VRF_ERR_ARG;
        for (int l = 0; l < p.n_landmarks; ++l) {
            const int cnt = p.lm_obs_ptr[l + 1] - p.lm_obs_ptr[l];
            if (cnt < 0 || p.lm_start_frame[l] < 0 || p.lm_start_frame[l] + cnt > VRF_NUM_FRAMES || p.lm_obs_ptr[l + 1] > p.n_obs) return VRF_ERR_ARG;
        }
        totM += p.n_landmarks; totO += p.n_obs;
    }
    if (totM == 0) return VRF_OK;
    if (totM > (1LL << 30) || totO > (1LL << 30)) return VRF_ERR_CAPACITY;
    const FmLayout L = fm_layout(n, (int)totM, (int)totO);
    int rc = fm_reserve(h, L.total);
    if (rc != VRF_OK) return rc;
    FCK(cudaSetDevice(h->device));
    unsigned char *hb = h->fm->h_buf, *db = h->fm->d_buf;
    FmHdr *hh = reinterpret_cast<FmHdr *>(hb + L.hdr);
    int *hprob = reinterpret_cast<int *>(hb + L.prob), *hstart = reinterpret_cast<int *>(hb + L.start);
    int *hobeg = reinterpret_cast<int *>(hb + L.obeg), *hocnt = reinterpret_cast<int *>(hb + L.ocnt);
    double *hpts = reinterpret_cast<double *>(hb + L.pts), *hdep = reinterpret_cast<double *>(hb + L.dep);
    double *hest = reinterpret_cast<double *>(hb + L.est);
    int *hflag = reinterpret_cast<int *>(hb + L.flag);
    uint8_t *hdyn = hb + L.dyn, *hrem = hb + L.rem;
    int g = 0, ob = 0;
    for (int i = 0; i < n; ++i) {
        const VrfFmProblem &p = probs[i];
        memcpy(hh[i].Ps, p.Ps, sizeof(p.Ps)); memcpy(hh[i].Rs, p.Rs, sizeof(p.Rs));
        memcpy(hh[i].tic, p.tic, sizeof(p.tic)); memcpy(hh[i].ric, p.ric, sizeof(p.ric));
        for (int l = 0; l < p.n_landmarks; ++l, ++g) {
            hprob[g] = i; hstart[g] = p.lm_start_frame[l];
            hobeg[g] = ob + p.lm_obs_ptr[l]; hocnt[g] = p.lm_obs_ptr[l + 1] - p.lm_obs_ptr[l];
            hest[g] = p.estimated_depth[l]; hflag[g] = p.estimate_flag ? p.estimate_flag[l] : 0;
            hdyn[g] = p.is_dynamic[l]; hrem[g] = 0;
        }
        if (p.n_obs > 0) {
            memcpy(hpts + 2 * (size_t)ob, p.obs_pts, sizeof(double) * 2 * p.n_obs);
            if (triangulate) memcpy(hdep + ob, p.obs_depth, sizeof(double) * p.n_obs);
        }
        ob += p.n_obs;
    }
    FCK(cudaMemcpyAsync(db, hb, L.total, cudaMemcpyHostToDevice, h->stream));
    LaunchCtx lc{h->stream, &h->launches, &h->prof};
    const int grid = (int)std::min<long long>((totM + FM_WPB - 1) / FM_WPB, (long long)h->sm_count * 16);
    if (triangulate) {
        lc.begin(K_FM_TRI);
        k_fm_triangulate<<<grid, FM_WPB * 32, 0, h->stream>>>(
            reinterpret_cast<const FmHdr *>(db + L.hdr), reinterpret_cast<const int *>(db + L.prob), reinterpret_cast<const int *>(db + L.start),
            reinterpret_cast<const int *>(db + L.obeg), reinterpret_cast<const int *>(db + L.ocnt), reinterpret_cast<const double *>(db + L.pts),
            reinterpret_cast<const double *>(db + L.dep), reinterpret_cast<double *>(db + L.est), reinterpret_cast<int *>(db + L.flag),
            db + L.dyn, (int)totM, h->cfg.depth_min_dist, h->cfg.depth_max_dist);
        lc.end();
    } else {
        lc.begin(K_FM_CHECK);
        k_fm_check<<<grid, FM_WPB * 32, 0, h->stream>>>(
            reinterpret_cast<const FmHdr *>(db + L.hdr), reinterpret_cast<const int *>(db + L.prob), reinterpret_cast<const int *>(db + L.start),
            reinterpret_cast<const int *>(db + L.obeg), reinterpret_cast<const int *>(db + L.ocnt), reinterpret_cast<const double *>(db + L.pts),
            reinterpret_cast<const double *>(db + L.est), db + L.dyn, db + L.rem, (int)totM, h->cfg.focal_length);
        lc.end();
    }
    FCK(cudaGetLastError());
    FCK(cudaMemcpyAsync(hb + L.out_begin, db + L.out_begin, L.out_end - L.out_begin, cudaMemcpyDeviceToHost, h->stream));
    FCK(cudaStreamSynchronize(h->stream));
    g = 0;
    for (int i = 0; i < n; ++i) {
        VrfFmProblem &p = probs[i];
        for (int l = 0; l < p.n_landmarks; ++l, ++g) {
            if (triangulate) { p.estimated_depth[l] = hest[g]; p.estimate_flag[l] = hflag[g]; }
            else { p.is_dynamic[l] = hdyn[g]; if (p.remove) p.remove[l] = hrem[g]; }
        }
    }
    return VRF_OK;
}

}  // namespace vrf

using namespace vrf;

extern "C" int vrf_fm_triangulate_with_depth_batch(vrf_handle *h, int n, VrfFmProblem *probs) { return fm_run(h, n, probs, true); }
extern "C" int vrf_fm_moving_consistency_check_batch(vrf_handle *h, int n, VrfFmProblem *probs) { return fm_run(h, n, probs, false); }

extern "C" int vrf_imu_preintegrate_batch(vrf_handle *h, int n, const VrfImuSegment *segs, VrfImuPreint *out)
{
    if (!h || n < 0 || (n > 0 && (!segs || !out))) return VRF_ERR_ARG;
    if (n == 0) return VRF_OK;
    long long tot = 0;
    for (int i = 0; i < n; ++i) {
        if (segs[i].n_samples < 0 || (segs[i].n_samples > 0 && (!segs[i].dt || !segs[i].acc || !segs[i].gyr))) return VRF_ERR_ARG;
        tot += segs[i].n_samples;
    }
    if (tot > (1LL << 28)) return VRF_ERR_CAPACITY;
    size_t o = 0;
    const size_t o_seg = o; o = al16(o + sizeof(PreintSeg) * n);
    const size_t o_dt = o; o = al16(o + sizeof(double) * tot);
    const size_t o_acc = o; o = al16(o + sizeof(double) * 3 * tot);
    const size_t o_gyr = o; o = al16(o + sizeof(double) * 3 * tot);
    const size_t o_in_end = o;
    const size_t o_out = o; o = al16(o + sizeof(VrfImuPreint) * n);
    int rc = fm_reserve(h, o);
    if (rc != VRF_OK) return rc;
    FCK(cudaSetDevice(h->device));
    unsigned char *hb = h->fm->h_buf, *db = h->fm->d_buf;
    PreintSeg *hs = reinterpret_cast<PreintSeg *>(hb + o_seg);
    double *hdt = reinterpret_cast<double *>(hb + o_dt), *hacc = reinterpret_cast<double *>(hb + o_acc), *hgyr = reinterpret_cast<double *>(hb + o_gyr);
    int off = 0;
    for (int i = 0; i < n; ++i) {
        const VrfImuSegment &s = segs[i];
        memcpy(hs[i].acc_0, s.acc_0, 24); memcpy(hs[i].gyr_0, s.gyr_0, 24);
        memcpy(hs[i].ba, s.linearized_ba, 24); memcpy(hs[i].bg, s.linearized_bg, 24);
        hs[i].n = s.n_samples; hs[i].off = off;
        if (s.n_samples > 0) {
            memcpy(hdt + off, s.dt, sizeof(double) * s.n_samples);
            memcpy(hacc + 3 * (size_t)off, s.acc, sizeof(double) * 3 * s.n_samples);
            memcpy(hgyr + 3 * (size_t)off, s.gyr, sizeof(double) * 3 * s.n_samples);
        }
        off += s.n_samples;
    }
    FCK(cudaMemcpyAsync(db, hb, o_in_end, cudaMemcpyHostToDevice, h->stream));
    LaunchCtx lc{h->stream, &h->launches, &h->prof};
    lc.begin(K_IMU_PREINT);
    k_imu_preint<<<(n + PI_WPB - 1) / PI_WPB, PI_WPB * 32, 0, h->stream>>>(
        reinterpret_cast<const PreintSeg *>(db + o_seg), reinterpret_cast<const double *>(db + o_dt), reinterpret_cast<const double *>(db + o_acc),
        reinterpret_cast<const double *>(db + o_gyr), reinterpret_cast<VrfImuPreint *>(db + o_out), n, h->cfg.acc_n, h->cfg.gyr_n, h->cfg.acc_w, h->cfg.gyr_w);
    lc.end();
    FCK(cudaGetLastError());
    FCK(cudaMemcpyAsync(hb + o_out, db + o_out, sizeof(VrfImuPreint) * n, cudaMemcpyDeviceToHost, h->stream));
    FCK(cudaStreamSynchronize(h->stream));
    memcpy(out, hb + o_out, sizeof(VrfImuPreint) * n);
    return VRF_OK;
}
