// ba_marg.cu -- construction of the marginalization prior on the GPU
// (reference estimator.cpp:1376-1574 + factor/marginalization_factor.cpp:
//  addResidualBlockInfo :92-112, preMarginalize :114-134, marginalize :181-315,
//  getParameterBlocks :317-338).
//
// One CTA per sequence:
//   1. canonical block order (dropped: pose0, speed-bias0, landmarks hosted at frame 0 |
//      kept: ex-pose, pose f.., speed-bias f..) -- the reference iterates an
//      unordered_map keyed by raw addresses; any order gives the same J0^T J0 / J0^T r0;
//   2. A = sum J^T J, b = sum J^T r over the prior factor, IMU factor (0,1) and every
//      projection factor hosted at frame 0 (with the Cauchy corrector), evaluated at the
//      gauge-fixed states (vector2double after double2vector);
//   3. A_mm^-1 as V diag(lambda > 1e-8 ? 1/lambda : 0) V^T from a parallel-order cyclic
//      Jacobi eigen-decomposition (Eigen::SelfAdjointEigenSolver in the reference);
//   4. Schur complement, second eigen-decomposition, J0 = sqrt(S) V^T, r0 = S^-1/2 V^T b;
//   5. kept block table with the reference's addr_shift (frame i -> i-1, or 10 -> 9).
#include "ba_math.cuh"

namespace vrf {

struct MargShared {
    int kind[64], index[64], lsize[64], gsize[64], idx[64], present[64], drop[64];
    int nb, m, n, pos, first_kept, go;
    int col_pose[BA_NF], col_sb[BA_NF], col_ex;
    double red[BA_THREADS / 32];
    double cs[2 * (BA_MAX_POS / 2 + 2)];
    int pq[2 * (BA_MAX_POS / 2 + 2)];
    double J[15 * 30], r[16];
    double dx[VRF_PRIOR_MAX_DIM], pr[VRF_PRIOR_MAX_DIM];
    double R[BA_NF * 9], ric[9];
    int flag;
};

// parallel-order cyclic Jacobi: A (n x n, row-major, global) -> eigenvalues on the diagonal, V eigenvectors (columns)
__device__ void jacobi_eig(double *A, double *V, int n, MargShared &sh)
{
    const int tid = threadIdx.x;
    for (int e = tid; e < n * n; e += BA_THREADS) V[e] = (e / n == e % n) ? 1.0 : 0.0;
    __syncthreads();
    if (n < 2) return;
    const int ne = (n + 1) & ~1;              // even number of "players"; index n (if any) is a bye
    const int npairs = ne / 2;
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0, dg = 0;
        for (int e = tid; e < n * n; e += BA_THREADS) {
            int i = e / n, j = e - i * n;
            double v = A[e];
            if (i == j) dg += v * v; else if (j > i) off += v * v;
        }
        off = block_sum(off, sh.red);
        dg = block_sum(dg, sh.red);
        if (off <= 1e-30 * dg || off == 0.0) break;
        for (int r = 0; r < ne - 1; ++r) {
            // rotation angles of this round's disjoint pairs
            for (int k = tid; k < npairs; k += BA_THREADS) {
                int p, q;
                if (k == 0) { p = ne - 1; q = r; }
                else { p = (r + k) % (ne - 1); q = (r - k + (ne - 1)) % (ne - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < n) {
                    double apq = A[p * n + q];
                    if (apq != 0.0) {
                        double app = A[p * n + p], aqq = A[q * n + q];
                        double tau = (aqq - app) / (2.0 * apq);
                        double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t); s = t * c;
                    }
                } else { q = -1; }
                sh.cs[2 * k] = c; sh.cs[2 * k + 1] = s; sh.pq[2 * k] = p; sh.pq[2 * k + 1] = q;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int e = tid; e < npairs * n; e += BA_THREADS) {
                int k = e / n, i = e - k * n;
                int p = sh.pq[2 * k], q = sh.pq[2 * k + 1];
                if (q < 0) continue;
                double c = sh.cs[2 * k], s = sh.cs[2 * k + 1];
                if (s == 0.0) continue;
                double a = A[i * n + p], b = A[i * n + q];
                A[i * n + p] = c * a - s * b; A[i * n + q] = s * a + c * b;
                a = V[i * n + p]; b = V[i * n + q];
                V[i * n + p] = c * a - s * b; V[i * n + q] = s * a + c * b;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int e = tid; e < npairs * n; e += BA_THREADS) {
                int k = e / n, j = e - k * n;
                int p = sh.pq[2 * k], q = sh.pq[2 * k + 1];
                if (q < 0) continue;
                double c = sh.cs[2 * k], s = sh.cs[2 * k + 1];
                if (s == 0.0) continue;
                double a = A[p * n + j], b = A[q * n + j];
                A[p * n + j] = c * a - s * b; A[q * n + j] = s * a + c * b;
            }
            __syncthreads();
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int find_block(const MargShared &sh, int kind, int index)
{
    for (int i = 0; i < sh.nb; ++i) if (sh.kind[i] == kind && sh.index[i] == index) return i;
    return -1;
}

#define MK_LM 100

__global__ void __launch_bounds__(BA_THREADS, 1)
k_ba_marg(const BaMeta *metas, const BaProbDev *probs, BaOutDev *outs, const BaMargDev *margs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MargShared &sh = *reinterpret_cast<MargShared *>(smem_raw);
    const BaMeta m = metas[blockIdx.x];
    const BaProbDev p = probs[blockIdx.x];
    BaOutDev &out = outs[blockIdx.x];
    const BaMargDev mg = margs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    const int M = m.M;
    const BaPriorStore *P = (p.prior && p.prior->valid) ? p.prior : nullptr;
    BaPriorStore *Q = p.prior_next;
    const double *pose = out.mpose, *sb = out.msb, *ex = out.mex;
    const double *lam = p.clam;
    const int flag = m.marg_flag;

    // ---- 1. block table (thread 0) ----
    if (tid == 0) {
        int go = (m.frame_count == VRF_WINDOW_SIZE);
        if (go && flag == VRF_MARGIN_SECOND_NEW) {
            int has = 0;
            if (P) for (int b = 0; b < P->n_blocks; ++b) if (P->kind[b] == VRF_BLK_POSE && P->index[b] == VRF_WINDOW_SIZE - 1) has = 1;
            go = has;
        }
        sh.go = go;
        sh.flag = 0;
        int nb = 0;
        auto add = [&](int kind, int index, int gs, int drop) {
            for (int i = 0; i < nb; ++i) if (sh.kind[i] == kind && sh.index[i] == index) return;
            sh.kind[nb] = kind; sh.index[nb] = index; sh.gsize[nb] = gs; sh.lsize[nb] = gs == 7 ? 6 : gs;
            sh.drop[nb] = drop; sh.present[nb] = 0; sh.idx[nb] = -1; ++nb;
        };
        if (go) {
            if (flag == VRF_MARGIN_OLD) { add(VRF_BLK_POSE, 0, 7, 1); if (m.use_imu) add(VRF_BLK_SPEEDBIAS, 0, 9, 1); }
            else add(VRF_BLK_POSE, VRF_WINDOW_SIZE - 1, 7, 1);
            sh.first_kept = nb;
            add(VRF_BLK_EXPOSE, 0, 7, 0);
            for (int f = 0; f < BA_NF; ++f) add(VRF_BLK_POSE, f, 7, 0);
            for (int f = 0; f < BA_NF; ++f) add(VRF_BLK_SPEEDBIAS, f, 9, 0);
            sh.nb = nb;
            if (P) for (int b = 0; b < P->n_blocks; ++b) { int i = find_block(sh, P->kind[b], P->index[b]); if (i >= 0) sh.present[i] = 1; }
            const int use_imu01 = (flag == VRF_MARGIN_OLD) && m.use_imu && p.imu[0].sum_dt < 10.0;
            if (use_imu01) {
                sh.present[find_block(sh, VRF_BLK_POSE, 0)] = 1; sh.present[find_block(sh, VRF_BLK_SPEEDBIAS, 0)] = 1;
                sh.present[find_block(sh, VRF_BLK_POSE, 1)] = 1; sh.present[find_block(sh, VRF_BLK_SPEEDBIAS, 1)] = 1;
            }
            int nl0 = 0;
            if (flag == VRF_MARGIN_OLD)
                for (int l = 0; l < M; ++l) {
                    const int nobs = p.obs_ptr[l + 1] - p.obs_ptr[l];
                    mg.lmcol[l] = -1;
                    if (p.start[l] != 0 || nobs < 2) continue;
                    mg.lmcol[l] = nl0++;          // provisional: ordinal among dropped landmarks
                    sh.present[find_block(sh, VRF_BLK_POSE, 0)] = 1;
                    sh.present[find_block(sh, VRF_BLK_EXPOSE, 0)] = 1;
                    for (int k = 1; k < nobs; ++k) sh.present[find_block(sh, VRF_BLK_POSE, k)] = 1;
                }
            else for (int l = 0; l < M; ++l) mg.lmcol[l] = -1;
            if (nl0 > BA_MAX_M0) { sh.go = 0; sh.flag = VRF_ERR_CAPACITY; }
            int pos = 0;
            for (int i = 0; i < sh.first_kept; ++i) if (sh.present[i]) { sh.idx[i] = pos; pos += sh.lsize[i]; }
            const int lm0 = pos;
            pos += nl0;
            sh.m = pos;
            for (int i = sh.first_kept; i < nb; ++i) if (sh.present[i] && !sh.drop[i]) { sh.idx[i] = pos; pos += sh.lsize[i]; }
            sh.pos = pos; sh.n = pos - sh.m;
            for (int l = 0; l < M; ++l) if (mg.lmcol[l] >= 0) mg.lmcol[l] += lm0;
            for (int f = 0; f < BA_NF; ++f) {
                int i = find_block(sh, VRF_BLK_POSE, f); sh.col_pose[f] = (i >= 0 && sh.present[i]) ? sh.idx[i] : -1;
                i = find_block(sh, VRF_BLK_SPEEDBIAS, f); sh.col_sb[f] = (i >= 0 && sh.present[i]) ? sh.idx[i] : -1;
            }
            { int i = find_block(sh, VRF_BLK_EXPOSE, 0); sh.col_ex = (i >= 0 && sh.present[i]) ? sh.idx[i] : -1; }
        }
    }
    __syncthreads();
    if (!sh.go) { if (tid == 0) { out.has_new_prior = 0; if (sh.flag) out.status = sh.flag; } return; }
    const int mm = sh.m, nn = sh.n, pos = sh.pos;
    double *A = mg.A, *bv = mg.b;
    for (int e = tid; e < pos * pos; e += BA_THREADS) A[e] = 0.0;
    for (int e = tid; e < pos; e += BA_THREADS) bv[e] = 0.0;
    for (int f = tid; f < BA_NF; f += BA_THREADS) d_q2R(pose + 7 * f + 3, sh.R + 9 * f);
    if (tid == 0) d_q2R(ex + 3, sh.ric);
    __syncthreads();

    // ---- 2a. prior factor ----
    if (P) {
        const int np = P->n;
        // residual at the current (gauge-fixed) states
        prior_dx(P, pose, sb, ex, sh.dx);
        __syncthreads();
        for (int rI = tid; rI < np; rI += BA_THREADS) {
            double a = P->r0[rI];
            const double *row = P->J0 + (size_t)rI * np;
            for (int k = 0; k < np; ++k) a += row[k] * sh.dx[k];
            sh.pr[rI] = a;
        }
        // prior column -> A column (reuse p.colmap scratch)
        for (int b = tid; b < P->n_blocks; b += BA_THREADS) {
            int i = find_block(sh, P->kind[b], P->index[b]);
            const int ls = P->size[b] == 7 ? 6 : P->size[b];
            for (int c = 0; c < ls; ++c) p.colmap[P->idx[b] + c] = (i >= 0) ? sh.idx[i] + c : -1;
        }
        __syncthreads();
        for (int a = tid; a < np; a += BA_THREADS) {
            double gsum = 0;
            for (int rI = 0; rI < np; ++rI) gsum += P->J0[(size_t)rI * np + a] * sh.pr[rI];
            if (p.colmap[a] >= 0) bv[p.colmap[a]] += gsum;
        }
        for (int e = tid; e < np * np; e += BA_THREADS) {
            int a = e / np, c = e - a * np;
            int ca = p.colmap[a], cc = p.colmap[c];
            if (ca < 0 || cc < 0) continue;
            double h = (c <= a) ? p.HP[(size_t)a * np + c] : p.HP[(size_t)c * np + a];     // J0^T J0 from the solve kernel
            A[(size_t)ca * pos + cc] += h;
        }
        __syncthreads();
    }
    // ---- 2b. IMU factor between frames 0 and 1 ----
    const bool use_imu01 = (flag == VRF_MARGIN_OLD) && m.use_imu && p.imu[0].sum_dt < 10.0;
    if (use_imu01 && warp == 0) {
        const VrfImuPreint *pre = p.imu;
        const double *S = p.imuS;
        double rr[15];
        ImuCtx cx;
        imu_residual_raw(pre, pose, sb, pose + 7, sb + 9, m.g_norm, rr, &cx);
        if (lane < 15) { double rw = 0; for (int k = lane; k < 15; ++k) rw += S[lane * 15 + k] * rr[k]; sh.r[lane] = rw; }
        if (lane < 30) {
            double col[15];
            imu_jac_col(pre, pose, sb, pose + 7, sb + 9, m.g_norm, &cx, lane, col);
            for (int rI = 0; rI < 15; ++rI) { double a = 0; for (int k = rI; k < 15; ++k) a += S[rI * 15 + k] * col[k]; sh.J[rI * 30 + lane] = a; }
        }
        __syncwarp();
        auto acol = [&](int c) { return c < 6 ? sh.col_pose[0] + c : c < 15 ? sh.col_sb[0] + (c - 6) : c < 21 ? sh.col_pose[1] + (c - 15) : sh.col_sb[1] + (c - 21); };
        for (int e = lane; e < 900; e += 32) {
            int a = e / 30, c = e - a * 30;
            double h = 0;
            for (int k = 0; k < 15; ++k) h += sh.J[k * 30 + a] * sh.J[k * 30 + c];
            atomicAdd(&A[(size_t)acol(a) * pos + acol(c)], h);
        }
        if (lane < 30) { double gsum = 0; for (int k = 0; k < 15; ++k) gsum += sh.J[k * 30 + lane] * sh.r[k]; atomicAdd(&bv[acol(lane)], gsum); }
    }
    __syncthreads();
    // ---- 2c. projection factors hosted at frame 0 (all four parameter blocks, Cauchy corrector) ----
    if (flag == VRF_MARGIN_OLD) {
        for (int l = warp; l < M; l += nwarp) {
            const int cl = mg.lmcol[l];
            if (cl < 0) continue;
            const int o0 = p.obs_ptr[l], nf = p.obs_ptr[l + 1] - o0 - 1;
            if (lane >= nf) continue;
            const int j = 1 + lane;
            double r[2], Ji[12], Jj[12], Jl[2], Je[12];
            proj_eval(pose, sh.R, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], p.obs[2 * o0], p.obs[2 * o0 + 1],
                      p.obs[2 * (o0 + j)], p.obs[2 * (o0 + j) + 1], true, false, r, Ji, Jj, Jl, Je);
            int cols[19]; double J0r[19], J1r[19];
            for (int c = 0; c < 6; ++c) {
                cols[c] = sh.col_pose[0] + c; J0r[c] = Ji[c]; J1r[c] = Ji[6 + c];
                cols[6 + c] = sh.col_pose[j] + c; J0r[6 + c] = Jj[c]; J1r[6 + c] = Jj[6 + c];
                cols[12 + c] = sh.col_ex + c; J0r[12 + c] = Je[c]; J1r[12 + c] = Je[6 + c];
            }
            cols[18] = cl; J0r[18] = Jl[0]; J1r[18] = Jl[1];
            for (int a = 0; a < 19; ++a) {
                atomicAdd(&bv[cols[a]], J0r[a] * r[0] + J1r[a] * r[1]);
                for (int c = 0; c < 19; ++c) atomicAdd(&A[(size_t)cols[a] * pos + cols[c]], J0r[a] * J0r[c] + J1r[a] * J1r[c]);
            }
        }
    }
    __syncthreads();
    __threadfence();
    // ---- 3. A_mm pseudo-inverse through its eigen-decomposition ----
    double *V = mg.V, *Ainv = mg.Ainv, *Amm = mg.Ainv;       // Amm lives in Ainv's buffer until the inverse is formed
    double *Tm = mg.T;
    // symmetrise into a separate m x m matrix (keep A intact for the Schur complement)
    double *Ms = mg.Ar;      // temporary? no: Ar is n x n.  Use T's buffer if large enough, else V2 -- sized on host for m*m
    (void)Ms;
    for (int e = tid; e < mm * mm; e += BA_THREADS) { int i = e / mm, j = e - i * mm; Amm[e] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]); }
    __syncthreads();
    jacobi_eig(Amm, V, mm, sh);
    // Ainv = V diag(winv) V^T ; eigenvalues are on the diagonal of Amm: stash them first
    double *wv = mg.br;      // br has room for n doubles only; use T's front (n*m >= m when n >= 1) for m eigenvalues
    wv = Tm;
    for (int k = tid; k < mm; k += BA_THREADS) { double w = Amm[(size_t)k * mm + k]; wv[k] = (w > 1e-8) ? 1.0 / w : 0.0; }
    __syncthreads();
    // scale V columns into Amm buffer is unsafe (aliasing): build Ainv row by row from V and wv into Ainv after copying wv to smem
    for (int k = tid; k < mm && k < 2 * (BA_MAX_POS / 2 + 2); k += BA_THREADS) sh.cs[k] = wv[k];
    __syncthreads();
    for (int e = tid; e < mm * mm; e += BA_THREADS) {
        int i = e / mm, j = e - i * mm;
        double a = 0;
        for (int k = 0; k < mm; ++k) a += V[(size_t)i * mm + k] * sh.cs[k] * V[(size_t)j * mm + k];
        Ainv[e] = a;
    }
    __syncthreads();
    // ---- 4. Schur complement ----
    for (int e = tid; e < nn * mm; e += BA_THREADS) {
        int i = e / mm, j = e - i * mm;
        double a = 0;
        for (int k = 0; k < mm; ++k) a += A[(size_t)(mm + i) * pos + k] * Ainv[(size_t)k * mm + j];
        Tm[e] = a;
    }
    __syncthreads();
    double *Ar = mg.Ar, *V2 = mg.V2, *br = mg.br;
    for (int e = tid; e < nn * nn; e += BA_THREADS) {
        int i = e / nn, j = e - i * nn;
        double a = A[(size_t)(mm + i) * pos + mm + j];
        for (int k = 0; k < mm; ++k) a -= Tm[(size_t)i * mm + k] * A[(size_t)k * pos + mm + j];
        Ar[e] = a;
    }
    for (int i = tid; i < nn; i += BA_THREADS) {
        double bb = bv[mm + i];
        for (int k = 0; k < mm; ++k) bb -= Tm[(size_t)i * mm + k] * bv[k];
        br[i] = bb;
    }
    __syncthreads();
    jacobi_eig(Ar, V2, nn, sh);
    // ---- 5. linearized_jacobians / residuals + kept blocks into the next prior store ----
    for (int k = tid; k < nn; k += BA_THREADS) {
        const double w = Ar[(size_t)k * nn + k];
        const double S = w > 1e-8 ? w : 0.0, Sinv = w > 1e-8 ? 1.0 / w : 0.0;
        const double ss = sqrt(S), sis = sqrt(Sinv);
        double vb = 0;
        for (int j = 0; j < nn; ++j) { Q->J0[(size_t)k * nn + j] = ss * V2[(size_t)j * nn + k]; vb += V2[(size_t)j * nn + k] * br[j]; }
        Q->r0[k] = sis * vb;
    }
    if (tid == 0) {
        int nk = 0;
        for (int i = sh.first_kept; i < sh.nb; ++i) {
            if (!sh.present[i] || sh.drop[i]) continue;
            Q->kind[nk] = sh.kind[i]; Q->size[nk] = sh.gsize[i]; Q->idx[nk] = sh.idx[i] - mm;
            const double *src = sh.kind[i] == VRF_BLK_POSE ? pose + 7 * sh.index[i] : sh.kind[i] == VRF_BLK_SPEEDBIAS ? sb + 9 * sh.index[i] : ex;
            for (int c = 0; c < sh.gsize[i]; ++c) Q->x0[9 * nk + c] = src[c];
            if (sh.kind[i] == VRF_BLK_EXPOSE) Q->index[nk] = 0;
            else if (flag == VRF_MARGIN_OLD) Q->index[nk] = sh.index[i] - 1;
            else Q->index[nk] = (sh.index[i] == VRF_WINDOW_SIZE) ? VRF_WINDOW_SIZE - 1 : sh.index[i];
            ++nk;
        }
        Q->n = nn; Q->n_blocks = nk; Q->valid = 1;
        out.has_new_prior = 1;
    }
}

size_t ba_marg_smem_bytes() { return sizeof(MargShared); }

int ba_marg_launch(const BaMeta *d_meta, const BaProbDev *d_prob, BaOutDev *d_out, BaMargDev *d_marg, int n, LaunchCtx &lc)
{
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_ba_marg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MargShared)) != cudaSuccess) return -1;
        configured = true;
    }
    lc.begin(K_BA_MARG);
    k_ba_marg<<<n, BA_THREADS, sizeof(MargShared), lc.st>>>(d_meta, d_prob, d_out, d_marg);
    lc.end();
    return 0;
}

}  // namespace vrf
