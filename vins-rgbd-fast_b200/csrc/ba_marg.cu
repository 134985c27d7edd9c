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
    BaMeta meta;                // this CTA's problem descriptors
    BaProbDev prob;
    BaMargDev mdev;
    int kind[64], index[64], lsize[64], gsize[64], idx[64], present[64], drop[64];
    int nb, m, n, pos, first_kept, go;
    int chunk[BA_MAX_LM / 32], maxobs, lm0, jdbg;
    int col_pose[BA_NF], col_sb[BA_NF], col_ex, col_td;
    double red[BA_THREADS / 32];
    __align__(16) double cs[2 * (BA_MAX_POS / 2 + 2)];
    __align__(16) double cs_b[2 * (BA_MAX_POS / 2 + 2)];
    __align__(8) int pq[2 * (BA_MAX_POS / 2 + 2)];
    double J[15 * 30], r[16];
    double dx[VRF_PRIOR_MAX_DIM], pr[VRF_PRIOR_MAX_DIM];
    double R[BA_NF * 9], ric[9];
    int flag;
    int pair_off[BA_NF], pair_cnt[BA_NF];       // factor lists of the pairs (host 0, observer j), phase 2c
};

// parallel-order cyclic Jacobi: A (n x n, row-major, global) -> eigenvalues on the diagonal, V eigenvectors (columns)
__device__ __noinline__ void jacobi_eig(double *A, double *V, int n, MargShared &sh)
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

// Same algorithm on matrices held in shared memory (n <= MARG_SMEM_N).  The round loop is bound by
// shared-memory bandwidth (every round reads and writes all of A and V), so the layout minimises bytes:
//  * matrices are padded to an even dimension ne with a zero row/column (its rotations are identities,
//    so no "bye" branches);
//  * A is kept as its upper triangle only (entry (i,j) lives at [min][max], odd leading dimension).  Row and
//    column rotations of a round are fused: the 2x2 block at rows {p1,q1} x cols {p2,q2} of J^T A J only
//    depends on the same 4 entries, and by symmetry only the blocks of the upper (pair x pair) triangle
//    exist: 4 loads + 4 stores per block, half the FP64 work of the two-sided update;
//  * V is stored transposed (VT[col][row]) so that a rotation of columns (p,q) streams two contiguous
//    rows of VT with 16-byte accesses, lanes along the row;
//  * the rotation needs no division: with d = aqq - app, e = 2 apq, h = hypot(d, e), g = |d| + h:
//    c = g / hypot(g, e), s = sign(d e) |e| / hypot(g, e)  (two rsqrt; FP64 div is ~10x an FMA here);
//  * every thread keeps its work items for the whole call in registers.
#define MARG_SMEM_N 92
#define MARG_NE(n) (((n) + 1) & ~1)
#define MARG_LDA(n) (MARG_NE(n) | 1)
#define MARG_A_ELEMS(n) (MARG_NE(n) * MARG_LDA(n))              /* even: VT stays 16-byte aligned */
#define MARG_V_ELEMS(n) (MARG_NE(n) * MARG_NE(n))
#define JAC_V_ITEMS 8      /* >= ceil(46 * 46 / 448), processed in groups of 4 */
// round-robin tournament: pair k of round r (ne players, player ne-1 fixed), returned with p < q
__device__ __forceinline__ int2 jac_pair(int r, int k, int ne)
{
    int p, q;
    if (k == 0) { p = ne - 1; q = r; }
    else {
        p = r + k; if (p >= ne - 1) p -= ne - 1;
        q = r - k; if (q < 0) q += ne - 1;
    }
    return p < q ? make_int2(p, q) : make_int2(q, p);
}
// rotation of pair k in round r from the current A: (c, s)
__device__ __forceinline__ double2 jac_angle(const double *A, int ld, int2 pq)
{
    double c = 1.0, s_ = 0.0;
    const double apq = A[pq.x * ld + pq.y];
    if (apq != 0.0) {
        const double d = A[pq.y * ld + pq.y] - A[pq.x * ld + pq.x], e2 = 2.0 * apq;
        const double w = d * d + e2 * e2;
        const double h = w * rsqrt(w);
        const double g = fabs(d) + h;
        const double rr = rsqrt(g * g + e2 * e2);
        const bool neg = (d != 0.0) && ((d < 0.0) != (e2 < 0.0));
        c = g * rr;
        s_ = (neg ? -fabs(e2) : fabs(e2)) * rr;
    }
    return make_double2(c, s_);
}

#define JAC_ANGLE_THREADS 64       /* warps 0-1 compute the next round's angles while the others rotate V */
// (Also measured and rejected: 256 threads per window with a < 113 KB frame, i.e. two windows or a window plus a
// front-end CTA per SM.  The Jacobi phase took the same 2.5 M cycles with half the threads -- it is bound by dependent
// shared-memory latency -- but the whole step got 7 % slower: SMs shared with front-end CTAs never drain, which starves
// the solve kernel that needs a whole SM.)
// (A table of the round-robin pairs in shared memory was tried and measured 7 % SLOWER: a round is bound by the
// latency of its dependent shared-memory accesses, not by instruction issue, and the lookup lengthens that chain.)
__device__ __noinline__ void jacobi_eig_smem(double *A, double *VT, int n, MargShared &sh)
{
    const int tid = threadIdx.x;
    const int ne = MARG_NE(n), ld = MARG_LDA(n), ldv = ne;
    const int npairs = ne / 2;
    for (int e = tid; e < ne * ldv; e += BA_THREADS) VT[e] = (e / ldv == e % ldv) ? 1.0 : 0.0;
    __syncthreads();
    if (n < 2) return;
    auto pair_at = [&](int r, int k) { return jac_pair(r, k, ne); };
    // A-block items: upper triangle of the (pair x pair) grid, dealt from the top thread down
    const int ntri = npairs * (npairs + 1) / 2;
    int bk1[4], bk2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int e = (BA_THREADS - 1 - tid) + u * BA_THREADS;
        bk1[u] = -1; bk2[u] = 0;
        if (e < ntri) {
            int k1 = 0, rowlen = npairs;
            while (e >= rowlen) { e -= rowlen; --rowlen; ++k1; }
            bk1[u] = k1; bk2[u] = k1 + e;
        }
    }
    // V items: (pair, row pair) on threads >= JAC_ANGLE_THREADS; consecutive lanes take consecutive row pairs
    const int nvit = npairs * npairs;             // npairs rotations x (ne / 2) row pairs
    const int nvthr = BA_THREADS - JAC_ANGLE_THREADS;
    short vk[JAC_V_ITEMS], vi[JAC_V_ITEMS];
#pragma unroll
    for (int u = 0; u < JAC_V_ITEMS; ++u) {
        const int e = (tid - JAC_ANGLE_THREADS) + u * nvthr;
        const bool ok = tid >= JAC_ANGLE_THREADS && e < nvit;
        vk[u] = (short)(ok ? e / npairs : -1);
        vi[u] = (short)(ok ? e % npairs : 0);
    }
    double2 *csb[2] = {reinterpret_cast<double2 *>(sh.cs), reinterpret_cast<double2 *>(sh.cs_b)};
    long long jt[4] = {0, 0, 0, 0}, jc = 0;
    int nsweep = 0;
#define JPROF(k) do { if (sh.jdbg) { long long t_ = clock64(); jt[k] += t_ - jc; jc = t_; } } while (0)
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0, dg = 0;
        for (int i = tid; i < n; i += BA_THREADS) {
            const double *row = A + i * ld;
            for (int j = i + 1; j < n; ++j) { const double v = row[j]; off += v * v; }
            dg += row[i] * row[i];
        }
        off = block_sum(off, sh.red);
        dg = block_sum(dg, sh.red);
        if (off <= 1e-26 * dg || off == 0.0) break;     // relative off-diagonal norm 1e-13
        ++nsweep;
        if (tid < npairs) csb[0][tid] = jac_angle(A, ld, pair_at(0, tid));
        __syncthreads();
        if (sh.jdbg) jc = clock64();
        for (int r = 0; r < ne - 1; ++r) {
            const double2 *cs2 = csb[r & 1];
            // ---- A <- J^T A J on the upper (pair x pair) triangle; loads of both blocks first ----
#pragma unroll
            for (int u0 = 0; u0 < 4; u0 += 2) {
                if (bk1[u0] < 0) break;
                double a11[2], a12[2], a21[2], a22[2];
                double2 r1[2], r2[2];
                int e11[2], e12[2], e21[2], e22[2];
                bool live[2];
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    const int k1 = bk1[u0 + uu], k2 = bk2[u0 + uu];
                    live[uu] = k1 >= 0;
                    const int kk1 = live[uu] ? k1 : 0, kk2 = live[uu] ? k2 : 0;
                    r1[uu] = cs2[kk1]; r2[uu] = cs2[kk2];
                    const int2 i1 = pair_at(r, kk1), i2 = pair_at(r, kk2);
                    // diagonal block (k1 == k2): entries (p,p) (p,q) (p,q) (q,q)
                    e11[uu] = min(i1.x, i2.x) * ld + max(i1.x, i2.x); e12[uu] = min(i1.x, i2.y) * ld + max(i1.x, i2.y);
                    e21[uu] = min(i1.y, i2.x) * ld + max(i1.y, i2.x); e22[uu] = min(i1.y, i2.y) * ld + max(i1.y, i2.y);
                    // (idle slots must not touch A: their stand-in block (0, 0) belongs to another thread, which rewrites it in this phase)
                    if (live[uu]) { a11[uu] = A[e11[uu]]; a12[uu] = A[e12[uu]]; a21[uu] = A[e21[uu]]; a22[uu] = A[e22[uu]]; }
                    else { a11[uu] = 0.0; a12[uu] = 0.0; a21[uu] = 0.0; a22[uu] = 0.0; }
                }
#pragma unroll
                for (int uu = 0; uu < 2; ++uu) {
                    const double c1 = r1[uu].x, s1 = r1[uu].y, c2 = r2[uu].x, s2 = r2[uu].y;
                    if (!live[uu] || (s1 == 0.0 && s2 == 0.0)) continue;
                    const double t11 = c2 * a11[uu] - s2 * a12[uu], t12 = s2 * a11[uu] + c2 * a12[uu];
                    const double t21 = c2 * a21[uu] - s2 * a22[uu], t22 = s2 * a21[uu] + c2 * a22[uu];
                    const double o11 = c1 * t11 - s1 * t21, o22 = s1 * t12 + c1 * t22;
                    if (bk1[u0 + uu] == bk2[u0 + uu]) {
                        // the rotation annihilates a_pq (e12 == e21 here)
                        A[e11[uu]] = o11; A[e22[uu]] = o22; A[e12[uu]] = 0.0;
                    } else {
                        A[e11[uu]] = o11; A[e12[uu]] = c1 * t12 - s1 * t22;
                        A[e21[uu]] = s1 * t11 + c1 * t21; A[e22[uu]] = o22;
                    }
                }
            }
            JPROF(0);
            __syncthreads();
            JPROF(1);
            if (tid < JAC_ANGLE_THREADS) {
                // next round's rotation angles (double-buffered) ...
                if (tid < npairs && r + 1 < ne - 1) csb[(r + 1) & 1][tid] = jac_angle(A, ld, pair_at(r + 1, tid));
            } else {
                // ... while the other warps apply this round's rotations to the eigenvectors
#pragma unroll
                for (int u0 = 0; u0 < JAC_V_ITEMS; u0 += 4) {
                    if (vk[u0] < 0) break;
                    double2 va[4], vb[4], rc[4];
                    double2 *colp[4], *colq[4];
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        const int k = vk[u0 + uu] < 0 ? 0 : vk[u0 + uu];
                        rc[uu] = cs2[k];
                        const int2 ip = pair_at(r, k);
                        colp[uu] = reinterpret_cast<double2 *>(VT + ip.x * ldv) + vi[u0 + uu];
                        colq[uu] = reinterpret_cast<double2 *>(VT + ip.y * ldv) + vi[u0 + uu];
                        if (vk[u0 + uu] >= 0) { va[uu] = *colp[uu]; vb[uu] = *colq[uu]; }
                        else { va[uu] = make_double2(0.0, 0.0); vb[uu] = va[uu]; }
                    }
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) {
                        if (vk[u0 + uu] < 0 || rc[uu].y == 0.0) continue;
                        const double c = rc[uu].x, s_ = rc[uu].y;
                        *colp[uu] = make_double2(c * va[uu].x - s_ * vb[uu].x, c * va[uu].y - s_ * vb[uu].y);
                        *colq[uu] = make_double2(s_ * va[uu].x + c * vb[uu].x, s_ * va[uu].y + c * vb[uu].y);
                    }
                }
            }
            JPROF(2);
            __syncthreads();
            JPROF(3);
        }
    }
    __syncthreads();
    if (sh.jdbg && blockIdx.x == 0 && (tid == 0 || tid == 300 || tid == 511))
        printf("jacobi n=%d sweeps=%d tid=%d ablock=%lld wait1=%lld angle|V=%lld wait2=%lld\n", n, nsweep, tid, jt[0], jt[1], jt[2], jt[3]);
#undef JPROF
}

__device__ __forceinline__ int find_block(const MargShared &sh, int kind, int index)
{
    for (int i = 0; i < sh.nb; ++i) if (sh.kind[i] == kind && sh.index[i] == index) return i;
    return -1;
}

#define MK_LM 100
#define MG_L 73                         // local columns of the on-chip projection system: pose f -> 6f, ex-pose -> 66, td -> 72
#define MG_LP (MG_L * (MG_L + 1) / 2)

__global__ void __launch_bounds__(BA_THREADS, 1)
k_ba_marg(const BaMeta *metas, const BaProbDev *probs, BaOutDev *outs, const BaMargDev *margs)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MargShared &sh = *reinterpret_cast<MargShared *>(smem_raw);
    // problem descriptors in shared memory, one copy per CTA (per-thread copies end up in local memory: see k_ba_solve)
    if (threadIdx.x < (int)(sizeof(BaMeta) / sizeof(int))) reinterpret_cast<int *>(&sh.meta)[threadIdx.x] = reinterpret_cast<const int *>(metas + blockIdx.x)[threadIdx.x];
    if (threadIdx.x < (int)(sizeof(BaProbDev) / sizeof(int))) reinterpret_cast<int *>(&sh.prob)[threadIdx.x] = reinterpret_cast<const int *>(probs + blockIdx.x)[threadIdx.x];
    if (threadIdx.x < (int)(sizeof(BaMargDev) / sizeof(int))) reinterpret_cast<int *>(&sh.mdev)[threadIdx.x] = reinterpret_cast<const int *>(margs + blockIdx.x)[threadIdx.x];
    __syncthreads();
    const BaMeta &m = sh.meta;
    const BaProbDev &p = sh.prob;
    BaOutDev &out = outs[blockIdx.x];
    const BaMargDev &mg = sh.mdev;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = BA_THREADS / 32;
    const int M = m.M;
    const BaPriorStore *P = (p.prior && p.prior->valid) ? p.prior : nullptr;
    BaPriorStore *Q = p.prior_next;
    const double *pose = out.mpose, *sb = out.msb, *ex = out.mex;
    const double *lam = p.clam;
    const int flag = m.marg_flag;

    long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tmark = clock64();
#define MPROF(k) do { long long t_ = clock64(); tp[k] += t_ - tmark; tmark = t_; } while (0)
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
        sh.jdbg = (m.debug & 8) != 0;
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
            add(VRF_BLK_TD, 0, 1, 0);
            for (int f = 0; f < BA_NF; ++f) add(VRF_BLK_POSE, f, 7, 0);
            for (int f = 0; f < BA_NF; ++f) add(VRF_BLK_SPEEDBIAS, f, 9, 0);
            sh.nb = nb;
            if (P) for (int b = 0; b < P->n_blocks; ++b) { int i = find_block(sh, P->kind[b], P->index[b]); if (i >= 0) sh.present[i] = 1; }
            const int use_imu01 = (flag == VRF_MARGIN_OLD) && m.use_imu && p.imu[0].sum_dt < 10.0;
            if (use_imu01) {
                sh.present[find_block(sh, VRF_BLK_POSE, 0)] = 1; sh.present[find_block(sh, VRF_BLK_SPEEDBIAS, 0)] = 1;
                sh.present[find_block(sh, VRF_BLK_POSE, 1)] = 1; sh.present[find_block(sh, VRF_BLK_SPEEDBIAS, 1)] = 1;
            }
        }
        sh.maxobs = 0;
    }
    __syncthreads();
    // landmarks hosted at frame 0 with >= 2 observations are dropped (estimator.cpp:1419-1460): flag them in
    // parallel; their column ordinal is a ballot prefix (landmark order = getDepthVector order)
    bool ldrop[BA_MAX_LM / BA_THREADS];
    int lrank[BA_MAX_LM / BA_THREADS];
#pragma unroll
    for (int u = 0; u < BA_MAX_LM / BA_THREADS; ++u) {
        const int l = u * BA_THREADS + tid;
        int nobs = 0;
        bool dflag = false;
        if (sh.go && flag == VRF_MARGIN_OLD && l < M) {
            nobs = p.obs_ptr[l + 1] - p.obs_ptr[l];
            dflag = (p.start[l] == 0 && nobs >= 2);
        }
        const unsigned ball = __ballot_sync(0xffffffffu, dflag);
        if (lane == 0) sh.chunk[u * (BA_THREADS / 32) + warp] = __popc(ball);
        if (dflag) atomicMax(&sh.maxobs, nobs);
        ldrop[u] = dflag;
        lrank[u] = __popc(ball & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (tid == 0 && sh.go) {
        {
            const int nb = sh.nb;
            int nl0 = 0;
            for (int c = 0; c < BA_MAX_LM / 32; ++c) nl0 += sh.chunk[c];
            if (nl0 > 0) {
                sh.present[find_block(sh, VRF_BLK_POSE, 0)] = 1;
                sh.present[find_block(sh, VRF_BLK_EXPOSE, 0)] = 1;
                if (m.td_factor) sh.present[find_block(sh, VRF_BLK_TD, 0)] = 1;      // ProjectionTdFactor keeps para_Td (estimator.cpp:1445-1458)
                for (int k = 1; k < sh.maxobs; ++k) sh.present[find_block(sh, VRF_BLK_POSE, k)] = 1;
            }
            if (nl0 > BA_MAX_M0) { sh.go = 0; sh.flag = VRF_ERR_CAPACITY; }
            int pos = 0;
            for (int i = 0; i < sh.first_kept; ++i) if (sh.present[i]) { sh.idx[i] = pos; pos += sh.lsize[i]; }
            const int lm0 = pos;
            sh.lm0 = lm0;
            pos += nl0;
            sh.m = pos;
            for (int i = sh.first_kept; i < nb; ++i) if (sh.present[i] && !sh.drop[i]) { sh.idx[i] = pos; pos += sh.lsize[i]; }
            sh.pos = pos; sh.n = pos - sh.m;
            for (int f = 0; f < BA_NF; ++f) {
                int i = find_block(sh, VRF_BLK_POSE, f); sh.col_pose[f] = (i >= 0 && sh.present[i]) ? sh.idx[i] : -1;
                i = find_block(sh, VRF_BLK_SPEEDBIAS, f); sh.col_sb[f] = (i >= 0 && sh.present[i]) ? sh.idx[i] : -1;
            }
            { int i = find_block(sh, VRF_BLK_EXPOSE, 0); sh.col_ex = (i >= 0 && sh.present[i]) ? sh.idx[i] : -1; }
            { int i = find_block(sh, VRF_BLK_TD, 0); sh.col_td = (i >= 0 && sh.present[i]) ? sh.idx[i] : -1; }
        }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < BA_MAX_LM / BA_THREADS; ++u) {
        const int l = u * BA_THREADS + tid;
        if (l >= M) continue;
        int col = -1;
        if (ldrop[u]) {
            int before = 0;
            for (int c = 0; c < u * (BA_THREADS / 32) + warp; ++c) before += sh.chunk[c];
            col = sh.lm0 + before + lrank[u];
        }
        mg.lmcol[l] = col;
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

    MPROF(0);
    // ---- 2a. prior factor, information form (the solve kernel has normalised the store): A += HP, b += gp + HP dx ----
    if (P) {
        const int np = P->n;
        const double *HPm = P->J0, *gp = P->r0;
        prior_dx(P, pose, sb, ex, out.mtd, sh.dx);          // at the current (gauge-fixed) states
        __syncthreads();
        for (int rI = warp; rI < np; rI += nwarp) {       // one warp per row, lanes along the row (coalesced)
            const double *row = HPm + (size_t)rI * np;
            double a = 0;
            for (int k = lane; k < np; k += 32) a += row[k] * sh.dx[k];
            a = warp_sum_d(a);
            if (lane == 0) sh.pr[rI] = a + gp[rI];
        }
        // prior column -> A column (reuse p.colmap scratch)
        for (int b = tid; b < P->n_blocks; b += BA_THREADS) {
            int i = find_block(sh, P->kind[b], P->index[b]);
            const int ls = P->size[b] == 7 ? 6 : P->size[b];
            for (int c = 0; c < ls; ++c) p.colmap[P->idx[b] + c] = (i >= 0) ? sh.idx[i] + c : -1;
        }
        __syncthreads();
        for (int a = tid; a < np; a += BA_THREADS) if (p.colmap[a] >= 0) bv[p.colmap[a]] += sh.pr[a];
        for (int e = tid; e < np * np; e += BA_THREADS) {
            int a = e / np, c = e - a * np;
            int ca = p.colmap[a], cc = p.colmap[c];
            if (ca < 0 || cc < 0) continue;
            A[(size_t)ca * pos + cc] += HPm[e];
        }
        __syncthreads();
    }
    MPROF(1);
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
            A[(size_t)acol(a) * pos + acol(c)] += h;            // distinct (a, c) -> distinct entries, this warp is the only writer
        }
        if (lane < 30) { double gsum = 0; for (int k = 0; k < 15; ++k) gsum += sh.J[k * 30 + lane] * sh.r[k]; bv[acol(lane)] += gsum; }
    }
    __syncthreads();
    // ---- 2c. projection factors hosted at frame 0 (all four parameter blocks, Cauchy corrector) ----
    // Accumulated on chip: the pose / ex-pose / td part in a packed MG_L x MG_L block in shared memory
    // (local index: pose f -> 6f, ex-pose -> 66, td -> 72), the landmark rows (coupling w[MG_L], h, g) in HBM.
    double *big0 = reinterpret_cast<double *>(smem_raw + ((sizeof(MargShared) + 15) & ~(size_t)15));
    double *H72 = big0;                     // MG_LP
    double *g72 = big0 + MG_LP;             // MG_L
    double *Wm = mg.Ainv;                   // [nl0][MG_L] landmark coupling rows (Ainv is only used by the slow path, later)
    double *hm = mg.V2;                     // [nl0] h, then [nl0] g   (V2 is only needed by the second decomposition)
    int lm0c = 0;
    for (int i = 0; i < sh.first_kept; ++i) if (sh.present[i]) lm0c += sh.lsize[i];
    const int nl0c = mm - lm0c;
    if (flag == VRF_MARGIN_OLD && nl0c > 0) {
        // Bit-reproducible accumulation without floating-point atomics, frame-pair major like the linearisation of k_ba_solve:
        // warp j - 1 owns the pair (host 0, observer j) and walks its factors (dropped landmarks seen in frame j, in landmark
        // order) in batches of 32, one lane per factor.  Per batch the sums over the lanes are formed with the reduce-scatter
        // butterfly and kept in registers over the batches:
        //   (b) the observer's own rows (gradient 6, diagonal 6 x 6 triangle 21, coupling with pose 0 / ex-pose / td 6 x 13):
        //       this pair is their only contributor -> stored straight into the shared system;
        //   (a) the 13 columns all factors share (pose 0, ex-pose, td: triangle 91 + gradient 13): per-pair partials in the
        //       L2 scratch, added over the pairs in order afterwards;
        //   landmark sums (coupling with the 13 shared columns, h, g): one slot per factor, added per landmark afterwards.
        for (int e = tid; e < MG_LP + MG_L; e += BA_THREADS) big0[e] = 0.0;
        for (int e = tid; e < nl0c * MG_L; e += BA_THREADS) Wm[e] = 0.0;
        int *flist = p.fac;                          // k_ba_solve's factor list is dead by now: [pair offsets] lists of landmarks
        for (int pass = 0; pass < 2; ++pass) {
            if (warp < BA_NF - 1) {
                const int j = warp + 1;
                int off = pass ? sh.pair_off[warp] : 0;
                for (int b0 = 0; b0 < M; b0 += 32) {
                    const int l = b0 + lane;
                    const bool a_ = l < M && mg.lmcol[l] >= 0 && (p.obs_ptr[l + 1] - p.obs_ptr[l] - 1) >= j;
                    const unsigned mask = __ballot_sync(0xffffffffu, a_);
                    if (pass && a_) flist[off + __popc(mask & ((1u << lane) - 1u))] = l;
                    off += __popc(mask);
                }
                if (!pass && lane == 0) sh.pair_cnt[warp] = off;
            }
            __syncthreads();
            if (!pass) {
                if (tid == 0) {
                    int acc = 0;
                    for (int q = 0; q < BA_NF - 1; ++q) { sh.pair_off[q] = acc; acc += sh.pair_cnt[q]; }
                    sh.pair_off[BA_NF - 1] = acc;
                }
                __syncthreads();
            }
        }
        __threadfence_block();
        auto pk72 = [](int a_, int b_) { return a_ >= b_ ? a_ * (a_ + 1) / 2 + b_ : b_ * (b_ + 1) / 2 + a_; };
        auto scol = [](int a_) { return a_ < 6 ? a_ : a_ < 12 ? 66 + (a_ - 6) : 72; };
        if (warp < BA_NF - 1) {
            const int j = warp + 1;
            double acca[7] = {0, 0, 0, 0, 0, 0, 0}, accb[7] = {0, 0, 0, 0, 0, 0, 0};
            const int f_end = sh.pair_off[warp + 1];
            for (int fb = sh.pair_off[warp]; fb < f_end; fb += 32) {
                const bool act = fb + lane < f_end;
                const int l = act ? flist[fb + lane] : 0;
                const int o0 = act ? p.obs_ptr[l] : 0;
                double r[2] = {0, 0}, Ji[12], Jj[12], Jl[2] = {0, 0}, Je[12], Jt[2] = {0, 0};
#pragma unroll
                for (int k = 0; k < 12; ++k) { Ji[k] = 0; Jj[k] = 0; Je[k] = 0; }
                if (act) {
                    double xi, yi, xj, yj;
                    obs_at(m, p, o0, out.mtd, xi, yi);
                    obs_at(m, p, o0 + j, out.mtd, xj, yj);
                    proj_eval(pose, sh.R, pose + 7 * j, sh.R + 9 * j, ex, sh.ric, lam[l], xi, yi, xj, yj, true, false, r, Ji, Jj, Jl, Je,
                              m.td_factor ? p.obs_vel + 2 * o0 : nullptr, m.td_factor ? p.obs_vel + 2 * (o0 + j) : nullptr,
                              m.td_factor ? Jt : nullptr);
                }
                double s0[13], s1[13];
#pragma unroll
                for (int c = 0; c < 6; ++c) { s0[c] = Ji[c]; s1[c] = Ji[6 + c]; s0[6 + c] = Je[c]; s1[6 + c] = Je[6 + c]; }
                s0[12] = Jt[0]; s1[12] = Jt[1];
                if (act) {
                    double *fp = p.fpart + (size_t)(o0 + j) * BA_FP_STRIDE;
#pragma unroll
                    for (int c = 0; c < 13; ++c) fp[c] = s0[c] * Jl[0] + s1[c] * Jl[1];
                    fp[13] = Jl[0] * Jl[0] + Jl[1] * Jl[1];
                    fp[14] = Jl[0] * r[0] + Jl[1] * r[1];
                    double *wr = Wm + (size_t)(mg.lmcol[l] - lm0c) * MG_L;
#pragma unroll
                    for (int a_ = 0; a_ < 6; ++a_) wr[6 * j + a_] = Jj[a_] * Jl[0] + Jj[6 + a_] * Jl[1];
                }
#pragma unroll
                for (int rd = 0; rd < 7; ++rd) {
                    double v[16];
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int e = rd * 16 + t;
                        if (e < 91) {
                            int a_ = 0;
                            while ((a_ + 1) * (a_ + 2) / 2 <= e) ++a_;
                            const int b_ = e - a_ * (a_ + 1) / 2;
                            v[t] = s0[a_] * s0[b_] + s1[a_] * s1[b_];
                        } else if (e < 104) v[t] = s0[e - 91] * r[0] + s1[e - 91] * r[1];
                        else v[t] = 0.0;
                    }
                    acca[rd] += reduce_scatter16(v, lane);
                }
#pragma unroll
                for (int rd = 0; rd < 7; ++rd) {
                    double v[16];
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int e = rd * 16 + t;
                        if (e < 6) v[t] = Jj[e] * r[0] + Jj[6 + e] * r[1];
                        else if (e < 27) {
                            int a_ = 0;
                            while ((a_ + 1) * (a_ + 2) / 2 <= e - 6) ++a_;
                            const int b_ = e - 6 - a_ * (a_ + 1) / 2;
                            v[t] = Jj[a_] * Jj[b_] + Jj[6 + a_] * Jj[6 + b_];
                        } else if (e < 105) {
                            const int a_ = (e - 27) / 13, c_ = (e - 27) % 13;
                            v[t] = Jj[a_] * s0[c_] + Jj[6 + a_] * s1[c_];
                        } else v[t] = 0.0;
                    }
                    accb[rd] += reduce_scatter16(v, lane);
                }
            }
            if (!(lane & 1)) {
                double *pp = p.pair_part + (size_t)warp * BA_PP_STRIDE + (lane >> 1);
#pragma unroll
                for (int rd = 0; rd < 7; ++rd) {
                    pp[16 * rd] = acca[rd];
                    const int e = rd * 16 + (lane >> 1);
                    if (e < 6) g72[6 * j + e] = accb[rd];
                    else if (e < 27) {
                        int a_ = 0;
                        while ((a_ + 1) * (a_ + 2) / 2 <= e - 6) ++a_;
                        H72[pk72(6 * j + a_, 6 * j + (e - 6 - a_ * (a_ + 1) / 2))] = accb[rd];
                    } else if (e < 105) H72[pk72(6 * j + (e - 27) / 13, scol((e - 27) % 13))] = accb[rd];
                }
            }
        }
        __syncthreads();
        // shared 13 x 13 block and gradient: sum of the per-pair partials in pair order; landmark rows: sum over the landmark's
        // factors in observation order
        for (int it = tid; it < 104 + M * 15; it += BA_THREADS) {
            if (it < 104) {
                double t = 0;
                for (int q = 0; q < BA_NF - 1; ++q) t += p.pair_part[(size_t)q * BA_PP_STRIDE + it];
                if (it < 91) {
                    int a_ = 0;
                    while ((a_ + 1) * (a_ + 2) / 2 <= it) ++a_;
                    H72[pk72(scol(a_), scol(it - a_ * (a_ + 1) / 2))] = t;
                } else g72[scol(it - 91)] = t;
                continue;
            }
            const int l = (it - 104) / 15, a_ = (it - 104) - 15 * l;
            const int cl = mg.lmcol[l];
            if (cl < 0) continue;
            const int li = cl - lm0c;
            const int o0 = p.obs_ptr[l], o1 = p.obs_ptr[l + 1];
            double t = 0;
            for (int o = o0 + 1; o < o1 && o - o0 <= BA_NF - 1; ++o) t += p.fpart[(size_t)o * BA_FP_STRIDE + a_];
            if (a_ < 13) Wm[(size_t)li * MG_L + scol(a_)] = t;
            else hm[(a_ == 13 ? 0 : nl0c) + li] = t;
        }
        __syncthreads();
        __threadfence();
        // scatter into A / b
        auto acol72 = [&](int q) { return q < 66 ? (sh.col_pose[q / 6] < 0 ? -1 : sh.col_pose[q / 6] + q % 6)
                                          : q < 72 ? (sh.col_ex < 0 ? -1 : sh.col_ex + (q - 66)) : sh.col_td; };
        for (int e = tid; e < MG_L * MG_L; e += BA_THREADS) {
            const int a_ = e / MG_L, c = e - a_ * MG_L;
            const int ca = acol72(a_), cc = acol72(c);
            if (ca < 0 || cc < 0) continue;
            const double v = H72[pk72(a_, c)];
            if (v != 0.0) A[(size_t)ca * pos + cc] += v;
        }
        for (int q = tid; q < MG_L; q += BA_THREADS) { const int ca = acol72(q); if (ca >= 0) bv[ca] += g72[q]; }
        for (int e = tid; e < nl0c * (MG_L + 1); e += BA_THREADS) {
            const int li = e / (MG_L + 1), q = e - li * (MG_L + 1);
            const int cl = lm0c + li;
            if (q == MG_L) { A[(size_t)cl * pos + cl] = hm[li]; bv[cl] = hm[nl0c + li]; }
            else {
                const int ca = acol72(q);
                const double v = Wm[(size_t)li * MG_L + q];
                if (ca >= 0 && v != 0.0) { A[(size_t)cl * pos + ca] = v; A[(size_t)ca * pos + cl] = v; }
            }
        }
    }
    __syncthreads();
    __threadfence();
    MPROF(2);
    // ---- 3./4. Schur complement A' = Arr - Arm Amm^+ Amr, b' = br - Arm Amm^+ bm ----
    // The reference forms Amm^+ = V diag(lambda > 1e-8 ? 1/lambda : 0) V^T (marginalization_factor.cpp:273-287).
    // Amm is an arrow matrix [[P (pose0 | speed-bias0, <= 15), C], [C^T, D]] with D diagonal (one entry per
    // dropped landmark; landmarks never couple to each other).  If Amm - 1e-8 I is positive definite, every
    // eigenvalue exceeds the threshold, the pseudo-inverse IS the inverse, and the Schur complement is formed
    // exactly by block elimination (landmarks, then the dense head).  Otherwise fall back to the explicit
    // eigen-decomposition (parallel Jacobi), as the reference does.
    double *big = reinterpret_cast<double *>(smem_raw + ((sizeof(MargShared) + 15) & ~(size_t)15));
    double *Ar = mg.Ar, *V2 = mg.V2, *br = mg.br;
    double *Tm = mg.T, *V = mg.V, *Ainv = mg.Ainv;
    const int nh = sh.first_kept > 0 ? (mm - 0) : 0;       // placeholder, head size computed below
    (void)nh;
    // head = dropped non-landmark columns [0, lm0), landmarks = [lm0, mm)
    int lm0 = 0;
    for (int i = 0; i < sh.first_kept; ++i) if (sh.present[i]) lm0 += sh.lsize[i];
    const int nl0 = mm - lm0;
    const double eps = 1e-8;
    // PD test of Amm - eps I
    if (tid == 0) sh.flag = 0;
    __syncthreads();
    for (int l = tid; l < nl0; l += BA_THREADS) if (!(A[(size_t)(lm0 + l) * pos + lm0 + l] - eps > 0.0)) sh.flag = 1;
    __syncthreads();
    double *Hs = big;                    // head system (lm0 x lm0), then its inverse
    if (!sh.flag && lm0 > 0) {
        for (int e = tid; e < lm0 * lm0; e += BA_THREADS) {
            int i = e / lm0, j = e - i * lm0;
            double a = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]) - (i == j ? eps : 0.0);
            for (int l = 0; l < nl0; ++l) a -= A[(size_t)i * pos + lm0 + l] * A[(size_t)j * pos + lm0 + l] / (A[(size_t)(lm0 + l) * pos + lm0 + l] - eps);
            Hs[e] = a;
        }
        __syncthreads();
        if (tid == 0) {        // tiny Cholesky (<= 15x15) as the PD test
            for (int j = 0; j < lm0 && !sh.flag; ++j) {
                double d = Hs[j * lm0 + j];
                for (int k = 0; k < j; ++k) d -= Hs[j * lm0 + k] * Hs[j * lm0 + k];
                if (!(d > 0.0)) { sh.flag = 1; break; }
                d = sqrt(d); Hs[j * lm0 + j] = d;
                for (int i = j + 1; i < lm0; ++i) { double t = Hs[i * lm0 + j]; for (int k = 0; k < j; ++k) t -= Hs[i * lm0 + k] * Hs[j * lm0 + k]; Hs[i * lm0 + j] = t / d; }
            }
        }
        __syncthreads();
    }
    const bool fast = (sh.flag == 0);
    MPROF(3);
    if (fast) {
        // (i) eliminate the landmarks from the [head | kept] system: X = head (lm0) + kept (nn) columns
        const int nx = lm0 + nn;
        double *Xs = mg.V;                       // nx*nx scratch in global (L2)
        double *bx = mg.T;                       // nx
        auto xcol = [&](int a) { return a < lm0 ? a : mm + (a - lm0); };
        // landmark elimination Xs = A_XX - sum_l w_l w_l^T / h_l with the coupling rows streamed through a
        // shared-memory tile (64 landmarks x MG_L, pre-scaled by 1/sqrt(h)); the head inverse scratch (Hs) is
        // rebuilt afterwards, so the tile may use the whole dynamic buffer behind it.
        double *tile = big + 4096;
        short *xm = reinterpret_cast<short *>(big + 4096 + 64 * MG_L);     // X index -> local index (or -1)
        for (int a_ = tid; a_ < nx; a_ += BA_THREADS) {
            const int ca = xcol(a_);
            int q = -1;
            for (int f = 0; f < BA_NF; ++f) if (sh.col_pose[f] >= 0 && ca >= sh.col_pose[f] && ca < sh.col_pose[f] + 6) q = 6 * f + (ca - sh.col_pose[f]);
            if (sh.col_ex >= 0 && ca >= sh.col_ex && ca < sh.col_ex + 6) q = 66 + (ca - sh.col_ex);
            if (sh.col_td >= 0 && ca == sh.col_td) q = 72;
            xm[a_] = (short)q;
        }
        for (int e = tid; e < nx * nx; e += BA_THREADS) {
            int i = e / nx, j = e - i * nx;
            const int ci = xcol(i), cj = xcol(j);
            Xs[e] = 0.5 * (A[(size_t)ci * pos + cj] + A[(size_t)cj * pos + ci]);
        }
        for (int i = tid; i < nx; i += BA_THREADS) bx[i] = bv[xcol(i)];
        __syncthreads();
        if (nl0 > 0) {
            double accs[72];      // up to ceil(nx*nx/512) entries per thread (nx <= 186 -> 68)
            int nacc = 0;
            for (int e = tid; e < nx * nx; e += BA_THREADS) if (nacc < 72) accs[nacc++] = 0.0;
            double bacc = 0.0;
            for (int t0 = 0; t0 < nl0; t0 += 64) {
                const int nt = min(64, nl0 - t0);
                for (int q = tid; q < nt * MG_L; q += BA_THREADS) {
                    const int li = t0 + q / MG_L;
                    tile[q] = Wm[(size_t)li * MG_L + (q % MG_L)] / sqrt(hm[li]);
                }
                __syncthreads();
                int k = 0;
                for (int e = tid; e < nx * nx && k < 72; e += BA_THREADS, ++k) {
                    const int i = e / nx, j = e - i * nx;
                    const int qi = xm[i], qj = xm[j];
                    if (qi < 0 || qj < 0) continue;
                    double s_ = 0;
                    for (int l = 0; l < nt; ++l) s_ += tile[l * MG_L + qi] * tile[l * MG_L + qj];
                    accs[k] += s_;
                }
                if (tid < nx && xm[tid] >= 0) {
                    double s_ = 0;
                    for (int l = 0; l < nt; ++l) s_ += tile[l * MG_L + xm[tid]] * hm[nl0 + t0 + l] / sqrt(hm[t0 + l]);
                    bacc += s_;
                }
                __syncthreads();
            }
            int k = 0;
            for (int e = tid; e < nx * nx && k < 72; e += BA_THREADS, ++k) Xs[e] -= accs[k];
            if (tid < nx) bx[tid] -= bacc;
        }
        __syncthreads();
        // (ii) invert the head block (Cholesky, one thread per column of the inverse) and eliminate it
        if (lm0 > 0) {
            if (tid == 0) {
                for (int e = 0; e < lm0 * lm0; ++e) Hs[e] = Xs[(e / lm0) * nx + (e % lm0)];
                for (int j = 0; j < lm0; ++j) {
                    double d = Hs[j * lm0 + j];
                    for (int k = 0; k < j; ++k) d -= Hs[j * lm0 + k] * Hs[j * lm0 + k];
                    d = sqrt(d); Hs[j * lm0 + j] = d;
                    for (int i = j + 1; i < lm0; ++i) { double t = Hs[i * lm0 + j]; for (int k = 0; k < j; ++k) t -= Hs[i * lm0 + k] * Hs[j * lm0 + k]; Hs[i * lm0 + j] = t / d; }
                }
            }
            __syncthreads();
            double *Hi = Hs + 256;               // inverse (lm0 x lm0)
            if (tid < lm0) {
                double x[16];
                for (int i = 0; i < lm0; ++i) x[i] = (i == tid) ? 1.0 : 0.0;
                for (int i = 0; i < lm0; ++i) { double t = x[i]; for (int k = 0; k < i; ++k) t -= Hs[i * lm0 + k] * x[k]; x[i] = t / Hs[i * lm0 + i]; }
                for (int i = lm0 - 1; i >= 0; --i) { double t = x[i]; for (int k = i + 1; k < lm0; ++k) t -= Hs[k * lm0 + i] * x[k]; x[i] = t / Hs[i * lm0 + i]; }
                for (int i = 0; i < lm0; ++i) Hi[i * lm0 + tid] = x[i];
            }
            __syncthreads();
            // T15 = X_r,head * Hi  (nn x lm0) in smem after Hi
            double *T15 = Hi + 256;
            for (int e = tid; e < nn * lm0; e += BA_THREADS) {
                int i = e / lm0, j = e - i * lm0;
                double a = 0;
                for (int k = 0; k < lm0; ++k) a += Xs[(size_t)(lm0 + i) * nx + k] * Hi[k * lm0 + j];
                T15[e] = a;
            }
            __syncthreads();
            for (int e = tid; e < nn * nn; e += BA_THREADS) {
                int i = e / nn, j = e - i * nn;
                double a = Xs[(size_t)(lm0 + i) * nx + lm0 + j];
                for (int k = 0; k < lm0; ++k) a -= T15[i * lm0 + k] * Xs[(size_t)k * nx + lm0 + j];
                Ar[e] = a;
            }
            for (int i = tid; i < nn; i += BA_THREADS) {
                double a = bx[lm0 + i];
                for (int k = 0; k < lm0; ++k) a -= T15[i * lm0 + k] * bx[k];
                br[i] = a;
            }
        } else {
            for (int e = tid; e < nn * nn; e += BA_THREADS) Ar[e] = Xs[e];
            for (int i = tid; i < nn; i += BA_THREADS) br[i] = bx[i];
        }
        __syncthreads();
    } else {
        // explicit pseudo-inverse through the eigen-decomposition of Amm (global-memory Jacobi)
        double *Amm = Ainv;
        for (int e = tid; e < mm * mm; e += BA_THREADS) { int i = e / mm, j = e - i * mm; Amm[e] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]); }
        __syncthreads();
        jacobi_eig(Amm, V, mm, sh);
        for (int k = tid; k < mm; k += BA_THREADS) { double w = Amm[(size_t)k * mm + k]; sh.cs[k] = (w > eps) ? 1.0 / w : 0.0; }
        __syncthreads();
        for (int e = tid; e < mm * mm; e += BA_THREADS) {
            int i = e / mm, j = e - i * mm;
            double a = 0;
            for (int k = 0; k < mm; ++k) a += V[(size_t)i * mm + k] * sh.cs[k] * V[(size_t)j * mm + k];
            Ainv[e] = a;
        }
        __syncthreads();
        for (int e = tid; e < nn * mm; e += BA_THREADS) {
            int i = e / mm, j = e - i * mm;
            double a = 0;
            for (int k = 0; k < mm; ++k) a += A[(size_t)(mm + i) * pos + k] * Ainv[(size_t)k * mm + j];
            Tm[e] = a;
        }
        __syncthreads();
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
    }
    MPROF(4);
    // ---- 5. the new prior in information form: HP = A' (symmetric, from its upper triangle -- what
    // SelfAdjointEigenSolver would read), gp = b'.  The reference factors A' = V S V^T, zeroes eigenvalues <= 1e-8 and
    // stores J0 = sqrt(S) V^T, r0 = S^-1/2 V^T b' (marginalization_factor.cpp:298-308); J0^T J0 and J0^T r0 equal A' and
    // b' up to the truncated part (|lambda| <= 1e-8 against ||A'|| ~ 1e7: below FP64 resolution of the products), so the
    // decomposition is only run when the factor form itself is asked for (k_ba_prior_factor).  The constant
    // c0 = 1/2 r0^T r0 = 1/2 b'^T A'^+ b' comes from a diagonally pivoted Cholesky that stops at pivots <= 1e-8. ----
    double c0acc = 0.0;
    {
        const int cap = MARG_A_ELEMS(MARG_SMEM_N) + MARG_V_ELEMS(MARG_SMEM_N);
        const bool insm = nn * nn + 3 * nn + 8 <= cap;
        double *Mw = insm ? big : V2;
        double *yv = insm ? big + nn * nn : Tm;
        double *colb = yv + nn;
        int *done = reinterpret_cast<int *>(colb + nn);
        for (int e = tid; e < nn * nn; e += BA_THREADS) {
            const int i = e / nn, j = e - i * nn;
            const double v = (j >= i) ? Ar[(size_t)i * nn + j] : Ar[(size_t)j * nn + i];
            Q->J0[e] = v; Mw[e] = v;
        }
        for (int i = tid; i < nn; i += BA_THREADS) { const double v = br[i]; Q->r0[i] = v; yv[i] = v; done[i] = 0; }
        __syncthreads();
        for (int k = 0; k < nn; ++k) {
            // pivot = largest remaining diagonal entry
            double v = -1.0;
            int bi = tid;
            for (int i = tid; i < nn; i += BA_THREADS) { const double d_ = done[i] ? -1.0 : Mw[(size_t)i * nn + i]; if (d_ > v) { v = d_; bi = i; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > v || (ov == v && oi < bi)) { v = ov; bi = oi; }
            }
            if (lane == 0) { sh.red[warp] = v; sh.chunk[warp] = bi; }
            __syncthreads();
            double dmax = -1.0;
            int pv = 0;
            for (int w = 0; w < nwarp; ++w) { const double ov = sh.red[w]; const int oi = sh.chunk[w]; if (ov > dmax || (ov == dmax && oi < pv)) { dmax = ov; pv = oi; } }
            if (!(dmax > 1e-8)) break;                         // uniform: every thread sees the same (dmax, pv)
            const double l_ = sqrt(dmax), yk = yv[pv] / l_;
            c0acc += yk * yk;
            __syncthreads();
            for (int i = tid; i < nn; i += BA_THREADS) {
                double cb = 0.0;
                if (!done[i] && i != pv) { cb = Mw[(size_t)i * nn + pv] / l_; yv[i] -= cb * yk; }
                colb[i] = cb;
            }
            if (tid == 0) done[pv] = 1;
            __syncthreads();
            for (int i = warp; i < nn; i += nwarp) {          // one warp per remaining row, lanes along the row
                const double ci = colb[i];
                if (ci == 0.0) continue;
                double *row = Mw + (size_t)i * nn;
                for (int j = lane; j < nn; j += 32) row[j] -= ci * colb[j];
            }
            __syncthreads();
        }
        __syncthreads();
    }
    MPROF(5);
    if (tid == 0) {
        int nk = 0;
        for (int i = sh.first_kept; i < sh.nb; ++i) {
            if (!sh.present[i] || sh.drop[i]) continue;
            Q->kind[nk] = sh.kind[i]; Q->size[nk] = sh.gsize[i]; Q->idx[nk] = sh.idx[i] - mm;
            const double *src = sh.kind[i] == VRF_BLK_POSE ? pose + 7 * sh.index[i] : sh.kind[i] == VRF_BLK_SPEEDBIAS ? sb + 9 * sh.index[i]
                                : sh.kind[i] == VRF_BLK_TD ? &out.mtd : ex;
            for (int c = 0; c < sh.gsize[i]; ++c) Q->x0[9 * nk + c] = src[c];
            if (sh.kind[i] == VRF_BLK_EXPOSE || sh.kind[i] == VRF_BLK_TD) Q->index[nk] = 0;
            else if (flag == VRF_MARGIN_OLD) Q->index[nk] = sh.index[i] - 1;
            else Q->index[nk] = (sh.index[i] == VRF_WINDOW_SIZE) ? VRF_WINDOW_SIZE - 1 : sh.index[i];
            ++nk;
        }
        Q->n = nn; Q->n_blocks = nk; Q->valid = 1; Q->form = 1; Q->c0 = 0.5 * c0acc;
        out.has_new_prior = 1;
        MPROF(6);
        for (int k = 0; k < 7; ++k) out.prof2[k] = tp[k];
        out.prof2[3] = fast ? tp[3] : -tp[3];          // negative => slow (explicit eigen) path was taken
    }
}

// 164 KB carve-out (minus the 1 KB reserve): leaves 92 KB of L1 for the global scratch of the elimination phases
static_assert(((sizeof(MargShared) + 15) & ~(size_t)15) + sizeof(double) * (MARG_A_ELEMS(MARG_SMEM_N) + MARG_V_ELEMS(MARG_SMEM_N)) <= 164 * 1024 - 1024,
              "marginalization frame must fit the 164 KB carve-out");
// ---------------------------------------------------------------------------
// k_ba_prior_factor: the reference's factor form of a stored (information-form) prior, on demand:
// A' = V S V^T (cyclic Jacobi, as Eigen::SelfAdjointEigenSolver in marginalization_factor.cpp:298),
// S = lambda > 1e-8 ? lambda : 0, linearized_jacobians = sqrt(S) V^T, linearized_residuals = S^-1/2 V^T b' (:299-308).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BA_THREADS, 1)
k_ba_prior_factor(const BaPriorStore *src, BaPriorStore *dst, double *scratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MargShared &sh = *reinterpret_cast<MargShared *>(smem_raw);
    double *big = reinterpret_cast<double *>(smem_raw + ((sizeof(MargShared) + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    const int nn = src->n;
    if (tid == 0) { sh.jdbg = 0; dst->n = nn; dst->n_blocks = src->n_blocks; dst->valid = 1; dst->form = 0; dst->c0 = src->c0; }
    for (int b = tid; b < src->n_blocks; b += BA_THREADS) { dst->kind[b] = src->kind[b]; dst->index[b] = src->index[b]; dst->size[b] = src->size[b]; dst->idx[b] = src->idx[b]; }
    for (int i = tid; i < src->n_blocks * 9; i += BA_THREADS) dst->x0[i] = src->x0[i];
    __syncthreads();
    const double *HPm = src->J0, *gp = src->r0;
    if (nn <= MARG_SMEM_N) {
        const int ne2 = MARG_NE(nn), ld2 = MARG_LDA(nn);
        double *As = big, *Vs = big + MARG_A_ELEMS(nn);
        for (int e = tid; e < ne2 * ld2; e += BA_THREADS) {
            const int i = e / ld2, j = e - i * ld2;
            As[e] = (i < nn && j < nn && j >= i) ? HPm[(size_t)i * nn + j] : 0.0;      // upper triangle
        }
        __syncthreads();
        jacobi_eig_smem(As, Vs, nn, sh);
        for (int k = tid; k < nn; k += BA_THREADS) {
            const double w = As[k * ld2 + k];
            const double S = w > 1e-8 ? w : 0.0, Sinv = w > 1e-8 ? 1.0 / w : 0.0;
            const double ss = sqrt(S), sis = sqrt(Sinv);
            double vb = 0;
            for (int j = 0; j < nn; ++j) { const double vjk = Vs[k * ne2 + j]; dst->J0[(size_t)k * nn + j] = ss * vjk; vb += vjk * gp[j]; }   // VT[col k][row j]
            dst->r0[k] = sis * vb;
        }
    } else {
        double *Ag = scratch, *Vg = scratch + (size_t)VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM;
        for (int e = tid; e < nn * nn; e += BA_THREADS) Ag[e] = HPm[e];
        __syncthreads();
        jacobi_eig(Ag, Vg, nn, sh);
        for (int k = tid; k < nn; k += BA_THREADS) {
            const double w = Ag[(size_t)k * nn + k];
            const double S = w > 1e-8 ? w : 0.0, Sinv = w > 1e-8 ? 1.0 / w : 0.0;
            const double ss = sqrt(S), sis = sqrt(Sinv);
            double vb = 0;
            for (int j = 0; j < nn; ++j) { const double vjk = Vg[(size_t)j * nn + k]; dst->J0[(size_t)k * nn + j] = ss * vjk; vb += vjk * gp[j]; }
            dst->r0[k] = sis * vb;
        }
    }
}

static size_t marg_smem() { return ((sizeof(MargShared) + 15) & ~(size_t)15) + sizeof(double) * (MARG_A_ELEMS(MARG_SMEM_N) + MARG_V_ELEMS(MARG_SMEM_N)); }
size_t ba_marg_smem_bytes() { return marg_smem(); }

// per-device shared-memory opt-in of the two marginalization kernels (called by ba_create() for the handle's device)
int ba_marg_configure()
{
    if (cudaFuncSetAttribute(k_ba_prior_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)marg_smem()) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_ba_marg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)marg_smem()) != cudaSuccess) return -1;
    return 0;
}

int ba_prior_factor_launch(const BaPriorStore *src, BaPriorStore *dst, double *scratch, LaunchCtx &lc)
{
    lc.begin(K_BA_PRIOR_FACTOR);
    k_ba_prior_factor<<<1, BA_THREADS, marg_smem(), lc.st>>>(src, dst, scratch);
    lc.end();
    return 0;
}

int ba_marg_launch(const BaMeta *d_meta, const BaProbDev *d_prob, BaOutDev *d_out, BaMargDev *d_marg, int n, LaunchCtx &lc)
{
    lc.begin(K_BA_MARG);
    k_ba_marg<<<n, BA_THREADS, marg_smem(), lc.st>>>(d_meta, d_prob, d_out, d_marg);
    lc.end();
    return 0;
}

}  // namespace vrf
