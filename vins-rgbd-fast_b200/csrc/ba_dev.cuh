// ba_dev.cuh -- device-side problem layout of the back end (shared by ba_kernels.cu,
// ba_marg.cu and ba_host.cu).
#pragma once
#include "common.cuh"
#include "handle.h"

#define BA_NF VRF_NUM_FRAMES
#define BA_NC 172                 // 11*6 pose + 11*9 speed-bias + 6 ex-pose + 1 td tangent columns
#define BA_COL_EX 165
#define BA_COL_TD 171
#define BA_WS 73                  // landmark coupling row: 66 pose columns + 6 ex-pose + 1 td
#define BA_THREADS 512
#define BA_MAX_LM 1024            // >= NUM_OF_F (parameters.h:14)
#define BA_MAX_OBS 8192
#define BA_MAX_M0 384             // landmarks hosted at frame 0 that one marginalization can drop
#define BA_MAX_POS (15 + BA_MAX_M0 + BA_NC)
// Deterministic accumulation (no floating-point atomics anywhere in the back end): every sum that several warps contribute
// to is first written as per-contributor partials and then added by ONE owner thread in a fixed order.
#define BA_PP_STRIDE 256          // per frame pair: [0..95] the 90 sums of pair_entry() | [96 + 16 q + e] Jg_q^T (Ji | Jj | r) | [208 + e] tri(Jg^T Jg)
#define BA_FP_STRIDE 16           // per projection factor: [0..5] host-frame part of the landmark's W row | [6] hll | [7] gl | [8..14] ex-pose / td part
#define BA_MAX_TASKS (BA_NF + BA_MAX_OBS / 32 + 64)

namespace vrf {

struct BaMeta {
    int M, nobs, nimu, np, nframes, use_imu, max_iter, marg_flag, has_prior, frame_count;
    int imu_j[BA_NF];
    int debug;            // VRF_BA_DEBUG env: bit0 = per-iteration trace (printf), bit1 = column Cholesky
    int ex_active;        // para_Ex_Pose variable (estimator.cpp:1191-1201)
    int td_factor;        // ESTIMATE_TD: ProjectionTdFactor instead of ProjectionFactor (:1270-1285)
    int td_active;        // para_Td variable (:1203-1212)
    int pad2;
    double g_norm;
    double tr_over_row;   // TR / ROW (projection_td_factor.cpp:52-53)
};

// prior as stored on device (MarginalizationInfo: marginalization_factor.h:62-74); produced by
// k_ba_marg or uploaded from a host VrfPrior.  Two forms of the same quadratic 1/2 |r0 + J0 dx|^2:
//   form 0 (factor, what the reference stores): J0 = linearized_jacobians (n x n), r0 = linearized_residuals
//   form 1 (information):                       J0 := J0^T J0 (symmetric), r0 := J0^T r0, c0 = 1/2 r0^T r0
// The kernels evaluate the prior in information form (cost = c0 + dx.(g + H dx / 2), gradient = g + H dx): k_ba_marg
// writes form 1 directly (no eigen-decomposition on the throughput path), k_ba_solve converts an uploaded form-0 prior
// in place, and k_ba_prior_factor produces the reference's factor form on demand (host copy-out of new_prior).
struct BaPriorStore {
    int n, n_blocks, valid, form;
    int kind[VRF_PRIOR_MAX_BLOCKS], index[VRF_PRIOR_MAX_BLOCKS], size[VRF_PRIOR_MAX_BLOCKS], idx[VRF_PRIOR_MAX_BLOCKS];
    double c0;
    double x0[VRF_PRIOR_MAX_BLOCKS * 9];
    double r0[VRF_PRIOR_MAX_DIM];
    double J0[VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM];
};

struct BaProbDev {
    // inputs
    const double *pose0, *sb0, *ex0;      // [11*7], [11*9], [7]
    const double *lam0;                   // [M]
    const int *start, *obs_ptr;
    const uint8_t *lm_const;
    const double *lm_ub;
    const double *obs;                    // [nobs][2]
    const double *obs_vel, *obs_td, *obs_row;   // [nobs][2], [nobs], [nobs]: ProjectionTdFactor inputs (td_factor only)
    const double *td0;                    // [1] para_Td
    const VrfImuPreint *imu;              // [10]
    const BaPriorStore *prior;            // NULL: no prior
    BaPriorStore *prior_next;             // written by the marginalization kernel
    double *HP;                           // [176*176] J0^T J0 scratch
    int *colmap;                          // [176] prior column -> tangent column (-1: constant block)
    // scratch / state
    double *lam, *clam;                   // current / candidate inverse depths
    double *lam_out;                      // optimised inverse depths of this batch item (result copy)
    double *W;                            // [M][BA_WS]
    double *hll, *gl, *jscale_l, *diag_l, *gd_l, *gn_l, *u_l, *y_l, *hinv_l, *shinv_l;   // shinv_l = sqrt(hinv_l)
    double *imuS;                         // [10][225] sqrt information (upper)
    int *fac;                             // [nobs] projection factors in frame-pair-major order: landmark | observer frame << 16
                                          //        (built once per solve by k_ba_solve; pair offsets in BaShared::pair_ptr)
    double *pair_part;                    // [45][BA_PP_STRIDE] per-pair partial sums of one linearisation
    double *fpart;                        // [nobs][BA_FP_STRIDE] per-factor contributions to the landmark sums
    double *task_cost;                    // [2][BA_MAX_TASKS] cost (and directional derivative) of each task of the dynamic queue, summed in task order
};

struct BaOutDev {
    int status, iterations, successful, termination;
    double initial_cost, final_cost;
    double pose[BA_NF * 7], sb[BA_NF * 9], ex[7];
    // double2vector (gauge fix) results
    double Ps[BA_NF * 3], Rs[BA_NF * 9], Vs[BA_NF * 3], Bas[BA_NF * 3], Bgs[BA_NF * 3];
    // states re-packed by vector2double after the gauge fix (marginalization linearises here)
    double mpose[BA_NF * 7], msb[BA_NF * 9], mex[7];
    double td, mtd;      // para_Td after the solve / as re-packed for marginalization
    int has_new_prior, armijo_failures;
    long long prof2[8];  // k_ba_marg: 0 table+zero, 1 prior, 2 imu+proj, 3 pd test, 4 schur (fast or eig), 5 jacobi, 6 output, 7 total kernel cycles of k_ba_solve
    long long prof[8];   // clock64 per phase: 0 linearise, 1 scale+grad, 2 cauchy, 3 schur, 4 cholesky, 5 solve tail, 6 dogleg, 7 candidate cost
};

// global scratch of the marginalization kernel (per problem)
struct BaMargDev {
    double *A, *b;          // [pos*pos], [pos]
    double *V, *Ainv;       // [m*m]
    double *T;              // [n*m]
    double *Ar, *V2, *br;   // [n*n], [n*n], [n]
    int *lmcol;             // [M] column of a dropped landmark (or -1)
};

// ba_kernels.cu / ba_marg.cu
size_t ba_solve_smem_bytes();
int ba_solve_configure();
int ba_marg_configure();
int ba_solve_launch(const BaMeta *d_meta, const BaProbDev *d_prob, BaOutDev *d_out, int n, LaunchCtx &lc);
int ba_marg_launch(const BaMeta *d_meta, const BaProbDev *d_prob, BaOutDev *d_out, BaMargDev *d_marg, int n, LaunchCtx &lc);
// factor form (J0, r0) of an information-form prior, on demand; scratch: 2 * VRF_PRIOR_MAX_DIM^2 doubles
int ba_prior_factor_launch(const BaPriorStore *src, BaPriorStore *dst, double *scratch, LaunchCtx &lc);

}  // namespace vrf
