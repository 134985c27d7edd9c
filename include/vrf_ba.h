/*
 * vrf_ba.h -- C ABI of the sliding-window visual-inertial bundle adjustment
 * (the back-end half of the hot path).  Included by vrf.h.
 *
 * Replaces Estimator::optimization() (vins_estimator/src/estimator/estimator.cpp:1161-1578)
 * = vector2double (:936-981) -> ceres::Problem assembly (:1166-1302) -> ceres::Solve
 * (DENSE_SCHUR + DOGLEG, :1348-1363) -> double2vector (:985-1111) -> marginalization
 * (:1376-1574, factor/marginalization_factor.cpp:3-338).
 *
 * The reference's optimisation reads/writes Estimator members
 * (para_Pose[11][7], para_SpeedBias[11][9], para_Feature[1000][1], para_Ex_Pose[1][7],
 * para_Td[1][1], estimator.h:159-164; f_manager.feature; pre_integrations[];
 * last_marginalization_info / last_marginalization_parameter_blocks, :167-168).
 * Here the same data crosses the boundary as POD arrays, gathered in exactly the
 * reference's order (FeatureManager::getDepthVector, feature_manager.cpp:302-324).
 */
#ifndef VRF_BA_H_
#define VRF_BA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRF_MARGIN_OLD          0   /* estimator.h: MARGIN_OLD */
#define VRF_MARGIN_SECOND_NEW   1   /* estimator.h: MARGIN_SECOND_NEW */

/* pass as VrfBaProblem::prior to reuse the prior the library kept on the device from the
 * previous vrf_ba_solve of this sequence (last_marginalization_info, estimator.h:167) */
#define VRF_PRIOR_DEVICE ((const VrfPrior *)(uintptr_t)1)

#define VRF_PRIOR_MAX_BLOCKS   40
#define VRF_PRIOR_MAX_DIM     176   /* 6 (ex) + 1 (td) + 11*6 + 11*9 = 172, padded */

/* kinds of parameter blocks (what the raw double* addresses of the reference encode) */
#define VRF_BLK_POSE        0   /* para_Pose[index]       size 7 (local 6) */
#define VRF_BLK_SPEEDBIAS   1   /* para_SpeedBias[index]  size 9 */
#define VRF_BLK_EXPOSE      2   /* para_Ex_Pose[0]        size 7 (local 6) */
#define VRF_BLK_TD          3   /* para_Td[0]             size 1 */

/* IntegrationBase state consumed by IMUFactor::Evaluate
 * (factor/integration_base.h:197-216, factor/imu_factor.h:20-205).  Matrices row-major. */
typedef struct VrfImuPreint {
    double sum_dt;
    double delta_p[3];
    double delta_q[4];          /* x, y, z, w */
    double delta_v[3];
    double linearized_ba[3];
    double linearized_bg[3];
    double jacobian[225];       /* 15x15 */
    double covariance[225];     /* 15x15 */
} VrfImuPreint;

/* One kept parameter block of a marginalization prior
 * (MarginalizationInfo::keep_block_{size,idx,data} + the address list
 *  last_marginalization_parameter_blocks, marginalization_factor.h:62-74). */
typedef struct VrfPriorBlock {
    int32_t kind;               /* VRF_BLK_* */
    int32_t index;              /* window frame index the block is attached to (after addr_shift) */
    int32_t size;               /* global size (7 / 9 / 1) */
    int32_t idx;                /* column offset inside the prior, i.e. keep_block_idx - m */
    double  x0[9];              /* keep_block_data: linearisation point (global parameterisation) */
} VrfPriorBlock;

/* MarginalizationInfo as consumed by MarginalizationFactor::Evaluate
 * (marginalization_factor.cpp:353-415). */
typedef struct VrfPrior {
    int32_t n;                  /* residual dimension (MarginalizationInfo::n) */
    int32_t n_blocks;
    VrfPriorBlock blocks[VRF_PRIOR_MAX_BLOCKS];
    double  linearized_jacobians[VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM]; /* n x n row-major, leading dim n */
    double  linearized_residuals[VRF_PRIOR_MAX_DIM];
} VrfPrior;

/* One optimization() call for one sequence. */
typedef struct VrfBaProblem {
    int32_t frame_count;        /* Estimator::frame_count (== VRF_WINDOW_SIZE in steady state) */
    int32_t use_imu;            /* USE_IMU */
    int32_t ex_constant;        /* 1: para_Ex_Pose constant (estimator.cpp:1191-1201: !(ESTIMATE_EXTRINSIC && frame_count ==
                                   WINDOW_SIZE && |Vs[0]| > 0.2) && !openExEstimation) */
    int32_t td_constant;        /* 1: para_Td constant (:1206-1211: !ESTIMATE_TD || |Vs[0]| < 0.2) */
    int32_t marginalization_flag;   /* VRF_MARGIN_OLD / VRF_MARGIN_SECOND_NEW */
    int32_t max_iterations;     /* NUM_ITERATIONS; 0 = use VrfConfig */
    /* states (vector2double order): [p(3), q(x,y,z,w)], [v, ba, bg] */
    double  para_Pose[VRF_NUM_FRAMES][7];
    double  para_SpeedBias[VRF_NUM_FRAMES][9];
    double  para_Ex_Pose[7];
    double  para_Td;
    /* landmarks in FeatureManager::getDepthVector order */
    int32_t n_landmarks;        /* f_manager.getFeatureCount() */
    int32_t n_obs;              /* total observations incl. the host observation of each landmark */
    const double  *para_Feature;    /* [M] inverse depth */
    const int32_t *lm_start_frame;  /* [M] FeaturePerId::start_frame (imu_i) */
    const int32_t *lm_estimate_flag;/* [M] FeaturePerId::estimate_flag (0/1/2) */
    const int32_t *lm_obs_ptr;      /* [M+1] CSR offsets into obs_pts; landmark l is seen in frames
                                       start_frame .. start_frame + (ptr[l+1]-ptr[l]) - 1 */
    const double  *obs_pts;         /* [n_obs][2] FeaturePerFrame::point.xy (normalised plane, z = 1) */
    /* IMU factors: pre_integrations[j], j = 1..frame_count, at imu[j-1] */
    const VrfImuPreint *imu;
    /* prior from the previous call (NULL: none) */
    const VrfPrior *prior;
    /* ProjectionTdFactor inputs, required iff VrfConfig::estimate_td (estimator.cpp:1270-1285,
     * factor/projection_td_factor.cpp:34-150); indexed like obs_pts */
    const double  *obs_velocity;    /* [n_obs][2] FeaturePerFrame::velocity.xy (z = 0, feature_manager.h:40-55) */
    const double  *obs_cur_td;      /* [n_obs]    FeaturePerFrame::cur_td (td at the time the frame was processed) */
    const double  *obs_row;         /* [n_obs]    FeaturePerFrame::uv.y() (pixel row, for the rolling-shutter term) */
} VrfBaProblem;

typedef struct VrfBaResult {
    int32_t status;             /* VRF_OK / VRF_SOFT_* */
    int32_t iterations;         /* trust-region iterations executed (summary.iterations.size()-1) */
    int32_t successful_steps;
    int32_t termination;        /* 0 max-iter, 1 function tol, 2 gradient tol, 3 parameter tol, 4 radius */
    double  initial_cost;
    double  final_cost;
    /* raw solver output (what ceres::Solve leaves in para_*) */
    double  para_Pose[VRF_NUM_FRAMES][7];
    double  para_SpeedBias[VRF_NUM_FRAMES][9];
    double  para_Ex_Pose[7];
    double  para_Td;
    double *para_Feature;       /* [M] caller-allocated */
    /* after double2vector (gauge fix, estimator.cpp:985-1111): world states */
    double  Ps[VRF_NUM_FRAMES][3];
    double  Rs[VRF_NUM_FRAMES][9];   /* row-major */
    double  Vs[VRF_NUM_FRAMES][3];
    double  Bas[VRF_NUM_FRAMES][3];
    double  Bgs[VRF_NUM_FRAMES][3];
    /* new prior (valid iff has_new_prior; produced when frame_count == WINDOW_SIZE) */
    int32_t has_new_prior;
    int32_t armijo_failures;    /* diagnostic: trust-region steps of a bound-constrained problem (estimate_flag == 2 landmarks,
                                   estimator.cpp:1293-1298) whose full step failed Ceres' sufficient-decrease test, i.e. for which
                                   the projected Armijo line search (TrustRegionMinimizer::DoLineSearch) had to contract the step */
    VrfPrior *new_prior;        /* caller-allocated, may be NULL to skip the copy-out */
} VrfBaResult;

/* Replaces Estimator::optimization() for one sequence (see file header). */
int vrf_ba_solve(vrf_handle *h, int seq, const VrfBaProblem *prob, VrfBaResult *res);
/* Batched: n independent sequences in the same kernel launches. */
int vrf_ba_solve_batch(vrf_handle *h, int n, const int32_t *seqs,
                       const VrfBaProblem *probs, VrfBaResult *res);

/* Pipelined form of vrf_ba_solve_batch for throughput over many sequences: submit() packs the problems, enqueues
 * their upload, the solve + marginalization kernels and the result copy, and returns without waiting; collect()
 * blocks until the OLDEST submitted batch has finished and fills `res` (same n / seqs as its submit).  Up to two
 * batches may be in flight; because a sequence's next window is built from the results of its previous one
 * (slideWindow + the new prior), the sequences of batches in flight must be disjoint (VRF_ERR_ARG otherwise) --
 * which is the natural order: different sequences publish on different frames.  VRF_ERR_CAPACITY when two batches
 * are already pending.  (Reference analogue: feature_buf between trackThread and processThread,
 * estimator_nodelet.cpp:380-384,462-479.) */
int vrf_ba_submit_batch(vrf_handle *h, int n, const int32_t *seqs, const VrfBaProblem *probs);
int vrf_ba_collect_batch(vrf_handle *h, int n, const int32_t *seqs, VrfBaResult *res);

/* Device-resident split form used for steady-state throughput measurement:
 * upload() packs and copies the problems to HBM once, enqueue() runs
 * solve + gauge fix + marginalization on the handle's stream without
 * synchronising, download() copies the results back. */
int vrf_ba_upload_batch(vrf_handle *h, int n, const int32_t *seqs, const VrfBaProblem *probs);
int vrf_ba_enqueue_batch(vrf_handle *h, int n, const int32_t *seqs);
int vrf_ba_download_batch(vrf_handle *h, int n, const int32_t *seqs, VrfBaResult *res);

#ifdef __cplusplus
}
#endif
#endif /* VRF_BA_H_ */
