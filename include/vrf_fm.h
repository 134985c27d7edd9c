/*
 * vrf_fm.h -- C ABI of the steps either side of Estimator::optimization() (SURVEY.md section 8f,
 * rows 3 and 4).  Included by vrf.h.
 *
 *   vrf_fm_triangulate_with_depth_batch      FeatureManager::triangulateWithDepth
 *                                            (vins_estimator/src/feature_manager/feature_manager.cpp:386-543),
 *                                            called right before optimization() (estimator.cpp processImage)
 *   vrf_fm_moving_consistency_check_batch    Estimator::movingConsistencyCheck (estimator.cpp:1965-2009) with
 *                                            reprojectionError / reprojectionError3D (:1944-1963), called right after
 *   vrf_imu_preintegrate_batch               IntegrationBase ctor + push_back()/propagate()/midPointIntegration
 *                                            (vins_estimator/src/factor/integration_base.h:13-162): produces the
 *                                            VrfImuPreint that vrf_ba_solve consumes
 *
 * All three are stateless batched calls: n independent sequences per launch, caller-owned host buffers in and out.
 */
#ifndef VRF_FM_H_
#define VRF_FM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRF_INIT_DEPTH 5.0      /* INIT_DEPTH (parameters.cpp:215) */

/* The landmark list of one sequence, in f_manager.feature (std::list) order.  Landmark l is observed in frames
 * lm_start_frame[l] .. lm_start_frame[l] + (lm_obs_ptr[l+1] - lm_obs_ptr[l]) - 1 (feature_per_frame). */
typedef struct VrfFmProblem {
    double  Ps[VRF_NUM_FRAMES][3];      /* Estimator::Ps */
    double  Rs[VRF_NUM_FRAMES][9];      /* Estimator::Rs, row-major */
    double  tic[3], ric[9];             /* tic[0], ric[0] */
    int32_t n_landmarks, n_obs;
    const int32_t *lm_start_frame;      /* [M]      FeaturePerId::start_frame */
    const int32_t *lm_obs_ptr;          /* [M+1]    CSR offsets into the obs_* arrays */
    const double  *obs_pts;             /* [n_obs][2] FeaturePerFrame::point.xy (normalised plane, z = 1) */
    const double  *obs_depth;           /* [n_obs]  FeaturePerFrame::depth in metres, 0 = no depth (feature_manager.h:40-65);
                                           only read by triangulate_with_depth */
    double  *estimated_depth;           /* [M] in/out FeaturePerId::estimated_depth (triangulate writes, check reads) */
    int32_t *estimate_flag;             /* [M] in/out FeaturePerId::estimate_flag (0 initial / 1 verified depth / 2 triangulated) */
    uint8_t *is_dynamic;                /* [M] in (triangulate skips dynamic landmarks) / out (consistency check) */
    uint8_t *remove;                    /* [M] out of the consistency check: 1 = feature_id goes into removeIndex; may be NULL */
} VrfFmProblem;

/* FeatureManager::triangulateWithDepth for n sequences.  Per landmark with estimated_depth <= 0, not dynamic,
 * >= 2 observations and start_frame < WINDOW_SIZE - 2: cross-validates every measured depth against every other
 * observation (reprojection residual < 10/460 on the normalised plane); the mean of the validated depths in the
 * host frame gives estimate_flag 1 (measured depth <= DEPTH_MAX_DIST) or 0 (only "rough" depths beyond it); a
 * landmark without any depth measurement is triangulated by SVD (flag 2, clamped at DEPTH_MIN_DIST ->
 * DEPTH_MAX_DIST); results < 0.1 fall back to INIT_DEPTH with flag 0. */
int vrf_fm_triangulate_with_depth_batch(vrf_handle *h, int n, VrfFmProblem *probs);

/* Estimator::movingConsistencyCheck for n sequences: per landmark (>= 2 observations, start_frame <
 * WINDOW_SIZE - 2, estimated_depth >= 0) the mean reprojection error and the mean 3-D error over its observers;
 * FOCAL_LENGTH * err > 10 or err3D > 2.0 marks it dynamic (is_dynamic = 1, remove = 1), else is_dynamic = 0. */
int vrf_fm_moving_consistency_check_batch(vrf_handle *h, int n, VrfFmProblem *probs);

/* One pre-integration interval: IntegrationBase(acc_0, gyr_0, ba, bg) followed by push_back(dt[k], acc[k], gyr[k]),
 * k = 0 .. n_samples-1 (estimator.cpp processIMU).  Noise densities come from VrfConfig (acc_n, gyr_n, acc_w, gyr_w). */
typedef struct VrfImuSegment {
    double  acc_0[3], gyr_0[3];         /* first sample (linearized_acc / linearized_gyr) */
    double  linearized_ba[3], linearized_bg[3];
    int32_t n_samples, reserved;
    const double *dt;                   /* [n_samples] */
    const double *acc;                  /* [n_samples][3] */
    const double *gyr;                  /* [n_samples][3] */
} VrfImuSegment;

/* IntegrationBase::propagate over every sample of n independent segments (sequential inside a segment, one warp per
 * segment across the batch): delta_p/q/v, sum_dt, the 15x15 jacobian and covariance. */
int vrf_imu_preintegrate_batch(vrf_handle *h, int n, const VrfImuSegment *segs, VrfImuPreint *out);

#ifdef __cplusplus
}
#endif
#endif /* VRF_FM_H_ */
