/*
 * vrf.h -- C ABI of the B200-native VINS-RGBD-FAST hot path.
 *
 * The reference (jianhengLiu/VINS-RGBD-FAST) has no FFI seam inside its hot
 * path: the seams are the C++ class surfaces `FeatureTracker`
 * (vins_estimator/src/feature_tracker/feature_tracker.h:31-97) and `Estimator`
 * (vins_estimator/src/estimator/estimator.h:35-202).  This header is the plain-C
 * boundary placed directly underneath those classes: every entry point names
 * the reference member function(s) it replaces.  Host shims that keep the
 * reference's class/member names live in vins-rgbd-fast_b200/host/, and
 * INTEGRATION.md shows the glue a maintainer adds to vins_estimator.
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types; all structs are POD.
 *   - the caller owns every host buffer for the duration of a call; the
 *     library owns all device memory and all per-sequence persistent state
 *     (previous pyramid, tracks, ids, marginalization prior) inside the handle.
 *   - return value: 0 = VRF_OK, <0 = hard error (bad argument, CUDA failure),
 *     >0 = soft numerical status (reference: Ceres failures are ignored and
 *     recovery happens in Estimator::failureDetection, estimator.cpp:1113-1159).
 *   - one handle drives `n_seq` independent RGB-D+IMU sequences on one GPU
 *     (the reference: one Estimator object per sequence); a sequence is not
 *     re-entrant, distinct sequences are batched into the same kernel launches.
 *   - there is NO CPU fallback: without a CUDA device vrf_create() fails.
 */
#ifndef VRF_H_
#define VRF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRF_OK                 0
#define VRF_ERR_ARG           (-1)
#define VRF_ERR_CUDA          (-2)
#define VRF_ERR_UNSUPPORTED   (-3)
#define VRF_ERR_CAPACITY      (-4)
#define VRF_ERR_NO_DEVICE     (-5)
#define VRF_SOFT_NONFINITE      1   /* non-finite state after a solve */
#define VRF_SOFT_NOT_SPD        2   /* reduced system not SPD even after regularisation */

#define VRF_WINDOW_SIZE        10   /* parameters.h:12  WINDOW_SIZE */
#define VRF_NUM_FRAMES         11   /* WINDOW_SIZE + 1 */
#define VRF_MAX_FEATURES     1000   /* parameters.h:14  NUM_OF_F */
#define VRF_TRACK_CAP        1024   /* per-sequence capacity of the track arrays */

typedef struct vrf_handle vrf_handle;

/* Image payload formats accepted by the tracker entry points. */
#define VRF_FMT_GRAY8   0   /* what cv_bridge MONO8 hands to readImage (estimator_nodelet.cpp:292-307) */
#define VRF_FMT_RGB8    1   /* raw sensor_msgs/Image rgb8 payload; gray = cv::cvtColor RGB2GRAY fixed point */

/* Depth payload formats (sensor_msgs/Image encodings handled at estimator_nodelet.cpp:512-533). */
#define VRF_DEPTH_NONE   0  /* no depth message: the reference substitutes an all-zero image (:513-516) */
#define VRF_DEPTH_16UC1  1  /* "16UC1" / "mono16": millimetres, used as is (:518-522) */
#define VRF_DEPTH_32FC1  2  /* "32FC1": metres; the reference converts the whole frame with
                               convertTo(CV_16UC1, 1000) (:523-527) -- here only the looked-up pixels are converted */

/*
 * Configuration: the reference's YAML -> extern globals (utility/parameters.h:17-77,
 * parameters.cpp:81-243) plus the PINHOLE intrinsics read by
 * FeatureTracker::readIntrinsicParameter (feature_tracker.cpp:497-501), restated
 * as one POD struct.
 */
typedef struct VrfConfig {
    /* front end */
    int32_t row, col;                 /* ROW, COL */
    int32_t max_cnt;                  /* MAX_CNT */
    int32_t min_dist;                 /* MIN_DIST */
    int32_t num_grid_rows, num_grid_cols;
    int32_t use_imu;                  /* USE_IMU: IMU-predicted LK, maxLevel 1 (else maxLevel 3) */
    int32_t equalize;                 /* EQUALIZE: cv::createCLAHE(3.0, Size(8,8)) on every incoming frame (feature_tracker.cpp:269-275) */
    int32_t fisheye;                  /* FISHEYE: setMask starts from fisheye_mask (feature_tracker.cpp:175-178); the mask image is handed
                                         over with vrf_set_fisheye_mask before the first frame */
    int32_t lk_max_level;             /* -1 = reference default; else explicit maxLevel (0..3) */
    int32_t use_ransac;               /* 1 = rejectWithF enabled (reference behaviour) */
    int32_t reserved0;
    double  f_threshold;              /* F_THRESHOLD */
    double  focal_length;             /* FOCAL_LENGTH = 460 (parameters.h:11) */
    double  fx, fy, cx, cy;           /* projection_parameters */
    double  k1, k2, p1, p2;           /* distortion_parameters */
    /* back end */
    int32_t num_iterations;           /* NUM_ITERATIONS (max_num_iterations) */
    int32_t estimate_extrinsic;       /* ESTIMATE_EXTRINSIC (0: ex-pose constant) */
    int32_t estimate_td;              /* ESTIMATE_TD: 1 = ProjectionTdFactor instead of ProjectionFactor (estimator.cpp:1270-1285) */
    int32_t fix_depth;                /* FIX_DEPTH */
    double  depth_max_dist;           /* DEPTH_MAX_DIST (upper bound 2/DEPTH_MAX_DIST for estimate_flag==2) */
    double  g_norm;                   /* G = (0,0,g_norm) (parameters.cpp:13,158) */
    double  acc_n, acc_w, gyr_n, gyr_w; /* IMU noise (IntegrationBase ctor, integration_base.h:24-31) */
    double  depth_min_dist;           /* DEPTH_MIN_DIST: features with 0 < depth < this are dropped
                                         (FeatureManager::addFeatureCheckParallax, feature_manager.cpp:76-80) */
    double  tr;                       /* TR: rolling-shutter read-out time per frame (rolling_shutter_tr, parameters.cpp);
                                         0 for a global shutter.  Used by ProjectionTdFactor (TR / ROW * row) */
} VrfConfig;

/* Fills `cfg` with the synthetic-benchmark defaults of SURVEY.md section 8(d). */
void vrf_config_default(VrfConfig *cfg);

/* Replaces: Estimator::setParameter + FeatureTracker ctor + initGridsDetector
 * (estimator.cpp:15-41, feature_tracker.cpp:27-94) for `n_seq` independent sequences. */
int  vrf_create(const VrfConfig *cfg, int n_seq, int device, vrf_handle **out);
void vrf_destroy(vrf_handle *h);
const char *vrf_strerror(int code);
/* Last CUDA error string seen by this handle (diagnostics). */
const char *vrf_last_cuda_error(const vrf_handle *h);
/* Number of kernel launches issued by this handle so far (bench.py `gpu_launches`). */
uint64_t vrf_launch_count(const vrf_handle *h);
/* Replaces: Estimator::clearState() + setParameter() for one sequence
 * (estimator_nodelet.cpp:255-258; failureDetection reboot estimator.cpp:345-353). */
int  vrf_reset_sequence(vrf_handle *h, int seq);

/* FISHEYE: the reference loads fisheye_mask = cv::imread(FISHEYE_MASK, 0) in Estimator::setParameter (estimator.cpp:31)
 * and uses it as the initial mask of setMask() (feature_tracker.cpp:175-178): tracked points and new corners are kept
 * only where the mask is 255, FAST only detects where it is non-zero.  mask: ROW x COL, 8-bit, `stride` bytes per row
 * (0 = COL).  Requires VrfConfig::fisheye = 1; tracker calls fail with VRF_ERR_ARG until the mask is set. */
int  vrf_set_fisheye_mask(vrf_handle *h, const uint8_t *mask, size_t stride);

/* ------------------------------------------------------------------------- */
/* Front end                                                                  */
/* ------------------------------------------------------------------------- */

/*
 * Per-sequence result of one readImage call: the public members the nodelet
 * reads after readImage + the updateID loop (estimator_nodelet.cpp:324-343):
 * cur_pts, cur_un_pts, pts_velocity, ids, track_cnt; plus parity/debug members.
 * All array pointers are caller-allocated with room for `capacity` features
 * (use VRF_TRACK_CAP); optional ones may be NULL.
 */
typedef struct VrfTrackOut {
    int32_t  capacity;         /* in */
    int32_t  n;                /* out: cur_pts.size() */
    float   *cur_pts;          /* out: 2*n (u,v)          FeatureTracker::cur_pts */
    float   *cur_un_pts;       /* out: 2*n (x,y)          FeatureTracker::cur_un_pts */
    float   *pts_velocity;     /* out: 2*n (vx,vy)        FeatureTracker::pts_velocity */
    int32_t *ids;              /* out: n, after updateID  FeatureTracker::ids */
    int32_t *track_cnt;        /* out: n                  FeatureTracker::track_cnt */
    int32_t  n_id;             /* out: FeatureTracker::n_id after the updateID loop */
    int32_t  n_predict;        /* out: number of LK inputs this frame (predict_pts.size()) */
    float   *predict_pts;      /* out (optional): 2*n_predict  FeatureTracker::predict_pts */
    float   *lk_pts;           /* out (optional): 2*n_predict  raw calcOpticalFlowPyrLK output */
    uint8_t *lk_status;        /* out (optional): n_predict    raw LK status */
    int32_t *grids_track_num;  /* out (optional): rows*cols    FeatureTracker::grids_track_num */
    uint8_t *grids_texture_status; /* out (optional): rows*cols */
    int32_t  n_unstable;       /* out: unstable_pts.size() */
    int32_t  status;           /* out: per-sequence soft status */
    /* depth of every feature at ((int)v, (int)u) of the depth frame handed in with this call, in millimetres
     * (depth_img.at<unsigned short>, feature_manager.cpp:71-74; pt_depth_m = depth_mm / 1000.0), and whether
     * addFeatureCheckParallax keeps the feature (0 = erased because 0 < depth < DEPTH_MIN_DIST, :76-80).
     * Zero / one when the call carried no depth frame for this sequence. */
    uint16_t *depth_mm;        /* out (optional): n */
    uint8_t  *depth_keep;      /* out (optional): n */
} VrfTrackOut;

/*
 * Replaces FeatureTracker::readImage(const cv::Mat&, double, const Matrix3d&)
 * (feature_tracker.h:36-37, feature_tracker.cpp:263-439) followed by the nodelet's
 * updateID loop (estimator_nodelet.cpp:324-330).
 *   img            host pointer, `fmt` payload, `stride` bytes per row
 *   cur_time       _cur_time
 *   relative_R     row-major 3x3, cam(prev)->cam(cur) (Estimator::predictMotion,
 *                  estimator.cpp:1790-1860); NULL = identity
 *   pub_this_frame the reference's global PUB_THIS_FRAME (parameters.cpp:39),
 *                  made an explicit argument
 */
int vrf_tracker_read_image(vrf_handle *h, int seq, const uint8_t *img, size_t stride, int fmt,
                           double cur_time, const double *relative_R, int pub_this_frame,
                           VrfTrackOut *out);

/* Batched form: the same call for `n` distinct sequences in one set of kernel
 * launches.  imgs[i] are host pointers (pinned memory recommended). */
int vrf_tracker_read_image_batch(vrf_handle *h, int n, const int32_t *seqs,
                                 const uint8_t *const *imgs, size_t stride, int fmt,
                                 const double *cur_times, const double *relative_Rs /* n*9 or NULL */,
                                 const int32_t *pub_flags, VrfTrackOut *outs);

/* RGB-D form of the batched call: additionally takes the depth frame that the nodelet pairs with every
 * image (estimator_nodelet.cpp:206-225,380-383) and performs, on the device, the depth decode of
 * process() (:512-533) and the per-feature lookup + DEPTH_MIN_DIST test of
 * FeatureManager::addFeatureCheckParallax (feature_manager.cpp:71-80); results in
 * VrfTrackOut::depth_mm / depth_keep.  depths[i] may be NULL (no depth message).  Depth frames are only
 * consumed on publish frames (the reference only forwards a frame to the back end when PUB_THIS_FRAME,
 * estimator_nodelet.cpp:336-384), so only those cross PCIe.  `depth_stride` in bytes (0 = tightly packed). */
int vrf_tracker_read_rgbd_batch(vrf_handle *h, int n, const int32_t *seqs,
                                const uint8_t *const *imgs, size_t stride, int fmt,
                                const void *const *depths, size_t depth_stride, int depth_fmt,
                                const double *cur_times, const double *relative_Rs,
                                const int32_t *pub_flags, VrfTrackOut *outs);

/* Pipelined form of vrf_tracker_read_rgbd_batch for throughput over many sequences: submit() enqueues the H2D
 * copies (pinned host memory recommended), every front-end kernel and the D2H copy of the results, and returns
 * without waiting; collect() blocks until the OLDEST submitted batch has finished and fills `outs` (same n / seqs
 * as its submit).  Up to three batches may be in flight (submit k+2, then collect k): the next batches' frames cross
 * PCIe while the current batch's kernels run.  submit() returns VRF_ERR_CAPACITY when three batches are already
 * pending.  The host frame buffers must stay valid until the batch has been collected.  The optional
 * parity/debug members of VrfTrackOut (predict_pts, lk_*, grids_*) reflect the latest submitted batch.
 * (Reference analogue: the img_buf / feature_buf queues between the ROS callbacks, trackThread and
 * processThread, estimator_nodelet.cpp:125-146,192-225,380-384.) */
int vrf_tracker_submit_rgbd_batch(vrf_handle *h, int n, const int32_t *seqs,
                                  const uint8_t *const *imgs, size_t stride, int fmt,
                                  const void *const *depths, size_t depth_stride, int depth_fmt,
                                  const double *cur_times, const double *relative_Rs,
                                  const int32_t *pub_flags);
int vrf_tracker_collect_batch(vrf_handle *h, int n, const int32_t *seqs, VrfTrackOut *outs);

/* Device-resident form: `d_imgs` is a device pointer to n contiguous frames
 * (frame i at d_imgs + i*frame_bytes, rows tightly packed) already in HBM.
 * Enqueues the whole front end on the handle's stream and returns without
 * synchronising; results are fetched with vrf_tracker_fetch_batch().
 * `d_depth` (may be NULL): n contiguous depth frames of format `depth_fmt` (VRF_DEPTH_*), frame i paired
 * with image i; read on publish frames only (see vrf_tracker_read_rgbd_batch). */
int vrf_tracker_enqueue_batch_dev(vrf_handle *h, int n, const int32_t *seqs,
                                  const uint8_t *d_imgs, int fmt, const void *d_depth, int depth_fmt,
                                  const double *cur_times, const double *relative_Rs,
                                  const int32_t *pub_flags);
int vrf_tracker_fetch_batch(vrf_handle *h, int n, const int32_t *seqs, VrfTrackOut *outs);
/* Blocks until all work enqueued on the handle's stream has finished. */
int vrf_synchronize(vrf_handle *h);
/* The CUDA stream (cudaStream_t) the handle launches on, for event timing. */
void *vrf_stream(vrf_handle *h);

/* Optional per-kernel profiling: when enabled every kernel launch is bracketed by
 * CUDA events on the handle's stream (the reference's analogue: TicToc running
 * averages, utility/tic_toc.h + feature_tracker.cpp:333-337,424-428).
 * vrf_profile_read() synchronises, returns the number of kernel kinds and fills
 * name / accumulated device milliseconds / launch count per kind. */
int vrf_profile_enable(vrf_handle *h, int on);
int vrf_profile_read(vrf_handle *h, int max_kernels, const char **names, double *total_ms,
                     uint64_t *launch_counts, int reset);

/* Test hooks for the tiny order-dependent host/device-shared routines. */
/* libstdc++ std::sort restatement used by setMask (feature_tracker.cpp:186-188):
 * writes the permutation that sorts `cnt` descending with the reference's tie order. */
void vrf_debug_sort_desc(const int32_t *cnt, int32_t n, int32_t *perm_out);
/* Copies an internal per-sequence device array to host (parity tests only).
 * what: "pyr<L>" (level L of the current image pyramid, rows*cols u8, tightly packed),
 *       "cand" (cells*kmax*3 float: x,y,response), "ncand" (cells int32),
 *       "cell_k" (cells int32), "maskpts" (2*n int32), "ba_prof" (8 int64 SM clock counts per
 *       phase of the last k_ba_solve in batch slot `seq`) -- returns bytes written or <0. */
long vrf_debug_read(vrf_handle *h, const char *what, int seq, void *dst, size_t dst_bytes);
/* Test entry: FeatureTracker::rejectWithF (feature_tracker.cpp:441-473: cv::findFundamentalMat(FM_RANSAC, F_THRESHOLD,
 * 0.99) on the lifted points) alone, on n caller-supplied pixel pairs (x, y interleaved); status_out[n] receives the
 * inlier mask.  Clobbers the per-call working arrays of `seq`; not part of the tracking API. */
int vrf_debug_reject_with_f(vrf_handle *h, int seq, int n, const float *cur_xy, const float *forw_xy, uint8_t *status_out);

/* ------------------------------------------------------------------------- */
/* Back end (declared in vrf_ba.h, included here for convenience)             */
/* ------------------------------------------------------------------------- */
#include "vrf_ba.h"
/* Neighbouring steps: triangulateWithDepth, movingConsistencyCheck, IMU pre-integration (vrf_fm.h) */
#include "vrf_fm.h"

#ifdef __cplusplus
}
#endif
#endif /* VRF_H_ */
