// feature_tracker.h -- drop-in replacement of the reference's FeatureTracker class surface
// (vins_estimator/src/feature_tracker/feature_tracker.h:31-97) on top of the vrf C ABI.
//
// The nodelet (vins_estimator/src/estimator_nodelet.cpp:313-343,404-441) only uses:
//   readImage(img, t, relative_R), updateID(i), readIntrinsicParameter(file),
//   initGridsDetector(), and the public members cur_pts, cur_un_pts, ids, track_cnt,
//   pts_velocity, predict_pts (+ fisheye_mask, grids_detector_img for visualisation).
// This shim keeps those names and fills the members from the device results after each call.
// It compiles against OpenCV/Eigen in the reference's catkin workspace; with
// -DVRF_SHIM_STANDALONE it compiles against the tiny stand-in types below so that this
// repository (no OpenCV C++ headers in the image) can at least syntax/ABI-check it.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vrf.h"

#ifdef VRF_SHIM_STANDALONE
namespace cv {
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct Mat { int rows = 0, cols = 0; size_t step = 0; const unsigned char *data = nullptr; int ch = 1; int channels() const { return ch; } bool empty() const { return !data; } };
}
namespace Eigen { struct Matrix3d { double m[9]; static Matrix3d Identity() { Matrix3d r{}; r.m[0] = r.m[4] = r.m[8] = 1; return r; } double operator()(int r, int c) const { return m[r * 3 + c]; } }; }
#else
#include <eigen3/Eigen/Dense>
#include <opencv2/opencv.hpp>
#endif

// The reference publishes on a frame iff the unsynchronised global PUB_THIS_FRAME is set (utility/parameters.h; written by
// the nodelet, estimator_nodelet.cpp:274-286, read inside readImage).  Inside the reference's tree the shim reads that very
// global; stand-alone it owns one.
#ifdef VRF_SHIM_STANDALONE
inline bool PUB_THIS_FRAME = false;
#else
extern bool PUB_THIS_FRAME;
#endif

class FeatureTracker
{
public:
    // Default-constructible like the reference's (`FeatureTracker featureTracker;` is a by-value member of Estimator,
    // estimator.h:117): the sequence slot inside a (possibly shared, batched) vrf_handle is bound later with attach(),
    // e.g. from Estimator::setParameter() where the reference loads the intrinsics (estimator.cpp:25-35).
    FeatureTracker() : n_id(0), h_(nullptr), seq_(-1), cfg_() {}
    // `cfg` replaces the extern globals of utility/parameters.h
    FeatureTracker(vrf_handle *handle, int seq, const VrfConfig &cfg) : n_id(0), h_(nullptr), seq_(-1), cfg_() { attach(handle, seq, cfg); }

    void attach(vrf_handle *handle, int seq, const VrfConfig &cfg)
    {
        h_ = handle; seq_ = seq; cfg_ = cfg;
        const size_t cap = VRF_TRACK_CAP;
        b_pts_.resize(2 * cap); b_un_.resize(2 * cap); b_vel_.resize(2 * cap); b_pred_.resize(2 * cap);
        b_ids_.resize(cap); b_cnt_.resize(cap); b_dmm_.resize(cap); b_dkeep_.resize(cap);
        grids_track_num.assign(cfg.num_grid_rows * cfg.num_grid_cols, 0);
    }

    // The reference's own signature (feature_tracker.h:36-37, call site estimator_nodelet.cpp:313): publish decision from the
    // global PUB_THIS_FRAME.
    void readImage(const cv::Mat &_img, double _cur_time, const Eigen::Matrix3d &_relative_R = Eigen::Matrix3d::Identity())
    {
        readImage(_img, _cur_time, _relative_R, PUB_THIS_FRAME);
    }

    // feature_tracker.cpp:263-439 (+ the nodelet's updateID loop, estimator_nodelet.cpp:324-330).
    // PUB_THIS_FRAME is an explicit argument instead of the reference's unsynchronised global.
    // `depth` (optional): the depth frame the nodelet pairs with this image (estimator_nodelet.cpp:206-225); its decode
    // (:512-533) and the per-feature lookup + DEPTH_MIN_DIST test of FeatureManager::addFeatureCheckParallax
    // (feature_manager.cpp:71-80) then run on the device: results in depth_mm / depth_keep.
    void readImage(const cv::Mat &_img, double _cur_time, const Eigen::Matrix3d &_relative_R, bool pub_this_frame,
                   const void *depth = nullptr, size_t depth_step = 0, int depth_fmt = VRF_DEPTH_NONE)
    {
        if (!h_) throw std::runtime_error("FeatureTracker::readImage before attach()");
        double R[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = _relative_R(r, c);
        VrfTrackOut o{};
        o.capacity = VRF_TRACK_CAP;
        o.cur_pts = b_pts_.data(); o.cur_un_pts = b_un_.data(); o.pts_velocity = b_vel_.data();
        o.ids = b_ids_.data(); o.track_cnt = b_cnt_.data(); o.predict_pts = b_pred_.data();
        o.grids_track_num = grids_track_num.data();
        o.depth_mm = b_dmm_.data(); o.depth_keep = b_dkeep_.data();
        const int fmt = _img.channels() == 3 ? VRF_FMT_RGB8 : VRF_FMT_GRAY8;
        const int32_t s = seq_, pub = pub_this_frame ? 1 : 0;
        const uint8_t *imgs[1] = {_img.data};
        const void *deps[1] = {depth};
        const int rc = vrf_tracker_read_rgbd_batch(h_, 1, &s, imgs, _img.step, fmt, depth ? deps : nullptr, depth_step, depth_fmt,
                                                   &_cur_time, R, &pub, &o);
        if (rc < 0) throw std::runtime_error(std::string("vrf_tracker_read_rgbd_batch: ") + vrf_strerror(rc));
        cur_time = _cur_time;
        cur_pts.resize(o.n); cur_un_pts.resize(o.n); pts_velocity.resize(o.n); ids.resize(o.n); track_cnt.resize(o.n);
        for (int i = 0; i < o.n; ++i) {
            cur_pts[i] = cv::Point2f(b_pts_[2 * i], b_pts_[2 * i + 1]);
            cur_un_pts[i] = cv::Point2f(b_un_[2 * i], b_un_[2 * i + 1]);
            pts_velocity[i] = cv::Point2f(b_vel_[2 * i], b_vel_[2 * i + 1]);
            ids[i] = b_ids_[i]; track_cnt[i] = b_cnt_[i];
        }
        depth_mm.assign(b_dmm_.begin(), b_dmm_.begin() + o.n);
        depth_keep.assign(b_dkeep_.begin(), b_dkeep_.begin() + o.n);
        predict_pts.resize(o.n_predict);
        for (int i = 0; i < o.n_predict; ++i) predict_pts[i] = cv::Point2f(b_pred_[2 * i], b_pred_[2 * i + 1]);
        n_id = o.n_id;
    }

    // ids are already assigned on the device in index order (feature_tracker.cpp:485-495):
    // the nodelet's `for (i = 0;; i++) if (!updateID(i)) break;` loop keeps working.
    bool updateID(unsigned int i) { return i < ids.size(); }

    void readIntrinsicParameter(const std::string &) {}   // intrinsics travel in VrfConfig
    void initGridsDetector() {}                            // grid table is built in vrf_create (feature_tracker.cpp:33-94)

    cv::Mat fisheye_mask, grids_detector_img;              // FISHEYE: hand fisheye_mask to the library once with setFisheyeMask()
    void setFisheyeMask() { if (!fisheye_mask.empty()) vrf_set_fisheye_mask(h_, fisheye_mask.data, fisheye_mask.step); }
    std::vector<cv::Point2f> cur_pts, predict_pts, cur_un_pts, pts_velocity;
    std::vector<int> ids, track_cnt, grids_track_num;
    std::vector<uint16_t> depth_mm;                        // depth_img.at<ushort>((int)v, (int)u) per feature (publish frames)
    std::vector<uint8_t> depth_keep;                       // 0: addFeatureCheckParallax erases it (0 < depth < DEPTH_MIN_DIST)
    double cur_time{};
    int n_id;

private:
    vrf_handle *h_;
    int seq_;
    VrfConfig cfg_;
    std::vector<float> b_pts_, b_un_, b_vel_, b_pred_;
    std::vector<int32_t> b_ids_, b_cnt_;
    std::vector<uint16_t> b_dmm_;
    std::vector<uint8_t> b_dkeep_;
};
