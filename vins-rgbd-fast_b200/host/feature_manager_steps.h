// feature_manager_steps.h -- the steps either side of Estimator::optimization(), re-hosted on the vrf C ABI
// (include/vrf_fm.h).  Free function templates over the reference's own types so that the bodies of
//   FeatureManager::triangulateWithDepth     (vins_estimator/src/feature_manager/feature_manager.cpp:386-543)
//   Estimator::movingConsistencyCheck        (vins_estimator/src/estimator/estimator.cpp:1965-2009)
//   the IntegrationBase::push_back loop of Estimator::processIMU (estimator.cpp, integration_base.h:32-38)
// become one call each; member names and side effects are the reference's.
#pragma once
#include <set>
#include <vector>

#include "../../include/vrf.h"

namespace vrf_host {

// gathers f_manager.feature (list order, every landmark) into the flat arrays of a VrfFmProblem
template <class EstimatorT>
struct FmGather {
    VrfFmProblem pb{};
    std::vector<int32_t> start, ptr{0}, flag;
    std::vector<double> pts, dep, est;
    std::vector<uint8_t> dyn, rem;
    explicit FmGather(EstimatorT &e)
    {
        for (auto &it : e.f_manager.feature) {
            start.push_back(it.start_frame);
            for (auto &f : it.feature_per_frame) { pts.push_back(f.point.x()); pts.push_back(f.point.y()); dep.push_back(f.depth); }
            ptr.push_back((int32_t)dep.size());
            est.push_back(it.estimated_depth); flag.push_back(it.estimate_flag); dyn.push_back(it.is_dynamic ? 1 : 0);
        }
        rem.assign(start.size(), 0);
        for (int i = 0; i < VRF_NUM_FRAMES; ++i)
            for (int r = 0; r < 3; ++r) {
                pb.Ps[i][r] = e.Ps[i](r);
                for (int c = 0; c < 3; ++c) pb.Rs[i][r * 3 + c] = e.Rs[i](r, c);
            }
        for (int r = 0; r < 3; ++r) {
            pb.tic[r] = e.tic[0](r);
            for (int c = 0; c < 3; ++c) pb.ric[r * 3 + c] = e.ric[0](r, c);
        }
        pb.n_landmarks = (int32_t)start.size(); pb.n_obs = (int32_t)dep.size();
        pb.lm_start_frame = start.data(); pb.lm_obs_ptr = ptr.data(); pb.obs_pts = pts.data(); pb.obs_depth = dep.data();
        pb.estimated_depth = est.data(); pb.estimate_flag = flag.data(); pb.is_dynamic = dyn.data(); pb.remove = rem.data();
    }
};

// f_manager.triangulateWithDepth(Ps, tic, ric)
template <class EstimatorT>
int triangulateWithDepth(EstimatorT &e, vrf_handle *h)
{
    FmGather<EstimatorT> g(e);
    const int rc = vrf_fm_triangulate_with_depth_batch(h, 1, &g.pb);
    if (rc < 0) return rc;
    size_t l = 0;
    for (auto &it : e.f_manager.feature) {
        it.used_num = (int)it.feature_per_frame.size();
        it.estimated_depth = g.est[l]; it.estimate_flag = g.flag[l];
        ++l;
    }
    return rc;
}

// movingConsistencyCheck(removeIndex)
template <class EstimatorT>
int movingConsistencyCheck(EstimatorT &e, vrf_handle *h, std::set<int> &removeIndex)
{
    FmGather<EstimatorT> g(e);
    const int rc = vrf_fm_moving_consistency_check_batch(h, 1, &g.pb);
    if (rc < 0) return rc;
    size_t l = 0;
    for (auto &it : e.f_manager.feature) {
        it.used_num = (int)it.feature_per_frame.size();
        it.is_dynamic = g.dyn[l] != 0;
        if (g.rem[l]) removeIndex.insert(it.feature_id);
        ++l;
    }
    return rc;
}

// Rebuilds *pre (IntegrationBase) from its own dt_buf / acc_buf / gyr_buf on the device: the equivalent of
// IntegrationBase::repropagate (integration_base.h:40-54) and of the push_back loop in processIMU when the samples of a
// whole keyframe interval are handed over at once.
template <class IntegrationBaseT>
int preintegrate(IntegrationBaseT &pre, vrf_handle *h)
{
    const int n = (int)pre.dt_buf.size();
    std::vector<double> acc(3 * n), gyr(3 * n);
    for (int k = 0; k < n; ++k)
        for (int r = 0; r < 3; ++r) { acc[3 * k + r] = pre.acc_buf[k](r); gyr[3 * k + r] = pre.gyr_buf[k](r); }
    VrfImuSegment sg{};
    for (int r = 0; r < 3; ++r) {
        sg.acc_0[r] = pre.linearized_acc(r); sg.gyr_0[r] = pre.linearized_gyr(r);
        sg.linearized_ba[r] = pre.linearized_ba(r); sg.linearized_bg[r] = pre.linearized_bg(r);
    }
    sg.n_samples = n; sg.dt = pre.dt_buf.data(); sg.acc = acc.data(); sg.gyr = gyr.data();
    VrfImuPreint o;
    const int rc = vrf_imu_preintegrate_batch(h, 1, &sg, &o);
    if (rc < 0) return rc;
    pre.sum_dt = o.sum_dt;
    for (int r = 0; r < 3; ++r) { pre.delta_p(r) = o.delta_p[r]; pre.delta_v(r) = o.delta_v[r]; }
    pre.delta_q.x() = o.delta_q[0]; pre.delta_q.y() = o.delta_q[1]; pre.delta_q.z() = o.delta_q[2]; pre.delta_q.w() = o.delta_q[3];
    for (int r = 0; r < 15; ++r)
        for (int c = 0; c < 15; ++c) { pre.jacobian(r, c) = o.jacobian[r * 15 + c]; pre.covariance(r, c) = o.covariance[r * 15 + c]; }
    if (n > 0) { pre.acc_0 = pre.acc_buf[n - 1]; pre.gyr_0 = pre.gyr_buf[n - 1]; }
    return rc;
}

}  // namespace vrf_host
