// estimator_optimization.h -- the body of Estimator::optimization() (reference
// vins_estimator/src/estimator/estimator.cpp:1161-1578) re-hosted on the vrf C ABI.
//
// It is written as a free function template over the reference's own `Estimator` type so
// that `void Estimator::optimization() { vrf_host::optimization(*this, handle, seq); }` is the
// only change in estimator.cpp.  It performs exactly the host-side steps of the reference:
//   vector2double()                      estimator.cpp:1164  (member of Estimator, reused)
//   gather landmarks/observations        :1243-1302 (same list order and filter as
//                                        FeatureManager::getDepthVector, feature_manager.cpp:302-324)
//   IMU factors = pre_integrations[1..]  :1226-1239
//   prior = the device-resident one      :1216-1223 (last_marginalization_info lives in the handle)
//   vrf_ba_solve                         replaces ceres::Solve + double2vector + marginalization
//   scatter results                      :985-1111 (Rs, Ps, Vs, Bas, Bgs, depths via setDepth)
#pragma once
#include <cmath>
#include <vector>

#include "../../include/vrf.h"

namespace vrf_host {

// `use_imu`, `estimate_extrinsic`, `estimate_td` are the reference's globals USE_IMU / ESTIMATE_EXTRINSIC /
// ESTIMATE_TD (utility/parameters.h); estimate_td must equal VrfConfig::estimate_td of the handle.
template <class EstimatorT>
int optimization(EstimatorT &e, vrf_handle *h, int seq, bool first_call_after_reset, int use_imu = 1,
                 int estimate_extrinsic = 0, int estimate_td = 0)
{
    e.vector2double();
    VrfBaProblem pb{};
    pb.frame_count = e.frame_count;
    pb.use_imu = use_imu;
    // estimator.cpp:1191-1201: the extrinsic becomes (and stays) variable once the window is full and the platform moves
    if ((estimate_extrinsic && e.frame_count == VRF_WINDOW_SIZE && e.Vs[0].norm() > 0.2) || e.openExEstimation)
        e.openExEstimation = true;
    pb.ex_constant = e.openExEstimation ? 0 : 1;
    // :1203-1212: td is fixed when it is not estimated or the platform is too slow
    pb.td_constant = (!estimate_td || e.Vs[0].norm() < 0.2) ? 1 : 0;
    pb.para_Td = e.para_Td[0][0];
    pb.marginalization_flag = (e.marginalization_flag == 0 /* MARGIN_OLD */) ? VRF_MARGIN_OLD : VRF_MARGIN_SECOND_NEW;
    for (int i = 0; i < VRF_NUM_FRAMES; ++i) {
        for (int k = 0; k < 7; ++k) pb.para_Pose[i][k] = e.para_Pose[i][k];
        for (int k = 0; k < 9; ++k) pb.para_SpeedBias[i][k] = e.para_SpeedBias[i][k];
    }
    for (int k = 0; k < 7; ++k) pb.para_Ex_Pose[k] = e.para_Ex_Pose[0][k];
    std::vector<double> lam, obs, vel, ctd, row;     // vel / ctd / row: ProjectionTdFactor inputs (estimator.cpp:1270-1285)
    std::vector<int32_t> start, flag, ptr(1, 0);
    int feature_index = -1;
    for (auto &it : e.f_manager.feature) {
        if (it.is_dynamic) continue;
        it.used_num = (int)it.feature_per_frame.size();
        if (!(it.used_num >= 2 && it.start_frame < VRF_WINDOW_SIZE - 2)) continue;
        ++feature_index;
        lam.push_back(e.para_Feature[feature_index][0]);
        start.push_back(it.start_frame);
        flag.push_back(it.estimate_flag);
        for (auto &f : it.feature_per_frame) {
            obs.push_back(f.point.x()); obs.push_back(f.point.y());
            if (estimate_td) { vel.push_back(f.velocity.x()); vel.push_back(f.velocity.y()); ctd.push_back(f.cur_td); row.push_back(f.uv.y()); }
        }
        ptr.push_back((int32_t)(obs.size() / 2));
    }
    pb.n_landmarks = (int32_t)lam.size();
    pb.n_obs = (int32_t)(obs.size() / 2);
    pb.para_Feature = lam.data(); pb.lm_start_frame = start.data(); pb.lm_estimate_flag = flag.data();
    pb.lm_obs_ptr = ptr.data(); pb.obs_pts = obs.data();
    if (estimate_td) { pb.obs_velocity = vel.data(); pb.obs_cur_td = ctd.data(); pb.obs_row = row.data(); }
    std::vector<VrfImuPreint> imu(VRF_WINDOW_SIZE);
    // without IMU the reference never allocates pre_integrations[] (processIMU is not called): nothing to gather
    for (int j = 1; use_imu && j <= e.frame_count; ++j) {
        const auto &p = *e.pre_integrations[j];
        VrfImuPreint &o = imu[j - 1];
        o.sum_dt = p.sum_dt;
        for (int k = 0; k < 3; ++k) { o.delta_p[k] = p.delta_p(k); o.delta_v[k] = p.delta_v(k); o.linearized_ba[k] = p.linearized_ba(k); o.linearized_bg[k] = p.linearized_bg(k); }
        o.delta_q[0] = p.delta_q.x(); o.delta_q[1] = p.delta_q.y(); o.delta_q[2] = p.delta_q.z(); o.delta_q[3] = p.delta_q.w();
        for (int r = 0; r < 15; ++r) for (int c = 0; c < 15; ++c) { o.jacobian[r * 15 + c] = p.jacobian(r, c); o.covariance[r * 15 + c] = p.covariance(r, c); }
    }
    pb.imu = use_imu ? imu.data() : nullptr;
    // Relocalisation residuals (estimator.cpp:1307-1346) are not part of the device problem: refuse instead of solving
    // a different problem (loop_closure is 0 in every shipped configuration of this path)
    if (e.relocalization_info) return VRF_ERR_UNSUPPORTED;
    pb.prior = first_call_after_reset ? nullptr : VRF_PRIOR_DEVICE;       // last_marginalization_info stays in HBM
    VrfBaResult res{};
    std::vector<double> lam_out(lam.size());
    res.para_Feature = lam_out.data();
    const int rc = vrf_ba_solve(h, seq, &pb, &res);
    if (rc < 0) return rc;
    // double2vector() results (estimator.cpp:985-1111) straight from the device
    for (int i = 0; i <= VRF_WINDOW_SIZE; ++i) {
        for (int r = 0; r < 3; ++r) {
            e.Ps[i](r) = res.Ps[i][r]; e.Vs[i](r) = res.Vs[i][r]; e.Bas[i](r) = res.Bas[i][r]; e.Bgs[i](r) = res.Bgs[i][r];
            for (int c = 0; c < 3; ++c) e.Rs[i](r, c) = res.Rs[i][r * 3 + c];
        }
    }
    for (size_t l = 0; l < lam_out.size(); ++l) e.para_Feature[l][0] = lam_out[l];
    // double2vector: tic/ric <- para_Ex_Pose, td <- para_Td (estimator.cpp:1033-1060)
    if (use_imu) {
        for (int k = 0; k < 7; ++k) e.para_Ex_Pose[0][k] = res.para_Ex_Pose[k];
        e.para_Td[0][0] = res.para_Td;
        for (int k = 0; k < 3; ++k) e.tic[0](k) = res.para_Ex_Pose[k];
        // ric = Quaterniond(w, x, y, z).normalized().toRotationMatrix()
        double q[4] = {res.para_Ex_Pose[3], res.para_Ex_Pose[4], res.para_Ex_Pose[5], res.para_Ex_Pose[6]};
        double nq = 0;
        for (double v : q) nq += v * v;
        nq = std::sqrt(nq);
        const double x = q[0] / nq, y = q[1] / nq, z = q[2] / nq, w = q[3] / nq;
        const double Rm[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                              2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                              2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) e.ric[0](r, c) = Rm[r * 3 + c];
        e.td = res.para_Td;
    }
    {   // f_manager.setDepth(dep) (feature_manager.cpp:197-223)
        auto dep = e.f_manager.getDepthVector();
        for (int i = 0; i < e.f_manager.getFeatureCount(); ++i) dep(i) = e.para_Feature[i][0];
        e.f_manager.setDepth(dep);
    }
    return rc;     // > 0: soft numerical status, handled by Estimator::failureDetection
}

}  // namespace vrf_host
