// Compile/link check of the host shims (vins-rgbd-fast_b200/host/*.h) against stand-in types that expose exactly the
// members of the reference's Estimator / FeatureManager / IntegrationBase the shims touch (Eigen and OpenCV are not in
// this image).  Built by tests/test_abi_c.py; the templates are instantiated, and the FeatureTracker shim is exercised
// up to the library call (which reports "no CUDA device" on a CPU-only machine).
#include <cmath>
#include <cstdio>
#include <list>
#include <vector>

#define VRF_SHIM_STANDALONE 1
#include "estimator_optimization.h"
#include "feature_manager_steps.h"
#include "feature_tracker.h"

struct V3 {
    double v[3] = {0, 0, 0};
    double &operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct M3 { double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; double &operator()(int r, int c) { return m[r * 3 + c]; } double operator()(int r, int c) const { return m[r * 3 + c]; } };
struct Quat { double q[4] = {0, 0, 0, 1}; double &x() { return q[0]; } double &y() { return q[1]; } double &z() { return q[2]; } double &w() { return q[3]; }
              double x() const { return q[0]; } double y() const { return q[1]; } double z() const { return q[2]; } double w() const { return q[3]; } };
struct M15 { double m[225] = {}; double &operator()(int r, int c) { return m[r * 15 + c]; } double operator()(int r, int c) const { return m[r * 15 + c]; } };
struct VecX { std::vector<double> d; double &operator()(int i) { return d[i]; } };
struct FeaturePerFrame { V3 point, velocity, uv; double cur_td = 0, depth = 0; };
struct FeaturePerId { int feature_id = 0, start_frame = 0, used_num = 0, estimate_flag = 0; bool is_dynamic = false; double estimated_depth = -1; std::vector<FeaturePerFrame> feature_per_frame; };
struct FeatureManager {
    std::list<FeaturePerId> feature;
    int getFeatureCount() { int n = 0; for (auto &it : feature) { it.used_num = (int)it.feature_per_frame.size(); if (it.used_num >= 2 && it.start_frame < 8 && !it.is_dynamic) ++n; } return n; }
    VecX getDepthVector() { VecX v; v.d.assign(getFeatureCount(), 0.0); return v; }
    void setDepth(const VecX &) {}
};
struct IntegrationBase {
    double sum_dt = 0; V3 delta_p, delta_v, linearized_ba, linearized_bg, linearized_acc, linearized_gyr, acc_0, gyr_0; Quat delta_q; M15 jacobian, covariance;
    std::vector<double> dt_buf; std::vector<V3> acc_buf, gyr_buf;
};
struct Estimator {
    int frame_count = 10, marginalization_flag = 0; bool openExEstimation = false, relocalization_info = false;
    double para_Pose[11][7] = {}, para_SpeedBias[11][9] = {}, para_Feature[1000][1] = {}, para_Ex_Pose[1][7] = {}, para_Td[1][1] = {};
    V3 Ps[11], Vs[11], Bas[11], Bgs[11], tic[1]; M3 Rs[11], ric[1]; double td = 0;
    FeatureManager f_manager; IntegrationBase *pre_integrations[11] = {};
    void vector2double() {}
};

int main()
{
    // force instantiation of every shim template against the stand-in types
    int (*f1)(Estimator &, vrf_handle *, int, bool, int, int, int) = &vrf_host::optimization<Estimator>;
    int (*f2)(Estimator &, vrf_handle *) = &vrf_host::triangulateWithDepth<Estimator>;
    int (*f3)(Estimator &, vrf_handle *, std::set<int> &) = &vrf_host::movingConsistencyCheck<Estimator>;
    int (*f4)(IntegrationBase &, vrf_handle *) = &vrf_host::preintegrate<IntegrationBase>;
    { FeatureTracker unattached; (void)unattached; }      // default-constructible
    std::printf("shims instantiated: %d\n", (f1 != nullptr) + (f2 != nullptr) + (f3 != nullptr) + (f4 != nullptr));
    VrfConfig cfg;
    vrf_config_default(&cfg);
    vrf_handle *h = nullptr;
    const int rc = vrf_create(&cfg, 1, 0, &h);
    std::printf("vrf_create rc %d\n", rc);
    if (rc == VRF_OK) {
        // the class surface of the reference: construct, feed one texture-less frame, read the public members
        // (default construction + attach(), as a by-value member of the reference's Estimator needs, estimator.h:117)
        FeatureTracker ft;
        ft.attach(h, 0, cfg);
        std::vector<unsigned char> img((size_t)cfg.row * cfg.col, 128);
        cv::Mat m; m.rows = cfg.row; m.cols = cfg.col; m.step = cfg.col; m.data = img.data();
        PUB_THIS_FRAME = true;
        ft.readImage(m, 0.0);                     // the reference's 3-argument signature
        std::printf("readImage: %zu features, updateID(0) %d\n", ft.cur_pts.size(), (int)ft.updateID(0));
        vrf_destroy(h);
    }
    return 0;
}
