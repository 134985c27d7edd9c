// ba_host.cu -- back-end entry points (placeholder until the BA kernels land).
#include "common.cuh"
#include "handle.h"
namespace vrf {
int ba_create(vrf_handle *) { return VRF_OK; }
void ba_destroy(vrf_handle *) {}
int ba_reset_sequence(vrf_handle *, int) { return VRF_OK; }
}
extern "C" int vrf_ba_solve(vrf_handle *, int, const VrfBaProblem *, VrfBaResult *) { return VRF_ERR_UNSUPPORTED; }
extern "C" int vrf_ba_solve_batch(vrf_handle *, int, const int32_t *, const VrfBaProblem *, VrfBaResult *) { return VRF_ERR_UNSUPPORTED; }
extern "C" int vrf_ba_upload_batch(vrf_handle *, int, const int32_t *, const VrfBaProblem *) { return VRF_ERR_UNSUPPORTED; }
extern "C" int vrf_ba_enqueue_batch(vrf_handle *, int, const int32_t *) { return VRF_ERR_UNSUPPORTED; }
extern "C" int vrf_ba_download_batch(vrf_handle *, int, const int32_t *, VrfBaResult *) { return VRF_ERR_UNSUPPORTED; }
extern "C" int vrf_debug_eval_projection(vrf_handle *, int, const double *, const double *, const double *, const double *, const double *, const double *, double *, double *, double *, double *, double *) { return VRF_ERR_UNSUPPORTED; }
extern "C" int vrf_debug_eval_imu(vrf_handle *, int, const VrfImuPreint *, const double *, const double *, const double *, const double *, double *, double *, double *, double *, double *) { return VRF_ERR_UNSUPPORTED; }
