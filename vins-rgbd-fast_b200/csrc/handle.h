// handle.h -- the opaque vrf_handle and internal cross-file entry points.
#pragma once
#include <vector>

#include "common.cuh"

#define VRF_CALL_SLOTS 4

namespace vrf {
struct BaState;   // ba_host.cu
}

struct vrf_handle {
    VrfConfig cfg;
    vrf::FrontCfg fc;
    int n_seq = 0, device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    vrf::FrontDev fd = {};
    // per-call descriptor ring (host pinned + device) so that enqueue calls can be pipelined
    vrf::SeqCall *h_calls_ring[VRF_CALL_SLOTS] = {};
    vrf::SeqCall *d_calls_ring[VRF_CALL_SLOTS] = {};
    cudaEvent_t call_ev[VRF_CALL_SLOTS] = {};
    unsigned call_ctr = 0;
    vrf::SeqCall *h_calls = nullptr, *d_calls = nullptr;   // slot in use by the current call
    uint8_t *d_stage = nullptr;
    size_t frame_bytes_max = 0;
    int *h_hdr = nullptr;
    void *h_out = nullptr;
    std::vector<int> cur_buf;
    std::vector<uint8_t> has_img;
    std::vector<double> prev_time;
    int last_n = 0;
    char errbuf[256];
    uint64_t launches = 0;
    vrf::BaState *ba = nullptr;
};

namespace vrf {
// frontend_kernels.cu
int front_configure_kernels(const FrontCfg &c);
int front_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, const uint8_t *d_frames,
                 size_t frame_bytes, int fmt, int any_pub, int sm_count, cudaStream_t st, uint64_t *launches);
int front_launch_tail(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, int any_pub,
                      cudaStream_t st, uint64_t *launches);
// ransac_kernels.cu
int ransac_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, cudaStream_t st,
                  uint64_t *launches);
// ba_host.cu
int ba_create(vrf_handle *h);
void ba_destroy(vrf_handle *h);
int ba_reset_sequence(vrf_handle *h, int seq);
}  // namespace vrf
