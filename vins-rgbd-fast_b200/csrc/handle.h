// handle.h -- the opaque vrf_handle and internal cross-file entry points.
#pragma once
#include <vector>

#include "common.cuh"
#include "tma.cuh"

#define VRF_CALL_SLOTS 32
#define VRF_COPY_CHUNKS 8
#define VRF_PIPE_DEPTH 3        // host-frame batches in flight (submit / collect): two ahead of the one being collected, so that the
                                // PCIe stream never waits for a collect() that is itself waiting for kernels held up by the BA stream
#define VRF_COPY_STREAMS 4      // H2D copies rotate over four streams (measured on B200, 444 frames/step: 31.8k -> 36.6k frames/s e2e vs two)

#define LK_WPB 8                // k_lk: warps (= features in flight) per CTA

namespace vrf {
// TMA tensor maps of the two pyramid buffers, one per level: u8 [n_seq][rows][cols] (strides pyr_bytes, pitch), box 48 x 32 x 1
struct LkMaps { CUtensorMap m[2][VRF_MAX_LEVELS]; };
struct BaState;   // ba_host.cu
struct FmState;   // fm_kernels.cu

// Kernel ids for launch counting and optional per-kernel CUDA-event profiling.
enum KernelId {
    K_INGEST = 0, K_CLAHE_LUT, K_CLAHE_APPLY, K_PYR, K_LK, K_POST_A, K_RANSAC, K_POST_B, K_FAST, K_FINISH,
    K_BA_SOLVE, K_BA_MARG, K_BA_PRIOR_FACTOR, K_FM_TRI, K_FM_CHECK, K_IMU_PREINT, K_COUNT
};
static const char *const kKernelNames[K_COUNT] = {
    "k_ingest", "k_clahe_lut", "k_clahe_apply", "k_pyr", "k_lk", "k_post_a", "k_ransac", "k_post_b", "k_fast", "k_finish",
    "k_ba_solve", "k_ba_marg", "k_ba_prior_factor", "k_fm_triangulate", "k_fm_check", "k_imu_preint"};

struct Prof {
    bool enabled = false;
    struct Rec { int id; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    double total_ms[K_COUNT] = {};
    uint64_t count[K_COUNT] = {};
};

// Passed to every launcher: stream + launch counter + optional event bracketing.
struct LaunchCtx {
    cudaStream_t st;
    uint64_t *launches;
    Prof *prof;
    cudaEvent_t pending_b = nullptr;
    void begin(int id)
    {
        ++*launches;
        if (prof && prof->enabled) {
            cudaEvent_t a, b;
            if (prof->pool.size() >= 2) {
                a = prof->pool.back(); prof->pool.pop_back();
                b = prof->pool.back(); prof->pool.pop_back();
            } else { cudaEventCreate(&a); cudaEventCreate(&b); }
            cudaEventRecord(a, st);
            prof->recs.push_back({id, a, b});
            pending_b = b;
        }
    }
    void end()
    {
        if (pending_b) { cudaEventRecord(pending_b, st); pending_b = nullptr; }
    }
};
}  // namespace vrf

struct vrf_handle {
    VrfConfig cfg;
    vrf::FrontCfg fc;
    int n_seq = 0, device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    // host-frame path: H2D copies run on their own stream, chunk by chunk, ahead of the kernels
    cudaStream_t copy_stream[VRF_COPY_STREAMS] = {};
    cudaEvent_t copy_ev[VRF_PIPE_DEPTH][VRF_COPY_CHUNKS][VRF_COPY_STREAMS] = {};
    vrf::FrontDev fd = {};
    vrf::LkMaps lk_maps;                    // built once in vrf_create (the pyramid buffers never move)
    // per-call descriptor ring (host pinned + device) so that enqueue calls can be pipelined
    vrf::SeqCall *h_calls_ring[VRF_CALL_SLOTS] = {};
    vrf::SeqCall *d_calls_ring[VRF_CALL_SLOTS] = {};
    cudaEvent_t call_ev[VRF_CALL_SLOTS] = {};
    unsigned call_ctr = 0;
    vrf::SeqCall *h_calls = nullptr, *d_calls = nullptr;   // slot in use by the current call
    // host-frame path (all lazily allocated): per pipeline slot a staging area for the frames and the publish
    // frames' depth planes, and pinned buffers that receive the fixed-width result copy of the batch
    uint8_t *d_stage[VRF_PIPE_DEPTH] = {};
    size_t frame_bytes_max = 0;
    uint8_t *d_stage_depth[VRF_PIPE_DEPTH] = {};
    size_t stage_depth_bytes[VRF_PIPE_DEPTH] = {};
    int *h_pipe_hdr[VRF_PIPE_DEPTH] = {};
    void *h_pipe_out[VRF_PIPE_DEPTH] = {};
    cudaEvent_t pipe_done[VRF_PIPE_DEPTH] = {};
    int pipe_n[VRF_PIPE_DEPTH] = {};
    bool pipe_busy[VRF_PIPE_DEPTH] = {};
    unsigned pipe_submit = 0, pipe_collect = 0;
    int out_w = 0;                          // entries per sequence in the fixed-width result copy
    int *h_hdr = nullptr;
    void *h_out = nullptr;
    std::vector<int> cur_buf;
    std::vector<uint8_t> has_img;
    std::vector<double> prev_time;
    int last_n = 0;
    char errbuf[256];
    uint64_t launches = 0;
    vrf::Prof prof;
    vrf::BaState *ba = nullptr;
    vrf::FmState *fm = nullptr;      // staging of the stateless feature-manager / pre-integration calls (lazily allocated)
};

namespace vrf {
// frontend_kernels.cu
int front_configure_kernels(const FrontCfg &c);
int front_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, const LkMaps &maps, const uint8_t *d_frames,
                 size_t frame_bytes, int fmt, int any_pub, int sm_count, LaunchCtx &lc);
// lk_kernels.cu
int lk_configure(const FrontCfg &c, const FrontDev &d, int n_seq, LkMaps *maps);
int lk_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, const LkMaps &maps, int sm_count, LaunchCtx &lc);
int front_launch_tail(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, int any_pub,
                      const uint8_t *d_depth, size_t depth_frame_bytes, int depth_fmt, LaunchCtx &lc);
// ransac_kernels.cu
int ransac_launch(const FrontCfg &c, const SeqCall *d_calls, int ncalls, const FrontDev &d, LaunchCtx &lc);
// ba_host.cu
int ba_create(vrf_handle *h);
void ba_destroy(vrf_handle *h);
int ba_reset_sequence(vrf_handle *h, int seq);
long ba_debug_prof(vrf_handle *h, int slot, void *dst, size_t bytes);
// fm_kernels.cu
void fm_destroy(vrf_handle *h);
}  // namespace vrf
