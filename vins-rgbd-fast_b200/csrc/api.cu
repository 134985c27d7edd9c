// api.cu -- the extern "C" boundary (include/vrf.h): handle lifetime, per-sequence
// host bookkeeping, and the front-end entry points.  Host code stays thin C++;
// all arithmetic of the hot path runs in the kernels of frontend_kernels.cu /
// ransac_kernels.cu / ba_kernels.cu.  There is no CPU fallback.
#include <stdlib.h>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"
#include "handle.h"
#include "introsort.h"

using namespace vrf;

#define CK(call)                                                              \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) {                                             \
            snprintf(h->errbuf, sizeof(h->errbuf), "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return VRF_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

template <class T>
static cudaError_t dmalloc(T **p, size_t n)
{
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}

extern "C" void vrf_config_default(VrfConfig *c)
{
    memset(c, 0, sizeof(*c));
    c->row = 480; c->col = 640;
    c->max_cnt = 150; c->min_dist = 25;
    c->num_grid_rows = 7; c->num_grid_cols = 8;
    c->use_imu = 1; c->equalize = 0; c->fisheye = 0;
    c->lk_max_level = -1; c->use_ransac = 1;
    c->f_threshold = 1.0; c->focal_length = 460.0;
    c->fx = 600.0; c->fy = 600.0; c->cx = 320.0; c->cy = 240.0;
    c->k1 = 0.1; c->k2 = -0.2; c->p1 = 1e-3; c->p2 = 1e-3;
    c->num_iterations = 8; c->estimate_extrinsic = 0; c->estimate_td = 0; c->fix_depth = 0;
    c->depth_max_dist = 10.0; c->g_norm = 9.81;
    c->acc_n = 0.1; c->acc_w = 0.001; c->gyr_n = 0.01; c->gyr_w = 0.0001;
    c->depth_min_dist = 0.3;      // config/realsense/vio.yaml: depth_min_dist
    c->tr = 0.0;
}

extern "C" const char *vrf_strerror(int code)
{
    switch (code) {
    case VRF_OK: return "ok";
    case VRF_ERR_ARG: return "invalid argument";
    case VRF_ERR_CUDA: return "CUDA error (see vrf_last_cuda_error)";
    case VRF_ERR_UNSUPPORTED: return "configuration not supported by this build";
    case VRF_ERR_CAPACITY: return "per-sequence feature capacity exceeded";
    case VRF_ERR_NO_DEVICE: return "no CUDA device available (the library has no CPU fallback)";
    case VRF_SOFT_NONFINITE: return "non-finite state after solve";
    case VRF_SOFT_NOT_SPD: return "reduced system not SPD";
    default: return "unknown vrf status";
    }
}

extern "C" const char *vrf_last_cuda_error(const vrf_handle *h) { return h ? h->errbuf : ""; }
extern "C" uint64_t vrf_launch_count(const vrf_handle *h) { return h ? h->launches : 0; }
extern "C" void *vrf_stream(vrf_handle *h) { return h ? (void *)h->stream : nullptr; }

extern "C" void vrf_debug_sort_desc(const int32_t *cnt, int32_t n, int32_t *perm_out)
{
    std::vector<SortItem> v(n > 0 ? n : 0);
    for (int i = 0; i < n; ++i) { v[i].key = cnt[i]; v[i].val = i; }
    std_sort_desc(v.data(), n);
    for (int i = 0; i < n; ++i) perm_out[i] = v[i].val;
}

static int build_front_cfg(const VrfConfig &cfg, FrontCfg &fc)
{
    memset(&fc, 0, sizeof(fc));
    if (cfg.row < 64 || cfg.col < 64 || (cfg.col & 15)) return VRF_ERR_ARG;   // uint4 row access
    if (cfg.max_cnt <= 0 || cfg.max_cnt > VRF_CAP / 2 || cfg.min_dist < 1) return VRF_ERR_ARG;
    // FOCAL_LENGTH is a compile-time constant of the reference (parameters.h:11) that enters rejectWithF, ProjectionFactor's
    // sqrt_info (= 460 / 1.5, estimator.cpp:23) and movingConsistencyCheck alike: the BA kernels fix it too, so a handle
    // whose front end would use another value is refused rather than silently inconsistent.
    if (cfg.focal_length != 460.0) return VRF_ERR_ARG;
    fc.rows = cfg.row; fc.cols = cfg.col;
    int maxLevel = cfg.lk_max_level < 0 ? (cfg.use_imu ? 1 : 3) : cfg.lk_max_level;
    if (maxLevel > VRF_MAX_LEVELS - 1) return VRF_ERR_ARG;
    fc.levels = maxLevel + 1;
    fc.max_cnt = cfg.max_cnt; fc.min_dist = cfg.min_dist;
    fc.grows = cfg.num_grid_rows; fc.gcols = cfg.num_grid_cols;
    fc.ncells = fc.grows * fc.gcols;
    if (fc.grows < 1 || fc.gcols < 1 || fc.ncells > VRF_MAX_CELLS) return VRF_ERR_ARG;
    // initGridsDetector (feature_tracker.cpp:33-94); ROW, COL are doubles in the reference
    fc.gh = (int)((double)cfg.row / fc.grows);
    fc.gw = (int)((double)cfg.col / fc.gcols);
    fc.thr = (int)(cfg.max_cnt / fc.ncells);
    if (fc.thr <= 0) return VRF_ERR_ARG;       // ROS_ASSERT_MSG(grids_threshold > 0) :89-93
    if (fc.gh < 8 || fc.gw < 8) return VRF_ERR_ARG;
    fc.kmax = fc.thr + 2;
    fc.use_imu = cfg.use_imu; fc.use_ransac = cfg.use_ransac;
    unsigned off = 0;
    int w = cfg.col, hh = cfg.row;
    for (int l = 0; l < fc.levels; ++l) {
        fc.lw[l] = w; fc.lh[l] = hh; fc.lp[l] = (w + 15) & ~15;
        fc.loff[l] = off;
        off += (unsigned)fc.lp[l] * hh;
        off = (off + 255u) & ~255u;
        if (w <= 2 * VRF_LK_WIN + 2 || hh <= 2 * VRF_LK_WIN + 2) return VRF_ERR_ARG;  // single reflection only
        w = (w + 1) / 2; hh = (hh + 1) / 2;
    }
    fc.pyr_bytes = off;
    fc.fx = cfg.fx; fc.fy = cfg.fy; fc.cx = cfg.cx; fc.cy = cfg.cy;
    fc.k1 = cfg.k1; fc.k2 = cfg.k2; fc.p1 = cfg.p1; fc.p2 = cfg.p2;
    // PinholeCamera::setParameters (PinholeCamera.cc:292-295,306-310)
    fc.ik11 = 1.0 / cfg.fx; fc.ik13 = -cfg.cx / cfg.fx;
    fc.ik22 = 1.0 / cfg.fy; fc.ik23 = -cfg.cy / cfg.fy;
    fc.nodist = (cfg.k1 == 0.0 && cfg.k2 == 0.0 && cfg.p1 == 0.0 && cfg.p2 == 0.0);
    fc.focal = cfg.focal_length; fc.f_thr = cfg.f_threshold;
    fc.depth_min_dist = cfg.depth_min_dist;
    fc.equalize = cfg.equalize ? 1 : 0;
    return VRF_OK;
}

extern "C" int vrf_create(const VrfConfig *cfg, int n_seq, int device, vrf_handle **out)
{
    if (!cfg || !out || n_seq < 1 || n_seq > VRF_MAX_BATCH) return VRF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return VRF_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return VRF_ERR_ARG;
    FrontCfg fc;
    int rc = build_front_cfg(*cfg, fc);
    if (rc != VRF_OK) return rc;
    vrf_handle *h = new (std::nothrow) vrf_handle();
    if (!h) return VRF_ERR_ARG;
    h->cfg = *cfg; h->fc = fc; h->n_seq = n_seq; h->device = device;
    h->errbuf[0] = 0; h->launches = 0;
    auto fail = [&](int code) { vrf_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail(VRF_ERR_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(VRF_ERR_CUDA);
    h->sm_count = prop.multiProcessorCount;
    {   // the handle's main stream (kernels, small packed uploads, result copies) outranks the bulk frame copies of any
        // handle's copy streams: a BA batch's 30 MB upload must not queue behind 500 MB of frames on the copy engines
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) return fail(VRF_ERR_CUDA);
    }
    for (int k = 0; k < VRF_COPY_STREAMS; ++k)
        if (cudaStreamCreateWithFlags(&h->copy_stream[k], cudaStreamNonBlocking) != cudaSuccess) return fail(VRF_ERR_CUDA);
    for (int p = 0; p < VRF_PIPE_DEPTH; ++p) {
        for (int k = 0; k < VRF_COPY_CHUNKS; ++k)
            for (int q = 0; q < VRF_COPY_STREAMS; ++q)
                if (cudaEventCreateWithFlags(&h->copy_ev[p][k][q], cudaEventDisableTiming) != cudaSuccess) return fail(VRF_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->pipe_done[p], cudaEventDisableTiming) != cudaSuccess) return fail(VRF_ERR_CUDA);
    }
    const size_t S = n_seq, SC = S * VRF_CAP, SG = S * VRF_MAX_CELLS;
    FrontDev &d = h->fd;
    cudaError_t e = cudaSuccess;
#define A(ptr, n) if (e == cudaSuccess) e = dmalloc(&(ptr), (n))
    A(d.pyr[0], S * fc.pyr_bytes); A(d.pyr[1], S * fc.pyr_bytes);
    A(d.cur_pts, SC); A(d.prev_un, SC); A(d.ids, SC); A(d.cnt, SC); A(d.n_pts, S); A(d.n_id, S);
    A(d.pred_pts, SC); A(d.lk_pts, SC); A(d.lk_status, SC); A(d.n_lk, S);
    A(d.t_prev, SC); A(d.t_forw, SC); A(d.t_prevun, SC); A(d.t_ids, SC); A(d.t_cnt, SC); A(d.t_n, S); A(d.t_keep, SC);
    A(d.unstable, SC); A(d.n_unstable, S); A(d.maskpts, 2 * SC); A(d.n_maskpts, S);
    A(d.grid_cnt, SG); A(d.tex_status, SG); A(d.cell_k, SG);
    A(d.cand, SG * fc.kmax * 3); A(d.ncand, SG);
    {   // result arrays [batch][out_pitch]: the publish logic adds at most kmax features per selected cell
        int w = 2 * cfg->max_cnt + fc.kmax * fc.ncells;
        h->out_w = d.out_pitch = w > VRF_CAP ? VRF_CAP : w;
    }
    const size_t SW = S * d.out_pitch;
    A(d.o_pts, SW); A(d.o_un, SW); A(d.o_vel, SW); A(d.o_ids, SW); A(d.o_cnt, SW); A(d.out_hdr, S * 8);
    A(d.o_depth, SW); A(d.o_dkeep, SW);
    A(d.work_prefix, VRF_MAX_BATCH + 1);
    if (fc.equalize) { A(d.clahe_lut, S * 64 * 256); }
    for (int k = 0; k < VRF_CALL_SLOTS; ++k) { A(h->d_calls_ring[k], S); }
    h->frame_bytes_max = (size_t)cfg->row * cfg->col * 3;
#undef A
    if (e != cudaSuccess) { snprintf(h->errbuf, sizeof(h->errbuf), "alloc: %s", cudaGetErrorString(e)); return fail(VRF_ERR_CUDA); }
    if (cudaMemset(d.tex_status, 1, SG) != cudaSuccess) return fail(VRF_ERR_CUDA);   // grids_texture_status = true
    for (int k = 0; k < VRF_CALL_SLOTS; ++k) {
        if (cudaMallocHost((void **)&h->h_calls_ring[k], S * sizeof(SeqCall)) != cudaSuccess) return fail(VRF_ERR_CUDA);
        if (cudaEventCreateWithFlags(&h->call_ev[k], cudaEventDisableTiming) != cudaSuccess) return fail(VRF_ERR_CUDA);
    }
    if (cudaMallocHost((void **)&h->h_hdr, S * 8 * sizeof(int)) != cudaSuccess) return fail(VRF_ERR_CUDA);
    if (cudaMallocHost((void **)&h->h_out, SW * (3 * sizeof(float2) + 2 * sizeof(int) + sizeof(uint16_t) + sizeof(uint8_t))) != cudaSuccess) return fail(VRF_ERR_CUDA);
    h->cur_buf.assign(n_seq, 0);
    h->has_img.assign(n_seq, 0);
    h->prev_time.assign(n_seq, 0.0);
    if (front_configure_kernels(fc) != 0) { snprintf(h->errbuf, sizeof(h->errbuf), "kernel attribute setup failed"); return fail(VRF_ERR_CUDA); }
    if (int lrc = lk_configure(fc, d, n_seq, &h->lk_maps)) { snprintf(h->errbuf, sizeof(h->errbuf), "k_lk setup (attributes / cuTensorMapEncodeTiled) failed: %d", lrc); return fail(VRF_ERR_CUDA); }
    rc = ba_create(h);
    if (rc != VRF_OK) return fail(rc);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(VRF_ERR_CUDA);
    *out = h;
    return VRF_OK;
}

extern "C" void vrf_destroy(vrf_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (int k = 0; k < VRF_COPY_STREAMS; ++k) if (h->copy_stream[k]) cudaStreamSynchronize(h->copy_stream[k]);
    if (h->stream) cudaStreamSynchronize(h->stream);
    ba_destroy(h);
    fm_destroy(h);
    FrontDev &d = h->fd;
    void *ptrs[] = {d.pyr[0], d.pyr[1], d.cur_pts, d.prev_un, d.ids, d.cnt, d.n_pts, d.n_id, d.pred_pts, d.lk_pts,
                    d.lk_status, d.n_lk, d.t_prev, d.t_forw, d.t_prevun, d.t_ids, d.t_cnt, d.t_n, d.t_keep,
                    d.unstable, d.n_unstable, d.maskpts, d.n_maskpts, d.grid_cnt, d.tex_status, d.cell_k, d.cand,
                    d.ncand, d.o_pts, d.o_un, d.o_vel, d.o_ids, d.o_cnt, d.out_hdr, d.work_prefix,
                    d.o_depth, d.o_dkeep, d.clahe_lut, const_cast<uint8_t *>(d.fisheye)};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int p = 0; p < VRF_PIPE_DEPTH; ++p) {
        if (h->d_stage[p]) cudaFree(h->d_stage[p]);
        if (h->d_stage_depth[p]) cudaFree(h->d_stage_depth[p]);
    }
    for (int k = 0; k < VRF_CALL_SLOTS; ++k) {
        if (h->d_calls_ring[k]) cudaFree(h->d_calls_ring[k]);
        if (h->h_calls_ring[k]) cudaFreeHost(h->h_calls_ring[k]);
        if (h->call_ev[k]) cudaEventDestroy(h->call_ev[k]);
    }
    if (h->h_hdr) cudaFreeHost(h->h_hdr);
    if (h->h_out) cudaFreeHost(h->h_out);
    for (int p = 0; p < VRF_PIPE_DEPTH; ++p) {
        for (int k = 0; k < VRF_COPY_CHUNKS; ++k)
            for (int q = 0; q < VRF_COPY_STREAMS; ++q) if (h->copy_ev[p][k][q]) cudaEventDestroy(h->copy_ev[p][k][q]);
        if (h->pipe_done[p]) cudaEventDestroy(h->pipe_done[p]);
        if (h->h_pipe_hdr[p]) cudaFreeHost(h->h_pipe_hdr[p]);
        if (h->h_pipe_out[p]) cudaFreeHost(h->h_pipe_out[p]);
    }
    for (int k = 0; k < VRF_COPY_STREAMS; ++k) if (h->copy_stream[k]) cudaStreamDestroy(h->copy_stream[k]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int vrf_synchronize(vrf_handle *h)
{
    if (!h) return VRF_ERR_ARG;
    CK(cudaStreamSynchronize(h->stream));
    return VRF_OK;
}

extern "C" int vrf_set_fisheye_mask(vrf_handle *h, const uint8_t *mask, size_t stride)
{
    if (!h || !mask || !h->cfg.fisheye) return VRF_ERR_ARG;
    if (stride == 0) stride = (size_t)h->cfg.col;
    if (stride < (size_t)h->cfg.col) return VRF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    FrontDev &d = h->fd;
    if (!d.fisheye) {
        uint8_t *p = nullptr;
        CK(cudaMalloc((void **)&p, (size_t)h->cfg.row * h->cfg.col));
        d.fisheye = p;
    }
    CK(cudaMemcpy2D(const_cast<uint8_t *>(d.fisheye), h->cfg.col, mask, stride, h->cfg.col, h->cfg.row, cudaMemcpyHostToDevice));
    return VRF_OK;
}

extern "C" int vrf_reset_sequence(vrf_handle *h, int seq)
{
    if (!h || seq < 0 || seq >= h->n_seq) return VRF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    FrontDev &d = h->fd;
    CK(cudaMemset(d.n_pts + seq, 0, sizeof(int)));
    CK(cudaMemset(d.n_id + seq, 0, sizeof(int)));
    CK(cudaMemset(d.grid_cnt + (size_t)seq * VRF_MAX_CELLS, 0, VRF_MAX_CELLS * sizeof(int)));
    CK(cudaMemset(d.tex_status + (size_t)seq * VRF_MAX_CELLS, 1, VRF_MAX_CELLS));
    h->cur_buf[seq] = 0; h->has_img[seq] = 0; h->prev_time[seq] = 0.0;
    return ba_reset_sequence(h, seq);
}

// Fill the per-call descriptors, upload them and enqueue every front-end kernel.
// `out_base`: batch position of item 0 (a host-frame batch is enqueued in chunks; outputs are indexed by
// batch position so that one fetch collects the whole batch).
// Depth: `d_depth` = batch of depth planes (`depth_frame_bytes` each, format `depth_fmt`); `dslots` (may be NULL =
// item i uses plane i) gives the plane index of every item, -1 = none.
static int enqueue_front(vrf_handle *h, int n, const int32_t *seqs, const uint8_t *d_frames, size_t frame_bytes,
                         int fmt, const double *times, const double *Rs, const int32_t *pubs, int out_base = 0,
                         const uint8_t *d_depth = nullptr, size_t depth_frame_bytes = 0, int depth_fmt = VRF_DEPTH_NONE,
                         const int *dslots = nullptr)
{
    if (n < 1 || out_base < 0 || out_base + n > h->n_seq || !seqs || !times) return VRF_ERR_ARG;
    if (fmt != VRF_FMT_GRAY8 && fmt != VRF_FMT_RGB8) return VRF_ERR_ARG;
    if (h->cfg.fisheye && !h->fd.fisheye) return VRF_ERR_ARG;        // FISHEYE without vrf_set_fisheye_mask: no silent fallback
    std::vector<uint8_t> seen(h->n_seq, 0);
    int any_pub = 0;
    const unsigned slot = h->call_ctr++ % VRF_CALL_SLOTS;
    CK(cudaEventSynchronize(h->call_ev[slot]));     // slot free again?
    h->h_calls = h->h_calls_ring[slot];
    h->d_calls = h->d_calls_ring[slot];
    for (int i = 0; i < n; ++i) {
        int s = seqs[i];
        if (s < 0 || s >= h->n_seq || seen[s]) return VRF_ERR_ARG;
        seen[s] = 1;
        SeqCall &c = h->h_calls[i];
        c.seq = s;
        c.pub = pubs ? (pubs[i] != 0) : 1;
        any_pub |= c.pub;
        c.first = h->has_img[s] ? 0 : 1;
        // cur_img lives in buffer cur_buf; forw_img goes to the other one
        // (first frame: cur_img = forw_img = img, feature_tracker.cpp:279-287)
        c.buf_prev = h->cur_buf[s];
        c.buf_cur = c.first ? h->cur_buf[s] : 1 - h->cur_buf[s];
        c.dslot = (d_depth && depth_fmt != VRF_DEPTH_NONE) ? (dslots ? dslots[i] : i) : -1;
        c.dt = times[i] - h->prev_time[s];
        for (int k = 0; k < 9; ++k) c.R[k] = Rs ? Rs[(size_t)i * 9 + k] : ((k % 4 == 0) ? 1.0 : 0.0);
    }
    CK(cudaMemcpyAsync(h->d_calls, h->h_calls, n * sizeof(SeqCall), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->call_ev[slot], h->stream));
    LaunchCtx lc{h->stream, &h->launches, &h->prof};
    FrontDev fd = h->fd;
    const size_t ob = (size_t)out_base * fd.out_pitch;
    fd.o_pts += ob; fd.o_un += ob; fd.o_vel += ob; fd.o_ids += ob; fd.o_cnt += ob; fd.o_depth += ob; fd.o_dkeep += ob;
    fd.out_hdr += (size_t)out_base * 8;
    front_launch(h->fc, h->d_calls, n, fd, h->lk_maps, d_frames, frame_bytes, fmt, any_pub, h->sm_count, lc);
    if (h->fc.use_ransac && any_pub) ransac_launch(h->fc, h->d_calls, n, fd, lc);
    front_launch_tail(h->fc, h->d_calls, n, fd, any_pub, d_depth, depth_frame_bytes, depth_fmt, lc);
    CK(cudaGetLastError());
    for (int i = 0; i < n; ++i) {
        int s = seqs[i];
        h->cur_buf[s] = h->h_calls[i].buf_cur;     // cur_img = forw_img
        h->has_img[s] = 1;
        h->prev_time[s] = times[i];
    }
    h->last_n = out_base + n;
    return VRF_OK;
}

// Result copy of a batch: headers + the [n][out_pitch] prefix of every output array, one contiguous transfer each,
// stream-ordered behind the kernels.  Host layout: the seven arrays back to back.
static int enqueue_result_copy(vrf_handle *h, int n, int *hdr_dst, void *out_dst)
{
    FrontDev &d = h->fd;
    const size_t nw = (size_t)n * h->out_w;
    char *hp = (char *)out_dst;
    CK(cudaMemcpyAsync(hdr_dst, d.out_hdr, (size_t)n * 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
#define OUT1D(src, T) CK(cudaMemcpyAsync(hp, src, nw * sizeof(T), cudaMemcpyDeviceToHost, h->stream)); hp += nw * sizeof(T)
    OUT1D(d.o_pts, float2); OUT1D(d.o_un, float2); OUT1D(d.o_vel, float2); OUT1D(d.o_ids, int); OUT1D(d.o_cnt, int);
    OUT1D(d.o_depth, uint16_t); OUT1D(d.o_dkeep, uint8_t);
#undef OUT1D
    return VRF_OK;
}

static int unpack_results(vrf_handle *h, int n, const int32_t *seqs, VrfTrackOut *outs, const int *hdrs, const void *out_src)
{
    FrontDev &d = h->fd;
    const int w = h->out_w;
    const size_t nw = (size_t)n * w;
    const char *hp = (const char *)out_src;
    const float2 *h_pts = (const float2 *)hp; hp += nw * sizeof(float2);
    const float2 *h_un = (const float2 *)hp; hp += nw * sizeof(float2);
    const float2 *h_vel = (const float2 *)hp; hp += nw * sizeof(float2);
    const int *h_ids = (const int *)hp; hp += nw * sizeof(int);
    const int *h_cnt = (const int *)hp; hp += nw * sizeof(int);
    const uint16_t *h_dep = (const uint16_t *)hp; hp += nw * sizeof(uint16_t);
    const uint8_t *h_dkeep = (const uint8_t *)hp;
    int worst = VRF_OK;
    for (int i = 0; i < n; ++i) {
        VrfTrackOut &o = outs[i];
        const int *hdr = hdrs + i * 8;
        const int s = seqs[i];
        o.n = hdr[0]; o.n_id = hdr[1]; o.n_predict = hdr[2]; o.n_unstable = hdr[3]; o.status = hdr[4];
        if (o.status != 0) worst = o.status;
        int m = o.n < o.capacity ? o.n : o.capacity;
        if (o.n > o.capacity) { o.status = VRF_ERR_CAPACITY; worst = VRF_ERR_CAPACITY; }
        if (m > w) m = w;
        if (o.cur_pts) memcpy(o.cur_pts, h_pts + (size_t)i * w, m * sizeof(float2));
        if (o.cur_un_pts) memcpy(o.cur_un_pts, h_un + (size_t)i * w, m * sizeof(float2));
        if (o.pts_velocity) memcpy(o.pts_velocity, h_vel + (size_t)i * w, m * sizeof(float2));
        if (o.ids) memcpy(o.ids, h_ids + (size_t)i * w, m * sizeof(int));
        if (o.track_cnt) memcpy(o.track_cnt, h_cnt + (size_t)i * w, m * sizeof(int));
        if (o.depth_mm) memcpy(o.depth_mm, h_dep + (size_t)i * w, m * sizeof(uint16_t));
        if (o.depth_keep) memcpy(o.depth_keep, h_dkeep + (size_t)i * w, m);
        // optional parity/debug members: read from the per-sequence state, i.e. only meaningful while no later
        // batch has been submitted (per-sequence copies; not on the fast path)
        int np = o.n_predict < o.capacity ? o.n_predict : o.capacity;
        if (o.predict_pts && np > 0) CK(cudaMemcpy(o.predict_pts, d.pred_pts + (size_t)s * VRF_CAP, np * sizeof(float2), cudaMemcpyDeviceToHost));
        if (o.lk_pts && np > 0) CK(cudaMemcpy(o.lk_pts, d.lk_pts + (size_t)s * VRF_CAP, np * sizeof(float2), cudaMemcpyDeviceToHost));
        if (o.lk_status && np > 0) CK(cudaMemcpy(o.lk_status, d.lk_status + (size_t)s * VRF_CAP, np, cudaMemcpyDeviceToHost));
        if (o.grids_track_num) CK(cudaMemcpy(o.grids_track_num, d.grid_cnt + (size_t)s * VRF_MAX_CELLS, h->fc.ncells * sizeof(int), cudaMemcpyDeviceToHost));
        if (o.grids_texture_status) CK(cudaMemcpy(o.grids_texture_status, d.tex_status + (size_t)s * VRF_MAX_CELLS, h->fc.ncells, cudaMemcpyDeviceToHost));
    }
    return worst;
}

static int fetch_front(vrf_handle *h, int n, const int32_t *seqs, VrfTrackOut *outs)
{
    if (!outs || n < 1 || n > h->n_seq) return VRF_ERR_ARG;
    int rc = enqueue_result_copy(h, n, h->h_hdr, h->h_out);
    if (rc != VRF_OK) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return unpack_results(h, n, seqs, outs, h->h_hdr, h->h_out);
}

// Bytes per result entry of the fixed-width copy: cur_pts, cur_un_pts, pts_velocity (float2), ids, track_cnt (int),
// depth_mm (u16), depth_keep (u8).
static const size_t kOutEntryBytes = 3 * sizeof(float2) + 2 * sizeof(int) + sizeof(uint16_t) + sizeof(uint8_t);

extern "C" int vrf_tracker_submit_rgbd_batch(vrf_handle *h, int n, const int32_t *seqs, const uint8_t *const *imgs,
                                             size_t stride, int fmt, const void *const *depths, size_t depth_stride,
                                             int depth_fmt, const double *cur_times, const double *relative_Rs,
                                             const int32_t *pub_flags)
{
    if (!h || !imgs || !seqs || !cur_times) return VRF_ERR_ARG;
    if (n < 1 || n > h->n_seq) return VRF_ERR_ARG;
    if (fmt != VRF_FMT_GRAY8 && fmt != VRF_FMT_RGB8) return VRF_ERR_ARG;
    if (depth_fmt != VRF_DEPTH_NONE && depth_fmt != VRF_DEPTH_16UC1 && depth_fmt != VRF_DEPTH_32FC1) return VRF_ERR_ARG;
    if (!depths) depth_fmt = VRF_DEPTH_NONE;
    const int slot = (int)(h->pipe_submit % VRF_PIPE_DEPTH);
    if (h->pipe_busy[slot]) return VRF_ERR_CAPACITY;          // VRF_PIPE_DEPTH batches already in flight: collect first
    CK(cudaSetDevice(h->device));
    const int bpp = (fmt == VRF_FMT_RGB8) ? 3 : 1;
    const size_t row_bytes = (size_t)h->cfg.col * bpp;
    const size_t frame_bytes = row_bytes * h->cfg.row;
    if (stride == 0) stride = row_bytes;
    if (stride < row_bytes) return VRF_ERR_ARG;
    {   // duplicates across the whole batch (each chunk re-checks its own part)
        std::vector<uint8_t> seen(h->n_seq, 0);
        for (int i = 0; i < n; ++i) {
            if (!imgs[i] || seqs[i] < 0 || seqs[i] >= h->n_seq || seen[seqs[i]]) return VRF_ERR_ARG;
            seen[seqs[i]] = 1;
        }
    }
    // ---- lazily allocated staging / result buffers of this pipeline slot ----
    if (!h->d_stage[slot]) CK(cudaMalloc((void **)&h->d_stage[slot], (size_t)h->n_seq * h->frame_bytes_max));
    if (!h->h_pipe_hdr[slot]) CK(cudaMallocHost((void **)&h->h_pipe_hdr[slot], (size_t)h->n_seq * 8 * sizeof(int)));
    if (!h->h_pipe_out[slot]) CK(cudaMallocHost(&h->h_pipe_out[slot], (size_t)h->n_seq * h->out_w * kOutEntryBytes));
    const size_t drow = (size_t)h->cfg.col * (depth_fmt == VRF_DEPTH_32FC1 ? 4 : 2);
    const size_t dframe = drow * h->cfg.row;
    if (depth_fmt != VRF_DEPTH_NONE) {
        if (depth_stride == 0) depth_stride = drow;
        if (depth_stride < drow) return VRF_ERR_ARG;
        if (h->stage_depth_bytes[slot] < (size_t)h->n_seq * dframe) {       // sized for the widest format seen
            CK(cudaStreamSynchronize(h->stream));
            if (h->d_stage_depth[slot]) CK(cudaFree(h->d_stage_depth[slot]));
            h->d_stage_depth[slot] = nullptr; h->stage_depth_bytes[slot] = 0;
            CK(cudaMalloc((void **)&h->d_stage_depth[slot], (size_t)h->n_seq * dframe));
            h->stage_depth_bytes[slot] = (size_t)h->n_seq * dframe;
        }
    }
    uint8_t *stage = h->d_stage[slot], *stage_d = h->d_stage_depth[slot];
    std::vector<int> dslots(n, -1);
    // Overlap of PCIe and kernels.  With another batch in flight (submit k+1 before collect k) the next batch's
    // frames already cross PCIe while this batch's kernels run, and the batch is enqueued in one piece (measured on
    // B200: cutting it up only multiplies the latency-bound small kernels, 3.5 -> 6.1 ms per 444 frames).  A lone
    // batch is cut into a few chunks so that chunk c+1's frames are copied while chunk c's kernels run.
    bool other_in_flight = false;
    for (int q = 0; q < VRF_PIPE_DEPTH; ++q) other_in_flight |= (q != slot && h->pipe_busy[q]);
    int nchunk = other_in_flight ? 1 : n / 96;
    nchunk = nchunk < 1 ? 1 : (nchunk > 4 ? 4 : nchunk);
    if (getenv("VRF_DEBUG_CHUNKS")) nchunk = atoi(getenv("VRF_DEBUG_CHUNKS"));
    auto bail = [&](int rc) {
        for (int q = 0; q < VRF_COPY_STREAMS; ++q) cudaStreamSynchronize(h->copy_stream[q]);
        cudaStreamSynchronize(h->stream);
        return rc;
    };
    static const int dbg_skip = getenv("VRF_DEBUG_SKIP") ? atoi(getenv("VRF_DEBUG_SKIP")) : 0;   // 1: no H2D, 2: no kernels
    for (int c = 0; c < nchunk; ++c) {
        const int i0 = (int)((long long)n * c / nchunk), i1 = (int)((long long)n * (c + 1) / nchunk);
        for (int i = i0; i < i1; ++i) {
            if (dbg_skip == 1) break;
            static const int dbg_ncs = getenv("VRF_DEBUG_COPY_STREAMS") ? atoi(getenv("VRF_DEBUG_COPY_STREAMS")) : VRF_COPY_STREAMS;
            cudaStream_t cs = h->copy_stream[i % dbg_ncs];
            if (stride == row_bytes)
                CK(cudaMemcpyAsync(stage + (size_t)i * frame_bytes, imgs[i], frame_bytes, cudaMemcpyHostToDevice, cs));
            else
                CK(cudaMemcpy2DAsync(stage + (size_t)i * frame_bytes, row_bytes, imgs[i], stride, row_bytes, h->cfg.row, cudaMemcpyHostToDevice, cs));
            // the depth plane is only consumed on publish frames: only those cross PCIe
            if (depth_fmt != VRF_DEPTH_NONE && depths[i] && (!pub_flags || pub_flags[i])) {
                dslots[i] = i;
                if (depth_stride == drow)
                    CK(cudaMemcpyAsync(stage_d + (size_t)i * dframe, depths[i], dframe, cudaMemcpyHostToDevice, cs));
                else
                    CK(cudaMemcpy2DAsync(stage_d + (size_t)i * dframe, drow, depths[i], depth_stride, drow, h->cfg.row, cudaMemcpyHostToDevice, cs));
            }
        }
        for (int q = 0; q < VRF_COPY_STREAMS; ++q) {
            CK(cudaEventRecord(h->copy_ev[slot][c][q], h->copy_stream[q]));
            CK(cudaStreamWaitEvent(h->stream, h->copy_ev[slot][c][q], 0));
        }
        if (dbg_skip == 2) continue;
        int rc = enqueue_front(h, i1 - i0, seqs + i0, stage + (size_t)i0 * frame_bytes, frame_bytes, fmt, cur_times + i0,
                               relative_Rs ? relative_Rs + (size_t)i0 * 9 : nullptr, pub_flags ? pub_flags + i0 : nullptr, i0,
                               depth_fmt != VRF_DEPTH_NONE ? stage_d : nullptr, dframe, depth_fmt, dslots.data() + i0);
        if (rc != VRF_OK) return bail(rc);
    }
    {   // ---- results, stream-ordered behind the kernels ----
        int rc = enqueue_result_copy(h, n, h->h_pipe_hdr[slot], h->h_pipe_out[slot]);
        if (rc != VRF_OK) return bail(rc);
    }
    CK(cudaEventRecord(h->pipe_done[slot], h->stream));
    h->pipe_busy[slot] = true; h->pipe_n[slot] = n;
    h->pipe_submit++;
    return VRF_OK;
}

extern "C" int vrf_tracker_collect_batch(vrf_handle *h, int n, const int32_t *seqs, VrfTrackOut *outs)
{
    if (!h || !outs || !seqs) return VRF_ERR_ARG;
    const int slot = (int)(h->pipe_collect % VRF_PIPE_DEPTH);
    if (!h->pipe_busy[slot] || h->pipe_n[slot] != n) return VRF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->pipe_done[slot]));
    h->pipe_busy[slot] = false;
    h->pipe_collect++;
    return unpack_results(h, n, seqs, outs, h->h_pipe_hdr[slot], h->h_pipe_out[slot]);
}

extern "C" int vrf_tracker_read_rgbd_batch(vrf_handle *h, int n, const int32_t *seqs, const uint8_t *const *imgs,
                                           size_t stride, int fmt, const void *const *depths, size_t depth_stride,
                                           int depth_fmt, const double *cur_times, const double *relative_Rs,
                                           const int32_t *pub_flags, VrfTrackOut *outs)
{
    if (!h || !outs) return VRF_ERR_ARG;
    if (h->pipe_submit != h->pipe_collect) return VRF_ERR_ARG;      // pipelined batches pending: collect them first
    int rc = vrf_tracker_submit_rgbd_batch(h, n, seqs, imgs, stride, fmt, depths, depth_stride, depth_fmt, cur_times,
                                           relative_Rs, pub_flags);
    if (rc != VRF_OK) return rc;
    return vrf_tracker_collect_batch(h, n, seqs, outs);
}

extern "C" int vrf_tracker_read_image_batch(vrf_handle *h, int n, const int32_t *seqs, const uint8_t *const *imgs,
                                            size_t stride, int fmt, const double *cur_times,
                                            const double *relative_Rs, const int32_t *pub_flags, VrfTrackOut *outs)
{
    return vrf_tracker_read_rgbd_batch(h, n, seqs, imgs, stride, fmt, nullptr, 0, VRF_DEPTH_NONE, cur_times, relative_Rs,
                                       pub_flags, outs);
}

extern "C" int vrf_tracker_read_image(vrf_handle *h, int seq, const uint8_t *img, size_t stride, int fmt,
                                      double cur_time, const double *relative_R, int pub_this_frame, VrfTrackOut *out)
{
    int32_t s = seq, p = pub_this_frame;
    const uint8_t *imgs[1] = {img};
    return vrf_tracker_read_image_batch(h, 1, &s, imgs, stride, fmt, &cur_time, relative_R, &p, out);
}

extern "C" int vrf_tracker_enqueue_batch_dev(vrf_handle *h, int n, const int32_t *seqs, const uint8_t *d_imgs, int fmt,
                                             const void *d_depth, int depth_fmt, const double *cur_times,
                                             const double *relative_Rs, const int32_t *pub_flags)
{
    if (!h || !d_imgs) return VRF_ERR_ARG;
    if (!d_depth) depth_fmt = VRF_DEPTH_NONE;
    if (depth_fmt != VRF_DEPTH_NONE && depth_fmt != VRF_DEPTH_16UC1 && depth_fmt != VRF_DEPTH_32FC1) return VRF_ERR_ARG;
    if (depth_fmt != VRF_DEPTH_NONE && ((uintptr_t)d_depth & 3) != 0) return VRF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    const int bpp = (fmt == VRF_FMT_RGB8) ? 3 : 1;
    const size_t frame_bytes = (size_t)h->cfg.col * bpp * h->cfg.row;
    if (((uintptr_t)d_imgs & 15) != 0) return VRF_ERR_ARG;
    const size_t dframe = (size_t)h->cfg.col * h->cfg.row * (depth_fmt == VRF_DEPTH_32FC1 ? 4 : 2);
    return enqueue_front(h, n, seqs, d_imgs, frame_bytes, fmt, cur_times, relative_Rs, pub_flags, 0,
                         (const uint8_t *)d_depth, dframe, depth_fmt, nullptr);
}

extern "C" int vrf_tracker_fetch_batch(vrf_handle *h, int n, const int32_t *seqs, VrfTrackOut *outs)
{
    if (!h || !seqs) return VRF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (n != h->last_n) return VRF_ERR_ARG;
    return fetch_front(h, n, seqs, outs);
}

extern "C" long vrf_debug_read(vrf_handle *h, const char *what, int seq, void *dst, size_t dst_bytes)
{
    if (!h || !what || !dst || seq < 0 || seq >= h->n_seq) return VRF_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return VRF_ERR_CUDA;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return VRF_ERR_CUDA;
    const FrontCfg &c = h->fc;
    FrontDev &d = h->fd;
    std::string w(what);
    if (w.rfind("pyr", 0) == 0 && w.size() == 4) {
        int l = w[3] - '0';
        if (l < 0 || l >= c.levels) return VRF_ERR_ARG;
        size_t need = (size_t)c.lw[l] * c.lh[l];
        if (dst_bytes < need) return VRF_ERR_ARG;
        const uint8_t *src = d.pyr[h->cur_buf[seq]] + (size_t)seq * c.pyr_bytes + c.loff[l];
        if (cudaMemcpy2D(dst, c.lw[l], src, c.lp[l], c.lw[l], c.lh[l], cudaMemcpyDeviceToHost) != cudaSuccess) return VRF_ERR_CUDA;
        return (long)need;
    }
    if (w == "ba_prof") return ba_debug_prof(h, seq, dst, dst_bytes);
    const void *src = nullptr;
    size_t need = 0;
    if (w == "cand") { src = d.cand + (size_t)seq * VRF_MAX_CELLS * c.kmax * 3; need = (size_t)c.ncells * c.kmax * 3 * sizeof(float); }
    else if (w == "ncand") { src = d.ncand + (size_t)seq * VRF_MAX_CELLS; need = c.ncells * sizeof(int); }
    else if (w == "cell_k") { src = d.cell_k + (size_t)seq * VRF_MAX_CELLS; need = c.ncells * sizeof(int); }
    else if (w == "maskpts") {
        int nm = 0;
        if (cudaMemcpy(&nm, d.n_maskpts + seq, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return VRF_ERR_CUDA;
        src = d.maskpts + (size_t)seq * 2 * VRF_CAP; need = (size_t)nm * sizeof(int2);
    } else return VRF_ERR_ARG;
    if (dst_bytes < need) return VRF_ERR_ARG;
    if (need && cudaMemcpy(dst, src, need, cudaMemcpyDeviceToHost) != cudaSuccess) return VRF_ERR_CUDA;
    return (long)need;
}

// Test entry: FeatureTracker::rejectWithF (feature_tracker.cpp:441-473) alone on caller-supplied point pairs (pixel
// coordinates of cur_pts / forw_pts) through the product kernel k_ransac; clobbers the working arrays of `seq`.
extern "C" int vrf_debug_reject_with_f(vrf_handle *h, int seq, int n, const float *cur_xy, const float *forw_xy, uint8_t *status_out)
{
    if (!h || !cur_xy || !forw_xy || !status_out || seq < 0 || seq >= h->n_seq || n < 0 || n > VRF_CAP) return VRF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    FrontDev &d = h->fd;
    const size_t base = (size_t)seq * VRF_CAP;
    std::vector<uint8_t> ones(n > 0 ? n : 1, 1);
    CK(cudaMemcpy(d.t_prev + base, cur_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.t_forw + base, forw_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.t_keep + base, ones.data(), n, cudaMemcpyHostToDevice));       // status starts as "keep" (n < 8: untouched)
    CK(cudaMemcpy(d.t_n + seq, &n, sizeof(int), cudaMemcpyHostToDevice));
    SeqCall call;
    memset(&call, 0, sizeof(call));
    call.seq = seq; call.pub = 1; call.dslot = -1;
    SeqCall *d_call = h->d_calls_ring[0];
    CK(cudaMemcpy(d_call, &call, sizeof(call), cudaMemcpyHostToDevice));
    LaunchCtx lc{h->stream, &h->launches, &h->prof};
    if (ransac_launch(h->fc, d_call, 1, d, lc) != 0) return VRF_ERR_CUDA;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(status_out, d.t_keep + base, n, cudaMemcpyDeviceToHost));
    return VRF_OK;
}

extern "C" int vrf_profile_enable(vrf_handle *h, int on)
{
    if (!h) return VRF_ERR_ARG;
    h->prof.enabled = on != 0;
    return VRF_OK;
}

extern "C" int vrf_profile_read(vrf_handle *h, int max_kernels, const char **names, double *total_ms,
                                uint64_t *launch_counts, int reset)
{
    if (!h) return VRF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    Prof &p = h->prof;
    for (auto &r : p.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { p.total_ms[r.id] += ms; p.count[r.id]++; }
        p.pool.push_back(r.a); p.pool.push_back(r.b);
    }
    p.recs.clear();
    int n = K_COUNT < max_kernels ? K_COUNT : max_kernels;
    for (int i = 0; i < n; ++i) {
        if (names) names[i] = kKernelNames[i];
        if (total_ms) total_ms[i] = p.total_ms[i];
        if (launch_counts) launch_counts[i] = p.count[i];
    }
    if (reset) for (int i = 0; i < K_COUNT; ++i) { p.total_ms[i] = 0; p.count[i] = 0; }
    return n;
}
