// ba_host.cu -- back-end entry points of the C ABI (include/vrf_ba.h): packs one
// Estimator::optimization() call per sequence (vector2double order, reference
// estimator.cpp:936-981 + problem assembly :1166-1302) into device buffers, launches the
// solve + marginalization kernels, and unpacks results.  Thin host code; no arithmetic of
// the hot path happens here.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <vector>

#include "ba_dev.cuh"

namespace vrf {

// pinned host staging of one packed problem
// Fixed part, then a variable part packed back to back for the problem's own M landmarks / O observations:
//   double lam[M], lm_ub[M], obs[2 O] | int start[M], obs_ptr[M + 1] | uint8 lm_const[M]
// so that the upload of a batch is ONE strided copy of only the used prefix of every slot (a 150-landmark window is 60 KB,
// the full-capacity pack 196 KB: the difference was 5 % of the PCIe traffic of the host-buffer path).
struct BaHostPack {
    double pose[BA_NF * 7], sb[BA_NF * 9], ex[7], td;
    VrfImuPreint imu[BA_NF - 1];
    double var[2 * BA_MAX_LM + 2 * BA_MAX_OBS + (2 * BA_MAX_LM + 2) / 2 + BA_MAX_LM / 8 + 2];
};
struct BaPackView { double *lam, *lm_ub, *obs; int *start, *obs_ptr; uint8_t *lm_const; size_t used_bytes; };
static inline BaPackView pack_view(BaHostPack *k, int M, int O)
{
    BaPackView v;
    v.lam = k->var; v.lm_ub = v.lam + M; v.obs = v.lm_ub + M;
    v.start = reinterpret_cast<int *>(v.obs + 2 * (size_t)O); v.obs_ptr = v.start + M;
    v.lm_const = reinterpret_cast<uint8_t *>(v.obs_ptr + M + 1);
    v.used_bytes = (size_t)(v.lm_const + M - reinterpret_cast<uint8_t *>(k));
    return v;
}
// ProjectionTdFactor inputs (only staged / uploaded when VrfConfig::estimate_td)
struct BaHostPackTd {
    double vel[BA_MAX_OBS * 2], ctd[BA_MAX_OBS], row[BA_MAX_OBS];
};

#define BA_PIPE 2      // batches in flight (vrf_ba_submit_batch / vrf_ba_collect_batch)

// Everything a batch in flight owns: pinned staging, its device mirror, the result buffers.  Two slots so that
// batch k+1 can be packed and uploaded while batch k runs.
struct BaSlot {
    BaHostPack *h_pack = nullptr, *d_pack = nullptr;       // [n_seq]
    BaHostPackTd *h_packtd = nullptr, *d_packtd = nullptr; // [n_seq], estimate_td only
    BaMeta *h_meta = nullptr, *d_meta = nullptr;
    BaProbDev *h_prob = nullptr, *d_prob = nullptr;
    BaOutDev *h_out = nullptr, *d_out = nullptr;
    BaMargDev *h_marg = nullptr, *d_marg = nullptr;
    BaPriorStore *h_prior = nullptr;                       // pinned upload staging [n_seq]
    double *h_lam = nullptr, *d_lam_out = nullptr;         // optimised inverse depths by batch position [n_seq][BA_MAX_LM]
    cudaEvent_t done = nullptr;
    size_t used_bytes = 0;                                 // widest used prefix of a BaHostPack in the batch being packed
    bool busy = false;
    std::vector<int> seqs;                                 // sequences of the batch this slot holds
};

struct BaState {
    int n_seq = 0;
    BaSlot slot[BA_PIPE];
    unsigned n_submit = 0, n_collect = 0;
    BaPriorStore *d_prior[2] = {nullptr, nullptr};   // [n_seq] each; cur index per sequence below
    BaPriorStore *h_prior_dl = nullptr;              // pinned download staging [1]
    BaPriorStore *d_prior_factor = nullptr;          // [1] factor form produced on demand for the host copy-out
    double *d_factor_scratch = nullptr;              // [2 * VRF_PRIOR_MAX_DIM^2]
    std::vector<int> prior_cur;                      // which store is "last_marginalization_info"
    std::vector<uint8_t> prior_valid;
    // per-sequence scratch
    double *d_lam = nullptr, *d_clam = nullptr, *d_W = nullptr, *d_vecs = nullptr, *d_imuS = nullptr, *d_HP = nullptr;
    int *d_colmap = nullptr;
    double *d_margbuf = nullptr;
    int *d_lmcol = nullptr;
    int *d_fac = nullptr;                            // [n_seq][BA_MAX_OBS] pair-major projection factor lists (k_ba_solve scratch)
    double *d_pair_part = nullptr, *d_fpart = nullptr, *d_task_cost = nullptr;   // fixed-order accumulation scratch (ba_dev.cuh)
    std::vector<int> last_M;
    std::vector<int> last_slots;
};

static const size_t kMargDoubles = (size_t)BA_MAX_POS * BA_MAX_POS + BA_MAX_POS + 2 * (size_t)(15 + BA_MAX_M0) * (15 + BA_MAX_M0) +
                                   (size_t)VRF_PRIOR_MAX_DIM * (15 + BA_MAX_M0) + 2 * (size_t)VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM +
                                   VRF_PRIOR_MAX_DIM;

#define BCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { snprintf(h->errbuf, sizeof(h->errbuf), "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e__)); return VRF_ERR_CUDA; } } while (0)

static int slot_alloc(vrf_handle *h, BaSlot &sl)
{
    if (sl.h_pack) return VRF_OK;
    const size_t S = h->n_seq;
    BCK(cudaMallocHost((void **)&sl.h_pack, S * sizeof(BaHostPack)));
    BCK(cudaMalloc((void **)&sl.d_pack, S * sizeof(BaHostPack)));
    if (h->cfg.estimate_td) {
        BCK(cudaMallocHost((void **)&sl.h_packtd, S * sizeof(BaHostPackTd)));
        BCK(cudaMalloc((void **)&sl.d_packtd, S * sizeof(BaHostPackTd)));
    }
    BCK(cudaMallocHost((void **)&sl.h_meta, S * sizeof(BaMeta)));
    BCK(cudaMalloc((void **)&sl.d_meta, S * sizeof(BaMeta)));
    BCK(cudaMallocHost((void **)&sl.h_prob, S * sizeof(BaProbDev)));
    BCK(cudaMalloc((void **)&sl.d_prob, S * sizeof(BaProbDev)));
    BCK(cudaMallocHost((void **)&sl.h_out, S * sizeof(BaOutDev)));
    BCK(cudaMalloc((void **)&sl.d_out, S * sizeof(BaOutDev)));
    BCK(cudaMallocHost((void **)&sl.h_marg, S * sizeof(BaMargDev)));
    BCK(cudaMalloc((void **)&sl.d_marg, S * sizeof(BaMargDev)));
    BCK(cudaMallocHost((void **)&sl.h_prior, S * sizeof(BaPriorStore)));
    BCK(cudaMallocHost((void **)&sl.h_lam, S * BA_MAX_LM * sizeof(double)));
    BCK(cudaMalloc((void **)&sl.d_lam_out, S * BA_MAX_LM * sizeof(double)));
    BCK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    return VRF_OK;
}

int ba_create(vrf_handle *h)
{
    BaState *b = new (std::nothrow) BaState();
    if (!b) return VRF_ERR_ARG;
    h->ba = b;
    const size_t S = h->n_seq;
    b->n_seq = h->n_seq;
    // kernel attributes are per device: vrf_create() has made the handle's device current
    if (ba_solve_configure() != 0 || ba_marg_configure() != 0) { snprintf(h->errbuf, sizeof(h->errbuf), "BA kernel attribute setup failed"); return VRF_ERR_CUDA; }
    if (int rc = slot_alloc(h, b->slot[0])) return rc;        // slot 1 is allocated by the first pipelined submit
    for (int k = 0; k < 2; ++k) {
        BCK(cudaMalloc((void **)&b->d_prior[k], S * sizeof(BaPriorStore)));
        BCK(cudaMemset(b->d_prior[k], 0, S * sizeof(BaPriorStore)));
    }
    BCK(cudaMallocHost((void **)&b->h_prior_dl, sizeof(BaPriorStore)));
    BCK(cudaMalloc((void **)&b->d_prior_factor, sizeof(BaPriorStore)));
    BCK(cudaMalloc((void **)&b->d_factor_scratch, 2 * (size_t)VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_lam, S * BA_MAX_LM * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_clam, S * BA_MAX_LM * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_W, S * BA_MAX_LM * BA_WS * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_vecs, S * 10 * BA_MAX_LM * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_imuS, S * (BA_NF - 1) * 225 * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_HP, S * VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_colmap, S * VRF_PRIOR_MAX_DIM * sizeof(int)));
    BCK(cudaMalloc((void **)&b->d_margbuf, S * kMargDoubles * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_lmcol, S * BA_MAX_LM * sizeof(int)));
    BCK(cudaMalloc((void **)&b->d_fac, S * BA_MAX_OBS * sizeof(int)));
    BCK(cudaMalloc((void **)&b->d_pair_part, S * (BA_NF * (BA_NF - 1) / 2) * BA_PP_STRIDE * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_fpart, S * (size_t)BA_MAX_OBS * BA_FP_STRIDE * sizeof(double)));
    BCK(cudaMalloc((void **)&b->d_task_cost, S * 2 * BA_MAX_TASKS * sizeof(double)));
    b->prior_cur.assign(S, 0);
    b->prior_valid.assign(S, 0);
    b->last_M.assign(S, 0);
    return VRF_OK;
}

void ba_destroy(vrf_handle *h)
{
    BaState *b = h->ba;
    if (!b) return;
    void *dev[] = {b->d_prior[0], b->d_prior[1], b->d_prior_factor, b->d_factor_scratch, b->d_lam, b->d_clam,
                   b->d_W, b->d_vecs, b->d_imuS, b->d_HP, b->d_colmap, b->d_margbuf, b->d_lmcol, b->d_fac,
                   b->d_pair_part, b->d_fpart, b->d_task_cost};
    for (void *p : dev) if (p) cudaFree(p);
    if (b->h_prior_dl) cudaFreeHost(b->h_prior_dl);
    for (BaSlot &sl : b->slot) {
        void *sdev[] = {sl.d_pack, sl.d_meta, sl.d_prob, sl.d_out, sl.d_marg, sl.d_lam_out, sl.d_packtd};
        for (void *p : sdev) if (p) cudaFree(p);
        void *host[] = {sl.h_pack, sl.h_meta, sl.h_prob, sl.h_out, sl.h_marg, sl.h_prior, sl.h_lam, sl.h_packtd};
        for (void *p : host) if (p) cudaFreeHost(p);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    delete b;
    h->ba = nullptr;
}

int ba_reset_sequence(vrf_handle *h, int seq)
{
    BaState *b = h->ba;
    if (!b) return VRF_OK;
    b->prior_valid[seq] = 0;
    return VRF_OK;
}

static int pack_problem(vrf_handle *h, BaSlot &sl, int slot, int seq, const VrfBaProblem *pb)
{
    BaState *b = h->ba;
    if (!pb || pb->n_landmarks < 0 || pb->n_obs < 0) return VRF_ERR_ARG;
    if (pb->n_landmarks > BA_MAX_LM || pb->n_obs > BA_MAX_OBS) return VRF_ERR_CAPACITY;
    if (pb->n_landmarks > 0 && pb->lm_obs_ptr && pb->lm_obs_ptr[0] != 0) return VRF_ERR_ARG;      // CSR offsets start at 0
    if (pb->frame_count < 1 || pb->frame_count > VRF_WINDOW_SIZE) return VRF_ERR_ARG;
    const int td_factor = h->cfg.estimate_td != 0;
    if (td_factor && pb->n_landmarks > 0 && (!pb->obs_velocity || !pb->obs_cur_td || !pb->obs_row)) return VRF_ERR_ARG;
    if (pb->n_landmarks > 0 && (!pb->para_Feature || !pb->lm_start_frame || !pb->lm_estimate_flag || !pb->lm_obs_ptr || !pb->obs_pts)) return VRF_ERR_ARG;
    if (pb->use_imu && !pb->imu) return VRF_ERR_ARG;
    BaHostPack &k = sl.h_pack[slot];
    const BaPackView kv = pack_view(&k, pb->n_landmarks, pb->n_obs);
    sl.used_bytes = std::max(sl.used_bytes, kv.used_bytes);
    memcpy(k.pose, pb->para_Pose, sizeof(k.pose));
    memcpy(k.sb, pb->para_SpeedBias, sizeof(k.sb));
    memcpy(k.ex, pb->para_Ex_Pose, sizeof(k.ex));
    k.td = pb->para_Td;
    if (pb->use_imu) memcpy(k.imu, pb->imu, sizeof(VrfImuPreint) * pb->frame_count);
    const int M = pb->n_landmarks;
    for (int l = 0; l < M; ++l) {
        kv.lam[l] = pb->para_Feature[l];
        kv.start[l] = pb->lm_start_frame[l];
        kv.obs_ptr[l] = pb->lm_obs_ptr[l];
        const int nobs = pb->lm_obs_ptr[l + 1] - pb->lm_obs_ptr[l];
        if (nobs < 1 || pb->lm_start_frame[l] < 0 || pb->lm_start_frame[l] + nobs - 1 > pb->frame_count) return VRF_ERR_ARG;
        // estimator.cpp:1291-1298: constant if (flag==1 && FIX_DEPTH); upper bound if flag==2
        kv.lm_const[l] = (pb->lm_estimate_flag[l] == 1 && h->cfg.fix_depth) ? 1 : 0;
        kv.lm_ub[l] = (pb->lm_estimate_flag[l] == 2) ? 2.0 / h->cfg.depth_max_dist : INFINITY;
    }
    kv.obs_ptr[M] = M ? pb->lm_obs_ptr[M] : 0;
    if (kv.obs_ptr[M] != pb->n_obs) return VRF_ERR_ARG;
    memcpy(kv.obs, pb->obs_pts, sizeof(double) * 2 * pb->n_obs);
    if (td_factor && pb->n_obs > 0) {
        BaHostPackTd &kt = sl.h_packtd[slot];
        memcpy(kt.vel, pb->obs_velocity, sizeof(double) * 2 * pb->n_obs);
        memcpy(kt.ctd, pb->obs_cur_td, sizeof(double) * pb->n_obs);
        memcpy(kt.row, pb->obs_row, sizeof(double) * pb->n_obs);
    }

    BaMeta &mt = sl.h_meta[slot];
    memset(&mt, 0, sizeof(mt));
    mt.M = M; mt.nobs = pb->n_obs; mt.nframes = pb->frame_count + 1; mt.use_imu = pb->use_imu;
    mt.frame_count = pb->frame_count;
    mt.max_iter = pb->max_iterations > 0 ? pb->max_iterations : h->cfg.num_iterations;
    mt.marg_flag = pb->marginalization_flag;
    mt.g_norm = h->cfg.g_norm;
    // estimator.cpp:1186-1212: ex-pose / td variable or constant; ESTIMATE_TD selects ProjectionTdFactor (:1270).
    // para_Td is only added (and possibly fixed) under USE_IMU; without IMU Ceres adds it implicitly as a variable.
    mt.ex_active = pb->ex_constant ? 0 : 1;
    mt.td_factor = td_factor;
    mt.td_active = (td_factor && !(pb->use_imu && pb->td_constant)) ? 1 : 0;
    mt.tr_over_row = h->cfg.tr / (double)h->cfg.row;
    { const char *e = getenv("VRF_BA_DEBUG"); mt.debug = e ? atoi(e) : 0; }
    mt.nimu = 0;
    if (pb->use_imu)
        for (int j = 1; j <= pb->frame_count; ++j) {
            if (pb->imu[j - 1].sum_dt > 10.0) continue;          // estimator.cpp:1231-1233
            mt.imu_j[mt.nimu++] = j;
        }
    // prior: host-supplied, device-resident from the previous call, or none
    int have_prior = 0;
    if (pb->prior == VRF_PRIOR_DEVICE) have_prior = b->prior_valid[seq];
    else if (pb->prior) {
        const VrfPrior *P = pb->prior;
        if (P->n < 0 || P->n > VRF_PRIOR_MAX_DIM || P->n_blocks > VRF_PRIOR_MAX_BLOCKS) return VRF_ERR_ARG;
        BaPriorStore *hp = sl.h_prior + slot;          // per-problem pinned staging (the slot is idle while it is packed)
        hp->n = P->n; hp->n_blocks = P->n_blocks; hp->valid = 1; hp->form = 0; hp->c0 = 0.0;
        for (int q = 0; q < P->n_blocks; ++q) {
            hp->kind[q] = P->blocks[q].kind; hp->index[q] = P->blocks[q].index; hp->size[q] = P->blocks[q].size; hp->idx[q] = P->blocks[q].idx;
            memcpy(hp->x0 + 9 * q, P->blocks[q].x0, sizeof(double) * 9);
        }
        memcpy(hp->r0, P->linearized_residuals, sizeof(double) * P->n);
        memcpy(hp->J0, P->linearized_jacobians, sizeof(double) * (size_t)P->n * P->n);
        BaPriorStore *dst = b->d_prior[b->prior_cur[seq]] + seq;
        const size_t bytes = offsetof(BaPriorStore, J0) + sizeof(double) * (size_t)P->n * P->n;
        if (cudaMemcpyAsync(dst, hp, bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return VRF_ERR_CUDA;
        b->prior_valid[seq] = 1;
        have_prior = 1;
    } else b->prior_valid[seq] = 0;
    mt.has_prior = have_prior;

    BaProbDev &pd = sl.h_prob[slot];
    BaHostPack *dp = sl.d_pack + slot;
    const BaPackView dv = pack_view(dp, pb->n_landmarks, pb->n_obs);           // same layout, device addresses
    pd.pose0 = dp->pose; pd.sb0 = dp->sb; pd.ex0 = dp->ex; pd.lam0 = dv.lam; pd.td0 = &dp->td;
    if (td_factor) { BaHostPackTd *dt = sl.d_packtd + slot; pd.obs_vel = dt->vel; pd.obs_td = dt->ctd; pd.obs_row = dt->row; }
    else { pd.obs_vel = nullptr; pd.obs_td = nullptr; pd.obs_row = nullptr; }
    pd.start = dv.start; pd.obs_ptr = dv.obs_ptr; pd.lm_const = dv.lm_const; pd.lm_ub = dv.lm_ub; pd.obs = dv.obs; pd.imu = dp->imu;
    pd.prior = have_prior ? b->d_prior[b->prior_cur[seq]] + seq : nullptr;
    pd.prior_next = b->d_prior[1 - b->prior_cur[seq]] + seq;
    pd.HP = b->d_HP + (size_t)seq * VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM;
    pd.colmap = b->d_colmap + (size_t)seq * VRF_PRIOR_MAX_DIM;
    pd.lam = b->d_lam + (size_t)seq * BA_MAX_LM;
    pd.lam_out = sl.d_lam_out + (size_t)slot * BA_MAX_LM;
    pd.clam = b->d_clam + (size_t)seq * BA_MAX_LM;
    pd.W = b->d_W + (size_t)seq * BA_MAX_LM * BA_WS;
    double *v = b->d_vecs + (size_t)seq * 10 * BA_MAX_LM;
    pd.hll = v; pd.gl = v + BA_MAX_LM; pd.jscale_l = v + 2 * BA_MAX_LM; pd.diag_l = v + 3 * BA_MAX_LM; pd.gd_l = v + 4 * BA_MAX_LM;
    pd.gn_l = v + 5 * BA_MAX_LM; pd.u_l = v + 6 * BA_MAX_LM; pd.y_l = v + 7 * BA_MAX_LM; pd.hinv_l = v + 8 * BA_MAX_LM;
    pd.shinv_l = v + 9 * BA_MAX_LM;
    pd.imuS = b->d_imuS + (size_t)seq * (BA_NF - 1) * 225;
    pd.fac = b->d_fac + (size_t)seq * BA_MAX_OBS;
    pd.pair_part = b->d_pair_part + (size_t)seq * (BA_NF * (BA_NF - 1) / 2) * BA_PP_STRIDE;
    pd.fpart = b->d_fpart + (size_t)seq * BA_MAX_OBS * BA_FP_STRIDE;
    pd.task_cost = b->d_task_cost + (size_t)seq * 2 * BA_MAX_TASKS;

    BaMargDev &mg = sl.h_marg[slot];
    double *mb = b->d_margbuf + (size_t)seq * kMargDoubles;
    const size_t mmx = 15 + BA_MAX_M0;
    mg.A = mb; mb += (size_t)BA_MAX_POS * BA_MAX_POS;
    mg.b = mb; mb += BA_MAX_POS;
    mg.V = mb; mb += mmx * mmx;
    mg.Ainv = mb; mb += mmx * mmx;
    mg.T = mb; mb += (size_t)VRF_PRIOR_MAX_DIM * mmx;
    mg.Ar = mb; mb += (size_t)VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM;
    mg.V2 = mb; mb += (size_t)VRF_PRIOR_MAX_DIM * VRF_PRIOR_MAX_DIM;
    mg.br = mb;
    mg.lmcol = b->d_lmcol + (size_t)seq * BA_MAX_LM;
    b->last_M[seq] = M;
    return VRF_OK;
}

// pack + upload one batch into pipeline slot `sl` (which must be idle)
static int upload(vrf_handle *h, BaSlot &sl, int n, const int32_t *seqs, const VrfBaProblem *probs)
{
    BaState *b = h->ba;
    if (!b || n < 1 || n > h->n_seq || !seqs || !probs) return VRF_ERR_ARG;
    BCK(cudaSetDevice(h->device));
    if (int rc = slot_alloc(h, sl)) return rc;
    std::vector<uint8_t> seen(h->n_seq, 0);
    for (int i = 0; i < n; ++i) {
        if (seqs[i] < 0 || seqs[i] >= h->n_seq || seen[seqs[i]]) return VRF_ERR_ARG;
        seen[seqs[i]] = 1;
    }
    // a sequence's window k+1 is built from the results of window k: batches in flight must be disjoint
    for (BaSlot &o : b->slot)
        if (&o != &sl && o.busy)
            for (int q : o.seqs) if (seen[q]) return VRF_ERR_ARG;
    sl.used_bytes = 0;
    for (int i = 0; i < n; ++i) {
        int rc = pack_problem(h, sl, i, seqs[i], &probs[i]);
        if (rc != VRF_OK) return rc;
    }
    // one strided copy of the used prefix of the n packed problems (pinned staging -> HBM)
    const size_t width = std::min(sizeof(BaHostPack), (sl.used_bytes + 15) & ~(size_t)15);
    BCK(cudaMemcpy2DAsync(sl.d_pack, sizeof(BaHostPack), sl.h_pack, sizeof(BaHostPack), width, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if (h->cfg.estimate_td)
        BCK(cudaMemcpyAsync(sl.d_packtd, sl.h_packtd, (size_t)n * sizeof(BaHostPackTd), cudaMemcpyHostToDevice, h->stream));
    BCK(cudaMemcpyAsync(sl.d_meta, sl.h_meta, n * sizeof(BaMeta), cudaMemcpyHostToDevice, h->stream));
    BCK(cudaMemcpyAsync(sl.d_prob, sl.h_prob, n * sizeof(BaProbDev), cudaMemcpyHostToDevice, h->stream));
    BCK(cudaMemcpyAsync(sl.d_marg, sl.h_marg, n * sizeof(BaMargDev), cudaMemcpyHostToDevice, h->stream));
    sl.seqs.assign(seqs, seqs + n);
    return VRF_OK;
}

static int enqueue(vrf_handle *h, BaSlot &sl, int n)
{
    LaunchCtx lc{h->stream, &h->launches, &h->prof};
    if (ba_solve_launch(sl.d_meta, sl.d_prob, sl.d_out, n, lc) != 0) return VRF_ERR_CUDA;
    if (ba_marg_launch(sl.d_meta, sl.d_prob, sl.d_out, sl.d_marg, n, lc) != 0) return VRF_ERR_CUDA;
    BCK(cudaGetLastError());
    return VRF_OK;
}

// result copy of a batch, stream-ordered behind its kernels; `done` marks its completion
static int enqueue_download(vrf_handle *h, BaSlot &sl, int n, bool want_lam)
{
    BCK(cudaMemcpyAsync(sl.h_out, sl.d_out, n * sizeof(BaOutDev), cudaMemcpyDeviceToHost, h->stream));
    if (want_lam)
        BCK(cudaMemcpyAsync(sl.h_lam, sl.d_lam_out, (size_t)n * BA_MAX_LM * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    BCK(cudaEventRecord(sl.done, h->stream));
    return VRF_OK;
}

static int finish_download(vrf_handle *h, BaSlot &sl, int n, const int32_t *seqs, VrfBaResult *res)
{
    BaState *b = h->ba;
    BCK(cudaEventSynchronize(sl.done));
    int worst = VRF_OK;
    for (int i = 0; i < n; ++i) {
        const BaOutDev &o = sl.h_out[i];
        const int seq = seqs[i];
        if (o.has_new_prior) { b->prior_cur[seq] = 1 - b->prior_cur[seq]; b->prior_valid[seq] = 1; }
        if (!res) continue;
        VrfBaResult &r = res[i];
        r.status = o.status; r.iterations = o.iterations; r.successful_steps = o.successful; r.termination = o.termination;
        r.initial_cost = o.initial_cost; r.final_cost = o.final_cost;
        memcpy(r.para_Pose, o.pose, sizeof(o.pose)); memcpy(r.para_SpeedBias, o.sb, sizeof(o.sb)); memcpy(r.para_Ex_Pose, o.ex, sizeof(o.ex));
        r.para_Td = o.td;
        memcpy(r.Ps, o.Ps, sizeof(o.Ps)); memcpy(r.Rs, o.Rs, sizeof(o.Rs)); memcpy(r.Vs, o.Vs, sizeof(o.Vs));
        memcpy(r.Bas, o.Bas, sizeof(o.Bas)); memcpy(r.Bgs, o.Bgs, sizeof(o.Bgs));
        r.has_new_prior = o.has_new_prior;
        r.armijo_failures = o.armijo_failures;
        if (r.para_Feature && b->last_M[seq] > 0)
            memcpy(r.para_Feature, sl.h_lam + (size_t)i * BA_MAX_LM, sizeof(double) * b->last_M[seq]);
        if (o.has_new_prior && r.new_prior) {
            // (the store just produced is only read, never written, by the one later batch that may be in flight)
            // the stored prior is in information form; the reference's (linearized_jacobians, linearized_residuals)
            // are produced here, on demand, by the eigen-decomposition kernel
            BaPriorStore *hp = b->h_prior_dl;
            {
                LaunchCtx lc{h->stream, &h->launches, &h->prof};
                if (ba_prior_factor_launch(b->d_prior[b->prior_cur[seq]] + seq, b->d_prior_factor, b->d_factor_scratch, lc) != 0) return VRF_ERR_CUDA;
                BCK(cudaStreamSynchronize(h->stream));
            }
            const BaPriorStore *src = b->d_prior_factor;
            BCK(cudaMemcpy(hp, src, offsetof(BaPriorStore, J0), cudaMemcpyDeviceToHost));
            const int nn = hp->n;
            BCK(cudaMemcpy(hp->J0, src->J0, sizeof(double) * (size_t)nn * nn, cudaMemcpyDeviceToHost));
            VrfPrior *P = r.new_prior;
            P->n = nn; P->n_blocks = hp->n_blocks;
            for (int q = 0; q < hp->n_blocks; ++q) {
                P->blocks[q].kind = hp->kind[q]; P->blocks[q].index = hp->index[q]; P->blocks[q].size = hp->size[q]; P->blocks[q].idx = hp->idx[q];
                memcpy(P->blocks[q].x0, hp->x0 + 9 * q, sizeof(double) * 9);
            }
            memcpy(P->linearized_residuals, hp->r0, sizeof(double) * nn);
            memcpy(P->linearized_jacobians, hp->J0, sizeof(double) * (size_t)nn * nn);
        }
        if (o.status != VRF_OK) worst = o.status;
    }
    return worst;
}

long ba_debug_prof(vrf_handle *h, int slot, void *dst, size_t bytes)
{
    BaState *b = h->ba;
    if (!b || slot < 0 || slot >= h->n_seq || bytes < sizeof(long long) * 20) return VRF_ERR_ARG;
    BaOutDev tmp;
    if (cudaMemcpy(&tmp, b->slot[0].d_out + slot, sizeof(BaOutDev), cudaMemcpyDeviceToHost) != cudaSuccess) return VRF_ERR_CUDA;
    long long o[20];
    memcpy(o, tmp.prof, sizeof(long long) * 8);
    memcpy(o + 8, tmp.prof2, sizeof(long long) * 8);
    o[16] = tmp.iterations; o[17] = tmp.successful; o[18] = tmp.termination; o[19] = tmp.status;
    memcpy(dst, o, sizeof(o));
    return (long)sizeof(o);
}

}  // namespace vrf

using namespace vrf;

static bool pipe_idle(const BaState *b) { return b->n_submit == b->n_collect; }

// The split device-resident form works on pipeline slot 0 and requires an empty pipeline.
extern "C" int vrf_ba_upload_batch(vrf_handle *h, int n, const int32_t *seqs, const VrfBaProblem *probs)
{
    if (!h || !h->ba || !pipe_idle(h->ba)) return VRF_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return VRF_ERR_CUDA;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return VRF_ERR_CUDA;      // pinned staging reuse
    return upload(h, h->ba->slot[0], n, seqs, probs);
}

extern "C" int vrf_ba_enqueue_batch(vrf_handle *h, int n, const int32_t *seqs)
{
    if (!h || !h->ba || !seqs || n < 1 || !pipe_idle(h->ba) || n != (int)h->ba->slot[0].seqs.size()) return VRF_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return VRF_ERR_CUDA;
    return enqueue(h, h->ba->slot[0], n);
}

extern "C" int vrf_ba_download_batch(vrf_handle *h, int n, const int32_t *seqs, VrfBaResult *res)
{
    if (!h || !h->ba || !seqs || n < 1 || !pipe_idle(h->ba) || n != (int)h->ba->slot[0].seqs.size()) return VRF_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return VRF_ERR_CUDA;
    int rc = enqueue_download(h, h->ba->slot[0], n, res != nullptr);
    if (rc != VRF_OK) return rc;
    return finish_download(h, h->ba->slot[0], n, seqs, res);
}

extern "C" int vrf_ba_submit_batch(vrf_handle *h, int n, const int32_t *seqs, const VrfBaProblem *probs)
{
    if (!h || !h->ba) return VRF_ERR_ARG;
    BaState *b = h->ba;
    BaSlot &sl = b->slot[b->n_submit % BA_PIPE];
    if (sl.busy) return VRF_ERR_CAPACITY;             // BA_PIPE batches already in flight: collect first
    int rc = upload(h, sl, n, seqs, probs);
    if (rc != VRF_OK) return rc;
    rc = enqueue(h, sl, n);
    if (rc != VRF_OK) return rc;
    rc = enqueue_download(h, sl, n, true);
    if (rc != VRF_OK) return rc;
    sl.busy = true;
    b->n_submit++;
    return VRF_OK;
}

extern "C" int vrf_ba_collect_batch(vrf_handle *h, int n, const int32_t *seqs, VrfBaResult *res)
{
    if (!h || !h->ba || !seqs) return VRF_ERR_ARG;
    BaState *b = h->ba;
    BaSlot &sl = b->slot[b->n_collect % BA_PIPE];
    if (!sl.busy || n != (int)sl.seqs.size()) return VRF_ERR_ARG;
    for (int i = 0; i < n; ++i) if (seqs[i] != sl.seqs[i]) return VRF_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return VRF_ERR_CUDA;
    int rc = finish_download(h, sl, n, seqs, res);
    sl.busy = false;
    b->n_collect++;
    return rc;
}

extern "C" int vrf_ba_solve_batch(vrf_handle *h, int n, const int32_t *seqs, const VrfBaProblem *probs, VrfBaResult *res)
{
    if (!h || !h->ba || !pipe_idle(h->ba)) return VRF_ERR_ARG;
    int rc = vrf_ba_submit_batch(h, n, seqs, probs);
    if (rc != VRF_OK) return rc;
    return vrf_ba_collect_batch(h, n, seqs, res);
}

extern "C" int vrf_ba_solve(vrf_handle *h, int seq, const VrfBaProblem *prob, VrfBaResult *res)
{
    int32_t s = seq;
    return vrf_ba_solve_batch(h, 1, &s, prob, res);
}
