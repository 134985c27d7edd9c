"""ctypes binding of the C ABI declared in include/vrf.h (harness only).

Fails loudly if libvrf.so is missing: the product has no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("VRF_LIB_PATH") or os.path.join(_PKG, "libvrf.so")      # override: A/B builds of the same CUDA library

TRACK_CAP = 1024
NUM_FRAMES = 11
PRIOR_MAX_BLOCKS = 40
PRIOR_MAX_DIM = 176

FMT_GRAY8, FMT_RGB8 = 0, 1
DEPTH_NONE, DEPTH_16UC1, DEPTH_32FC1 = 0, 1, 2
MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1
BLK_POSE, BLK_SPEEDBIAS, BLK_EXPOSE, BLK_TD = 0, 1, 2, 3


class VrfConfig(C.Structure):
    _fields_ = [
        ("row", C.c_int32), ("col", C.c_int32), ("max_cnt", C.c_int32), ("min_dist", C.c_int32),
        ("num_grid_rows", C.c_int32), ("num_grid_cols", C.c_int32), ("use_imu", C.c_int32),
        ("equalize", C.c_int32), ("fisheye", C.c_int32), ("lk_max_level", C.c_int32),
        ("use_ransac", C.c_int32), ("reserved0", C.c_int32),
        ("f_threshold", C.c_double), ("focal_length", C.c_double),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("k1", C.c_double), ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
        ("num_iterations", C.c_int32), ("estimate_extrinsic", C.c_int32), ("estimate_td", C.c_int32),
        ("fix_depth", C.c_int32), ("depth_max_dist", C.c_double), ("g_norm", C.c_double),
        ("acc_n", C.c_double), ("acc_w", C.c_double), ("gyr_n", C.c_double), ("gyr_w", C.c_double),
        ("depth_min_dist", C.c_double), ("tr", C.c_double),
    ]


class VrfTrackOut(C.Structure):
    _fields_ = [
        ("capacity", C.c_int32), ("n", C.c_int32),
        ("cur_pts", C.c_void_p), ("cur_un_pts", C.c_void_p), ("pts_velocity", C.c_void_p),
        ("ids", C.c_void_p), ("track_cnt", C.c_void_p),
        ("n_id", C.c_int32), ("n_predict", C.c_int32),
        ("predict_pts", C.c_void_p), ("lk_pts", C.c_void_p), ("lk_status", C.c_void_p),
        ("grids_track_num", C.c_void_p), ("grids_texture_status", C.c_void_p),
        ("n_unstable", C.c_int32), ("status", C.c_int32),
        ("depth_mm", C.c_void_p), ("depth_keep", C.c_void_p),
    ]


class VrfImuPreint(C.Structure):
    _fields_ = [
        ("sum_dt", C.c_double), ("delta_p", C.c_double * 3), ("delta_q", C.c_double * 4),
        ("delta_v", C.c_double * 3), ("linearized_ba", C.c_double * 3), ("linearized_bg", C.c_double * 3),
        ("jacobian", C.c_double * 225), ("covariance", C.c_double * 225),
    ]


class VrfPriorBlock(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_int32), ("size", C.c_int32), ("idx", C.c_int32),
                ("x0", C.c_double * 9)]


class VrfPrior(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_blocks", C.c_int32),
        ("blocks", VrfPriorBlock * PRIOR_MAX_BLOCKS),
        ("linearized_jacobians", C.c_double * (PRIOR_MAX_DIM * PRIOR_MAX_DIM)),
        ("linearized_residuals", C.c_double * PRIOR_MAX_DIM),
    ]


class VrfBaProblem(C.Structure):
    _fields_ = [
        ("frame_count", C.c_int32), ("use_imu", C.c_int32), ("ex_constant", C.c_int32),
        ("td_constant", C.c_int32), ("marginalization_flag", C.c_int32), ("max_iterations", C.c_int32),
        ("para_Pose", (C.c_double * 7) * NUM_FRAMES), ("para_SpeedBias", (C.c_double * 9) * NUM_FRAMES),
        ("para_Ex_Pose", C.c_double * 7), ("para_Td", C.c_double),
        ("n_landmarks", C.c_int32), ("n_obs", C.c_int32),
        ("para_Feature", C.c_void_p), ("lm_start_frame", C.c_void_p), ("lm_estimate_flag", C.c_void_p),
        ("lm_obs_ptr", C.c_void_p), ("obs_pts", C.c_void_p),
        ("imu", C.POINTER(VrfImuPreint)), ("prior", C.POINTER(VrfPrior)),
        ("obs_velocity", C.c_void_p), ("obs_cur_td", C.c_void_p), ("obs_row", C.c_void_p),
    ]


class VrfBaResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("iterations", C.c_int32), ("successful_steps", C.c_int32),
        ("termination", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("para_Pose", (C.c_double * 7) * NUM_FRAMES), ("para_SpeedBias", (C.c_double * 9) * NUM_FRAMES),
        ("para_Ex_Pose", C.c_double * 7), ("para_Td", C.c_double), ("para_Feature", C.c_void_p),
        ("Ps", (C.c_double * 3) * NUM_FRAMES), ("Rs", (C.c_double * 9) * NUM_FRAMES),
        ("Vs", (C.c_double * 3) * NUM_FRAMES), ("Bas", (C.c_double * 3) * NUM_FRAMES),
        ("Bgs", (C.c_double * 3) * NUM_FRAMES),
        ("has_new_prior", C.c_int32), ("armijo_failures", C.c_int32), ("new_prior", C.POINTER(VrfPrior)),
    ]


class VrfFmProblem(C.Structure):
    """include/vrf_fm.h"""
    _fields_ = [
        ("Ps", (C.c_double * 3) * NUM_FRAMES), ("Rs", (C.c_double * 9) * NUM_FRAMES),
        ("tic", C.c_double * 3), ("ric", C.c_double * 9),
        ("n_landmarks", C.c_int32), ("n_obs", C.c_int32),
        ("lm_start_frame", C.c_void_p), ("lm_obs_ptr", C.c_void_p), ("obs_pts", C.c_void_p), ("obs_depth", C.c_void_p),
        ("estimated_depth", C.c_void_p), ("estimate_flag", C.c_void_p), ("is_dynamic", C.c_void_p), ("remove", C.c_void_p),
    ]


class VrfImuSegment(C.Structure):
    _fields_ = [
        ("acc_0", C.c_double * 3), ("gyr_0", C.c_double * 3), ("linearized_ba", C.c_double * 3), ("linearized_bg", C.c_double * 3),
        ("n_samples", C.c_int32), ("reserved", C.c_int32),
        ("dt", C.c_void_p), ("acc", C.c_void_p), ("gyr", C.c_void_p),
    ]


class FmProblem:
    """Owns the numpy buffers a VrfFmProblem points to (one sequence's f_manager.feature list + window states)."""

    def __init__(self, Ps, Rs, tic, ric, start, obs_ptr, obs_pts, obs_depth, est_depth, est_flag=None, is_dynamic=None):
        self.start = np.ascontiguousarray(start, np.int32)
        self.obs_ptr = np.ascontiguousarray(obs_ptr, np.int32)
        self.obs_pts = np.ascontiguousarray(obs_pts, np.float64).reshape(-1, 2)
        self.obs_depth = np.ascontiguousarray(obs_depth, np.float64)
        M = len(self.start)
        self.est_depth = np.ascontiguousarray(est_depth, np.float64).copy()
        self.est_flag = np.zeros(M, np.int32) if est_flag is None else np.ascontiguousarray(est_flag, np.int32).copy()
        self.is_dynamic = np.zeros(M, np.uint8) if is_dynamic is None else np.ascontiguousarray(is_dynamic, np.uint8).copy()
        self.remove = np.zeros(M, np.uint8)
        self.Ps = np.ascontiguousarray(Ps, np.float64).reshape(NUM_FRAMES, 3)
        self.Rs = np.ascontiguousarray(Rs, np.float64).reshape(NUM_FRAMES, 3, 3)
        self.tic = np.ascontiguousarray(tic, np.float64); self.ric = np.ascontiguousarray(ric, np.float64).reshape(3, 3)
        c = self.c = VrfFmProblem()
        C.memmove(c.Ps, self.Ps.ctypes.data, 8 * 3 * NUM_FRAMES); C.memmove(c.Rs, self.Rs.ctypes.data, 8 * 9 * NUM_FRAMES)
        C.memmove(c.tic, self.tic.ctypes.data, 24); C.memmove(c.ric, self.ric.ctypes.data, 72)
        c.n_landmarks, c.n_obs = M, len(self.obs_depth)
        c.lm_start_frame, c.lm_obs_ptr = self.start.ctypes.data, self.obs_ptr.ctypes.data
        c.obs_pts, c.obs_depth = self.obs_pts.ctypes.data, self.obs_depth.ctypes.data
        c.estimated_depth, c.estimate_flag = self.est_depth.ctypes.data, self.est_flag.ctypes.data
        c.is_dynamic, c.remove = self.is_dynamic.ctypes.data, self.remove.ctypes.data


# every symbol include/vrf.h + include/vrf_ba.h + include/vrf_fm.h declare
EXPORTS = [
    "vrf_config_default", "vrf_create", "vrf_destroy", "vrf_strerror", "vrf_last_cuda_error",
    "vrf_launch_count", "vrf_reset_sequence", "vrf_set_fisheye_mask", "vrf_tracker_read_image", "vrf_tracker_read_image_batch",
    "vrf_tracker_read_rgbd_batch", "vrf_tracker_submit_rgbd_batch", "vrf_tracker_collect_batch",
    "vrf_tracker_enqueue_batch_dev", "vrf_tracker_fetch_batch", "vrf_synchronize", "vrf_stream",
    "vrf_profile_enable", "vrf_profile_read", "vrf_debug_sort_desc", "vrf_debug_read", "vrf_debug_reject_with_f", "vrf_ba_solve", "vrf_ba_solve_batch", "vrf_ba_upload_batch",
    "vrf_ba_enqueue_batch", "vrf_ba_download_batch", "vrf_ba_submit_batch", "vrf_ba_collect_batch",
    "vrf_fm_triangulate_with_depth_batch", "vrf_fm_moving_consistency_check_batch", "vrf_imu_preintegrate_batch",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C vins-rgbd-fast_b200` "
                           "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.vrf_config_default.argtypes = [C.POINTER(VrfConfig)]
    lib.vrf_config_default.restype = None
    lib.vrf_create.argtypes = [C.POINTER(VrfConfig), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.vrf_destroy.argtypes = [C.c_void_p]
    lib.vrf_set_fisheye_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.vrf_destroy.restype = None
    lib.vrf_strerror.argtypes = [C.c_int]
    lib.vrf_strerror.restype = C.c_char_p
    lib.vrf_last_cuda_error.argtypes = [C.c_void_p]
    lib.vrf_last_cuda_error.restype = C.c_char_p
    lib.vrf_launch_count.argtypes = [C.c_void_p]
    lib.vrf_launch_count.restype = C.c_uint64
    lib.vrf_reset_sequence.argtypes = [C.c_void_p, C.c_int]
    lib.vrf_tracker_read_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_double,
                                           C.c_void_p, C.c_int, C.POINTER(VrfTrackOut)]
    lib.vrf_tracker_read_image_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(VrfTrackOut)]
    lib.vrf_tracker_read_rgbd_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                                C.c_void_p, C.c_size_t, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(VrfTrackOut)]
    lib.vrf_tracker_submit_rgbd_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                                  C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vrf_tracker_collect_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VrfTrackOut)]
    lib.vrf_tracker_enqueue_batch_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vrf_tracker_fetch_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VrfTrackOut)]
    lib.vrf_synchronize.argtypes = [C.c_void_p]
    lib.vrf_stream.argtypes = [C.c_void_p]
    lib.vrf_stream.restype = C.c_void_p
    lib.vrf_debug_sort_desc.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.vrf_debug_sort_desc.restype = None
    lib.vrf_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.vrf_profile_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.vrf_debug_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.vrf_debug_read.restype = C.c_long
    lib.vrf_debug_reject_with_f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vrf_ba_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(VrfBaProblem), C.POINTER(VrfBaResult)]
    lib.vrf_ba_solve_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VrfBaProblem), C.POINTER(VrfBaResult)]
    lib.vrf_ba_submit_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VrfBaProblem)]
    lib.vrf_ba_collect_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VrfBaResult)]
    lib.vrf_ba_upload_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VrfBaProblem)]
    lib.vrf_ba_enqueue_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.vrf_ba_download_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(VrfBaResult)]
    lib.vrf_fm_triangulate_with_depth_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(VrfFmProblem)]
    lib.vrf_fm_moving_consistency_check_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(VrfFmProblem)]
    lib.vrf_imu_preintegrate_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(VrfImuSegment), C.POINTER(VrfImuPreint)]
    _lib = lib
    return lib


def default_config(**over):
    cfg = VrfConfig()
    load().vrf_config_default(C.byref(cfg))
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def check(rc, h=None, allow_soft=True):
    if rc < 0 or (rc > 0 and not allow_soft):
        lib = load()
        msg = lib.vrf_strerror(rc).decode()
        if h is not None:
            msg += " | " + lib.vrf_last_cuda_error(h).decode()
        raise RuntimeError(f"vrf error {rc}: {msg}")
    return rc


def sort_desc_perm(cnt):
    c = np.ascontiguousarray(cnt, np.int32)
    out = np.zeros(len(c), np.int32)
    load().vrf_debug_sort_desc(c.ctypes.data, len(c), out.ctypes.data)
    return out


class TrackResult:
    """numpy views of one VrfTrackOut."""

    def __init__(self, ncells, debug=True):
        cap = TRACK_CAP
        self._cur = np.zeros((cap, 2), np.float32)
        self._un = np.zeros((cap, 2), np.float32)
        self._vel = np.zeros((cap, 2), np.float32)
        self._ids = np.zeros(cap, np.int32)
        self._cnt = np.zeros(cap, np.int32)
        self._pred = np.zeros((cap, 2), np.float32)
        self._lk = np.zeros((cap, 2), np.float32)
        self._lkst = np.zeros(cap, np.uint8)
        self._grid = np.zeros(ncells, np.int32)
        self._tex = np.zeros(ncells, np.uint8)
        self._dmm = np.zeros(cap, np.uint16)
        self._dkeep = np.zeros(cap, np.uint8)
        self.debug = debug

    def fill(self, o: VrfTrackOut):
        o.capacity = TRACK_CAP
        o.cur_pts = self._cur.ctypes.data
        o.cur_un_pts = self._un.ctypes.data
        o.pts_velocity = self._vel.ctypes.data
        o.ids = self._ids.ctypes.data
        o.track_cnt = self._cnt.ctypes.data
        o.depth_mm = self._dmm.ctypes.data
        o.depth_keep = self._dkeep.ctypes.data
        if self.debug:
            o.predict_pts = self._pred.ctypes.data
            o.lk_pts = self._lk.ctypes.data
            o.lk_status = self._lkst.ctypes.data
            o.grids_track_num = self._grid.ctypes.data
            o.grids_texture_status = self._tex.ctypes.data
        self._o = o

    def finish(self):
        o = self._o
        n = min(o.n, TRACK_CAP)
        self.n = o.n
        self.cur_pts = self._cur[:n]
        self.cur_un_pts = self._un[:n]
        self.pts_velocity = self._vel[:n]
        self.ids = self._ids[:n]
        self.track_cnt = self._cnt[:n]
        self.depth_mm = self._dmm[:n]
        self.depth_keep = self._dkeep[:n]
        self.n_id = o.n_id
        self.n_predict = o.n_predict
        self.n_unstable = o.n_unstable
        self.status = o.status
        npd = min(o.n_predict, TRACK_CAP)
        self.predict_pts = self._pred[:npd]
        self.lk_pts = self._lk[:npd]
        self.lk_status = self._lkst[:npd]
        self.grids_track_num = self._grid
        self.grids_texture_status = self._tex
        return self


class Handle:
    """RAII wrapper of vrf_handle for `n_seq` sequences on one GPU."""

    def __init__(self, cfg: VrfConfig, n_seq=1, device=0):
        self.lib = load()
        self.cfg = cfg
        self.n_seq = n_seq
        self.ncells = cfg.num_grid_rows * cfg.num_grid_cols
        hp = C.c_void_p()
        rc = self.lib.vrf_create(C.byref(cfg), n_seq, device, C.byref(hp))
        if rc != 0:
            raise RuntimeError(f"vrf_create failed: {rc} {self.lib.vrf_strerror(rc).decode()}")
        self.h = hp

    def close(self):
        if getattr(self, "h", None):
            self.lib.vrf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.vrf_launch_count(self.h))

    def stream(self):
        return self.lib.vrf_stream(self.h)

    def synchronize(self):
        check(self.lib.vrf_synchronize(self.h), self.h)

    def profile(self, on):
        check(self.lib.vrf_profile_enable(self.h, int(on)), self.h)

    def profile_read(self, reset=True):
        names = (C.c_char_p * 16)()
        ms = np.zeros(16, np.float64)
        cnt = np.zeros(16, np.uint64)
        n = self.lib.vrf_profile_read(self.h, 16, C.cast(names, C.c_void_p), ms.ctypes.data, cnt.ctypes.data, int(reset))
        check(n, self.h)
        return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def debug_read(self, what, seq, dtype, count):
        buf = np.zeros(count, dtype)
        rc = self.lib.vrf_debug_read(self.h, what.encode(), seq, buf.ctypes.data, buf.nbytes)
        if rc < 0:
            raise RuntimeError(f"vrf_debug_read({what}) failed: {rc}")
        return buf[: rc // buf.itemsize]

    def pyramid_level(self, seq, level):
        w, hh = self.cfg.col, self.cfg.row
        for _ in range(level):
            w, hh = (w + 1) // 2, (hh + 1) // 2
        return self.debug_read(f"pyr{level}", seq, np.uint8, w * hh).reshape(hh, w)

    def reject_with_f(self, seq, cur_pts, forw_pts):
        """rejectWithF alone (test entry): inlier mask of cv::findFundamentalMat(FM_RANSAC) on the lifted points."""
        a = np.ascontiguousarray(cur_pts, np.float32).reshape(-1, 2)
        b = np.ascontiguousarray(forw_pts, np.float32).reshape(-1, 2)
        st = np.zeros(len(a), np.uint8)
        check(self.lib.vrf_debug_reject_with_f(self.h, seq, len(a), a.ctypes.data, b.ctypes.data, st.ctypes.data), self.h, allow_soft=False)
        return st

    def reset(self, seq):
        check(self.lib.vrf_reset_sequence(self.h, seq), self.h)

    def read_image(self, seq, img, t, R=None, pub=True, debug=True):
        img = np.ascontiguousarray(img)
        fmt = FMT_RGB8 if img.ndim == 3 else FMT_GRAY8
        out = VrfTrackOut()
        res = TrackResult(self.ncells, debug)
        res.fill(out)
        Rp = None
        if R is not None:
            Rm = np.ascontiguousarray(R, np.float64)
            Rp = Rm.ctypes.data
        rc = self.lib.vrf_tracker_read_image(self.h, seq, img.ctypes.data, 0, fmt, float(t), Rp, int(bool(pub)), C.byref(out))
        check(rc, self.h)
        return res.finish()

    def read_rgbd(self, seq, img, depth, t, R=None, pub=True, debug=True):
        """readImage + the depth decode / per-feature depth lookup of the back end's ingest (one sequence)."""
        return self.read_image_batch([seq], [img], [t], None if R is None else [R], [int(bool(pub))], debug=debug,
                                     depths=[depth])[0]

    def read_image_batch(self, seqs, imgs, times, Rs=None, pubs=None, debug=False, depths=None):
        n = len(seqs)
        imgs = [np.ascontiguousarray(im) for im in imgs]
        fmt = FMT_RGB8 if imgs[0].ndim == 3 else FMT_GRAY8
        dfmt, dptr = DEPTH_NONE, None
        if depths is not None:
            depths = [None if dm is None else np.ascontiguousarray(dm) for dm in depths]
            kinds = {dm.dtype for dm in depths if dm is not None}
            if kinds:
                (kind,) = kinds
                dfmt = DEPTH_32FC1 if kind == np.float32 else DEPTH_16UC1
                assert kind in (np.dtype(np.float32), np.dtype(np.uint16))
                dptrs = (C.c_void_p * n)(*[None if dm is None else dm.ctypes.data for dm in depths])
                dptr = C.cast(dptrs, C.c_void_p)
        seq_a = np.asarray(seqs, np.int32)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        t_a = np.asarray(times, np.float64)
        R_a = None if Rs is None else np.ascontiguousarray(np.asarray(Rs, np.float64).reshape(n, 9))
        p_a = None if pubs is None else np.asarray(pubs, np.int32)
        outs = (VrfTrackOut * n)()
        results = [TrackResult(self.ncells, debug) for _ in range(n)]
        for r, o in zip(results, outs):
            r.fill(o)
        rc = self.lib.vrf_tracker_read_rgbd_batch(
            self.h, n, seq_a.ctypes.data, C.cast(ptrs, C.c_void_p), 0, fmt, dptr, 0, dfmt, t_a.ctypes.data,
            None if R_a is None else R_a.ctypes.data, None if p_a is None else p_a.ctypes.data, outs)
        check(rc, self.h)
        return [r.finish() for r in results]

    # ---- back end ----
    def ba_solve_batch(self, seqs, problems, want_prior=True):
        """problems: list of vrf_b200.ba_problem.BaProblem (finalized). Returns list of BaSolution.
        want_prior=False passes new_prior = NULL: the new prior stays on the device in information form and the
        eigen-decomposition that produces (linearized_jacobians, linearized_residuals) is not run."""
        from .ba_problem import BaSolution
        n = len(seqs)
        seq_a = np.asarray(seqs, np.int32)
        probs = (VrfBaProblem * n)()
        sols = [BaSolution(pb.M) for pb in problems]
        res = (VrfBaResult * n)()
        for i, pb in enumerate(problems):
            C.memmove(C.byref(probs[i]), C.byref(pb.c), C.sizeof(VrfBaProblem))
            if not want_prior:
                sols[i].c.new_prior = None
            C.memmove(C.byref(res[i]), C.byref(sols[i].c), C.sizeof(VrfBaResult))
        rc = self.lib.vrf_ba_solve_batch(self.h, n, seq_a.ctypes.data, probs, res)
        check(rc, self.h)
        for i, s in enumerate(sols):
            C.memmove(C.byref(s.c), C.byref(res[i]), C.sizeof(VrfBaResult))
            s.rc = res[i].status
        return sols

    def ba_solve(self, seq, problem, want_prior=True):
        return self.ba_solve_batch([seq], [problem], want_prior)[0]

    def ba_upload(self, seqs, problems):
        n = len(seqs)
        seq_a = np.asarray(seqs, np.int32)
        probs = (VrfBaProblem * n)()
        for i, pb in enumerate(problems):
            C.memmove(C.byref(probs[i]), C.byref(pb.c), C.sizeof(VrfBaProblem))
        check(self.lib.vrf_ba_upload_batch(self.h, n, seq_a.ctypes.data, probs), self.h, allow_soft=False)

    def ba_enqueue(self, seqs):
        seq_a = np.asarray(seqs, np.int32)
        check(self.lib.vrf_ba_enqueue_batch(self.h, len(seqs), seq_a.ctypes.data), self.h, allow_soft=False)

    def ba_download(self, seqs):
        seq_a = np.asarray(seqs, np.int32)
        check(self.lib.vrf_ba_download_batch(self.h, len(seqs), seq_a.ctypes.data, None), self.h)

    def set_fisheye_mask(self, mask):
        m = np.ascontiguousarray(mask, np.uint8)
        check(self.lib.vrf_set_fisheye_mask(self.h, m.ctypes.data, m.shape[1]), self.h, allow_soft=False)

    # ---- steps either side of optimization() (include/vrf_fm.h) ----
    def _fm_call(self, fn, problems):
        arr = (VrfFmProblem * len(problems))()
        for i, pb in enumerate(problems):
            C.memmove(C.byref(arr[i]), C.byref(pb.c), C.sizeof(VrfFmProblem))
        check(fn(self.h, len(problems), arr), self.h, allow_soft=False)

    def fm_triangulate_with_depth(self, problems):
        """FeatureManager::triangulateWithDepth on a list of FmProblem (results land in .est_depth / .est_flag)."""
        self._fm_call(self.lib.vrf_fm_triangulate_with_depth_batch, problems)

    def fm_moving_consistency_check(self, problems):
        """Estimator::movingConsistencyCheck (results in .is_dynamic / .remove)."""
        self._fm_call(self.lib.vrf_fm_moving_consistency_check_batch, problems)

    def imu_preintegrate(self, segments):
        """segments: list of (acc0, gyr0, ba, bg, dt[n], acc[n,3], gyr[n,3]).  Returns a ctypes array of VrfImuPreint."""
        n = len(segments)
        segs = (VrfImuSegment * n)()
        keep = []
        for i, (a0, g0, ba, bg, dt, acc, gyr) in enumerate(segments):
            dt = np.ascontiguousarray(dt, np.float64); acc = np.ascontiguousarray(acc, np.float64).reshape(-1, 3)
            gyr = np.ascontiguousarray(gyr, np.float64).reshape(-1, 3)
            keep += [dt, acc, gyr]
            for k in range(3):
                segs[i].acc_0[k], segs[i].gyr_0[k], segs[i].linearized_ba[k], segs[i].linearized_bg[k] = a0[k], g0[k], ba[k], bg[k]
            segs[i].n_samples = len(dt)
            segs[i].dt, segs[i].acc, segs[i].gyr = dt.ctypes.data, acc.ctypes.data, gyr.ctypes.data
        out = (VrfImuPreint * n)()
        check(self.lib.vrf_imu_preintegrate_batch(self.h, n, segs, out), self.h, allow_soft=False)
        return out

    # ---- preallocated batch calls (no per-call python allocations; used by bench.py's e2e arm) ----
    def make_track_batch(self, n, debug=False):
        outs = (VrfTrackOut * n)()
        results = [TrackResult(self.ncells, debug) for _ in range(n)]
        for r, o in zip(results, outs):
            r.fill(o)
        return outs, results

    def read_image_batch_into(self, seq_a, ptrs, fmt, t_a, R_a, p_a, outs, dptrs=None, dfmt=DEPTH_NONE):
        """seq_a/t_a/R_a/p_a: contiguous numpy arrays, ptrs / dptrs: (c_void_p * n) of host frame / depth-frame pointers."""
        rc = self.lib.vrf_tracker_read_rgbd_batch(self.h, len(seq_a), seq_a.ctypes.data, C.cast(ptrs, C.c_void_p), 0, fmt,
                                                  None if dptrs is None else C.cast(dptrs, C.c_void_p), 0, dfmt,
                                                  t_a.ctypes.data, R_a.ctypes.data, p_a.ctypes.data, outs)
        return check(rc, self.h)

    def submit_batch_into(self, seq_a, ptrs, fmt, t_a, R_a, p_a, dptrs=None, dfmt=DEPTH_NONE):
        """Pipelined host-frame call (vrf_tracker_submit_rgbd_batch); pair with collect_batch_into()."""
        rc = self.lib.vrf_tracker_submit_rgbd_batch(self.h, len(seq_a), seq_a.ctypes.data, C.cast(ptrs, C.c_void_p), 0, fmt,
                                                    None if dptrs is None else C.cast(dptrs, C.c_void_p), 0, dfmt,
                                                    t_a.ctypes.data, R_a.ctypes.data, p_a.ctypes.data)
        return check(rc, self.h, allow_soft=False)

    def collect_batch_into(self, seq_a, outs):
        return check(self.lib.vrf_tracker_collect_batch(self.h, len(seq_a), seq_a.ctypes.data, outs), self.h)

    def make_ba_batch(self, problems):
        from .ba_problem import BaSolution
        n = len(problems)
        probs = (VrfBaProblem * n)()
        res = (VrfBaResult * n)()
        sols = [BaSolution(pb.M) for pb in problems]
        for i, pb in enumerate(problems):
            C.memmove(C.byref(probs[i]), C.byref(pb.c), C.sizeof(VrfBaProblem))
            C.memmove(C.byref(res[i]), C.byref(sols[i].c), C.sizeof(VrfBaResult))
        return probs, res, sols

    def ba_solve_batch_into(self, seq_a, probs, res):
        return check(self.lib.vrf_ba_solve_batch(self.h, len(seq_a), seq_a.ctypes.data, probs, res), self.h)

    def ba_submit_into(self, seq_a, probs):
        return check(self.lib.vrf_ba_submit_batch(self.h, len(seq_a), seq_a.ctypes.data, probs), self.h, allow_soft=False)

    def ba_collect_into(self, seq_a, res):
        return check(self.lib.vrf_ba_collect_batch(self.h, len(seq_a), seq_a.ctypes.data, res), self.h)

    def enqueue_dev(self, seqs, d_ptr, fmt, times, Rs=None, pubs=None, d_depth=None, depth_fmt=DEPTH_16UC1):
        n = len(seqs)
        seq_a = np.asarray(seqs, np.int32)
        t_a = np.asarray(times, np.float64)
        R_a = None if Rs is None else np.ascontiguousarray(np.asarray(Rs, np.float64).reshape(n, 9))
        p_a = None if pubs is None else np.asarray(pubs, np.int32)
        rc = self.lib.vrf_tracker_enqueue_batch_dev(
            self.h, n, seq_a.ctypes.data, d_ptr, fmt, d_depth, depth_fmt if d_depth else DEPTH_NONE, t_a.ctypes.data,
            None if R_a is None else R_a.ctypes.data, None if p_a is None else p_a.ctypes.data)
        check(rc, self.h)

    def fetch(self, seqs, debug=False):
        n = len(seqs)
        seq_a = np.asarray(seqs, np.int32)
        outs = (VrfTrackOut * n)()
        results = [TrackResult(self.ncells, debug) for _ in range(n)]
        for r, o in zip(results, outs):
            r.fill(o)
        rc = self.lib.vrf_tracker_fetch_batch(self.h, n, seq_a.ctypes.data, outs)
        check(rc, self.h)
        return [r.finish() for r in results]
