"""CPU oracle for the front end: a restatement of FeatureTracker (control flow
only) on top of the REAL OpenCV 4.13 (`cv2`).  TEST INFRASTRUCTURE ONLY.

Follows vins_estimator/src/feature_tracker/feature_tracker.{h,cpp} and
camera_model/src/camera_models/PinholeCamera.cc of the reference:

  readImage               feature_tracker.cpp:263-439
  initGridsDetector       :33-94
  inBorder                :96-103
  gridDetect              :105-171
  setMask                 :173-208
  addPoints(KeyPoint&)    :220-233
  rejectWithF             :441-473
  updateID                :485-495 (+ loop estimator_nodelet.cpp:324-330)
  undistortedPoints       :542-593
  predictPtsInNextFrame   :595-608
  liftProjective          PinholeCamera.cc:450-510
  spaceToPlane            PinholeCamera.cc:520-543
  distortion              PinholeCamera.cc:646-663

All image arithmetic (pyramids, LK, FAST, RANSAC, circles) is executed by
OpenCV itself through cv2 -- the same library the reference links -- so this
oracle is the reference's arithmetic with only the glue restated.

Deterministic-semantics decisions (documented in DESIGN.md):
  * the reference races per-cell FAST (reading `mask`) against addPoints
    (drawing circles into `mask`); the oracle defines the deterministic
    semantics "every cell sees the mask snapshot taken right after setMask()"
    (feature_tracker.cpp:397-409 with all futures finishing before the first
    addPoints) -- what the reference computes whenever the worker threads win
    the race.
  * std::sort tie order: taken from the real libstdc++ std::sort through
    oracle/stdsort.cpp.
  * PUB_THIS_FRAME (a global in the reference) is an explicit argument.
  * LK maxLevel: reference hard-codes 1 (IMU) / 3 (no IMU); `lk_max_level`
    overrides it when the benchmark config asks for 3/4-level pyramids
    (SURVEY.md section 8d).
"""
import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import cv2
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SORT_SO = os.path.join(_HERE, "_build", "libstdsort.so")
_sort_lib = None


def build_stdsort():
    os.makedirs(os.path.dirname(_SORT_SO), exist_ok=True)
    src = os.path.join(_HERE, "stdsort.cpp")
    if (not os.path.exists(_SORT_SO)) or os.path.getmtime(_SORT_SO) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", _SORT_SO, src])
    return _SORT_SO


def stdsort_desc_perm(cnt):
    """Permutation produced by the reference's std::sort (descending count)."""
    global _sort_lib
    if _sort_lib is None:
        _sort_lib = ctypes.CDLL(build_stdsort())
        _sort_lib.oracle_stdsort_desc_by_cnt.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    c = np.ascontiguousarray(cnt, dtype=np.int32)
    out = np.zeros(len(c), np.int32)
    if len(c):
        _sort_lib.oracle_stdsort_desc_by_cnt(c.ctypes.data, len(c), out.ctypes.data)
    return out


def cv_round(v):
    """cvRound: round half to even (SSE cvtss2si) -- what Point2f->Point does."""
    return int(np.rint(v))


@dataclass
class FrontendConfig:
    row: int = 480
    col: int = 640
    max_cnt: int = 150
    min_dist: int = 25
    f_threshold: float = 1.0
    num_grid_rows: int = 7
    num_grid_cols: int = 8
    use_imu: int = 1
    focal_length: float = 460.0     # parameters.h:11
    fx: float = 600.0
    fy: float = 600.0
    cx: float = 320.0
    cy: float = 240.0
    k1: float = 0.1
    k2: float = -0.2
    p1: float = 1e-3
    p2: float = 1e-3
    lk_max_level: int = -1          # -1 => reference default (1 with IMU, 3 without)
    use_ransac: int = 1             # 0 skips rejectWithF (kernel bring-up only)
    equalize: int = 0               # EQUALIZE: CLAHE on every incoming frame (feature_tracker.cpp:269-275)
    fisheye: int = 0                # FISHEYE: setMask starts from fisheye_mask (feature_tracker.cpp:175-178)


class PinholeCamera:
    """PinholeCamera.cc:450-543,646-663 (the three functions on the path)."""

    def __init__(self, cfg):
        self.fx, self.fy, self.cx, self.cy = cfg.fx, cfg.fy, cfg.cx, cfg.cy
        self.k1, self.k2, self.p1, self.p2 = cfg.k1, cfg.k2, cfg.p1, cfg.p2
        self.inv_K11 = 1.0 / cfg.fx
        self.inv_K13 = -cfg.cx / cfg.fx
        self.inv_K22 = 1.0 / cfg.fy
        self.inv_K23 = -cfg.cy / cfg.fy
        self.no_distortion = (cfg.k1 == 0.0 and cfg.k2 == 0.0 and cfg.p1 == 0.0 and cfg.p2 == 0.0)

    def distortion(self, x, y):
        k1, k2, p1, p2 = self.k1, self.k2, self.p1, self.p2
        mx2 = x * x
        my2 = y * y
        mxy = x * y
        rho2 = mx2 + my2
        rad = k1 * rho2 + k2 * rho2 * rho2
        return (x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2),
                y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2))

    def lift_projective(self, u, v):
        mx_d = self.inv_K11 * u + self.inv_K13
        my_d = self.inv_K22 * v + self.inv_K23
        if self.no_distortion:
            return mx_d, my_d, 1.0
        dx, dy = self.distortion(mx_d, my_d)
        mx_u = mx_d - dx
        my_u = my_d - dy
        for _ in range(1, 8):
            dx, dy = self.distortion(mx_u, my_u)
            mx_u = mx_d - dx
            my_u = my_d - dy
        return mx_u, my_u, 1.0

    def space_to_plane(self, X, Y, Z):
        x = X / Z
        y = Y / Z
        if not self.no_distortion:
            dx, dy = self.distortion(x, y)
            x = x + dx
            y = y + dy
        return self.fx * x + self.cx, self.fy * y + self.cy

    # The three functions above are written with elementwise float64 arithmetic only, so they accept numpy
    # arrays as well as Python floats and give bit-identical results either way (IEEE double, no contraction);
    # the tracker below calls them on whole point arrays so that the timed CPU baseline is OpenCV + arithmetic,
    # not interpreter loops.  tests/test_oracle_frontend.py pins array == per-point evaluation.
    def lift_projective_pts(self, pts):
        """liftProjective of an (n, 2) float32 point array -> (x, y) float64 arrays (z == 1)."""
        p = np.asarray(pts, np.float32).reshape(-1, 2).astype(np.float64)
        x, y, _ = self.lift_projective(p[:, 0], p[:, 1])
        return x, y


class FeatureTrackerRef:
    def __init__(self, cfg: FrontendConfig):
        self.cfg = cfg
        self.cam = PinholeCamera(cfg)
        self.fast = cv2.FastFeatureDetector_create()   # feature_tracker.cpp:29
        self.n_id = 0
        self.cur_img = None
        self.forw_img = None
        self.cur_pts = np.zeros((0, 2), np.float32)
        self.forw_pts = np.zeros((0, 2), np.float32)
        self.predict_pts = np.zeros((0, 2), np.float32)
        self.unstable_pts = np.zeros((0, 2), np.float32)
        self.cur_un_pts = np.zeros((0, 2), np.float32)
        self.prev_un_pts = np.zeros((0, 2), np.float32)
        self.pts_velocity = np.zeros((0, 2), np.float32)
        self.ids = []
        self.track_cnt = []
        self.cur_un_pts_map = None       # (sorted ids, points): the std::map<int, Point2f> of the reference
        self.prev_un_pts_map = None
        self.cur_time = 0.0
        self.prev_time = 0.0
        self.mask = None
        self.fisheye_mask = None         # set by the caller when cfg.fisheye (Estimator::setParameter, estimator.cpp:31)
        self.last_status = None
        self.last_ransac_status = None
        self.last_new_keypoints = []
        self.init_grids_detector()

    # feature_tracker.cpp:33-94
    def init_grids_detector(self):
        c = self.cfg
        ROW, COL = float(c.row), float(c.col)
        R, C = c.num_grid_rows, c.num_grid_cols
        gh = int(ROW / R)
        gw = int(COL / C)
        grh = int(ROW - (R - 1) * gh)
        grw = int(COL - (C - 1) * gw)
        self.grid_height, self.grid_width = gh, gw
        rects = []
        for i in range(R):
            for j in range(C):
                if j == 0:
                    x, w = 0, gw + 3
                elif j < C - 1:
                    x, w = j * gw - 3, gw + 6
                else:
                    x, w = j * gw - 3, grw + 3
                if i == 0:
                    y, h = 0, gh + 3
                elif i < R - 1:
                    y, h = i * gh - 3, gh + 6
                else:
                    y, h = i * gh - 3, grh + 3
                rects.append((x, y, w, h))
        self.grids_rect = rects
        self.grids_track_num = [0] * len(rects)
        self.grids_texture_status = [True] * len(rects)
        self.grids_threshold = int(c.max_cnt / len(rects))
        assert self.grids_threshold > 0

    # feature_tracker.cpp:96-103
    def in_border(self, pt):
        x = cv_round(pt[0])
        y = cv_round(pt[1])
        return 1 <= x < self.cfg.col - 1 and 1 <= y < self.cfg.row - 1

    # feature_tracker.cpp:595-608
    def predict_pts_in_next_frame(self, R):
        x, y = self.cam.lift_projective_pts(self.cur_pts)
        xyz = np.stack([x, y, np.ones_like(x)], 1)
        # batched (3,3)@(3,1): the same product kernel as the per-point `R @ [x, y, z]` (pinned in the tests)
        P = np.matmul(np.asarray(R, np.float64)[None], xyz[:, :, None])[:, :, 0] if len(x) else np.zeros((0, 3))
        u, v = self.cam.space_to_plane(P[:, 0], P[:, 1], P[:, 2])
        self.predict_pts = np.stack([u, v], 1).astype(np.float32)

    def in_border_pts(self, pts):
        """inBorder for an (n, 2) array (cvRound = rint, half to even)."""
        x = np.rint(pts[:, 0]).astype(np.int64)
        y = np.rint(pts[:, 1]).astype(np.int64)
        return (1 <= x) & (x < self.cfg.col - 1) & (1 <= y) & (y < self.cfg.row - 1)

    # feature_tracker.cpp:105-171 (deterministic mask-snapshot semantics)
    def grid_detect(self, grid_id, mask_snapshot):
        x, y, w, h = self.grids_rect[grid_id]
        kps = self.fast.detect(self.forw_img[y:y + h, x:x + w], mask_snapshot[y:y + h, x:x + w])
        kps = [(k.pt[0], k.pt[1], k.response) for k in kps]
        if not kps:
            self.grids_texture_status[grid_id] = False
            return []
        num_to_add = self.grids_threshold - self.grids_track_num[grid_id] + 2
        if len(kps) <= num_to_add:
            return [(np.float32(px + x), np.float32(py + y), r) for (px, py, r) in kps]
        slots = [None] * num_to_add
        K = num_to_add
        min_id = 0
        for j, (px, py, r) in enumerate(kps):
            if num_to_add > 0:
                slots[j] = (np.float32(px + x), np.float32(py + y), r)
                num_to_add -= 1
                if r < slots[min_id][2]:
                    min_id = j
            elif r > slots[min_id][2]:
                slots[min_id] = (np.float32(px + x), np.float32(py + y), r)
                for k in range(K):
                    if slots[k][2] < slots[min_id][2]:
                        min_id = k
        return slots

    # feature_tracker.cpp:173-208
    def set_mask(self):
        c = self.cfg
        if c.fisheye:                                       # feature_tracker.cpp:175-178
            self.mask = self.fisheye_mask.copy()
        else:
            self.mask = np.full((c.row, c.col), 255, np.uint8)
        perm = stdsort_desc_perm(self.track_cnt)
        pts, ids, cnt = [], [], []
        for k in perm:
            p = self.forw_pts[k]
            px, py = cv_round(p[0]), cv_round(p[1])
            if self.mask[py, px] == 255:
                pts.append(p)
                ids.append(self.ids[k])
                cnt.append(self.track_cnt[k])
                cv2.circle(self.mask, (px, py), c.min_dist, 0, -1)
        self.forw_pts = np.array(pts, np.float32).reshape(-1, 2)
        self.ids = ids
        self.track_cnt = cnt
        for p in self.unstable_pts:
            cv2.circle(self.mask, (cv_round(p[0]), cv_round(p[1])), c.min_dist, 0, -1)

    # feature_tracker.cpp:220-233
    def add_points(self, kps):
        c = self.cfg
        add = []
        for (px, py, r) in kps:
            ix, iy = cv_round(px), cv_round(py)
            if self.mask[iy, ix] == 255:
                add.append((px, py))
                self.ids.append(-1)
                self.track_cnt.append(1)
                cv2.circle(self.mask, (ix, iy), c.min_dist, 0, -1)
        if add:
            self.forw_pts = np.concatenate([self.forw_pts.reshape(-1, 2), np.array(add, np.float32)], 0)

    def _reduce(self, status):
        keep = np.asarray(status).astype(bool)
        self.cur_pts = self.cur_pts[keep]
        self.forw_pts = self.forw_pts[keep]
        self.ids = [v for v, k in zip(self.ids, keep) if k]
        self.cur_un_pts = self.cur_un_pts[keep]
        self.track_cnt = [v for v, k in zip(self.track_cnt, keep) if k]

    # feature_tracker.cpp:441-473
    def reject_with_f(self):
        c = self.cfg
        self.last_ransac_status = None
        if len(self.forw_pts) >= 8 and c.use_ransac:
            n = len(self.cur_pts)
            x, y = self.cam.lift_projective_pts(self.cur_pts)
            un_cur = np.stack([c.focal_length * x / 1.0 + c.col / 2.0, c.focal_length * y / 1.0 + c.row / 2.0], 1).astype(np.float32)
            x, y = self.cam.lift_projective_pts(self.forw_pts)
            un_forw = np.stack([c.focal_length * x / 1.0 + c.col / 2.0, c.focal_length * y / 1.0 + c.row / 2.0], 1).astype(np.float32)
            self.last_un_cur, self.last_un_forw = un_cur, un_forw
            _, status = cv2.findFundamentalMat(un_cur, un_forw, cv2.FM_RANSAC, c.f_threshold, 0.99)
            if status is None:
                status = np.zeros((n, 1), np.uint8)   # C++: empty status => reduceVector drops all
            status = status.ravel()
            self.last_ransac_status = status.copy()
            self._reduce(status)

    # feature_tracker.cpp:542-593
    def undistorted_points(self):
        n = len(self.cur_pts)
        x, y = self.cam.lift_projective_pts(self.cur_pts)
        self.cur_un_pts = np.stack([x / 1.0, y / 1.0], 1).astype(np.float32)      # (x / z, y / z) with z == 1
        ids = np.asarray(self.ids, np.int64).reshape(-1)
        # cur_un_pts_map: std::map::insert keeps the first entry of an id (only -1 can repeat)
        uniq, first = np.unique(ids, return_index=True) if n else (np.zeros(0, np.int64), np.zeros(0, np.int64))
        vel = np.zeros((n, 2), np.float32)
        if self.prev_un_pts_map is not None and len(self.prev_un_pts_map[0]):
            dt = self.cur_time - self.prev_time
            pids, ppts = self.prev_un_pts_map
            pos = np.searchsorted(pids, ids)
            pos_c = np.minimum(pos, len(pids) - 1)
            hit = (ids != -1) & (pids[pos_c] == ids)
            d = (self.cur_un_pts[hit] - ppts[pos_c[hit]]).astype(np.float32)
            vel[hit] = (d.astype(np.float64) / dt).astype(np.float32)
        self.pts_velocity = vel
        self.cur_un_pts_map = (uniq, self.cur_un_pts[first])                      # sorted ids -> point
        self.prev_un_pts_map = self.cur_un_pts_map

    # feature_tracker.cpp:485-495 + estimator_nodelet.cpp:324-330
    def update_ids(self):
        for i in range(len(self.ids)):
            if self.ids[i] == -1:
                self.ids[i] = self.n_id
                self.n_id += 1

    # feature_tracker.cpp:263-439
    def read_image(self, img, cur_time, relative_R=None, pub_this_frame=True):
        c = self.cfg
        if relative_R is None:
            relative_R = np.eye(3)
        self.cur_time = cur_time
        img = np.ascontiguousarray(img)
        if c.equalize:                                      # feature_tracker.cpp:269-275 (the real cv::CLAHE)
            img = cv2.createCLAHE(3.0, (8, 8)).apply(img)
        if self.forw_img is None:
            self.cur_img = self.forw_img = img
        else:
            self.forw_img = img
        self.forw_pts = np.zeros((0, 2), np.float32)
        self.unstable_pts = np.zeros((0, 2), np.float32)
        self.last_status = None
        self.last_lk_pts = None
        if len(self.cur_pts):
            crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
            if c.use_imu:
                self.predict_pts_in_next_frame(relative_R)
                ml = 1 if c.lk_max_level < 0 else c.lk_max_level
                fp, status, _ = cv2.calcOpticalFlowPyrLK(
                    self.cur_img, self.forw_img, self.cur_pts, self.predict_pts.copy(),
                    winSize=(21, 21), maxLevel=ml, criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
            else:
                ml = 3 if c.lk_max_level < 0 else c.lk_max_level
                fp, status, _ = cv2.calcOpticalFlowPyrLK(
                    self.cur_img, self.forw_img, self.cur_pts, None, winSize=(21, 21), maxLevel=ml)
            self.forw_pts = fp.reshape(-1, 2).astype(np.float32)
            status = status.ravel().copy()
            self.last_lk_pts = self.forw_pts.copy()
            self.last_lk_status = status.copy()
            ib = self.in_border_pts(self.forw_pts)
            ok = status.astype(bool)
            self.unstable_pts = self.forw_pts[~ok & ib].astype(np.float32).reshape(-1, 2)
            status[ok & ~ib] = 0
            self.last_status = status.copy()
            self._reduce(status)
        self.track_cnt = [n + 1 for n in self.track_cnt]
        self.last_new_keypoints = []
        self.last_grids_id = []
        if pub_this_frame:
            self.reject_with_f()
            self.set_mask()
            n_max_cnt = c.max_cnt - len(self.forw_pts)
            if n_max_cnt > 0:
                R, C = c.num_grid_rows, c.num_grid_cols
                fp = self.forw_pts.reshape(-1, 2)
                col_id = fp[:, 0].astype(np.int64) // self.grid_width        # (int)p.x truncates; coordinates are > 0 here
                row_id = fp[:, 1].astype(np.int64) // self.grid_height
                col_id[col_id == C] -= 1
                row_id[row_id == R] -= 1
                self.grids_track_num = np.bincount(col_id + C * row_id, minlength=R * C).astype(int).tolist()
                grids_id = []
                for i in range(R * C):
                    if self.grids_track_num[i] < self.grids_threshold and self.grids_texture_status[i]:
                        grids_id.append(i)
                    else:
                        self.grids_texture_status[i] = True
                self.last_grids_id = grids_id
                snapshot = self.mask.copy()
                per_cell = [self.grid_detect(g, snapshot) for g in grids_id]
                for kps in per_cell:
                    self.last_new_keypoints.append(kps)
                    self.add_points(kps)
        self.prev_un_pts = self.cur_un_pts
        self.cur_img = self.forw_img
        self.cur_pts = self.forw_pts
        self.undistorted_points()
        self.prev_time = self.cur_time
        # nodelet: update all ids right after readImage (estimator_nodelet.cpp:324-330)
        self.update_ids()

    def feature_map(self):
        """estimator_nodelet.cpp:336-363: {id: [x,y,1,u,v,vx,vy]} for track_cnt>1."""
        out = {}
        for j, fid in enumerate(self.ids):
            if self.track_cnt[j] > 1:
                out[fid] = np.array([self.cur_un_pts[j, 0], self.cur_un_pts[j, 1], 1.0,
                                     self.cur_pts[j, 0], self.cur_pts[j, 1],
                                     self.pts_velocity[j, 0], self.pts_velocity[j, 1]], np.float64)
        return out


# ---------------------------------------------------------------------------
# Depth ingest: the step right after the front end.
# ---------------------------------------------------------------------------
def decode_depth(depth_msg, rows, cols):
    """estimator_nodelet.cpp:512-533: no message -> zeros; 16UC1/mono16 -> shared as is;
    32FC1 -> depth_32fc1.convertTo(depth_img, CV_16UC1, 1000).

    cv2's Python API has no Mat::convertTo binding.  OpenCV's 32F->16U cvtScale works in float32:
    saturate_cast<ushort>(cvRound(src * 1000.f)) with round-half-even, x86 out-of-int-range / NaN -> INT_MIN -> 0.
    The same float pipeline is reachable as cv2.addWeighted(src, 1000, src, 0, 0, dtype=CV_16U), used here so that
    the arithmetic is still executed by the real OpenCV; tests/test_oracle_frontend.py pins the numpy formula
    against it (incl. NaN / inf / negative / > 65.535 m values)."""
    if depth_msg is None:
        return np.zeros((rows, cols), np.uint16)
    if depth_msg.dtype == np.uint16:
        return depth_msg
    if depth_msg.dtype == np.float32:
        return cv2.addWeighted(depth_msg, 1000.0, depth_msg, 0.0, 0.0, dtype=cv2.CV_16U)
    raise ValueError("Unknown depth encoding!")


def decode_depth_numpy(depth_32f):
    """Inspectable spec of the 32FC1 branch of decode_depth (what the kernel implements per looked-up pixel)."""
    t = depth_32f.astype(np.float32) * np.float32(1000.0)
    with np.errstate(invalid="ignore"):
        r = np.where((t >= np.float32(-2147483648.0)) & (t < np.float32(2147483648.0)), np.rint(t), -2147483648.0)
    return np.clip(r, 0, 65535).astype(np.uint16)


def depth_lookup(depth_img, cur_pts, depth_min_dist):
    """FeatureManager::addFeatureCheckParallax, feature_manager.cpp:71-80, for every feature (u, v) = cur_pts[j]
    (the map entries 3 and 4, estimator_nodelet.cpp:344-352):
        pt_depth_mm = depth_img.at<unsigned short>((int)v, (int)u);  pt_depth_m = pt_depth_mm / 1000.0
        erase the feature iff 0 < pt_depth_m < DEPTH_MIN_DIST
    Returns (depth_mm u16 [n], keep u8 [n])."""
    p = np.asarray(cur_pts, np.float32).reshape(-1, 2)
    mm = depth_img[p[:, 1].astype(np.int64), p[:, 0].astype(np.int64)].astype(np.uint16)
    d_m = mm.astype(np.float64) / 1000.0
    keep = np.where((0 < d_m) & (d_m < depth_min_dist), 0, 1).astype(np.uint8)
    return mm, keep
