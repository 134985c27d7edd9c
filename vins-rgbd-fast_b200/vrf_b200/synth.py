"""Seeded synthetic RGB-D + IMU sequences (harness utility, not product code).

There are no datasets in the reference repo and no
network, so tests and bench.py render their own sequences:

* a band-limited noise texture (uniform u8 at 1/4 resolution, bicubic
  upsample, Gaussian sigma=1) painted on a smooth height field z = H(x, y)
  (depth 1.5 .. 4.0 m) in front of the camera;
* a smooth 6-DoF trajectory (sum of sinusoids; <= ~0.5 deg/frame rotation,
  <= ~2 cm/frame translation at 30 Hz);
* per frame: RGB8 (H x W x 3), gray (cv::cvtColor RGB2GRAY fixed point, what
  cv_bridge MONO8 conversion produces, estimator_nodelet.cpp:292-307),
  depth 16UC1 in mm with 5 % invalid (0) pixels;
* a pinhole + radial-tangential camera (PinholeCamera.cc model);
* 200 Hz IMU samples (gyro/acc + white noise + constant bias), world frame
  z-up with G = (0, 0, 9.81) (parameters.cpp:13).

Everything is a pure function of (seed, config) so CPU oracle and GPU path
consume identical inputs.
"""
from dataclasses import dataclass

import cv2
import numpy as np

# camera axes (x right, y down, z forward) -> world axes (x forward, y left, z up)
R_W_S = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
G_WORLD = np.array([0.0, 0.0, 9.81])


@dataclass
class CamModel:
    fx: float = 600.0
    fy: float = 600.0
    cx: float = 320.0
    cy: float = 240.0
    k1: float = 0.1
    k2: float = -0.2
    p1: float = 1e-3
    p2: float = 1e-3
    width: int = 640
    height: int = 480


def band_limited_texture(h, w, seed, channels=1):
    r = np.random.default_rng(seed)
    chans = []
    for _ in range(channels):
        small = r.integers(0, 256, (h // 4 + 2, w // 4 + 2)).astype(np.uint8)
        big = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
        chans.append(cv2.GaussianBlur(big, (0, 0), 1.0))
    return chans[0] if channels == 1 else np.stack(chans, -1)


def rgb_to_gray(rgb):
    """cv::cvtColor(COLOR_RGB2GRAY) for 8U (OpenCV 4.x, 15-bit fixed point):
    (R*9798 + G*19235 + B*3735 + 16384) >> 15  -- pinned against cv2 4.13 in tests."""
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    return ((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)


def so3_exp(v):
    th = np.linalg.norm(v)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / (th * th) * (K @ K)


def so3_log(R):
    c = (np.trace(R) - 1) * 0.5
    c = min(1.0, max(-1.0, c))
    th = np.arccos(c)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-9:
        return 0.5 * w
    return th / (2 * np.sin(th)) * w


class Trajectory:
    """Smooth 6-DoF motion expressed in the scene frame S (camera nominal frame)."""

    def __init__(self, seed, fps=30.0, trans_per_frame=0.02, rot_deg_per_frame=0.5):
        r = np.random.default_rng(seed + 7919)
        self.fps = fps
        vmax = trans_per_frame * fps          # m/s
        wmax = np.deg2rad(rot_deg_per_frame) * fps
        self.w_p = r.uniform(0.6, 1.6, (3, 2))
        self.ph_p = r.uniform(0, 2 * np.pi, (3, 2))
        # amplitude so that |v| <= vmax/sqrt(3) per axis (two harmonics)
        self.A_p = (vmax / np.sqrt(3.0)) / (2.0 * self.w_p) * r.uniform(0.5, 1.0, (3, 2))
        self.A_p[2] *= 0.5
        self.w_r = r.uniform(0.5, 1.4, (3, 2))
        self.ph_r = r.uniform(0, 2 * np.pi, (3, 2))
        self.A_r = (wmax / np.sqrt(3.0)) / (2.0 * self.w_r) * r.uniform(0.5, 1.0, (3, 2))

    def pos_s(self, t):
        return (self.A_p * np.sin(self.w_p * t + self.ph_p)).sum(1) - (self.A_p * np.sin(self.ph_p)).sum(1)

    def vel_s(self, t):
        return (self.A_p * self.w_p * np.cos(self.w_p * t + self.ph_p)).sum(1)

    def acc_s(self, t):
        return (-self.A_p * self.w_p ** 2 * np.sin(self.w_p * t + self.ph_p)).sum(1)

    def rot_s(self, t):
        th = (self.A_r * np.sin(self.w_r * t + self.ph_r)).sum(1) - (self.A_r * np.sin(self.ph_r)).sum(1)
        return so3_exp(th)

    # world-frame (z-up) body pose; body == camera (RIC = I, TIC = 0)
    def R_wb(self, t):
        return R_W_S @ self.rot_s(t)

    def p_w(self, t):
        return R_W_S @ self.pos_s(t)

    def v_w(self, t):
        return R_W_S @ self.vel_s(t)

    def a_w(self, t):
        return R_W_S @ self.acc_s(t)

    def gyro_body(self, t, h=1e-4):
        return so3_log(self.rot_s(t - h).T @ self.rot_s(t + h)) / (2 * h)

    def acc_body(self, t):
        return self.R_wb(t).T @ (self.a_w(t) + G_WORLD)

    def relative_R(self, t0, t1):
        """cam(t0) -> cam(t1): what Estimator::predictMotion hands to readImage
        (estimator.cpp:1790-1860) when the gyro integration is exact."""
        return self.rot_s(t1).T @ self.rot_s(t0)


class Scene:
    def __init__(self, seed, cam: CamModel, texel=0.004, extent=(9.0, 7.0)):
        self.cam = cam
        self.texel = texel
        self.ex, self.ey = extent
        tw, th = int(self.ex / texel), int(self.ey / texel)
        self.tex = band_limited_texture(th, tw, seed, channels=3)
        r = np.random.default_rng(seed + 104729)
        ctrl = r.uniform(1.5, 4.0, (int(self.ey / 2.5) + 3, int(self.ex / 2.5) + 3)).astype(np.float32)
        self.hres = 0.02
        hw, hh = int(self.ex / self.hres), int(self.ey / self.hres)
        self.H = np.clip(cv2.resize(ctrl, (hw, hh), interpolation=cv2.INTER_CUBIC), 1.2, 4.5)
        # per-pixel undistorted rays (PinholeCamera::liftProjective, 8 fixed-point iterations)
        u, v = np.meshgrid(np.arange(cam.width, dtype=np.float64), np.arange(cam.height, dtype=np.float64))
        mx_d = (u - cam.cx) / cam.fx
        my_d = (v - cam.cy) / cam.fy
        mx, my = mx_d.copy(), my_d.copy()
        for _ in range(8):
            dx, dy = self._dist(mx, my)
            mx, my = mx_d - dx, my_d - dy
        self.rays = np.stack([mx, my, np.ones_like(mx)], -1)

    def _dist(self, x, y):
        c = self.cam
        r2 = x * x + y * y
        rad = c.k1 * r2 + c.k2 * r2 * r2
        return (x * rad + 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x),
                y * rad + 2 * c.p2 * x * y + c.p1 * (r2 + 2 * y * y))

    def _height(self, X, Y):
        mx = ((X + self.ex / 2) / self.hres).astype(np.float32)
        my = ((Y + self.ey / 2) / self.hres).astype(np.float32)
        return cv2.remap(self.H, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)

    def render(self, R_sc, p_s, noise_seed=None, noise_sigma=1.0, invalid_frac=0.05):
        """Returns rgb (H,W,3 u8), gray (H,W u8), depth_mm (H,W u16)."""
        d = self.rays @ R_sc.T
        lam = np.full(d.shape[:2], 2.7)
        for _ in range(10):
            X = p_s[0] + lam * d[..., 0]
            Y = p_s[1] + lam * d[..., 1]
            lam = (self._height(X, Y) - p_s[2]) / d[..., 2]
        X = p_s[0] + lam * d[..., 0]
        Y = p_s[1] + lam * d[..., 1]
        tx = ((X + self.ex / 2) / self.texel).astype(np.float32)
        ty = ((Y + self.ey / 2) / self.texel).astype(np.float32)
        rgb = cv2.remap(self.tex, tx, ty, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        depth_m = lam  # rays have z == 1 in the camera frame => lambda is camera-frame depth
        depth = np.clip(np.rint(depth_m * 1000.0), 0, 65535).astype(np.uint16)
        if noise_seed is not None:
            r = np.random.default_rng(noise_seed)
            if noise_sigma > 0:
                rgb = np.clip(np.rint(rgb.astype(np.float32) + r.normal(0, noise_sigma, rgb.shape)), 0, 255).astype(np.uint8)
            if invalid_frac > 0:
                depth[r.random(depth.shape) < invalid_frac] = 0
        return rgb, rgb_to_gray(rgb), depth


class Sequence:
    """One independent RGB-D + IMU sequence (lazy per-frame rendering)."""

    def __init__(self, seed, cam: CamModel = None, fps=30.0, imu_rate=200.0):
        self.seed = seed
        self.cam = cam or CamModel()
        self.fps = fps
        self.imu_rate = imu_rate
        self.traj = Trajectory(seed, fps)
        self.scene = Scene(seed, self.cam)
        r = np.random.default_rng(seed + 15485863)
        self.gyr_bias = r.uniform(-0.01, 0.01, 3)
        self.acc_bias = r.uniform(-0.01, 0.01, 3)
        self.gyr_n = 0.01
        self.acc_n = 0.1

    def time(self, k):
        return 1.0 + k / self.fps

    def frame(self, k):
        t = self.time(k)
        return self.scene.render(self.traj.rot_s(t), self.traj.pos_s(t), noise_seed=self.seed * 100003 + k)

    def relative_R(self, k):
        """Rotation cam(k-1) -> cam(k) (identity for k == 0)."""
        if k == 0:
            return np.eye(3)
        return self.traj.relative_R(self.time(k - 1), self.time(k))

    def imu_samples(self, t0, t1):
        """(t, gyro, acc) rows for t0 <= t < t1 at imu_rate with noise + bias."""
        k0 = int(np.ceil(t0 * self.imu_rate - 1e-9))
        k1 = int(np.ceil(t1 * self.imu_rate - 1e-9))
        out = []
        for k in range(k0, k1):
            t = k / self.imu_rate
            r = np.random.default_rng(self.seed * 7 + 13 * k)
            g = self.traj.gyro_body(t) + self.gyr_bias + r.normal(0, self.gyr_n, 3)
            a = self.traj.acc_body(t) + self.acc_bias + r.normal(0, self.acc_n, 3)
            out.append((t, g, a))
        return out


def render_gray_frames(seed, n_frames, cam: CamModel = None):
    s = Sequence(seed, cam)
    frames, rels = [], []
    for k in range(n_frames):
        _, g, _ = s.frame(k)
        frames.append(g)
        rels.append(s.relative_R(k))
    return s, frames, rels
