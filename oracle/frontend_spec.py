"""Bit-level numpy restatement of the third-party (OpenCV) arithmetic on the
front-end hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls these OpenCV entry points (not vendored under
/root/reference):

* cv::calcOpticalFlowPyrLK            vins_estimator/src/feature_tracker/feature_tracker.cpp:302-310
  (-> buildOpticalFlowPyramid / pyrDown, calcScharrDeriv, LKTrackerInvoker)
* cv::FastFeatureDetector::detect     feature_tracker.cpp:109-110 (default ctor :29
  => threshold 10, nonmaxSuppression, TYPE_9_16)
* cv::circle(mask, pt, MIN_DIST,0,-1) feature_tracker.cpp:201,206,230

OpenCV version pinned for parity: opencv-python-headless 4.13.0 (this image).
These restatements are the *inspectable spec* the sm_100a kernels implement;
tests/test_oracle_frontend.py pins each of them against real cv2 4.13
(bit-exact for pyrDown / Scharr / FAST / circle, <=1e-3 px + identical status
for LK).  They are slow (python loops) and meant for small cases only.
"""
import numpy as np

# ----------------------------------------------------------------------------
# pyramid: cv::pyrDown as used by cv::buildOpticalFlowPyramid
# ----------------------------------------------------------------------------

def reflect101(i, n):
    """cv::borderInterpolate(i, n, BORDER_REFLECT_101) for |overshoot| < n."""
    if n == 1:
        return 0
    while i < 0 or i >= n:
        if i < 0:
            i = -i
        else:
            i = 2 * n - 2 - i
    return i


def pyr_down(img):
    """cv::pyrDown(u8): separable [1 4 6 4 1], (sum+128)>>8, REFLECT_101,
    dst size ((w+1)/2, (h+1)/2)."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    src = img.astype(np.int32)
    k = (1, 4, 6, 4, 1)
    # horizontal pass at even columns -> (h, ow)
    cols = np.array([[reflect101(2 * x + d - 2, w) for d in range(5)] for x in range(ow)])
    hor = np.zeros((h, ow), np.int32)
    for d in range(5):
        hor += k[d] * src[:, cols[:, d]]
    rows = np.array([[reflect101(2 * y + d - 2, h) for d in range(5)] for y in range(oh)])
    out = np.zeros((oh, ow), np.int32)
    for d in range(5):
        out += k[d] * hor[rows[:, d], :]
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(img, max_level):
    pyr = [np.ascontiguousarray(img)]
    for _ in range(max_level):
        pyr.append(pyr_down(pyr[-1]))
    return pyr


# ----------------------------------------------------------------------------
# Scharr derivative image (cv::calcScharrDeriv inside lkpyramid.cpp)
# ----------------------------------------------------------------------------

def scharr_deriv(img):
    """Returns (dx, dy) int16, un-normalised 3x3 Scharr, REFLECT_101 at the
    image border (row -1 -> row 1, col -1 -> col 1)."""
    h, w = img.shape
    s = img.astype(np.int32)
    ru = np.array([reflect101(y - 1, h) for y in range(h)])
    rd = np.array([reflect101(y + 1, h) for y in range(h)])
    t0 = 3 * (s[ru, :] + s[rd, :]) + 10 * s          # smoothing along y
    t1 = s[rd, :] - s[ru, :]                           # derivative along y
    cl = np.array([reflect101(x - 1, w) for x in range(w)])
    cr = np.array([reflect101(x + 1, w) for x in range(w)])
    dx = t0[:, cr] - t0[:, cl]
    dy = 3 * (t1[:, cl] + t1[:, cr]) + 10 * t1
    return dx.astype(np.int16), dy.astype(np.int16)


# ----------------------------------------------------------------------------
# pyramidal LK (LKTrackerInvoker), winSize 21x21
# ----------------------------------------------------------------------------
W_BITS = 14
FLT_SCALE = np.float32(1.0 / (1 << 20))
FLT_EPSILON = np.float32(1.1920929e-07)


def _cv_round_f32(v):
    """cvRound on a float32 value (round half to even)."""
    return int(np.rint(np.float32(v)))


def _pad_img(img, win):
    """Pyramid level as stored by buildOpticalFlowPyramid: REFLECT_101 border
    of `win` pixels on each side."""
    return np.pad(img, win, mode="reflect")


def _pad_deriv(d, win):
    """Derivative level: BORDER_CONSTANT(0) of `win` pixels."""
    return np.pad(d, win, mode="constant")


def _sum_a_opencv(P):
    """float32 accumulation of the 21x21 integer products exactly as OpenCV's SSE path
    (lkpyramid.cpp, CV_SIMD128): per row, pixels 0..15 go to 4 lane accumulators
    (lane = x mod 4, float adds in x order), pixels 16..20 to a scalar accumulator;
    result = scalar + ((l0 + l2) + (l1 + l3))  (v_reduce_sum).  Pinned empirically
    against cv2 4.13: bit-exact tracks (tests/test_oracle_frontend.py)."""
    f32 = np.float32
    vq = np.zeros(4, np.float32)
    sc = f32(0)
    Pf = P.astype(np.float32)          # |products| < 2^24: exact
    for y in range(P.shape[0]):
        for x0 in (0, 4, 8, 12):
            vq = (vq + Pf[y, x0:x0 + 4]).astype(np.float32)
        for x in range(16, P.shape[1]):
            sc = f32(sc + Pf[y, x])
    return f32(sc + f32(f32(vq[0] + vq[2]) + f32(vq[1] + vq[3])))


def _sum_b_opencv(P):
    """float32 accumulation of diff*I{x,y} as OpenCV's SSE path: per row and per 8-pixel
    group, v_dotprod forms the exact int32 pair sums p[k]+p[k+4] (k=0..3), converts them
    to float and adds them to 4 accumulators; pixels 16..20 go to a scalar float
    accumulator (each product converted to float first);
    result = scalar + ((a0 + a2) + (a1 + a3))."""
    f32 = np.float32
    acc = np.zeros(4, np.float32)
    sc = f32(0)
    for y in range(P.shape[0]):
        for x0 in (0, 8):
            p = P[y, x0:x0 + 8]
            acc = (acc + (p[0:4] + p[4:8]).astype(np.float32)).astype(np.float32)
        for x in range(16, P.shape[1]):
            sc = f32(sc + f32(int(P[y, x])))
    return f32(sc + f32(f32(acc[0] + acc[2]) + f32(acc[1] + acc[3])))


def lk_track(prev_pyr, next_pyr, prev_pts, next_pts_init, max_level,
             use_initial_flow=True, win=21, max_count=30, epsilon=0.01,
             min_eig_threshold=1e-4, accum="opencv"):
    """Restatement of cv::calcOpticalFlowPyrLK for 8UC1 images given prebuilt
    pyramids.  Returns (next_pts float32 Nx2, status uint8 N).

    `epsilon` is the TermCriteria epsilon; OpenCV squares it internally
    (criteria.epsilon *= criteria.epsilon), and clamps maxCount to [0,100] and
    epsilon to [0,10].
    """
    f32 = np.float32
    n = len(prev_pts)
    eps2 = np.float64(min(max(epsilon, 0.0), 10.0)) ** 2
    max_count = min(max(max_count, 0), 100)
    half = f32((win - 1) * 0.5)
    nextp = np.array(next_pts_init if use_initial_flow else prev_pts, dtype=np.float32).reshape(-1, 2).copy()
    prev_pts = np.asarray(prev_pts, dtype=np.float32).reshape(-1, 2)
    status = np.ones(n, np.uint8)

    for level in range(max_level, -1, -1):
        I = prev_pyr[level]
        J = next_pyr[level]
        rows, cols = I.shape
        dx, dy = scharr_deriv(I)
        Ip = _pad_img(I, win).astype(np.int64)
        Jp = _pad_img(J, win).astype(np.int64)
        dxp = _pad_deriv(dx, win).astype(np.int64)
        dyp = _pad_deriv(dy, win).astype(np.int64)
        scale = f32(1.0 / (1 << level))
        for p in range(n):
            prevPt = prev_pts[p] * scale
            if level == max_level:
                if use_initial_flow:
                    nextPt = nextp[p] * scale
                else:
                    nextPt = prevPt.copy()
            else:
                nextPt = nextp[p] * f32(2.0)
            nextp[p] = nextPt
            prevPt = prevPt - half
            ix = int(np.floor(prevPt[0])); iy = int(np.floor(prevPt[1]))
            if ix < -win or ix >= cols or iy < -win or iy >= rows:
                if level == 0:
                    status[p] = 0
                continue
            a = f32(prevPt[0] - f32(ix)); b = f32(prevPt[1] - f32(iy))
            one = f32(1.0)
            iw00 = _cv_round_f32((one - a) * (one - b) * f32(1 << W_BITS))
            iw01 = _cv_round_f32(a * (one - b) * f32(1 << W_BITS))
            iw10 = _cv_round_f32((one - a) * b * f32(1 << W_BITS))
            iw11 = (1 << W_BITS) - iw00 - iw01 - iw10
            y0 = iy + win; x0 = ix + win
            def interp(A, yy, xx, w00, w01, w10, w11):
                return (A[yy:yy + win, xx:xx + win] * w00 + A[yy:yy + win, xx + 1:xx + win + 1] * w01 +
                        A[yy + 1:yy + win + 1, xx:xx + win] * w10 + A[yy + 1:yy + win + 1, xx + 1:xx + win + 1] * w11)
            Iw = (interp(Ip, y0, x0, iw00, iw01, iw10, iw11) + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)
            Ix = (interp(dxp, y0, x0, iw00, iw01, iw10, iw11) + (1 << (W_BITS - 1))) >> W_BITS
            Iy = (interp(dyp, y0, x0, iw00, iw01, iw10, iw11) + (1 << (W_BITS - 1))) >> W_BITS
            # OpenCV accumulates in float32 in a fixed SIMD lane order; accum="opencv"
            # reproduces that order bit-exactly, accum="exact" rounds the exact integer
            # sum once (differs by ~1e-6 relative).
            if accum == "opencv" and win == 21:
                A11 = f32(_sum_a_opencv(Ix * Ix) * FLT_SCALE)
                A12 = f32(_sum_a_opencv(Ix * Iy) * FLT_SCALE)
                A22 = f32(_sum_a_opencv(Iy * Iy) * FLT_SCALE)
            else:
                A11 = f32(f32(int((Ix * Ix).sum())) * FLT_SCALE)
                A12 = f32(f32(int((Ix * Iy).sum())) * FLT_SCALE)
                A22 = f32(f32(int((Iy * Iy).sum())) * FLT_SCALE)
            D = f32(A11 * A22 - A12 * A12)
            minEig = f32((A22 + A11 - np.sqrt(f32((A11 - A22) * (A11 - A22) + f32(4.0) * A12 * A12))) / f32(2 * win * win))
            if minEig < f32(min_eig_threshold) or D < FLT_EPSILON:
                if level == 0:
                    status[p] = 0
                continue
            D = f32(one / D)
            nextPt = nextPt - half
            prevDelta = np.zeros(2, np.float32)
            for j in range(max_count):
                jx = int(np.floor(nextPt[0])); jy = int(np.floor(nextPt[1]))
                if jx < -win or jx >= cols or jy < -win or jy >= rows:
                    if level == 0:
                        status[p] = 0
                    break
                a = f32(nextPt[0] - f32(jx)); b = f32(nextPt[1] - f32(jy))
                w00 = _cv_round_f32((one - a) * (one - b) * f32(1 << W_BITS))
                w01 = _cv_round_f32(a * (one - b) * f32(1 << W_BITS))
                w10 = _cv_round_f32((one - a) * b * f32(1 << W_BITS))
                w11 = (1 << W_BITS) - w00 - w01 - w10
                Jw = (interp(Jp, jy + win, jx + win, w00, w01, w10, w11) + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)
                diff = Jw - Iw
                if accum == "opencv" and win == 21:
                    b1 = f32(_sum_b_opencv(diff * Ix) * FLT_SCALE)
                    b2 = f32(_sum_b_opencv(diff * Iy) * FLT_SCALE)
                else:
                    b1 = f32(f32(int((diff * Ix).sum())) * FLT_SCALE)
                    b2 = f32(f32(int((diff * Iy).sum())) * FLT_SCALE)
                delta = np.array([f32(f32(A12 * b2 - A22 * b1) * D), f32(f32(A12 * b1 - A11 * b2) * D)], np.float32)
                nextPt = nextPt + delta
                nextp[p] = nextPt + half
                if np.float64(delta[0]) * np.float64(delta[0]) + np.float64(delta[1]) * np.float64(delta[1]) <= eps2:
                    break
                if j > 0 and abs(delta[0] + prevDelta[0]) < 0.01 and abs(delta[1] + prevDelta[1]) < 0.01:
                    nextp[p] = nextp[p] - delta * f32(0.5)
                    break
                prevDelta = delta
            # `err` is requested by the reference (feature_tracker.cpp:301) so
            # the final bounds check at level 0 is active.
            if status[p] and level == 0:
                q = nextp[p] - half
                qx = int(np.floor(q[0])); qy = int(np.floor(q[1]))
                if qx < -win or qx >= cols or qy < -win or qy >= rows:
                    status[p] = 0
    return nextp, status


# ----------------------------------------------------------------------------
# FAST-9/16, threshold 10, NMS (cv::FastFeatureDetector default)
# ----------------------------------------------------------------------------
RING16 = ((0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
          (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3))


def fast_score_map(roi, threshold=10):
    """cornerScore<16> - 1 >= threshold ... as a dense map over the ROI.

    score(y,x) = max over the 16 arcs of 9 contiguous ring pixels of
                 max(min_k d_k, -max_k d_k) - 1, d_k = I_c - I_ring[k];
    kept iff >= threshold, only for the ROI interior [3,h-3)x[3,w-3)."""
    h, w = roi.shape
    sc = np.zeros((h, w), np.int32)
    if h < 7 or w < 7:
        return sc
    I = roi.astype(np.int32)
    c = I[3:h - 3, 3:w - 3]
    d = np.stack([c - I[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for (dx, dy) in RING16], 0)  # 16 x H x W
    d = np.concatenate([d, d[:8]], 0)   # cyclic
    best = np.full(c.shape, -10 ** 9, np.int32)
    for s in range(16):
        arc = d[s:s + 9]
        best = np.maximum(best, np.maximum(arc.min(0), -arc.max(0)))
    s = best - 1
    sc[3:h - 3, 3:w - 3] = np.where(s >= threshold, s, 0)
    return sc


def fast_detect(roi, threshold=10, mask=None):
    """Returns list of (x, y, score) in row-major order: strict-maximum NMS
    over the 8 neighbours, then the mask filter (mask!=0 keeps)."""
    sc = fast_score_map(roi, threshold)
    h, w = roi.shape
    out = []
    if h < 7 or w < 7:
        return out
    p = np.pad(sc, 1)
    ctr = p[1:-1, 1:-1]
    keep = ctr > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            keep &= ctr > p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    ys, xs = np.nonzero(keep)
    for y, x in zip(ys, xs):
        if mask is not None and mask[y, x] == 0:
            continue
        out.append((int(x), int(y), int(sc[y, x])))
    return out


# ----------------------------------------------------------------------------
# filled circle (cv::circle thickness=-1) as a predicate
# ----------------------------------------------------------------------------

def circle_covers(cx, cy, r, x, y):
    """True iff cv::circle(mask,(cx,cy),r,0,-1) zeroes pixel (x,y)."""
    return (x - cx) * (x - cx) + (y - cy) * (y - cy) <= r * r


def clahe(img, clip=3.0, tiles=8):
    """cv::createCLAHE(clip, Size(tiles, tiles))->apply(img) (OpenCV imgproc/clahe.cpp: CLAHE_CalcLut_Body +
    CLAHE_Interpolation_Body; call site feature_tracker.cpp:269-275), incl. the border extension for frame sizes that are
    not multiples of the tile grid: copyMakeBorder(src, ext, 0, tiles - h % tiles, 0, tiles - w % tiles, REFLECT_101) --
    a dimension that already divides still receives `tiles` extra pixels.  Pinned bit-exactly against cv2 4.13 in
    tests/test_oracle_frontend.py."""
    h, w = img.shape
    if h % tiles == 0 and w % tiles == 0:
        ext = img
    else:
        eh, ew = h + (tiles - h % tiles), w + (tiles - w % tiles)
        ry, rx = np.arange(eh), np.arange(ew)
        ry = np.where(ry >= h, 2 * h - 2 - ry, ry); rx = np.where(rx >= w, 2 * w - 2 - rx, rx)      # REFLECT_101
        ext = img[ry][:, rx]
    th, tw = ext.shape[0] // tiles, ext.shape[1] // tiles
    tot = th * tw
    lut_scale = np.float32(255.0) / np.float32(tot)
    cl = max(int(clip * tot / 256), 1)
    luts = np.zeros((tiles, tiles, 256), np.uint8)
    for ty in range(tiles):
        for tx in range(tiles):
            hist = np.bincount(ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw].ravel(), minlength=256).astype(np.int64)
            clipped = int(np.maximum(hist - cl, 0).sum())
            hist = np.minimum(hist, cl)
            batch = clipped // 256
            resid = clipped - batch * 256
            hist += batch
            if resid != 0:
                step = max(256 // resid, 1)
                i = 0
                while i < 256 and resid > 0:
                    hist[i] += 1; i += step; resid -= 1
            s = np.cumsum(hist).astype(np.float32) * lut_scale
            luts[ty, tx] = np.clip(np.rint(s), 0, 255).astype(np.uint8)
    one, half = np.float32(1), np.float32(0.5)
    inv_tw, inv_th = one / np.float32(tw), one / np.float32(th)
    txf = np.arange(w, dtype=np.float32) * inv_tw - half
    tyf = np.arange(h, dtype=np.float32) * inv_th - half
    tx1 = np.floor(txf).astype(np.int32); ty1 = np.floor(tyf).astype(np.int32)
    xa = (txf - tx1.astype(np.float32)).astype(np.float32); ya = (tyf - ty1.astype(np.float32)).astype(np.float32)
    xa1, ya1 = one - xa, one - ya
    tx2 = np.minimum(tx1 + 1, tiles - 1); tx1 = np.maximum(tx1, 0)
    ty2 = np.minimum(ty1 + 1, tiles - 1); ty1 = np.maximum(ty1, 0)
    v = img.astype(np.int64)
    Y1, Y2, X1, X2 = ty1[:, None], ty2[:, None], tx1[None, :], tx2[None, :]
    a = luts[Y1, X1, v].astype(np.float32); b = luts[Y1, X2, v].astype(np.float32)
    c = luts[Y2, X1, v].astype(np.float32); d = luts[Y2, X2, v].astype(np.float32)
    res = (a * xa1[None, :] + b * xa[None, :]) * ya1[:, None] + (c * xa1[None, :] + d * xa[None, :]) * ya[:, None]
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)
