"""numpy restatement of cv::findFundamentalMat(FM_RANSAC) as called by
FeatureTracker::rejectWithF (reference feature_tracker.cpp:441-473, call at :462).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

OpenCV is not vendored in the reference; this restates OpenCV 4.13's
modules/calib3d ptsetreg.cpp (RANSACPointSetRegistrator::run, getSubset,
RANSACUpdateNumIters, LMeDSPointSetRegistrator::run) and fundam.cpp
(FMEstimatorCallback::runKernel/run7Point/computeError, haveCollinearPoints,
cv::RNG, cv::solveCubic) and is pinned against the real cv2.findFundamentalMat in
tests/test_oracle_ransac.py (identical inlier masks for n >= 15).  It is the
inspectable spec of csrc/ransac_kernels.cu.

Note: for 8 <= n < 15 OpenCV switches to LMedS (lmeds_F below).  At n = 14 the median (index 7 of the sorted residuals) is
the best residual outside the minimal sample: pinned against cv2 (identical masks).  For n <= 13 the median index falls
inside the 7 residuals of the minimal sample itself, i.e. on rounding noise (~1e-25) of OpenCV's LAPACK SVD: cv2 returns a
different mask when one input coordinate moves by 1 ulp (tests/test_oracle_ransac.py), so only the semantics (LMedS runs,
the survivors are the points within 0.001 px of one minimal-sample model) can be restated there, not the bits.
"""
import numpy as np, cv2, math
M32 = 0xFFFFFFFF
class CvRNG:
    def __init__(self, state=0xFFFFFFFFFFFFFFFF): self.state = state
    def next(self):
        self.state = ((self.state & M32) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & M32
    def uniform(self, a, b): return a if a == b else int(self.next() % (b - a) + a)

FLT_EPS = 1.1920929e-07
def have_collinear(pts, count):
    i = count - 1
    for j in range(i):
        dx1 = float(pts[j][0]) - float(pts[i][0]); dy1 = float(pts[j][1]) - float(pts[i][1])
        for k in range(j):
            dx2 = float(pts[k][0]) - float(pts[i][0]); dy2 = float(pts[k][1]) - float(pts[i][1])
            if abs(dx2*dy1 - dy2*dx1) <= FLT_EPS*(abs(dx1)+abs(dy1)+abs(dx2)+abs(dy2)): return True
    return False

def get_subset(m1, m2, rng, model_points=7, max_attempts=10000):
    count = len(m1)
    for it in range(max_attempts):
        idx = []
        i = 0
        while i < model_points:
            v = rng.uniform(0, count)
            while v in idx: v = rng.uniform(0, count)
            idx.append(v); i += 1
        ms1 = m1[idx]; ms2 = m2[idx]
        if not have_collinear(ms1, model_points) and not have_collinear(ms2, model_points):
            return idx
    return None

def solve_cubic(c):
    a0,a1,a2,a3 = c
    if a0 == 0:
        if a1 == 0:
            if a2 == 0: return [0.0] if a3 == 0 else []   # n=-1 / 0
            return [-a3/a2]
        d = a2*a2 - 4*a1*a3
        if d >= 0:
            d = math.sqrt(d); q1 = (-a2 + d)*0.5; q2 = (a2 + d)*-0.5
            if abs(q1) > abs(q2): return [q1/a1, a3/q1]
            else: return [q2/a1, a3/q2] if q2 != 0 else [0.0, 0.0]   # approx
        return []
    a0 = 1./a0; a1 *= a0; a2 *= a0; a3 *= a0
    Q = (a1*a1 - 3*a2)*(1./9); R = (2*a1*a1*a1 - 9*a1*a2 + 27*a3)*(1./54)
    Qc = Q*Q*Q; d = Qc - R*R
    if d > 0:
        theta = math.acos(R/math.sqrt(Qc)); sq = math.sqrt(Q); t0 = -2*sq; t1 = theta*(1./3); t2 = a1*(1./3)
        return [t0*math.cos(t1)-t2, t0*math.cos(t1+(2.*math.pi/3))-t2, t0*math.cos(t1+(4.*math.pi/3))-t2]
    elif d == 0:
        if R >= 0: x0 = -2*(R**(1./3)) - a1/3; x1 = (R**(1./3)) - a1/3
        else: x0 = 2*((-R)**(1./3)) - a1/3; x1 = -((-R)**(1./3)) - a1/3
        return [x0, x1]
    else:
        d = math.sqrt(-d); e = (d + abs(R))**(1./3)
        if R > 0: e = -e
        return [(e + Q/e) - a1*(1./3)]

def null_basis(A, method):
    if method == 'svd':
        u,w,vt = np.linalg.svd(A, full_matrices=True); return vt[7].copy(), vt[8].copy()
    if method == 'qr':
        q,r = np.linalg.qr(A.T, mode='complete'); return q[:,7].copy(), q[:,8].copy()

def run7point(m1, m2, normalize=False, method='svd'):
    m1 = m1.astype(np.float64); m2 = m2.astype(np.float64)
    T1 = T2 = None
    if normalize:
        c1 = m1.mean(0); c2 = m2.mean(0)
        s1 = np.sqrt(((m1-c1)**2).sum(1)).mean(); s2 = np.sqrt(((m2-c2)**2).sum(1)).mean()
        if s1 < FLT_EPS or s2 < FLT_EPS: return []
        s1 = math.sqrt(2.)/s1; s2 = math.sqrt(2.)/s2
        m1 = (m1-c1)*s1; m2 = (m2-c2)*s2
        T1 = np.array([[s1,0,-c1[0]*s1],[0,s1,-c1[1]*s1],[0,0,1]]); T2 = np.array([[s2,0,-c2[0]*s2],[0,s2,-c2[1]*s2],[0,0,1]])
    A = np.zeros((7,9))
    x0,y0 = m1[:,0],m1[:,1]; x1,y1 = m2[:,0],m2[:,1]
    A[:,0]=x1*x0; A[:,1]=x1*y0; A[:,2]=x1; A[:,3]=y1*x0; A[:,4]=y1*y0; A[:,5]=y1; A[:,6]=x0; A[:,7]=y0; A[:,8]=1
    f1,f2 = null_basis(A, method)
    f1 = f1 - f2
    c = [0]*4
    t0 = f2[4]*f2[8]-f2[5]*f2[7]; t1 = f2[3]*f2[8]-f2[5]*f2[6]; t2 = f2[3]*f2[7]-f2[4]*f2[6]
    c[3] = f2[0]*t0 - f2[1]*t1 + f2[2]*t2
    c[2] = (f1[0]*t0 - f1[1]*t1 + f1[2]*t2 - f1[3]*(f2[1]*f2[8]-f2[2]*f2[7]) + f1[4]*(f2[0]*f2[8]-f2[2]*f2[6]) - f1[5]*(f2[0]*f2[7]-f2[1]*f2[6])
            + f1[6]*(f2[1]*f2[5]-f2[2]*f2[4]) - f1[7]*(f2[0]*f2[5]-f2[2]*f2[3]) + f1[8]*(f2[0]*f2[4]-f2[1]*f2[3]))
    t0 = f1[4]*f1[8]-f1[5]*f1[7]; t1 = f1[3]*f1[8]-f1[5]*f1[6]; t2 = f1[3]*f1[7]-f1[4]*f1[6]
    c[1] = (f2[0]*t0 - f2[1]*t1 + f2[2]*t2 - f2[3]*(f1[1]*f1[8]-f1[2]*f1[7]) + f2[4]*(f1[0]*f1[8]-f1[2]*f1[6]) - f2[5]*(f1[0]*f1[7]-f1[1]*f1[6])
            + f2[6]*(f1[1]*f1[5]-f1[2]*f1[4]) - f2[7]*(f1[0]*f1[5]-f1[2]*f1[3]) + f2[8]*(f1[0]*f1[4]-f1[1]*f1[3]))
    c[0] = f1[0]*t0 - f1[1]*t1 + f1[2]*t2
    roots = solve_cubic(c)
    Fs = []
    for r in roots:
        lam = r; mu = 1.0
        s = f1[8]*r + f2[8]
        F = np.zeros(9)
        if abs(s) > 2.220446049250313e-16:
            mu = 1./s; lam *= mu; F[8] = 1.0
        else: F[8] = 0.0
        F[:8] = f1[:8]*lam + f2[:8]*mu
        F = F.reshape(3,3)
        if normalize:
            F = T2.T @ F @ T1
            if abs(F[2,2]) > 2.220446049250313e-16: F = F/F[2,2]
        Fs.append(F)
    return Fs

def compute_error(m1, m2, F):
    F = F.ravel()
    x1 = m1[:,0].astype(np.float64); y1 = m1[:,1].astype(np.float64); x2 = m2[:,0].astype(np.float64); y2 = m2[:,1].astype(np.float64)
    a = F[0]*x1 + F[1]*y1 + F[2]; b = F[3]*x1 + F[4]*y1 + F[5]; c = F[6]*x1 + F[7]*y1 + F[8]
    s2 = 1./(a*a + b*b); d2 = x2*a + y2*b + c
    a = F[0]*x2 + F[3]*y2 + F[6]; b = F[1]*x2 + F[4]*y2 + F[7]; c = F[2]*x2 + F[5]*y2 + F[8]
    s1 = 1./(a*a + b*b); d1 = x1*a + y1*b + c
    return np.maximum(d1*d1*s1, d2*d2*s2).astype(np.float32)

def update_niters(p, ep, model_points, max_iters):
    p = min(max(p,0.),1.); ep = min(max(ep,0.),1.)
    num = max(1.-p, 2.2250738585072014e-308); denom = 1. - (1.-ep)**model_points
    if denom < 2.2250738585072014e-308: return 0
    num = math.log(num); denom = math.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters*(-denom)) else int(np.rint(num/denom))

def ransac_F(m1, m2, thr=1.0, conf=0.99, max_iters=1000, normalize=False, method='svd', trace=None):
    count = len(m1)
    rng = CvRNG()
    niters = max(max_iters, 1); max_good = 0; best_mask = np.zeros(count, np.uint8); best_F = None
    it = 0
    thr2 = np.float32(thr*thr) if False else thr*thr
    while it < niters:
        idx = get_subset(m1, m2, rng)
        if idx is None:
            if it == 0: return None, None
            break
        Fs = run7point(m1[idx], m2[idx], normalize, method)
        for F in Fs:
            err = compute_error(m1, m2, F)
            mask = (err <= thr2).astype(np.uint8)
            good = int(mask.sum())
            if good > max(max_good, 6):
                best_mask = mask; best_F = F; max_good = good
                niters = update_niters(conf, (count - good)/count, 7, niters)
        if trace is not None: trace.append((it, idx, len(Fs), max_good, niters))
        it += 1
    if max_good > 0: return best_F, best_mask
    return None, None

def lmeds_F(m1, m2, conf=0.99, max_iters=1000, normalize=False, method='svd'):
    """LMeDSPointSetRegistrator::run (ptsetreg.cpp), the path cv::findFundamentalMat(FM_RANSAC) takes for npoints < 15:
    RANSACUpdateNumIters(conf, 0.45, 7, maxIters) (>= 3) iterations, getSubset with the default 1000 attempts, median =
    std::nth_element at count / 2 of the float residuals, strict minimum in iteration order,
    sigma = max(2.5 * 1.4826 * (1 + 5 / (count - 7)) * sqrt(median), 0.001), mask = err <= (float)(sigma^2)."""
    count = len(m1)
    rng = CvRNG()
    niters = max(update_niters(conf, 0.45, 7, max_iters), 3)
    min_median = 1.7976931348623157e308; best_F = None
    for it in range(niters):
        idx = get_subset(m1, m2, rng, 7, 1000)
        if idx is None:
            if it == 0: return None, None
            break
        for F in run7point(m1[idx], m2[idx], normalize, method):
            err = np.sort(compute_error(m1, m2, F))
            med = float(err[count//2])
            if med < min_median: min_median = med; best_F = F
    if min_median < 1.7976931348623157e308:
        sigma = 2.5*1.4826*(1 + 5./(count - 7))*math.sqrt(min_median)
        sigma = max(sigma, 0.001)
        err = compute_error(m1, m2, best_F)
        mask = (err <= np.float32(sigma*sigma)).astype(np.uint8)
        return (best_F if mask.sum() >= 7 else None), mask
    return None, None
