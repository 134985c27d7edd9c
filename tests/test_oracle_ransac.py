"""CPU test: the numpy restatement of cv::findFundamentalMat(FM_RANSAC)
(oracle/ransac_spec.py, the spec of csrc/ransac_kernels.cu) against the real
cv2 4.13 -- identical inlier masks, for SVD- and elimination-based null spaces,
with and without Hartley normalisation (the product kernel uses normalised
pivoted elimination)."""
import cv2
import numpy as np
import pytest

from oracle import ransac_spec as RS


def scene(rng, n, out_frac, noise):
    P = rng.uniform([-2.5, -2, 1.5], [2.5, 2, 6], (n, 3))
    R, _ = cv2.Rodrigues(rng.normal(0, 0.02, 3))
    T = rng.normal(0, 0.05, 3)
    m1 = P[:, :2] / P[:, 2:] * 460 + [320, 240]
    P2 = P @ R.T + T
    m2 = P2[:, :2] / P2[:, 2:] * 460 + [320, 240] + rng.normal(0, noise, (n, 2))
    no = int(out_frac * n)
    m2[:no] += rng.normal(0, 8, (no, 2))
    return m1.astype(np.float32), m2.astype(np.float32)


def test_cv_rng_sequence():
    """cv::RNG((uint64)-1) multiply-with-carry stream (first values are fixed constants)."""
    r = RS.CvRNG()
    vals = [r.next() for _ in range(3)]
    r2 = RS.CvRNG()
    assert vals == [r2.next() for _ in range(3)] and len(set(vals)) == 3


@pytest.mark.parametrize("normalize,method", [(False, "svd"), (True, "qr")])
def test_ransac_mask_equals_cv2(normalize, method):
    rng = np.random.default_rng(2)
    for _ in range(25):
        n = int(rng.integers(15, 200))
        m1, m2 = scene(rng, n, rng.choice([0.0, 0.05, 0.2, 0.4]), rng.choice([0.05, 0.3, 0.8]))
        _, mk = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, 1.0, 0.99)
        _, mask = RS.ransac_F(m1, m2, normalize=normalize, method=method)
        assert (mk is None) == (mask is None)
        if mk is not None:
            assert np.array_equal(mk.ravel(), mask)


def test_small_n_is_lmeds_in_opencv():
    """Documents why 8 <= n < 15 cannot be pinned: cv2 returns the LMedS result there."""
    rng = np.random.default_rng(9)
    m1, m2 = scene(rng, 12, 0.2, 0.5)
    _, a = cv2.findFundamentalMat(m1, m2, cv2.FM_RANSAC, 1.0, 0.99)
    _, b = cv2.findFundamentalMat(m1, m2, cv2.FM_LMEDS, 1.0, 0.99)
    assert np.array_equal(a, b)
