"""TEST INFRASTRUCTURE ONLY -- CPU restatements of third-party arithmetic the
reference calls but which is absent from /root/reference and from this image.

PARITY UNPINNED against the packages themselves: the reference's
requirements.txt names ``scikit-image`` and ``lsq-ellipse`` with no version pin
(/root/reference/requirements.txt:8-9), neither package is installed here,
there is no network, and the reference has no tests or golden vectors.  Each
function restates the package's published algorithm and is anchored on the
reference call site that uses it:

  downscale_local_mean   /root/reference/ellipse_to_circle.py:301
  canny                  /root/reference/ellipse_to_circle.py:245-250
  ProjectiveTransform/warp  /root/reference/ellipse_to_circle.py:112-114
  LsqEllipse             /root/reference/ellipse_to_circle.py:57-59

CROSS-CHECKED (tests/test_thirdparty_pins.py) against independent
implementations of the same operations that ARE in this image:

  warp                   scipy.ndimage.map_coordinates(order=1, mode='grid-constant', cval=img[0,0]) + the same
                         clip / *2**16 / truncation: <= 1 DN on <= 1 pixel in 10^5 (1 pixel of 4.0 M over 35 warps,
                         a last-ulp truncation), and the committed circ_* fixtures themselves likewise
  LsqEllipse             a 6x6 generalised-eigenvalue solve of the same constrained conic problem (1e-6: centre,
                         axes, angle, both phi branches) and cv2.fitEllipseDirect (median 6e-6, 90 % < 5e-5,
                         max 4e-3 -- OpenCV's own arithmetic is partly single precision), 200 random arcs
  downscale_local_mean   cv2.resize(INTER_AREA) to 1e-12 on multiples of 4, zero padding worked by hand
  canny                  its gaussian stage against cv2.sepFilter2D / cv2.GaussianBlur (BORDER_CONSTANT) and its
                         sobel stage against cv2.Sobel (BORDER_REFLECT) to 1e-12 relative; the whole detector
                         against cv2.Canny on a fixture's flood image (same boundary within 2 px for >= 95 %).
                         The interpolated non-maximum suppression and the hysteresis have no bit-level
                         counterpart here and stay restatements.

Acceptance check (tests/test_oracle.py): a synthetic elliptical Sun pushed
through the *reference's own* two_step + correct_image with these stand-ins
comes out circular with the row count unchanged.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.  The product never does.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi


# --------------------------------------------------------------------------
# skimage.transform.downscale_local_mean  (block_reduce with np.mean, cval=0)
# --------------------------------------------------------------------------
def downscale_local_mean(image, factors, cval=0, clip=True):
    image = np.asarray(image)
    pad = []
    for n, f in zip(image.shape, factors):
        pad.append((0, (-n) % f))
    image = np.pad(image, pad, mode='constant', constant_values=cval)
    fr, fc = factors
    h, w = image.shape
    blocks = image.reshape(h // fr, fr, w // fc, fc)
    return blocks.mean(axis=(1, 3))


# --------------------------------------------------------------------------
# skimage.feature.canny (float image, no mask, mode='constant')
# --------------------------------------------------------------------------
def _gaussian(img, sigma):
    # skimage.filters.gaussian -> scipy.ndimage.gaussian_filter(truncate=4.0)
    return ndi.gaussian_filter(img, sigma, mode='constant', cval=0.0, truncate=4.0)


def _nms_bilinear(isobel, jsobel, magnitude, eroded_mask, low_threshold):
    """Interpolated non-maximum suppression (skimage/feature/_canny_cy.pyx)."""
    rows, cols = magnitude.shape
    out = np.zeros_like(magnitude)
    m = magnitude
    pm = np.pad(m, 1, mode='constant')          # so x+-1 / y+-1 never index outside

    def nb(dx, dy):
        return pm[1 + dx:1 + dx + rows, 1 + dy:1 + dy + cols]

    active = eroded_mask & (m >= low_threshold)
    is_down = isobel <= 0
    is_up = isobel >= 0
    is_left = jsobel <= 0
    is_right = jsobel >= 0
    cond1 = (is_up & is_right) | (is_down & is_left)
    cond2 = (is_down & is_right) | (is_up & is_left)
    ai = np.abs(isobel)
    aj = np.abs(jsobel)
    with np.errstate(divide='ignore', invalid='ignore'):
        w_ji = aj / ai
        w_ij = ai / aj
    # cond1 branch
    c1a = cond1 & (ai > aj)
    c1b = cond1 & ~(ai > aj)
    c2a = ~cond1 & cond2 & (ai < aj)
    c2b = ~cond1 & cond2 & ~(ai < aj)
    w = np.zeros_like(m)
    n11 = np.zeros_like(m); n12 = np.zeros_like(m); n21 = np.zeros_like(m); n22 = np.zeros_like(m)
    for sel, ww, a11, a12, a21, a22 in (
        (c1a, w_ji, nb(1, 0), nb(1, 1), nb(-1, 0), nb(-1, -1)),
        (c1b, w_ij, nb(0, 1), nb(1, 1), nb(0, -1), nb(-1, -1)),
        (c2a, w_ij, nb(0, 1), nb(-1, 1), nb(0, -1), nb(1, -1)),
        (c2b, w_ji, nb(-1, 0), nb(-1, 1), nb(1, 0), nb(1, -1)),
    ):
        w = np.where(sel, ww, w)
        n11 = np.where(sel, a11, n11); n12 = np.where(sel, a12, n12)
        n21 = np.where(sel, a21, n21); n22 = np.where(sel, a22, n22)
    with np.errstate(invalid='ignore'):
        c_plus = (n12 * w + n11 * (1.0 - w)) <= m
        c_minus = (n22 * w + n21 * (1.0 - w)) <= m
    keep = active & (cond1 | cond2) & c_plus & c_minus
    out[keep] = m[keep]
    return out


def canny(image, sigma=1.0, low_threshold=None, high_threshold=None):
    image = np.asarray(image, dtype=np.float64)
    if low_threshold is None:
        low_threshold = 0.1
    if high_threshold is None:
        high_threshold = 0.2
    # float image: dtype_limits -> (-1, 1), thresholds are used as given.
    mask = np.ones(image.shape, dtype=np.float64)
    eroded = np.ones(image.shape, dtype=bool)
    eroded[:1, :] = False
    eroded[-1:, :] = False
    eroded[:, :1] = False
    eroded[:, -1:] = False
    bleed_over = _gaussian(mask, sigma) + np.finfo(np.float64).eps
    smoothed = _gaussian(image, sigma) / bleed_over
    jsobel = ndi.sobel(smoothed, axis=1)
    isobel = ndi.sobel(smoothed, axis=0)
    magnitude = np.sqrt(isobel * isobel + jsobel * jsobel)
    low_masked = _nms_bilinear(isobel, jsobel, magnitude, eroded, low_threshold)
    low_mask = low_masked > 0
    labels, count = ndi.label(low_mask, np.ones((3, 3), bool))
    if count == 0:
        return low_mask
    high_mask = low_mask & (low_masked >= high_threshold)
    nonzero = np.unique(labels[high_mask])
    good = np.zeros((count + 1,), bool)
    good[nonzero] = True
    return good[labels]


# --------------------------------------------------------------------------
# skimage.transform.ProjectiveTransform / warp (order 1, mode 'constant')
# --------------------------------------------------------------------------
class ProjectiveTransform:
    def __init__(self, matrix=None):
        self.params = np.eye(3) if matrix is None else np.asarray(matrix, dtype=np.float64)


def warp(image, inverse_map, output_shape=None, cval=0.0):
    """Bilinear inverse warp, constant mode, result clipped to the input range
    (skimage.transform.warp -> _warp_fast + bilinear_interpolation).

    ``c = (M00*x + M01*y + M02) / z``, ``r = (M10*x + M11*y + M12) / z``,
    ``z = M20*x + M21*y + M22``; taps floor/ceil; ``dr = r - floor(r)``;
    out-of-image taps read ``cval``.
    """
    M = inverse_map.params if hasattr(inverse_map, 'params') else np.asarray(inverse_map)
    image = np.asarray(image, dtype=np.float64)
    rows, cols = image.shape
    oh, ow = int(output_shape[0]), int(output_shape[1])
    out = np.empty((oh, ow), dtype=np.float64)
    x = np.arange(ow, dtype=np.float64)
    cval = float(cval)

    def pix(rr, cc):
        inside = (rr >= 0) & (rr < rows) & (cc >= 0) & (cc < cols)
        v = image[np.clip(rr, 0, rows - 1), np.clip(cc, 0, cols - 1)]
        return np.where(inside, v, cval)

    for tfr in range(oh):
        y = float(tfr)
        z = M[2, 0] * x + M[2, 1] * y + M[2, 2]
        c = (M[0, 0] * x + M[0, 1] * y + M[0, 2]) / z
        r = (M[1, 0] * x + M[1, 1] * y + M[1, 2]) / z
        minr = np.floor(r).astype(np.int64); minc = np.floor(c).astype(np.int64)
        maxr = np.ceil(r).astype(np.int64); maxc = np.ceil(c).astype(np.int64)
        dr = r - minr
        dc = c - minc
        top = (1 - dc) * pix(minr, minc) + dc * pix(minr, maxc)
        bottom = (1 - dc) * pix(maxr, minc) + dc * pix(maxr, maxc)
        out[tfr] = (1 - dr) * top + dr * bottom
    lo, hi = image.min(), image.max()
    if not (lo <= cval <= hi):
        lo, hi = min(lo, cval), max(hi, cval)
    np.clip(out, lo, hi, out=out)
    return out


# --------------------------------------------------------------------------
# ellipse.LsqEllipse  (Halir & Flusser direct least squares)
# --------------------------------------------------------------------------
class LsqEllipse:
    def fit(self, X):
        X = np.asarray(X, dtype=float)
        x, y = X.T
        D1 = np.vstack([x ** 2, x * y, y ** 2]).T
        D2 = np.vstack([x, y, np.ones_like(x)]).T
        S1 = D1.T @ D1
        S2 = D1.T @ D2
        S3 = D2.T @ D2
        C1 = np.array([[0., 0., 2.], [0., -1., 0.], [2., 0., 0.]])
        M = np.linalg.inv(C1) @ (S1 - S2 @ np.linalg.inv(S3) @ S2.T)
        _, eigvec = np.linalg.eig(M)
        cond = 4 * np.multiply(eigvec[0, :], eigvec[2, :]) - np.power(eigvec[1, :], 2)
        a1 = eigvec[:, np.nonzero(cond > 0)[0]]
        a2 = np.linalg.inv(-S3) @ S2.T @ a1
        self.coef_ = np.vstack([a1, a2])
        return self

    @property
    def coefficients(self):
        return np.asarray(self.coef_).ravel()

    def as_parameters(self):
        """centre, width, height, phi with ``width`` the semi-axis lying at angle
        ``phi`` from the first coordinate axis (the convention the reference's
        get_correction_matrix / two_step need)."""
        a = self.coefficients[0]
        b = self.coefficients[1] / 2.
        c = self.coefficients[2]
        d = self.coefficients[3] / 2.
        f = self.coefficients[4] / 2.
        g = self.coefficients[5]
        x0 = (c * d - b * f) / (b ** 2 - a * c)
        y0 = (a * f - b * d) / (b ** 2 - a * c)
        numerator = 2 * (a * f ** 2 + c * d ** 2 + g * b ** 2 - 2 * b * d * f - a * c * g)
        root = np.sqrt((a - c) ** 2 + 4 * b ** 2)
        denominator1 = (b ** 2 - a * c) * (root - (c + a))
        denominator2 = (b ** 2 - a * c) * (-root - (c + a))
        width = np.sqrt(numerator / denominator1)
        height = np.sqrt(numerator / denominator2)
        if b == 0 and a < c:
            phi = 0.0
        elif b == 0 and a > c:
            phi = np.pi / 2
        elif b != 0 and a < c:
            phi = 0.5 * np.arctan(2 * b / (a - c))
        elif b != 0 and a > c:
            phi = 0.5 * (np.pi + np.arctan(2 * b / (a - c)))
        else:
            phi = 0.0
        return (x0, y0), width, height, phi

    def return_fit(self, n_points=None, t=None):
        if t is None:
            t = np.linspace(0, 2 * np.pi, n_points)
        center, width, height, phi = self.as_parameters()
        x = center[0] + width * np.cos(t) * np.cos(phi) - height * np.sin(t) * np.sin(phi)
        y = center[1] + width * np.cos(t) * np.sin(phi) + height * np.sin(t) * np.cos(phi)
        return np.c_[x, y]
