"""Limb detection and ellipse fit that precede the circularisation warp
(reference ellipse_to_circle.py:53-91 two_step, :148-291 get_flood_image /
get_edge_list, :294-314 ellipse_to_circle).  SURVEY.md 8(f)#1: this front end
works on a 4x-downscaled image and is host-side (NumPy / SciPy / OpenCV, the
reference's own dependencies); the 4x4 block sums come from the device
(shg_downscale4_sum) so the full-resolution disk never leaves HBM.

The reference calls scikit-image (canny, downscale_local_mean) and lsq-ellipse
(LsqEllipse) here; neither is a dependency of this package, so their published
algorithms are implemented below:
  canny       Gaussian smoothing normalised by the smoothed mask, Sobel
              gradients, interpolated non-maximum suppression, hysteresis by
              8-connected components (skimage.feature.canny, float image);
  LsqEllipse  Halir & Flusser numerically stable direct least squares.
"""
from __future__ import annotations

import math

import cv2
import numpy as np
from numpy.polynomial import Polynomial
from scipy import ndimage as ndi
from scipy.spatial import ConvexHull

from .geometry import correction_matrix

NUM_REG = 2       # largest edge regions kept (reference ellipse_to_circle.py:31)
EDGE_CROP = 0.017  # fraction of the edge extent dropped at top and bottom (:277)


# ------------------------------------------------------------------ threshold
def flood_image(image):
    """Binarise the blurred image at a threshold taken from the valley of its
    brightness histogram (reference ellipse_to_circle.py:148-228)."""
    rows, cols = image.shape
    fallback = 0.9 * np.sum(image) / (rows * cols)
    bw = int(rows * 0.01)
    blurred = cv2.blur(image, ksize=(bw, bw))
    ceiling = np.percentile(blurred, 99)
    counts, edges = np.histogram(blurred[blurred < ceiling], bins=20)
    c0, c1, c2, c3 = Polynomial.fit(edges[1:], counts, 3).convert().coef
    disc = 4 * c2 ** 2 - 12 * c3 * c1
    valley = (-2 * c2 + np.sqrt(disc)) / (6 * c3) if disc >= 0 else fallback
    start = -1
    for i in range(len(edges) - 1):
        if edges[i] <= valley < edges[i + 1]:
            start = i
    if start < 0:
        level = fallback
    else:
        i = start
        while 0 < i < len(edges) - 2:           # walk downhill to the local minimum of the histogram
            if counts[i - 1] < counts[i]:
                i -= 1
            elif counts[i + 1] < counts[i]:
                i += 1
            else:
                break
        level = edges[i - 1] if i >= 1 else edges[i]
    return np.where(blurred < level, 0.0, 65000.0)


# ---------------------------------------------------------------------- canny
def _smooth(img, sigma):
    return ndi.gaussian_filter(img, sigma, mode='constant', cval=0.0, truncate=4.0)


def canny_edges(image, sigma, low, high):
    image = np.asarray(image, dtype=np.float64)
    rows, cols = image.shape
    weight = _smooth(np.ones_like(image), sigma) + np.finfo(np.float64).eps
    smoothed = _smooth(image, sigma) / weight
    gj = ndi.sobel(smoothed, axis=1)
    gi = ndi.sobel(smoothed, axis=0)
    mag = np.sqrt(gi * gi + gj * gj)

    interior = np.zeros(image.shape, bool)
    interior[1:-1, 1:-1] = True
    cand = interior & (mag >= low)
    ii, jj = np.nonzero(cand)                   # work on the candidate pixels only
    a, b, m = gi[ii, jj], gj[ii, jj], mag[ii, jj]
    aa, ab = np.abs(a), np.abs(b)
    same = ((a >= 0) & (b >= 0)) | ((a <= 0) & (b <= 0))
    opp = ((a <= 0) & (b >= 0)) | ((a >= 0) & (b <= 0))
    steep = aa > ab                              # gradient closer to the row axis
    flat = aa < ab
    # neighbour offsets (di, dj) of the two pixels that bracket the gradient direction
    # on the + side; the - side is the mirror image.
    case = np.where(same, np.where(steep, 0, 1), np.where(flat, 2, 3))
    d1i = np.array([1, 0, 0, -1])[case]
    d1j = np.array([0, 1, 1, 0])[case]
    d2i = np.array([1, 1, -1, -1])[case]
    d2j = np.array([1, 1, 1, 1])[case]
    with np.errstate(divide='ignore', invalid='ignore'):
        w = np.where((case == 0) | (case == 3), ab / aa, aa / ab)
    plus = mag[ii + d2i, jj + d2j] * w + mag[ii + d1i, jj + d1j] * (1.0 - w)
    minus = mag[ii - d2i, jj - d2j] * w + mag[ii - d1i, jj - d1j] * (1.0 - w)
    with np.errstate(invalid='ignore'):
        keep = (same | opp) & (plus <= m) & (minus <= m)
    thin = np.zeros_like(mag)
    thin[ii[keep], jj[keep]] = m[keep]

    weak = thin > 0
    labels, count = ndi.label(weak, np.ones((3, 3), bool))
    if count == 0:
        return weak
    strong = np.zeros(count + 1, bool)
    strong[np.unique(labels[weak & (thin >= high)])] = True
    strong[0] = False
    return strong[labels]


# ----------------------------------------------------------------- edge points
def limb_points(image, sigma=2.0):
    """Edge pixels (row, col) of the solar limb in a float image scaled to [0, 1)
    (reference get_edge_list, ellipse_to_circle.py:231-291)."""
    while True:
        if sigma <= 0:
            raise Exception('ERROR: could not find any edges')
        low = np.median(cv2.blur(image, ksize=(5, 5))) / 10
        edges = canny_edges(flood_image(image), sigma, low, low * 1.5)
        labelled, n_regions = ndi.label(edges, structure=np.ones((3, 3)))
        if n_regions:
            break
        sigma -= 0.5                               # retry with less blur (:254-256)
    sizes = np.bincount(labelled.ravel(), minlength=n_regions + 1)
    sizes[0] = -1
    # the reference picks labels by list.index of the sorted sizes: ties resolve to the lowest label
    ranked = sorted(sizes.tolist(), reverse=True)[:min(n_regions, NUM_REG)]
    chosen = [sizes.tolist().index(s) for s in ranked]
    pts = np.argwhere(np.isin(labelled, chosen))
    hull = pts[ConvexHull(pts).vertices]
    on_hull = np.zeros(edges.shape, bool)
    on_hull[hull[:, 0], hull[:, 1]] = True
    kept = np.zeros(edges.shape, bool)
    for lab in chosen:
        region = labelled == lab
        if np.any(region & on_hull):
            kept |= region
    lo, hi = pts[:, 0].min(), pts[:, 0].max()
    span = hi - lo
    band = np.zeros(edges.shape, bool)
    band[int(lo + span * EDGE_CROP):int(hi - span * EDGE_CROP), :] = True
    return np.argwhere(kept & band).astype(float), np.argwhere(edges)


# ---------------------------------------------------------------- ellipse fit
def fit_ellipse(points):
    """Halir-Flusser direct least-squares ellipse through (row, col) points.
    Returns (centre, width, height, phi) in LsqEllipse.as_parameters() terms:
    `width` is the semi-axis lying at angle phi from the first coordinate axis."""
    x, y = np.asarray(points, dtype=float).T
    D1 = np.stack([x * x, x * y, y * y], axis=1)
    D2 = np.stack([x, y, np.ones_like(x)], axis=1)
    S1, S2, S3 = D1.T @ D1, D1.T @ D2, D2.T @ D2
    C1 = np.array([[0., 0., 2.], [0., -1., 0.], [2., 0., 0.]])
    M = np.linalg.inv(C1) @ (S1 - S2 @ np.linalg.inv(S3) @ S2.T)
    _, vec = np.linalg.eig(M)
    good = 4 * vec[0] * vec[2] - vec[1] ** 2 > 0
    a1 = vec[:, np.nonzero(good)[0]]
    a2 = np.linalg.inv(-S3) @ S2.T @ a1
    A, B, C, D, F, G = np.vstack([a1, a2]).ravel()[:6]
    b, d, f = B / 2., D / 2., F / 2.
    den = b * b - A * C
    x0, y0 = (C * d - b * f) / den, (A * f - b * d) / den
    num = 2 * (A * f * f + C * d * d + G * b * b - 2 * b * d * f - A * C * G)
    root = np.sqrt((A - C) ** 2 + 4 * b * b)
    width = np.sqrt(num / (den * (root - (C + A))))
    height = np.sqrt(num / (den * (-root - (C + A))))
    if b == 0:
        phi = 0.0 if A < C else (np.pi / 2 if A > C else 0.0)
    elif A < C:
        phi = 0.5 * np.arctan(2 * b / (A - C))
    elif A > C:
        phi = 0.5 * (np.pi + np.arctan(2 * b / (A - C)))
    else:
        phi = 0.0
    return (x0, y0), width, height, phi


def ellipse_outline(center, width, height, phi, n_points=100):
    t = np.linspace(0, 2 * np.pi, n_points)
    return np.c_[center[0] + width * np.cos(t) * np.cos(phi) - height * np.sin(t) * np.sin(phi),
                 center[1] + width * np.cos(t) * np.sin(phi) + height * np.sin(t) * np.cos(phi)]


def two_step(points):
    """Fit, drop the points that lie well inside the first ellipse, refit, then
    relabel the axes so that |phi| <= pi/4 (reference ellipse_to_circle.py:62-91).
    Returns (centre, height, phi, ratio, kept points, outline)."""
    center, width, height, phi = fit_ellipse(points)
    mat, _ = correction_matrix(phi, height / width)
    resid = np.linalg.norm(mat @ (points - np.array(center)).T * height, axis=0) - 1
    kept = points[resid > -np.max(resid)]
    center, width, height, phi = fit_ellipse(kept)
    outline = ellipse_outline(center, width, height, phi)
    ratio = width / height
    for _ in range(2):
        if phi > math.pi / 4:
            phi -= math.pi / 2
            ratio = 1 / ratio
            height = height / ratio
        if phi < -math.pi / 4:
            phi += math.pi / 2
            ratio = 1 / ratio
            height = height / ratio
    return np.array(center), height, phi, ratio, kept, outline


# ------------------------------------------------- device-assisted limb search
def gaussian_weights(sigma, truncate=4.0):
    """scipy.ndimage._gaussian_kernel1d: weights at offsets 0..radius (symmetric)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius:]


def percentile_from_pair(a, b, n, q):
    """np.percentile(..., q) (method 'linear') given the two order statistics that
    bracket it: a = sorted[floor((n-1)q/100)], b = the next one."""
    virtual = (n - 1) * np.true_divide(q, 100)
    gamma = virtual - np.floor(virtual)
    diff = b - a
    out = a + diff * gamma
    if gamma >= 0.5:
        out = b - diff * (1 - gamma)
    return out


def flood_level(counts, edges, fallback):
    """Threshold of get_flood_image from the 20-bin histogram (reference :176-217)."""
    c0, c1, c2, c3 = Polynomial.fit(edges[1:], counts, 3).convert().coef
    disc = 4 * c2 ** 2 - 12 * c3 * c1
    valley = (-2 * c2 + np.sqrt(disc)) / (6 * c3) if disc >= 0 else fallback
    start = -1
    for i in range(len(edges) - 1):
        if edges[i] <= valley < edges[i + 1]:
            start = i
    if start < 0:
        return fallback
    i = start
    while 0 < i < len(edges) - 2:
        if counts[i - 1] < counts[i]:
            i -= 1
        elif counts[i + 1] < counts[i]:
            i += 1
        else:
            break
    return edges[i - 1] if i >= 1 else edges[i]


def _components(flat, cols):
    """8-connected components of a sorted array of flat pixel indices; labels are
    numbered 1.. in raster order of each component's first pixel (scipy.ndimage.label).
    A union-find over the sorted list in libshg (host code; ~10^4 points)."""
    import ctypes as C
    from ._lib import call
    flat = np.ascontiguousarray(flat, dtype=np.int64)
    labels = np.empty(len(flat), dtype=np.int32)
    count = C.c_int32(0)
    call('shg_label_points', flat.ctypes.data, len(flat), int(cols), labels.ctypes.data, C.byref(count))
    return int(count.value), labels.astype(np.int64)


def limb_points_device(eng, sums, sigma=2.0):
    """limb_points() with the image-sized work on the GPU.  `sums` is the int32
    (rows, cols) device tensor of 4x4 block sums; results are bit-identical to
    limb_points(sums * 2**-20) because every device step mirrors the operation
    order of the library call it replaces (csrc/limb.cu)."""
    rows, cols = sums.shape
    n = rows * cols
    scale_img = 2.0 ** -20
    fallback = 0.9 * (eng.sum_u32(sums) * scale_img) / (rows * cols)
    bw = int(rows * 0.01)
    scale = 1.0 / (bw * bw)
    box = eng.box_sum_u32(sums, bw, bw)

    def blurred(b, s):
        return (float(b) * scale_img) * s

    prev = int(np.floor((n - 1) * np.true_divide(99, 100)))
    b_lo, b_hi = eng.select_u32(box, [prev, min(prev + 1, n - 1)])
    ceiling = percentile_from_pair(blurred(b_lo, scale), blurred(b_hi, scale), n, 99)
    r_lo, r_hi = eng.blur_range(box, scale, ceiling)
    first, last = blurred(r_lo, scale), blurred(r_hi, scale)
    if first == last:
        first, last = first - 0.5, last + 0.5
    edges = np.linspace(first, last, 21)
    counts = eng.blur_hist(box, scale, ceiling, edges)
    level = flood_level(counts, edges, fallback)

    box5 = eng.box_sum_u32(sums, 5, 5)
    m_lo, m_hi = eng.select_u32(box5, [(n - 1) // 2, n // 2])
    s5 = 1.0 / 25
    median = blurred(m_lo, s5) if n % 2 else np.mean([blurred(m_lo, s5), blurred(m_hi, s5)])
    low = median / 10
    high = low * 1.5
    while True:
        if sigma <= 0:
            raise Exception('ERROR: could not find any edges')
        with eng.stage('ellipse_fit:canny(device)'):
            flat, mag = eng.canny_candidates(box, scale, level, gaussian_weights(sigma), low)
        if len(flat):
            count, lab = _components(flat, cols)
            strong = np.zeros(count + 1, bool)
            strong[np.unique(lab[mag >= high])] = True
            keep = strong[lab]
            flat_e = flat[keep]                                   # canny's edge pixels, raster order
            if len(flat_e):
                break
        sigma -= 0.5
    # labels of the surviving components, renumbered 1.. in raster order of their first pixel: what
    # scipy.ndimage.label(edges) would give (a component survives hysteresis whole, so its pixels and
    # the relative order of first pixels are unchanged)
    kept_labels = np.flatnonzero(strong)
    renum = np.zeros(count + 1, dtype=np.int64)
    renum[kept_labels] = np.arange(1, len(kept_labels) + 1)
    n_regions, lab = len(kept_labels), renum[lab[keep]]
    sizes = np.bincount(lab, minlength=n_regions + 1)
    sizes[0] = -1
    ranked = sorted(sizes.tolist(), reverse=True)[:min(n_regions, NUM_REG)]
    chosen = [sizes.tolist().index(s) for s in ranked]
    sel = np.isin(lab, chosen)
    flat_sel = flat_e[sel]
    pts = np.stack([flat_sel // cols, flat_sel % cols], axis=1)
    # hull vertices are extreme points, and every extreme point is the first or last pixel of its row:
    # the hull of those (<= 2 per row) has the same vertices as the hull of all points
    row_start = np.flatnonzero(np.diff(pts[:, 0], prepend=-1))
    row_last = np.append(row_start[1:] - 1, len(pts) - 1)
    ends = np.unique(np.concatenate([row_start, row_last]))
    hull_flat = flat_sel[ends[ConvexHull(pts[ends]).vertices]] if len(ends) >= 3 else \
        flat_sel[ConvexHull(pts).vertices]
    kept = np.zeros(len(flat_e), bool)
    for c in chosen:
        region = lab == c
        if np.isin(flat_e[region], hull_flat).any():
            kept |= region
    lo, hi = pts[:, 0].min(), pts[:, 0].max()
    span = hi - lo
    r_all = flat_e // cols
    band = (r_all >= int(lo + span * EDGE_CROP)) & (r_all < int(hi - span * EDGE_CROP))
    out = flat_e[kept & band]
    return (np.stack([out // cols, out % cols], axis=1).astype(float),
            np.stack([flat_e // cols, flat_e % cols], axis=1))


def fit_from_device(eng, sums):
    """fit_from_block_sums with the limb search on the GPU."""
    with eng.stage('ellipse_fit:limb_search'):
        pts, raw = limb_points_device(eng, sums)
    pts, raw = pts * 4, raw * 4
    with eng.stage('ellipse_fit:two_step(host)'):
        center, height, phi, ratio, kept, outline = two_step(pts)
    return np.array([center[1], center[0]]), height, phi, ratio, kept, raw, outline


def fit_from_block_sums(block_sums):
    """block_sums: (ceil(ih/4), ceil(N/4)) integer sums of 4x4 pixel blocks
    (DN units).  Returns (centre_xy, height, phi, ratio, kept points, raw edge
    points, outline) in full-resolution pixel coordinates."""
    small = np.asarray(block_sums, dtype=np.float64) * (1.0 / (16 * 65536))    # mean of image/65536: exact
    pts, raw = limb_points(small)
    pts, raw = pts * 4, raw * 4
    center, height, phi, ratio, kept, outline = two_step(pts)
    return np.array([center[1], center[0]]), height, phi, ratio, kept, raw, outline
