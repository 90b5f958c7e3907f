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


def thin_edges(image, sigma, low):
    """Canny up to and including the interpolated non-maximum suppression: (flat indices ascending,
    magnitudes) of the thin-edge pixels with magnitude >= low -- the list csrc/limb.cu produces on the GPU."""
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
        keep = (same | opp) & (plus <= m) & (minus <= m) & (m > 0)
    return ii[keep] * cols + jj[keep], m[keep]


def canny_edges(image, sigma, low, high):
    image = np.asarray(image, dtype=np.float64)
    flat, m = thin_edges(image, sigma, low)
    thin = np.zeros(image.shape)
    thin.ravel()[flat] = m

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
_C1_INV = np.linalg.inv(np.array([[0., 0., 2.], [0., -1., 0.], [2., 0., 0.]]))


def fit_ellipse(points):
    """Halir-Flusser direct least-squares ellipse through (row, col) points.
    Returns (centre, width, height, phi) in LsqEllipse.as_parameters() terms:
    `width` is the semi-axis lying at angle phi from the first coordinate axis."""
    from ._lib import call
    pts = np.ascontiguousarray(points, dtype=np.float64)
    S = np.empty((6, 6))
    call('shg_conic_scatter', pts.ctypes.data, len(pts), S.ctypes.data)    # sum of d d^T, d = [x^2, xy, y^2, x, y, 1]
    S1, S2, S3 = S[:3, :3], S[:3, 3:], S[3:, 3:]
    M = _C1_INV @ (S1 - S2 @ np.linalg.inv(S3) @ S2.T)
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


class _LazyOutline:
    """The fitted ellipse as 100 points, computed when somebody looks at it (only the diagnostic plot does)."""

    def __init__(self, *params):
        self.params, self._pts = params, None

    def __array__(self, dtype=None, copy=None):
        if self._pts is None:
            self._pts = ellipse_outline(*self.params)
        return self._pts if dtype is None else self._pts.astype(dtype)

    def __getitem__(self, idx):
        return np.asarray(self)[idx]


def two_step(points):
    """Fit, drop the points that lie well inside the first ellipse, refit, then
    relabel the axes so that |phi| <= pi/4 (reference ellipse_to_circle.py:62-91).
    Returns (centre, height, phi, ratio, kept points, outline)."""
    center, width, height, phi = fit_ellipse(points)
    mat, _ = correction_matrix(phi, height / width)
    resid = np.linalg.norm(mat @ (points - np.array(center)).T * height, axis=0) - 1
    kept = points[resid > -np.max(resid)]
    center, width, height, phi = fit_ellipse(kept)
    outline = _LazyOutline(center, width, height, phi)
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


def _hull_vertices(x, y):
    """Indices of the convex-hull vertices of integer points (the set scipy.spatial.ConvexHull(...).vertices
    reports; raises like Qhull when the points do not span a plane).  Exact integer monotone chain in libshg:
    Qhull took 1.4 ms on the ~1600 row ends of a config-5 limb, this takes ~20 us."""
    import ctypes as C
    from ._lib import call
    xy = np.ascontiguousarray(np.stack([x, y], axis=1), dtype=np.int64)
    out = np.empty(len(xy), dtype=np.int64)
    m = C.c_int64(0)
    call('shg_hull_vertices', xy.ctypes.data, len(xy), out.ctypes.data, C.byref(m))
    if m.value < 3:
        raise Exception('ERROR: the limb pixels found do not span an area (convex hull is degenerate)')
    return out[:m.value]


def limb_points_device(eng, sums, sigma=2.0, chained=None):
    """limb_points() with the image-sized work on the GPU.  `sums` is the int32
    (rows, cols) device tensor of 4x4 block sums; results are bit-identical to
    limb_points(sums * 2**-20) because every device step mirrors the operation
    order of the library call it replaces (csrc/limb.cu).

    chained (default; SHG_LIMB_STEPWISE=1 selects the other): the threshold search runs as one queue of
    kernels with its scalars kept on the device and two blocking read-backs in total; otherwise every
    scalar is read back as soon as it exists (~15 round trips; the original formulation, kept as the
    cross-check of the chained one)."""
    import os
    if chained is None:
        chained = not os.environ.get('SHG_LIMB_STEPWISE')
    rows, cols = sums.shape
    n = rows * cols
    scale_img = 2.0 ** -20
    bw = int(rows * 0.01)
    scale = 1.0 / (bw * bw)
    s5 = 1.0 / 25

    def blurred(b, s):
        return (float(b) * scale_img) * s

    prev = int(np.floor((n - 1) * np.true_divide(99, 100)))
    ranks = [prev, min(prev + 1, n - 1), (n - 1) // 2, n // 2]
    if chained:
        virtual = (n - 1) * np.true_divide(99, 100)
        gamma = float(virtual - np.floor(virtual))
        box, f = eng.limb_front(sums, bw, ranks, gamma)
        b_lo, b_hi, m_lo, m_hi = f['stats']
        ceiling = percentile_from_pair(blurred(b_lo, scale), blurred(b_hi, scale), n, 99)
        r_lo, r_hi = f['range']
        first, last = blurred(r_lo, scale), blurred(r_hi, scale)
        if first == last:
            first, last = first - 0.5, last + 0.5
        edges = np.linspace(first, last, 21)
        counts = f['counts']
        if ceiling != f['ceiling'] or not np.array_equal(edges, f['edges']):
            # the device's percentile interpolation / np.linspace disagrees with NumPy's in the last bit
            # (never observed): redo the histogram with the host's numbers
            r_lo, r_hi = eng.blur_range(box, scale, ceiling)
            first, last = blurred(r_lo, scale), blurred(r_hi, scale)
            if first == last:
                first, last = first - 0.5, last + 0.5
            edges = np.linspace(first, last, 21)
            counts = eng.blur_hist(box, scale, ceiling, edges)
        fallback = 0.9 * (f['total'] * scale_img) / (rows * cols)
    else:
        fallback = 0.9 * (eng.sum_u32(sums) * scale_img) / (rows * cols)
        box = eng.box_sum_u32(sums, bw, bw)
        b_lo, b_hi = eng.select_u32(box, ranks[:2])
        ceiling = percentile_from_pair(blurred(b_lo, scale), blurred(b_hi, scale), n, 99)
        r_lo, r_hi = eng.blur_range(box, scale, ceiling)
        first, last = blurred(r_lo, scale), blurred(r_hi, scale)
        if first == last:
            first, last = first - 0.5, last + 0.5
        edges = np.linspace(first, last, 21)
        counts = eng.blur_hist(box, scale, ceiling, edges)
        box5 = eng.box_sum_u32(sums, 5, 5)
        m_lo, m_hi = eng.select_u32(box5, ranks[2:])
    level = flood_level(counts, edges, fallback)
    median = blurred(m_lo, s5) if n % 2 else np.mean([blurred(m_lo, s5), blurred(m_hi, s5)])
    low = median / 10
    high = low * 1.5
    while True:
        if sigma <= 0:
            raise Exception('ERROR: could not find any edges')
        with eng.stage('ellipse_fit:canny(device)'):
            if chained:
                flat, mag = eng.limb_canny(box, scale, level, gaussian_weights(sigma), low)
            else:
                flat, mag = eng.canny_candidates(box, scale, level, gaussian_weights(sigma), low)
        sel = select_limb_pixels(flat, mag, high, cols)
        if sel is not None:
            return sel
        sigma -= 0.5


def select_limb_pixels(flat, mag, high, cols):
    """Host half of the limb search, on the sparse list of thin-edge pixels the device returns
    (`flat` ascending flat indices, `mag` their gradient magnitudes): canny's hysteresis, the two largest
    8-connected regions, the convex-hull test, the top / bottom crop (reference ellipse_to_circle.py:250-291).
    Returns (kept (row, col) float points, all edge points) or None when canny leaves no edge."""
    if not len(flat):
        return None
    count, lab = _components(flat, cols)
    strong = np.zeros(count + 1, bool)
    strong[lab[mag >= high]] = True
    keep = strong[lab]
    flat_e = flat[keep]                                       # canny's edge pixels, raster order
    if not len(flat_e):
        return None
    # labels of the surviving components, renumbered 1.. in raster order of their first pixel: what
    # scipy.ndimage.label(edges) would give (a component survives hysteresis whole, so its pixels and
    # the relative order of first pixels are unchanged)
    renum = np.cumsum(strong)                                 # strong[0] is False: kept labels become 1..
    n_regions = int(renum[-1])
    lab = renum[lab[keep]]
    sizes = np.bincount(lab, minlength=n_regions + 1)
    sizes[0] = -1
    # the reference picks labels by list.index of the sorted sizes: ties resolve to the lowest label (twice)
    top = np.sort(sizes)[::-1][:min(n_regions, NUM_REG)]
    chosen = [int(np.argmax(sizes == s_)) for s_ in top]
    sel = lab == chosen[0]
    for c in chosen[1:]:
        sel |= lab == c
    flat_sel = flat_e[sel]
    lab_sel = lab[sel]
    rows_sel = flat_sel // cols
    cols_sel = flat_sel - rows_sel * cols
    # hull vertices are extreme points, and every extreme point is the first or last pixel of its row:
    # the hull of those (<= 2 per row) has the same vertices as the hull of all points
    first = np.empty(len(rows_sel), bool)
    first[0] = True
    np.not_equal(rows_sel[1:], rows_sel[:-1], out=first[1:])
    last = np.empty(len(rows_sel), bool)
    last[-1] = True
    last[:-1] = first[1:]
    ends = np.flatnonzero(first | last)
    on_hull = ends[_hull_vertices(rows_sel[ends], cols_sel[ends])]
    touching = np.unique(lab_sel[on_hull])                     # regions with a pixel among the hull vertices
    lo, hi = int(rows_sel[0]), int(rows_sel[-1])               # raster order: first / last row of the selection
    span = hi - lo
    r_all = flat_e // cols
    band = (r_all >= int(lo + span * EDGE_CROP)) & (r_all < int(hi - span * EDGE_CROP))
    kept = np.isin(lab, touching) if len(touching) > 1 else lab == touching[0]
    out = flat_e[kept & band]
    r_out = out // cols
    return (np.stack([r_out, out - r_out * cols], axis=1).astype(float),
            np.stack([r_all, flat_e - r_all * cols], axis=1))


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
