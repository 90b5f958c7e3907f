"""CPU tests: independent cross-checks of the third-party restatements in
oracle/thirdparty.py (scikit-image's warp / canny stages / downscale_local_mean
and lsq-ellipse's LsqEllipse are not installable here).

Each restatement is checked against a DIFFERENT implementation of the same
published operation that IS present in this image:

  warp (ellipse_to_circle.py:112-114)      scipy.ndimage.map_coordinates(order=1, mode='grid-constant')
  LsqEllipse (ellipse_to_circle.py:57-59)  cv2.fitEllipseDirect (OpenCV's Fitzgibbon direct fit) and a
                                           generalised-eigenvalue solve of the same constrained conic problem
  downscale_local_mean (:301)              cv2.resize(INTER_AREA) + a hand-padded case
  canny's gaussian / sobel (:245-250)      cv2.GaussianBlur / cv2.Sobel / cv2.sepFilter2D with matched borders
"""
import math

import cv2
import numpy as np
import pytest
import scipy.linalg
from scipy import ndimage as ndi

from oracle import shg_oracle as O
from oracle import thirdparty as T
from helpers import ALL_CASES, golden

PHI_RATIO = [(0.0, 1.0), (0.0, 1.22), (0.12, 1.22), (-0.2, 0.85), (0.5, 1.6), (-0.7, 2.3)]


# ------------------------------------------------------------------ (a) warp
def _warp_via_map_coordinates(img_u16, mat3, out_shape):
    """The reference's correct_image pixel work (ellipse_to_circle.py:112-118) with scipy's own
    linear interpolator standing in for skimage's: image / 65536 -> inverse map through mat3 ->
    order-1 interpolation with out-of-image taps = cval -> clip to the input range -> * 2**16 -> uint16."""
    image = img_u16 / 65536
    oh, ow = out_shape
    cc, rr = np.meshgrid(np.arange(ow, dtype=np.float64), np.arange(oh, dtype=np.float64))
    z = mat3[2, 0] * cc + mat3[2, 1] * rr + mat3[2, 2]
    x = (mat3[0, 0] * cc + mat3[0, 1] * rr + mat3[0, 2]) / z
    y = (mat3[1, 0] * cc + mat3[1, 1] * rr + mat3[1, 2]) / z
    cval = image[0, 0]
    out = ndi.map_coordinates(image, [y, x], order=1, mode='grid-constant', cval=cval, prefilter=False)
    out = np.clip(out, min(image.min(), cval), max(image.max(), cval))
    return (2 ** 16 * out).astype(np.uint16)


def _warp_via_restatement(img_u16, mat3, out_shape):
    image = img_u16 / 65536
    fixed = T.warp(image, T.ProjectiveTransform(matrix=mat3), output_shape=out_shape, cval=image[0, 0])
    return (2 ** 16 * fixed).astype(np.uint16)


@pytest.mark.parametrize('name', ALL_CASES)
def test_warp_restatement_matches_scipy_map_coordinates(name):
    """0 differing pixels between thirdparty.warp (what made the circ_* fixtures), the oracle's 1-D row
    form, and scipy's independent order-1 interpolator, for the fixture's own geometry and six more."""
    g = golden(name)
    disk = g['disk1']                                   # shift 0
    phi0 = 0.0 if math.isnan(float(g['slant'])) else math.radians(float(g['slant']))
    pairs = [(phi0, float(g['ratio']))] + PHI_RATIO
    for phi, ratio in pairs:
        _, mat3, out_shape, _, _ = O.warp_geometry(disk.shape, phi, ratio)
        want = _warp_via_map_coordinates(disk, mat3, out_shape)
        got2d = _warp_via_restatement(disk, mat3, out_shape)
        got1d, _ = O.warp_rows(disk, phi, ratio)
        assert got2d.shape == want.shape == got1d.shape
        # the 2-D restatement and the 1-D row form are the same arithmetic: identical
        assert np.array_equal(got2d, got1d), (name, phi, ratio)
        # scipy evaluates the same two-tap interpolation with its own operation order, so a value that lands
        # within an ulp of an integer can truncate the other way: at most 1 DN, on at most 1 pixel in 10^5
        # (measured over these 35 warps: 1 pixel of 4.0 M)
        delta = np.abs(got2d.astype(np.int32) - want.astype(np.int32))
        assert delta.max() <= 1, (name, phi, ratio)
        assert np.count_nonzero(delta) <= max(1, delta.size // 100000), (name, phi, ratio, np.count_nonzero(delta))


def test_warp_fixture_is_what_scipy_gives():
    """The committed circ_* fixtures themselves (not just today's restatement) against scipy."""
    total = pixels = 0
    for name in ALL_CASES:
        g = golden(name)
        phi = 0.0 if math.isnan(float(g['slant'])) else math.radians(float(g['slant']))
        ratio = float(g['ratio'])
        shifts = [int(s) for s in g['shift']]
        for sh in (int(s) for s in g['shift_requested']):
            disk = g['disk%d' % shifts.index(sh)]
            _, mat3, out_shape, _, _ = O.warp_geometry(disk.shape, phi, ratio)
            via_scipy = _warp_via_map_coordinates(disk, mat3, out_shape)
            delta = np.abs(via_scipy.astype(np.int32) - g['circ_%d' % sh].astype(np.int32))
            assert delta.max() <= 1 and np.count_nonzero(delta) <= 1, (name, sh, np.count_nonzero(delta))
            total += np.count_nonzero(delta)
            pixels += delta.size
    assert total <= 2, (total, pixels)            # measured: 1 pixel of ~1 M (a last-ulp truncation, see above)


# ------------------------------------------------------------ (b) LsqEllipse
def _random_ellipse_points(rng, n=None):
    n = n or int(rng.integers(40, 400))
    cx, cy = rng.uniform(-300, 1500, size=2)
    a = rng.uniform(40, 900)
    b = a * rng.uniform(0.25, 0.97)
    th = rng.uniform(-math.pi / 2, math.pi / 2)
    t0 = rng.uniform(0, 2 * math.pi)
    arc = rng.uniform(1.2 * math.pi, 2 * math.pi)          # limb arcs, not always a full ellipse
    t = t0 + arc * rng.random(n)
    noise = rng.normal(0, 0.002 * b, size=(2, n))
    x = cx + a * np.cos(t) * math.cos(th) - b * np.sin(t) * math.sin(th) + noise[0]
    y = cy + a * np.cos(t) * math.sin(th) + b * np.sin(t) * math.cos(th) + noise[1]
    return np.stack([x, y], axis=1), (cx, cy, a, b, th)


def _canonical(center, width, height, phi):
    """(cx, cy, major, minor, angle of the major axis in [0, pi)) of an as_parameters() result."""
    if width >= height:
        major, minor, ang = width, height, phi
    else:
        major, minor, ang = height, width, phi + math.pi / 2
    return center[0], center[1], major, minor, ang % math.pi


def _angle_close(a, b, tol):
    d = abs(a - b) % math.pi
    return min(d, math.pi - d) < tol


def _conic_generalised_eig(points):
    """min a'Sa subject to a'Ca = 1 (4ac - b^2 = 1) as the generalised eigenproblem S a = l C a on
    centred / scaled data -- the textbook Fitzgibbon form, solved without the Halir-Flusser block split."""
    p = np.asarray(points, dtype=float)
    m = p.mean(axis=0)
    s = np.abs(p - m).max()
    x, y = ((p - m) / s).T
    D = np.stack([x * x, x * y, y * y, x, y, np.ones_like(x)], axis=1)
    S = D.T @ D
    Cm = np.zeros((6, 6))
    Cm[0, 2] = Cm[2, 0] = 2
    Cm[1, 1] = -1
    w, v = scipy.linalg.eig(S, Cm)
    best = None
    for k in range(6):
        a = np.real(v[:, k])
        if not np.isfinite(w[k]) or abs(np.imag(w[k])) > 1e-9:
            continue
        if 4 * a[0] * a[2] - a[1] ** 2 > 0 and np.real(w[k]) > 0:
            if best is None or np.real(w[k]) < best[0]:
                best = (np.real(w[k]), a)
    A, B, Cc, Dd, E, F = best[1]
    # undo the normalisation: x = (X - mx) / s
    mx, my = m
    a0 = A / s ** 2
    b0 = B / s ** 2
    c0 = Cc / s ** 2
    d0 = Dd / s - 2 * A * mx / s ** 2 - B * my / s ** 2
    e0 = E / s - 2 * Cc * my / s ** 2 - B * mx / s ** 2
    f0 = F + A * mx ** 2 / s ** 2 + Cc * my ** 2 / s ** 2 + B * mx * my / s ** 2 - Dd * mx / s - E * my / s
    return np.array([a0, b0, c0, d0, e0, f0])


def _params_from_conic(co):
    reg = T.LsqEllipse()
    reg.coef_ = np.asarray(co, dtype=float).reshape(6, 1)
    return reg.as_parameters()


def test_lsq_ellipse_matches_generalised_eigen_solution_and_opencv():
    rng = np.random.default_rng(20261017)
    checked_cv = 0
    dev_cv = []
    for trial in range(200):
        pts, truth = _random_ellipse_points(rng)
        center, width, height, phi = T.LsqEllipse().fit(pts).as_parameters()
        got = _canonical(center, float(width), float(height), float(phi))
        # (1) the same constrained least-squares problem solved as one 6x6 generalised eigenproblem
        ref = _canonical(*_params_from_conic(_conic_generalised_eig(pts)))
        scale = ref[2]
        assert abs(got[0] - ref[0]) < 1e-6 * scale and abs(got[1] - ref[1]) < 1e-6 * scale, trial
        assert abs(got[2] / ref[2] - 1) < 1e-6 and abs(got[3] / ref[3] - 1) < 1e-6, trial
        assert _angle_close(got[4], ref[4], 1e-5 / max(1e-3, 1 - ref[3] / ref[2])), trial
        # (2) the truth the points were drawn from (noise 0.2 % of the minor axis)
        cx, cy, a, b, th = truth
        assert abs(got[0] - cx) < 0.02 * a and abs(got[1] - cy) < 0.02 * a, trial
        assert abs(got[2] / a - 1) < 0.02 and abs(got[3] / b - 1) < 0.02, trial
        assert _angle_close(got[4], th % math.pi, 0.02 / max(0.02, 1 - b / a)), trial
        # (3) OpenCV's direct ellipse fit: an independent implementation of the same algebraic fit.  It
        # returns ((cx, cy), (full axis along `angle`, full axis across it), angle in degrees).
        (ox, oy), (d1, d2), ang = cv2.fitEllipseDirect(pts.astype(np.float32))
        if not (np.isfinite([ox, oy, d1, d2]).all() and d1 > 0 and d2 > 0):
            continue
        ocv = _canonical((ox, oy), d1 / 2, d2 / 2, math.radians(ang))
        # OpenCV solves the same algebraic problem in its own (partly single-precision) arithmetic: on short
        # noisy arcs the two drift apart by up to a few 1e-3; the bulk agrees to ~1e-5 (checked below)
        assert abs(got[0] - ocv[0]) < 1e-2 * scale and abs(got[1] - ocv[1]) < 1e-2 * scale, trial
        assert abs(got[2] / ocv[2] - 1) < 1e-2 and abs(got[3] / ocv[3] - 1) < 1e-2, trial
        assert _angle_close(got[4], ocv[4], 1e-2 / max(1e-2, 1 - ocv[3] / ocv[2])), trial
        dev_cv.append(max(abs(got[2] / ocv[2] - 1), abs(got[3] / ocv[3] - 1)))
        checked_cv += 1
    assert np.median(dev_cv) < 1e-4 and np.percentile(dev_cv, 90) < 1e-3, np.percentile(dev_cv, [50, 90, 100])
    assert checked_cv >= 150


def test_lsq_ellipse_phi_branch_convention():
    """`width` is the semi-axis lying at angle `phi` from the FIRST coordinate axis -- the convention the
    reference's two_step / get_correction_matrix rely on: rebuilding the outline from
    (centre, width, height, phi) must put the points back on the ellipse, for both a<c and a>c branches."""
    rng = np.random.default_rng(7)
    seen = set()
    for trial in range(60):
        pts, truth = _random_ellipse_points(rng, n=300)
        reg = T.LsqEllipse().fit(pts)
        center, width, height, phi = reg.as_parameters()
        a, c = reg.coefficients[0], reg.coefficients[2]
        seen.add(bool(a < c))
        # distance of every input point to the fitted outline, in the frame of the fitted axes
        d = pts - np.asarray(center)
        u = d[:, 0] * math.cos(phi) + d[:, 1] * math.sin(phi)          # along `width`
        v = -d[:, 0] * math.sin(phi) + d[:, 1] * math.cos(phi)         # along `height`
        r = (u / width) ** 2 + (v / height) ** 2
        assert np.abs(r - 1).max() < 0.05, trial
        out = reg.return_fit(n_points=50)
        d = out - np.asarray(center)
        u = d[:, 0] * math.cos(phi) + d[:, 1] * math.sin(phi)
        v = -d[:, 0] * math.sin(phi) + d[:, 1] * math.cos(phi)
        assert np.abs((u / width) ** 2 + (v / height) ** 2 - 1).max() < 1e-9
    assert seen == {True, False}


def test_two_step_normalises_phi_and_ratio():
    """After two_step's swap loop phi is within +-pi/4 and ratio = (axis along rows) / (axis along frames),
    so stretching the frame axis by `ratio` makes the disk round (ellipse_to_circle.py:79-90)."""
    rng = np.random.default_rng(3)
    for trial in range(40):
        ry, rx = rng.uniform(300, 600), rng.uniform(300, 600)      # semi-axes along rows / along frames
        tilt = rng.uniform(-0.2, 0.2)
        t = rng.uniform(0, 2 * math.pi, 500)
        rows = 700 + ry * np.cos(t) * math.cos(tilt) - rx * np.sin(t) * math.sin(tilt)
        cols = 650 + ry * np.cos(t) * math.sin(tilt) + rx * np.sin(t) * math.cos(tilt)
        pts = np.stack([rows, cols], axis=1) + rng.normal(0, 0.3, size=(500, 2))
        center, height, phi, ratio, kept = O.two_step(pts)
        assert -math.pi / 4 <= phi <= math.pi / 4
        assert abs(ratio / (ry / rx) - 1) < 0.01, (trial, ratio, ry / rx)
        assert abs(center[0] - 700) < 2 and abs(center[1] - 650) < 2


# --------------------------------------------------- (c) downscale_local_mean
def test_downscale_local_mean_matches_inter_area():
    rng = np.random.default_rng(5)
    for shape in ((64, 48), (400, 1000), (128, 4), (4, 4)):
        img = rng.random(shape)
        got = T.downscale_local_mean(img, (4, 4))
        want = cv2.resize(img, (shape[1] // 4, shape[0] // 4), interpolation=cv2.INTER_AREA)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    # integer-valued input: the block mean is exact in both
    img = rng.integers(0, 65536, size=(96, 80)).astype(np.float64)
    assert np.array_equal(T.downscale_local_mean(img, (4, 4)),
                          img.reshape(24, 4, 20, 4).sum(axis=(1, 3)) / 16)


def test_downscale_local_mean_pads_with_zeros():
    """Shapes that are not multiples of 4: skimage's block_reduce pads with cval=0 at the bottom / right and
    still divides by 16 (so edge blocks are darker) -- a padded case worked by hand."""
    img = np.arange(1, 7 * 6 + 1, dtype=np.float64).reshape(7, 6)
    got = T.downscale_local_mean(img, (4, 4))
    assert got.shape == (2, 2)
    assert got[0, 0] == img[0:4, 0:4].sum() / 16
    assert got[0, 1] == img[0:4, 4:6].sum() / 16           # 2 real columns + 2 columns of zeros
    assert got[1, 0] == img[4:7, 0:4].sum() / 16           # 3 real rows + 1 row of zeros
    assert got[1, 1] == img[4:7, 4:6].sum() / 16
    rng = np.random.default_rng(6)
    big = rng.random((203, 101))
    pad = np.zeros((204, 104))
    pad[:203, :101] = big
    np.testing.assert_allclose(T.downscale_local_mean(big, (4, 4)),
                               cv2.resize(pad, (26, 51), interpolation=cv2.INTER_AREA), rtol=0, atol=1e-12)


# ------------------------------------------------ (d) canny: gaussian + sobel
def _gauss_kernel(sigma, truncate=4.0):
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


@pytest.mark.parametrize('sigma', [2.0, 1.5, 1.0, 0.5])
def test_canny_gaussian_stage_matches_opencv(sigma):
    """skimage's canny smooths with scipy.ndimage.gaussian_filter(mode='constant') (zero padding, kernel
    radius int(4 sigma + 0.5)); OpenCV's separable filter with the same taps and BORDER_CONSTANT agrees to
    1e-12, and so does its own GaussianBlur kernel generator."""
    rng = np.random.default_rng(8)
    img = np.where(rng.random((90, 130)) < 0.4, 0.0, 65000.0)
    img[30:60, 40:90] = 65000.0
    got = T._gaussian(img, sigma)
    k = _gauss_kernel(sigma)
    want = cv2.sepFilter2D(img, cv2.CV_64F, k, k, borderType=cv2.BORDER_CONSTANT)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * 65000)
    size = len(k)
    want2 = cv2.GaussianBlur(img, (size, size), sigmaX=sigma, sigmaY=sigma, borderType=cv2.BORDER_CONSTANT)
    np.testing.assert_allclose(got, want2, rtol=0, atol=1e-12 * 65000)
    np.testing.assert_allclose(cv2.getGaussianKernel(size, sigma).ravel(), k, rtol=0, atol=1e-15)


def test_canny_sobel_stage_matches_opencv():
    """ndi.sobel (mode 'reflect' = OpenCV's BORDER_REFLECT, the edge pixel repeated) against cv2.Sobel."""
    rng = np.random.default_rng(9)
    img = rng.random((70, 110)) * 65000
    np.testing.assert_allclose(ndi.sobel(img, axis=1), cv2.Sobel(img, cv2.CV_64F, 1, 0, ksize=3,
                                                                  borderType=cv2.BORDER_REFLECT),
                               rtol=0, atol=1e-12 * 65000 * 8)
    np.testing.assert_allclose(ndi.sobel(img, axis=0), cv2.Sobel(img, cv2.CV_64F, 0, 1, ksize=3,
                                                                  borderType=cv2.BORDER_REFLECT),
                               rtol=0, atol=1e-12 * 65000 * 8)


def test_canny_restatement_finds_the_same_limb_as_opencv_canny():
    """End-to-end sanity of the canny restatement on the binarised flood image of a fixture: the edge
    pixels lie on the 0 / 65000 boundary (within the 2-sigma smoothing radius), form a thin closed curve,
    and OpenCV's own Canny of the same smoothed image marks the same boundary (every restated edge pixel
    has an OpenCV edge pixel within 2 px and vice versa for >= 95 %)."""
    g = golden('ser16_rot')
    image = g['disk0'] / 65536
    small = T.downscale_local_mean(image, (4, 4))
    flood = O.flood_image(small)
    low = np.median(cv2.blur(small, ksize=(5, 5))) / 10
    edges = T.canny(flood, sigma=2.0, low_threshold=low, high_threshold=low * 1.5)
    assert edges.any()
    boundary = ndi.binary_dilation(flood > 0, iterations=1) & ~ndi.binary_erosion(flood > 0, iterations=1,
                                                                                  border_value=1)
    near = ndi.binary_dilation(boundary, iterations=3)
    assert np.all(near[edges]), 'restated canny marks pixels away from the flood boundary'
    # thin: at most ~2 pixels per boundary pixel row/column crossing
    assert edges.sum() <= 2.5 * boundary.sum()
    k = _gauss_kernel(2.0)
    sm = cv2.sepFilter2D(flood, cv2.CV_64F, k, k, borderType=cv2.BORDER_CONSTANT)
    sm8 = np.clip(sm / 65000 * 255, 0, 255).astype(np.uint8)
    ocv = cv2.Canny(sm8, 5, 10, L2gradient=True) > 0
    ocv[:1] = ocv[-1:] = False
    ocv[:, :1] = ocv[:, -1:] = False
    d_to_ocv = ndi.distance_transform_edt(~ocv)
    d_to_ours = ndi.distance_transform_edt(~edges)
    inner = np.zeros_like(edges)
    inner[4:-4, 4:-4] = True                               # zero padding vs replicate differ at the frame border
    assert np.mean(d_to_ocv[edges & inner] <= 2.0) >= 0.95
    assert np.mean(d_to_ours[ocv & inner] <= 2.0) >= 0.95
