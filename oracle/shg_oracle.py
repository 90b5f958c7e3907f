"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's frame-stack
reconstruction path, used as the parity checker.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module; the product (solex_ser_recon_en_b200/) never does
and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors.
This restatement is pinned by fixtures generated from the UNMODIFIED reference
modules imported in the build container (oracle/make_golden.py ->
tests/golden/*.npz, checked by tests/test_oracle.py), for every function except
the scikit-image / lsq-ellipse call sites, which are PARITY UNPINNED (those
packages are absent; see oracle/thirdparty.py).

Everything is stated in *raw file coordinates* where that is the natural form
for the device kernels: a raw frame is ``(H, W)``; when ``W > H`` the reference
rotates it (np.rot90) so that image pixel ``img[i, j] == raw[j, W-1-i]`` with
``ih = W`` slit positions and ``iw = H`` dispersion pixels.
"""
from __future__ import annotations

import math
import struct

import numpy as np

from . import thirdparty

SER_HEADER_BYTES = 178


# --------------------------------------------------------------------------
# a1/a2  frame source            /root/reference/video_reader.py:12-126
# --------------------------------------------------------------------------
def ser_info(path):
    """SER header fields the reference uses (video_reader.py:31-66): u32 LE at
    bytes 26/30/34/38 = Width/Height/PixelDepthPerPlane/FrameCount, payload at
    178, depth 8 -> uint8 else uint16, endianness flag ignored."""
    with open(path, 'rb') as f:
        hdr = f.read(SER_HEADER_BYTES)
    width, height, depth, count = struct.unpack_from('<4I', hdr, 26)
    dtype = np.uint8 if depth == 8 else np.uint16
    rotated = width > height                       # video_reader.py:84-91
    ih, iw = (width, height) if rotated else (height, width)
    return dict(width=width, height=height, depth=depth, n_frames=count, dtype=dtype,
                rotated=rotated, ih=ih, iw=iw)


def ser_stack(path):
    """Raw payload as an ``(N, H, W)`` memmap (no orientation, no scaling)."""
    info = ser_info(path)
    return np.memmap(path, dtype=info['dtype'], mode='r', offset=SER_HEADER_BYTES,
                     shape=(info['n_frames'], info['height'], info['width']))


def orient(raw_frame):
    """Raw frame -> the (ih, iw) uint16 image the reference hands out
    (video_reader.py:117-122): rot90 when W > H, 8-bit scaled by 256."""
    img = raw_frame
    if raw_frame.shape[1] > raw_frame.shape[0]:
        img = np.rot90(img)
    if img.dtype == np.uint8:
        img = img.astype(np.uint16) * 256
    return img


# --------------------------------------------------------------------------
# a3  mean / max frame           /root/reference/solex_util.py:174-188
# --------------------------------------------------------------------------
def raw_sum_max(stack, k0=0, k1=None):
    """Integer sum (uint64) and max of raw frames ``[k0, k1)`` in raw layout,
    in file units (8-bit data NOT yet scaled).  Partial results from frame
    ranges add / max together exactly -- this is what ranks all-reduce."""
    k1 = stack.shape[0] if k1 is None else k1
    s = np.zeros(stack.shape[1:], dtype=np.uint64)
    m = np.zeros(stack.shape[1:], dtype=stack.dtype)
    for a in range(k0, k1, 64):
        blk = np.asarray(stack[a:min(k1, a + 64)])
        s += blk.sum(axis=0, dtype=np.uint64)
        np.maximum(m, blk.max(axis=0), out=m)
    return s, m


def finalize_mean_max(raw_sum, raw_max, n_frames, eight_bit):
    """``(sum / N).astype(uint16)`` (solex_util.py:188) is an integer floor
    division for every N < 2**36; 8-bit frames are scaled by 256 *before*
    accumulation (video_reader.py:121-122), i.e. mean = floor(256*sum/N)."""
    scale = 256 if eight_bit else 1
    mean_raw = (raw_sum * np.uint64(scale)) // np.uint64(n_frames)
    max_raw = raw_max.astype(np.uint64) * np.uint64(scale)
    mean_raw = mean_raw.astype(np.uint16)
    max_raw = max_raw.astype(np.uint16)
    if raw_sum.shape[1] > raw_sum.shape[0]:
        return np.ascontiguousarray(np.rot90(mean_raw)), np.ascontiguousarray(np.rot90(max_raw))
    return mean_raw, max_raw


def mean_max(stack):
    s, m = raw_sum_max(stack)
    return finalize_mean_max(s, m, stack.shape[0], stack.dtype == np.uint8)


# --------------------------------------------------------------------------
# cv2.blur on uint16 as built in this image (opencv 4.13 box_filter.simd.hpp,
# ColumnSum<int, ushort>): exact integer box sum S with BORDER_REFLECT_101 and
# anchor k//2, then  rint(float32(S) * float32(1/(kw*kh)))  in the SIMD body
# (columns < W - W%8) and  rint(S * (1.0/(kw*kh)))  in double in the scalar tail.
# Checked against cv2.blur itself in tests/test_oracle.py.
# --------------------------------------------------------------------------
def _reflect101(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def box_sum(img, kw, kh):
    H, W = img.shape
    xs = _reflect101(np.arange(-(kw // 2), W - (kw // 2) + kw - 1), W)
    ys = _reflect101(np.arange(-(kh // 2), H - (kh // 2) + kh - 1), H)
    p = img[np.ix_(ys, xs)].astype(np.int64)
    c = np.zeros((p.shape[0] + 1, p.shape[1] + 1), np.int64)
    c[1:, 1:] = p.cumsum(0).cumsum(1)
    return c[kh:kh + H, kw:kw + W] - c[0:H, kw:kw + W] - c[kh:kh + H, 0:W] + c[0:H, 0:W]


def box_blur_u16(img, kw, kh):
    S = box_sum(img, kw, kh)
    scale = 1.0 / (kw * kh)
    out = np.rint(S.astype(np.float32) * np.float32(scale)).astype(np.int64)
    tail = img.shape[1] % 8
    if tail:
        out[:, -tail:] = np.rint(S[:, -tail:] * scale).astype(np.int64)
    return np.clip(out, 0, 65535).astype(np.uint16)


# --------------------------------------------------------------------------
# a4  slit extent                /root/reference/solex_util.py:165-172,223-227
# --------------------------------------------------------------------------
def slit_extent(max_img):
    blur = box_blur_u16(max_img, 5, 5)
    ymean = blur.sum(axis=1, dtype=np.int64) / blur.shape[1]      # == np.mean(blur, 1)
    where_sun = ymean > np.median(ymean) / 5
    lb = int(np.argmax(where_sun))
    ub = int(max_img.shape[0] - 1 - np.argmax(where_sun[::-1]))
    clip = int((ub - lb) * 0.05)
    return min(max_img.shape[0] - 1, lb + clip), max(0, ub - clip)


# --------------------------------------------------------------------------
# a5  line minima                /root/reference/solex_util.py:228-231,242
# --------------------------------------------------------------------------
def line_minima(mean_img, y1, y2):
    bwy = int((y2 - y1) * 0.01)
    if bwy < 1:
        raise ValueError('slit extent too small for the 1 % vertical blur (needs y2-y1 >= 100)')
    blur = box_blur_u16(mean_img, 25, bwy)
    min_intensity = 12 + np.argmin(blur[:, 12:-13], axis=1)       # -25//2 == -13
    min_sharp = np.argmin(mean_img, axis=1)
    return min_intensity.astype(np.int64), min_sharp.astype(np.int64)


# --------------------------------------------------------------------------
# a6  cubic fit                  /root/reference/solex_util.py:233-259
# --------------------------------------------------------------------------
def polyfit3(x, y):
    """Ascending-order cubic coefficients, as ``np.flip(np.polyfit(x, y, 3))``."""
    return np.flip(np.asarray(np.polyfit(x, y, 3), dtype='d'))


def polyval_asc(x, p):
    """Horner, ascending coefficients (numpy.polynomial.polynomial.polyval)."""
    x = np.asarray(x, dtype='d')
    c = p[3] + 0.0 * x
    for k in (2, 1, 0):
        c = p[k] + c * x
    return c


def mode_of_tenths(delta):
    """``values[argpartition(-counts, 2)[:2][0]]`` of ``around(delta, 1)``
    (solex_util.py:245-247): needs >= 3 bins; returns one of the two most
    populated bins as numpy's introselect leaves them."""
    values, counts = np.unique(np.around(delta, 1), return_counts=True)
    ind = np.argpartition(-counts, kth=2)[:2]
    return values[ind[0]]


def line_fit(min_intensity, min_sharp, y1, y2, ih):
    xs = np.arange(y1, y2)
    xd = np.asarray(xs, dtype='d')
    p1 = polyfit3(xs, min_intensity[y1:y2])
    delta = polyval_asc(xd, p1) - min_intensity[y1:y2]
    keep = np.abs(delta / np.std(delta)) < 3
    p2 = polyfit3(xs[keep], min_intensity[y1:y2][keep])
    delta_sharp = polyval_asc(xd, p2) - min_sharp[y1:y2]
    shift = mode_of_tenths(delta_sharp)
    mask_good = np.abs(delta_sharp - shift) < 5
    p3 = polyfit3(xs[mask_good], min_sharp[y1:y2][mask_good])
    curve = polyval_asc(np.arange(ih), p3)
    fl = np.floor(curve)
    fit = np.stack([fl, curve - fl, np.arange(ih, dtype='d'), curve], axis=1)
    return dict(p1=p1, p2=p2, p3=p3, keep=keep, mask_good=mask_good, shift=shift, fit=fit)


def mean_and_fit(stack):
    """compute_mean_return_fit (solex_util.py:191-274) without its side effects."""
    mean_img, max_img = mean_max(stack)
    y1, y2 = slit_extent(max_img)
    mi, ms = line_minima(mean_img, y1, y2)
    lf = line_fit(mi, ms, y1, y2, mean_img.shape[0])
    lf.update(mean_img=mean_img, max_img=max_img, y1=y1, y2=y2, min_intensity=mi, min_sharp=ms)
    return lf


# --------------------------------------------------------------------------
# a7  per-frame reconstruction   /root/reference/solex_util.py:93-144
# --------------------------------------------------------------------------
def recon_tables(fit, shifts, iw):
    """Left tap index per shift and the two weights (solex_util.py:113-123).
    Clipped indices keep the unclipped weights."""
    fit = np.asarray(fit)
    ih = fit.shape[0]
    il = np.empty((len(shifts), ih), dtype=np.int64)
    for s, sh in enumerate(shifts):
        v = (fit[:, 0] + np.ones(ih) * sh).astype(int)
        v[v < 0] = 0
        v[v > iw - 2] = iw - 2
        il[s] = v
    lw = np.ones(ih) - fit[:, 1]
    rw = np.ones(ih) - lw
    return il, lw, rw


def recon(stack, fit, shifts, k0=0, k1=None):
    """``disk[s][i, k] = trunc(L*lw[i] + R*rw[i])`` in float64, separate
    multiply / multiply / add (solex_util.py:131-134).  Evaluated in raw
    coordinates: for rotated scans L = raw_k[il, W-1-i]."""
    k1 = stack.shape[0] if k1 is None else k1
    N, H, W = stack.shape
    rotated = W > H
    ih, iw = (W, H) if rotated else (H, W)
    il, lw, rw = recon_tables(fit, shifts, iw)
    scale = 256 if stack.dtype == np.uint8 else 1
    out = [np.zeros((ih, k1 - k0), dtype=np.uint16) for _ in shifts]
    rows = np.arange(ih)
    xcol = W - 1 - rows
    for a in range(k0, k1, 32):
        blk = np.asarray(stack[a:min(k1, a + 32)]).astype(np.uint16) * np.uint16(scale)
        for s in range(len(shifts)):
            if rotated:
                L = blk[:, il[s], xcol]
                R = blk[:, il[s] + 1, xcol]
            else:
                L = blk[:, rows, il[s]]
                R = blk[:, rows, il[s] + 1]
            v = L * lw[None, :] + R * rw[None, :]
            out[s][:, a - k0:a - k0 + blk.shape[0]] = v.T.astype(np.uint16)
    return out


def shift_list(options_shift, ellipse_fit_shift=10):
    """Solex_recon.py:55 -- the two implicit shifts come first, de-duplicated."""
    return list(dict.fromkeys([ellipse_fit_shift, 0] + list(options_shift)))


# --------------------------------------------------------------------------
# a9/a10  circularisation warp   /root/reference/ellipse_to_circle.py:39-50,94-145
# --------------------------------------------------------------------------
def _rot(x):
    return np.array([[np.cos(x), np.sin(x)], [-np.sin(x), np.cos(x)]])


def correction_matrix(phi, r):
    stretch = _rot(phi) @ np.array([[r, 0], [0, 1]]) @ _rot(-phi)
    theta = np.arctan(stretch[1, 0] / stretch[0, 0])
    corr = _rot(theta) @ stretch
    corr[1, 0] = 0
    corr /= corr[1, 1]
    return np.linalg.inv(corr), theta


def warp_geometry(shape, phi, ratio):
    """mat3 and output shape of correct_image (ellipse_to_circle.py:100-114)."""
    mat, theta = correction_matrix(phi, ratio)
    h, w = shape
    corners = np.array([[0, 0], [0, h], [w, 0], [w, h]])
    new_corners = (np.linalg.inv(mat) @ corners.T).T
    new_h = np.max(new_corners[:, 1]) - np.min(new_corners[:, 1])
    new_w = np.max(new_corners[:, 0]) - np.min(new_corners[:, 0])
    mat3 = np.zeros((3, 3))
    mat3[:2, :2] = mat
    mat3[2, 2] = 1
    mat3 = mat3 @ np.array([[1, 0, np.min(new_corners[:, 0])], [0, 1, np.min(new_corners[:, 1])], [0, 0, 1]])
    return mat, mat3, (int(np.ceil(new_h)), int(np.ceil(new_w))), new_corners, theta


def warp_rows(img_u16, phi, ratio):
    """correct_image's pixel work as a per-row 1-D resample in DN units.

    mat3 rows 1 and 2 are exactly [0,1,t] and [0,0,1] with t == 0, so the
    bilinear warp degenerates to ``x = (m00*c + m01*r) + m02``,
    ``out = (1-d)*in[r, floor x] + d*in[r, ceil x]`` with out-of-range taps
    reading ``in[0, 0]``; skimage clips to the input range; the reference
    scales by 2**16 (exact) and truncates (ellipse_to_circle.py:115-118)."""
    mat, mat3, (oh, ow), _, _ = warp_geometry(img_u16.shape, phi, ratio)
    assert mat3[1, 0] == 0 and mat3[1, 1] == 1 and mat3[1, 2] == 0
    h, w = img_u16.shape
    img = img_u16.astype(np.float64)
    cval = img[0, 0]
    lo, hi = img.min(), img.max()
    out = np.empty((oh, ow), dtype=np.uint16)
    c = np.arange(ow, dtype=np.float64)
    for r in range(oh):
        x = (mat3[0, 0] * c + mat3[0, 1] * float(r)) + mat3[0, 2]
        x0 = np.floor(x)
        x1 = np.ceil(x)
        d = x - x0
        i0 = x0.astype(np.int64)
        i1 = x1.astype(np.int64)
        if r < h:
            row = img[r]
            L = np.where((i0 >= 0) & (i0 < w), row[np.clip(i0, 0, w - 1)], cval)
            R = np.where((i1 >= 0) & (i1 < w), row[np.clip(i1, 0, w - 1)], cval)
        else:
            L = np.full(ow, cval)
            R = np.full(ow, cval)
        v = (1 - d) * L + d * R
        out[r] = np.clip(v, lo, hi).astype(np.uint16)
    return out, mat3


def warped_circle(center_xy, height, phi, ratio, shape):
    """new centre / radius of correct_image (ellipse_to_circle.py:119-122)."""
    mat, _, _, new_corners, _ = warp_geometry(shape, phi, ratio)
    c = (np.linalg.inv(mat) @ np.asarray(center_xy, dtype='d').T).T - \
        np.array([np.min(new_corners[:, 0]), np.min(new_corners[:, 1])])
    rad = height * np.sqrt(np.abs(ratio / np.linalg.det(mat)))
    return c[0], c[1], rad


# --------------------------------------------------------------------------
# a11  ellipse fit front end     /root/reference/ellipse_to_circle.py:148-314
#      (PARITY UNPINNED: canny / LsqEllipse / downscale are restatements)
# --------------------------------------------------------------------------
def flood_image(image):
    import cv2
    from numpy import polynomial
    thresh = 0.9 * np.sum(image) / (image.shape[0] * image.shape[1])
    bw = int(image.shape[0] * 0.01)
    blurred = cv2.blur(image, ksize=(bw, bw))
    very_bright = np.percentile(blurred, 99)
    data = blurred.flatten()
    data = data[data < very_bright]
    n, bins = np.histogram(data, bins=20)
    d, c, b, a = polynomial.polynomial.Polynomial.fit(bins[1:], n, 3).convert().coef
    disc = 4 * b ** 2 - 12 * a * c
    thresh2 = (-2 * b + np.sqrt(disc)) / (6 * a) if disc >= 0 else thresh
    start = -1
    for i in range(len(bins) - 1):
        if bins[i] <= thresh2 < bins[i + 1]:
            start = i
    if start == -1:
        thresh3 = thresh
    else:
        i = start
        while 0 < i < len(bins) - 2:
            if n[i - 1] < n[i]:
                i -= 1
            elif n[i + 1] < n[i]:
                i += 1
            else:
                break
        if i >= 1:
            i -= 1
        thresh3 = bins[i]
    return np.where(blurred < thresh3, 0.0, 65000.0)


def edge_points(image, sigma=2.0, num_reg=2):
    import cv2
    from scipy import ndimage as ndi
    from scipy.spatial import ConvexHull
    if sigma <= 0:
        raise ValueError('no edges found')
    low = np.median(cv2.blur(image, ksize=(5, 5))) / 10
    edges = thirdparty.canny(flood_image(image), sigma=sigma, low_threshold=low, high_threshold=low * 1.5)
    labelled, nf = ndi.label(edges, structure=np.ones((3, 3)))
    if nf == 0:
        return edge_points(image, sigma - 0.5, num_reg)
    sizes = [-1] + [int(np.sum(labelled == i)) for i in range(1, nf + 1)]
    biggest = [sizes.index(s) for s in sorted(sizes, reverse=True)[:min(nf, num_reg)]]
    filt = np.isin(labelled, biggest)
    X = np.argwhere(filt)
    hull = np.zeros(edges.shape, bool)
    Xc = X[ConvexHull(X).vertices]
    hull[Xc[:, 0], Xc[:, 1]] = True
    filt = np.zeros(edges.shape, bool)
    for lab in biggest:
        if np.any((labelled == lab) & hull):
            filt |= labelled == lab
    x_min, x_max = X[:, 0].min(), X[:, 0].max()
    dx = x_max - x_min
    crop = 0.017
    mask = np.zeros(edges.shape, bool)
    mask[int(x_min + dx * crop):int(x_max - dx * crop), :] = True
    return np.argwhere(filt & mask).astype(float)


def _ellipse_params(points):
    reg = thirdparty.LsqEllipse().fit(points)
    return reg.as_parameters()


def two_step(points):
    """ellipse_to_circle.py:62-91."""
    center, width, height, phi = _ellipse_params(points)
    mat, _ = correction_matrix(phi, height / width)
    values = np.linalg.norm(mat @ (points - np.array(center)).T * height, axis=0) - 1
    kept = points[values > -max(values)]
    center, width, height, phi = _ellipse_params(kept)
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
    return np.array(center), height, phi, ratio, kept


def ellipse_fit(img_u16):
    """Everything of ellipse_to_circle (ellipse_to_circle.py:294-314) except the
    warp: returns (centre_xy, height, phi, ratio, kept edge points (row, col))."""
    image = img_u16 / 65536
    X = edge_points(thirdparty.downscale_local_mean(image, (4, 4))) * 4
    center, height, phi, ratio, kept = two_step(X)
    return np.array([center[1], center[0]]), height, phi, ratio, kept


def ellipse_to_circle(img_u16):
    center, height, phi, ratio, kept = ellipse_fit(img_u16)
    fix, mat3 = warp_rows(img_u16, phi, ratio)
    circle = warped_circle(center, height, phi, ratio, img_u16.shape)
    P = np.ones((kept.shape[0], 3))
    P[:, 0] = kept[:, 1]
    P[:, 1] = kept[:, 0]
    Pt = (np.linalg.inv(mat3) @ P.T).T
    borders = [np.min(Pt[:, 0]), np.min(Pt[:, 1]), np.max(Pt[:, 0]), np.max(Pt[:, 1])]
    return fix, circle, ratio, phi, borders


# --------------------------------------------------------------------------
# a12  transversalium            /root/reference/solex_util.py:76-86,383-516
# --------------------------------------------------------------------------
def reject_outliers_mean(rat, m=2.0):
    med = np.median(rat)
    d = np.abs(rat - med)
    mdev = np.median(d)
    s = d / mdev if mdev else np.zeros(len(d))
    return np.mean(rat[s < m])


def transversalium_rows(img, circle, borders):
    """Row range and per-row robust mean log-ratio (solex_util.py:384-395)."""
    cx, cy, rad = circle
    y1 = math.ceil(max(cy - rad, borders[1]))
    y2 = math.floor(min(cy + rad, borders[3]))
    ratios = [0.0]
    for y in range(y1 + 1, y2):
        dx = math.floor((rad ** 2 - (y - cy) ** 2) ** 0.5)
        xa = math.ceil(max(cx - dx, borders[0]))
        xb = math.floor(min(cx + dx, borders[2]))
        with np.errstate(divide='ignore', invalid='ignore'):
            rat = np.log(img[y, xa:xb] / img[y - 1, xa:xb])
            ratios.append(reject_outliers_mean(rat))
    return y1, y2, np.array(ratios)


def tukey_taper(n, a=0.05):
    def t(x):
        if 0 <= x < a * n / 2:
            return 1 / 2 * (1 - math.cos(2 * math.pi * x / (a * n)))
        elif a * n / 2 <= x <= n / 2:
            return 1
        elif n / 2 <= x <= n:
            return t(n - x)
        return 1
    return np.array([t(x) for x in range(n)])


def transversalium_gain(ratios, y1, y2, n_rows, strength=301):
    """solex_util.py:400-404,456-479."""
    from scipy.signal import savgol_filter
    trend = savgol_filter(ratios, min(strength, len(ratios) // 2 * 2 - 1), 3)
    detrended = ratios - trend
    detrended -= np.mean(detrended)
    correction = np.exp(-np.cumsum(detrended))
    n = correction.shape[0]
    corr_t = np.ones(n) + (correction - np.ones(n)) * tukey_taper(n)
    c = np.ones(n_rows)
    c[y1:y2] = corr_t
    return c


def apply_row_gain(img, c):
    """solex_util.py:489,515-516: multiply rows, clip at 65535, truncate."""
    ret = (img.T * c).T
    ret[ret > 65535] = 65535
    return np.array(ret, dtype='uint16')


def correct_transversalium(img, circle, borders, strength=301):
    y1, y2, ratios = transversalium_rows(img, circle, borders)
    c = transversalium_gain(ratios, y1, y2, img.shape[0], strength)
    return apply_row_gain(img, c), c


# --------------------------------------------------------------------------
# a8/a13  glue                   /root/reference/Solex_recon.py:49-174
# --------------------------------------------------------------------------
def solex_read(stack, shifts_requested, flip_x=False, ellipse_fit_shift=10):
    lf = mean_and_fit(stack)
    shifts = shift_list(shifts_requested, ellipse_fit_shift)
    disks = recon(stack, lf['fit'], shifts)
    if flip_x:
        disks = [np.flip(d, axis=1) for d in disks]
    return disks, shifts, lf


def solex_process(disks, shifts, shifts_requested, bounds, ratio_fixe=None, slant_fix=None,
                  transversalium=True, strength=301):
    """Per requested shift: circularise then detransversalium
    (Solex_recon.py:93-152); returns {shift: (circular, detrans)} + geometry."""
    out = {}
    circle = (-1, -1, -1)
    borders = [0, 0, 0, 0]
    phi = None
    for i, sh in enumerate(shifts):
        requested = sh in shifts_requested
        if ratio_fixe is None and slant_fix is None:
            circ, circle, ratio_fixe, phi, borders = ellipse_to_circle(disks[i])
            slant_fix = math.degrees(phi)
        else:
            ratio = ratio_fixe if ratio_fixe is not None else 1.0
            phi = math.radians(slant_fix) if slant_fix is not None else 0.0
            if requested:
                circ, _ = warp_rows(disks[i], phi, ratio)
        if not requested:
            continue
        if transversalium:
            if circle != (-1, -1, -1):
                det, _ = correct_transversalium(circ, circle, borders, strength)
            else:
                det, _ = correct_transversalium(
                    circ, (0, 0, 99999), [0, bounds[0] + 20, circ.shape[1] - 1, bounds[1] - 20], strength)
        else:
            det = circ
        out[sh] = (circ, det)
    return out, dict(circle=circle, borders=borders, ratio=ratio_fixe, slant=slant_fix)
