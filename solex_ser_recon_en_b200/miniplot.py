"""Diagnostic figures without matplotlib.

The reference draws three diagnostic PNGs per scan with matplotlib
(/root/reference/solex_util.py:263-273 `_spectral_line_data.png`, :482-488
`_transversalium_correction.png`; /root/reference/ellipse_to_circle.py:316-341
`_ellipse_fit.png`).  matplotlib is an optional dependency of this package: when
it is missing the same files are still produced -- a drop-in must leave the
same set of outputs behind -- as plain OpenCV drawings with the same content
(curve, points, axes box), not pixel-identical to matplotlib's rendering.
"""
from __future__ import annotations

import cv2
import numpy as np


def have_matplotlib() -> bool:
    try:
        import matplotlib.figure    # noqa: F401
        import matplotlib.pyplot    # noqa: F401
        return True
    except Exception:
        return False


def _canvas(w=1200, h=900):
    img = np.full((h, w, 3), 255, np.uint8)
    box = (90, 40, w - 40, h - 80)                         # x0, y0, x1, y1 of the axes
    cv2.rectangle(img, (box[0], box[1]), (box[2], box[3]), (0, 0, 0), 1)
    return img, box


def _to_px(x, y, xr, yr, box, flip_y=False):
    x0, y0, x1, y1 = box
    u = x0 + (np.asarray(x, float) - xr[0]) / max(xr[1] - xr[0], 1e-300) * (x1 - x0)
    t = (np.asarray(y, float) - yr[0]) / max(yr[1] - yr[0], 1e-300)
    v = y0 + t * (y1 - y0) if flip_y else y1 - t * (y1 - y0)
    return np.stack([u, v], axis=-1).round().astype(np.int32)


def _labels(img, box, xr, yr, xlabel, ylabel):
    f = cv2.FONT_HERSHEY_SIMPLEX
    cv2.putText(img, '%.6g' % xr[0], (box[0] - 10, box[3] + 25), f, 0.5, (0, 0, 0), 1, cv2.LINE_AA)
    cv2.putText(img, '%.6g' % xr[1], (box[2] - 60, box[3] + 25), f, 0.5, (0, 0, 0), 1, cv2.LINE_AA)
    cv2.putText(img, '%.6g' % yr[0], (5, box[3]), f, 0.5, (0, 0, 0), 1, cv2.LINE_AA)
    cv2.putText(img, '%.6g' % yr[1], (5, box[1] + 12), f, 0.5, (0, 0, 0), 1, cv2.LINE_AA)
    cv2.putText(img, xlabel, ((box[0] + box[2]) // 2 - 40, box[3] + 55), f, 0.6, (0, 0, 0), 1, cv2.LINE_AA)
    cv2.putText(img, ylabel, (5, box[1] - 15), f, 0.6, (0, 0, 0), 1, cv2.LINE_AA)


def line_plot(path, y, xlabel='', ylabel=''):
    """ax.plot(y) with axis labels."""
    y = np.asarray(y, float)
    img, box = _canvas()
    xr = (0.0, float(max(1, len(y) - 1)))
    lo, hi = (float(np.nanmin(y)), float(np.nanmax(y))) if len(y) else (0.0, 1.0)
    pad = 0.05 * (hi - lo) or 0.5
    yr = (lo - pad, hi + pad)
    if len(y) > 1:
        cv2.polylines(img, [_to_px(np.arange(len(y)), y, xr, yr, box)], False, (180, 90, 30), 1, cv2.LINE_AA)
    _labels(img, box, xr, yr, xlabel, ylabel)
    cv2.imwrite(path, img)


def _gray(image, size):
    a = np.asarray(image).astype(np.float64)
    lo, hi = float(a.min()), float(a.max())
    a = ((a - lo) / max(hi - lo, 1e-300) * 255).astype(np.uint8)
    return cv2.cvtColor(cv2.resize(a, size, interpolation=cv2.INTER_AREA), cv2.COLOR_GRAY2BGR)


def image_overlay(path, image, points_xy=None, curve_xy=None, hlines=(), vlines=(), title=''):
    """imshow(image) with optional red points, a blue curve and axis-parallel lines (image coordinates)."""
    h, w = np.asarray(image).shape[:2]
    scale = min(1400.0 / w, 1000.0 / h)
    size = (max(1, int(w * scale)), max(1, int(h * scale)))
    img = _gray(image, size)
    if points_xy is not None and len(points_xy):
        for x, y in (np.asarray(points_xy, float) * scale).round().astype(np.int32):
            cv2.circle(img, (int(x), int(y)), 2, (0, 0, 255), -1)
    if curve_xy is not None and len(curve_xy) > 1:
        cv2.polylines(img, [(np.asarray(curve_xy, float) * scale).round().astype(np.int32)], False, (255, 80, 0), 1,
                      cv2.LINE_AA)
    for y in hlines:
        cv2.line(img, (0, int(y * scale)), (size[0] - 1, int(y * scale)), (255, 160, 0), 1)
    for x in vlines:
        cv2.line(img, (int(x * scale), 0), (int(x * scale), size[1] - 1), (255, 160, 0), 1)
    if title:
        cv2.putText(img, title, (10, 20), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (0, 255, 255), 1, cv2.LINE_AA)
    cv2.imwrite(path, img)


def panels(path, tiles):
    """Stack already rendered BGR tiles (same width after resize) vertically into one figure."""
    w = max(t.shape[1] for t in tiles)
    rows = [cv2.copyMakeBorder(t, 0, 4, 0, w - t.shape[1], cv2.BORDER_CONSTANT, value=(255, 255, 255)) for t in tiles]
    cv2.imwrite(path, np.vstack(rows))
