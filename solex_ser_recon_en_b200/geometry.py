"""Host-side geometry of the circularisation (a handful of 2x2 / 3x3 products;
the pixel work is the shg_warp_rows kernel).

Follows the reference's conventions so the numbers in the log and the canvas
size agree with it:
  get_correction_matrix   /root/reference/ellipse_to_circle.py:39-50
  correct_image canvas    /root/reference/ellipse_to_circle.py:100-111
  new centre / radius     /root/reference/ellipse_to_circle.py:119-122
"""
from __future__ import annotations

import numpy as np


def rotation(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, s], [-s, c]])


def correction_matrix(phi, ratio):
    """Shear + x-scale that turns the fitted ellipse into a circle while
    mapping image rows onto themselves.  Returns (inverse map 2x2, unrotation angle)."""
    stretch = rotation(phi) @ np.diag([ratio, 1.0]) @ rotation(-phi)
    theta = np.arctan(stretch[1, 0] / stretch[0, 0])
    forward = rotation(theta) @ stretch
    forward[1, 0] = 0
    forward = forward / forward[1, 1]
    return np.linalg.inv(forward), theta


def warp_plan(shape, phi, ratio):
    """(mat 2x2, mat3 3x3 output->input map, (out_rows, out_cols), translated corners, theta)
    for an (h, w) image."""
    mat, theta = correction_matrix(phi, ratio)
    h, w = shape
    corners = np.array([[0, 0], [0, h], [w, 0], [w, h]])
    moved = (np.linalg.inv(mat) @ corners.T).T
    x_min, y_min = moved[:, 0].min(), moved[:, 1].min()
    new_w = moved[:, 0].max() - x_min
    new_h = moved[:, 1].max() - y_min
    mat3 = np.zeros((3, 3))
    mat3[:2, :2] = mat
    mat3[2, 2] = 1
    mat3 = mat3 @ np.array([[1, 0, x_min], [0, 1, y_min], [0, 0, 1]])
    return mat, mat3, (int(np.ceil(new_h)), int(np.ceil(new_w))), moved, theta


def moved_circle(center_xy, height, phi, ratio, shape):
    """Centre and radius of the disk after the warp."""
    mat, _, _, moved, _ = warp_plan(shape, phi, ratio)
    c = (np.linalg.inv(mat) @ np.asarray(center_xy, dtype='d').T).T - \
        np.array([moved[:, 0].min(), moved[:, 1].min()])
    radius = height * np.sqrt(np.abs(ratio / np.linalg.det(mat)))
    return c, radius
