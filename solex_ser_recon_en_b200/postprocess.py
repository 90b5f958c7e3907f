"""Batched circularisation + transversalium for every requested shift of one
scan.  Same arithmetic as calling correct_image / correct_transversalium2 per
image (reference Solex_recon.py:103-152); batching only removes the per-image
device -> host round trips (one min/max read-back and one row-statistics
read-back for the whole set instead of one per shift)."""
from __future__ import annotations

import numpy as np
import torch

from . import geometry
from .device_image import DeviceImage
from .engine import get_engine


def frames_of(eng, image):
    """(frame-major (N, ih) uint16 tensor, flip) of a DeviceImage or an (ih, N) array."""
    if isinstance(image, DeviceImage):
        if image.layout == 'frames':
            return image.tensor, image.flip
        host = image.numpy()
    else:
        host = np.asarray(image)
    if host.dtype.kind == 'f' and host.size and host.max() < 1.0 + 1e-9:     # the reference passes disk / 65536
        host = np.rint(host * 65536.0)
    dn = np.ascontiguousarray(host.T).astype(np.uint16)
    return torch.from_numpy(dn).to(eng.device), False


def circularise_many(images, phi, ratio):
    """Warp every image with the same (phi, ratio).  Returns
    (list of row-major DeviceImage, mat 2x2, mat3, theta)."""
    eng = get_engine()
    if not images:
        return [], None, None, None
    pairs = [frames_of(eng, im) for im in images]
    n, ih = pairs[0][0].shape
    mat, mat3, out_shape, _, theta = geometry.warp_plan((ih, n), phi, ratio)
    with eng.stage('minmax'):
        mm = torch.tensor([[65535, 0]] * len(pairs), dtype=torch.int32, device=eng.device)
        corners = eng.empty((len(pairs),), torch.uint16)
        from ._lib import call
        for i, (fr, flip) in enumerate(pairs):
            call('shg_minmax_u16', fr.data_ptr(), fr.numel(), mm[i].data_ptr(), eng.stream)
            corners[i:i + 1].copy_(fr[n - 1 if flip else 0, 0:1])                # image[0, 0]
        eng.n_launches += len(pairs)
        mm_h = mm.cpu().numpy()
        corners_h = corners.cpu().numpy()
    out = []
    with eng.stage('warp'):
        for i, (fr, flip) in enumerate(pairs):
            t = eng.warp(fr, flip, mat3, out_shape, float(corners_h[i]), float(mm_h[i, 0]), float(mm_h[i, 1]))
            out.append(DeviceImage(eng, t))
    return out, mat, mat3, theta


def detransversalium_many(images, circle, borders, strength):
    """Transversalium correction of several row-major device images that share
    the disk geometry.  Returns (list of DeviceImage, gains (S, rows))."""
    from .solex_util import transversalium_gain
    eng = get_engine()
    if not images:
        return [], np.zeros((0, 0))
    tens = [im.rows_tensor() if isinstance(im, DeviceImage) else
            torch.from_numpy(np.ascontiguousarray(np.asarray(im), dtype=np.uint16)).to(eng.device) for im in images]
    y1, y2, rows, xa, xb = eng.transversalium_chords(circle, borders)
    with eng.stage('transv_stats'):
        stats = eng.transversalium_row_stats_many(tens, rows, xa, xb)
    h = tens[0].shape[0]
    with eng.stage('transv_gain_host'):
        gains = np.stack([transversalium_gain(np.concatenate([[0.0], s]), y1, y2, h, strength) for s in stats])
        gains_d = torch.from_numpy(gains).to(eng.device)
    with eng.stage('row_scale'):
        out = [DeviceImage(eng, eng.row_scale(t, gains_d[i])) for i, t in enumerate(tens)]
    return out, gains
