"""Drop-in for the reference's ellipse_to_circle module
(/root/reference/ellipse_to_circle.py): ellipse_to_circle, correct_image,
get_correction_matrix, two_step keep their signatures and return values; the
warp runs on the GPU (shg_warp_rows) and the limb fit works from 4x4 block
sums computed on the GPU.
"""
from __future__ import annotations

import math

import numpy as np

from . import ellipse_fit, geometry
from .device_image import DeviceImage
from .engine import get_engine
from .solex_util import logme, output_path

NUM_REG = ellipse_fit.NUM_REG
rot = geometry.rotation
get_correction_matrix = geometry.correction_matrix
two_step = ellipse_fit.two_step


def _as_frames(eng, image):
    """(frame-major tensor, flip) for a DeviceImage or an (ih, N) array.  Arrays
    in [0, 1) (the reference passes disk / 65536) are scaled back to DN."""
    import torch
    if isinstance(image, DeviceImage):
        if image.layout == 'frames':
            return image.tensor, image.flip
        host = image.numpy()
    else:
        host = np.asarray(image)
    if host.dtype.kind == 'f':
        host = np.rint(host * 65536.0) if host.size and host.max() < 1.0 + 1e-9 else host
    dn = np.ascontiguousarray(host.T).astype(np.uint16)
    return torch.from_numpy(dn).to(eng.device), False


def _log_geometry(options, mat, theta, phi, ratio, new_center, new_radius, height):
    """The block correct_image prints / logs when print_log is set (reference :123-144)."""
    print('unrotation angle theta = ' + '{:.3f}'.format(math.degrees(theta)) + ' degrees')
    print('Y/X ratio : ' + '{:.3f}'.format(ratio))
    if '_nolog' in options:                                   # logme would drop every line: do not format them
        return
    basefich0 = options['basefich0']
    log = basefich0 + '_log.txt'
    np.set_printoptions(suppress=True)
    logme(log, options, 'Y/X ratio : ' + '{:.3f}'.format(ratio))
    logme(log, options, 'Tilt angle : ' + '{:.3f}'.format(math.degrees(phi)) + ' degrees')
    logme(log, options, 'Linear transform correction matrix : \n' + str(mat))
    where = (str(new_center) + ', ' + '{:.3f}'.format(new_radius)) if not height == -1.0 else 'UNKNOWN'
    logme(log, options, 'Disk position, radius : ' + where)
    logme(log, options, 'Unrotation : ' + '{:.3f}'.format(math.degrees(theta)) + ' degrees')
    np.set_printoptions(suppress=False)


def correct_image(image, phi, ratio, center, height, options, print_log=False):
    """Shear / scale the image along the scan axis so that the Sun is round
    (reference ellipse_to_circle.py:94-145).  Returns
    (uint16 image, (cx, cy, radius), mat3); the image is a DeviceImage when the
    input was one, else an ndarray."""
    eng = get_engine()
    frames, flip = _as_frames(eng, image)
    n, ih = frames.shape
    mat, mat3, out_shape, _, theta = geometry.warp_plan((ih, n), phi, ratio)
    with eng.stage('minmax'):
        mm = eng.minmax_device(frames)
    with eng.stage('warp'):
        out = eng.warp_batch(frames, None, flip, mat3, out_shape, mm)[0]
    new_center, new_radius = geometry.moved_circle(np.asarray(center, dtype='d'), height, phi, ratio, (ih, n))
    if print_log:
        _log_geometry(options, mat, theta, phi, ratio, new_center, new_radius, height)
    result = DeviceImage(eng, out) if isinstance(image, DeviceImage) else out.cpu().numpy()
    return result, (new_center[0], new_center[1], new_radius), mat3


def _fit_points(eng, frames, flip):
    with eng.stage('ellipse_fit'):
        sums = eng.downscale4(frames, flip)
        return ellipse_fit.fit_from_device(eng, sums)


_fit_pool = None


def start_fit(image, ready=None):
    """Start the limb search + ellipse fit of a frame-major DeviceImage in a helper
    thread, on a side stream that first waits for the kernel that produced the
    image (`ready`, or everything queued on the current stream).  Returns a future of
    _fit_points' result; fit_geometry picks it up through image.fit_future."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    global _fit_pool
    eng = get_engine()
    if _fit_pool is None:
        _fit_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix='shg-fit')
    if not hasattr(eng, '_fit_stream'):
        eng._fit_stream = torch.cuda.Stream(device=eng.device)
    side = eng._fit_stream
    if ready is not None:
        side.wait_event(ready)                  # only the kernel that produced this image, not what was queued after it
    else:
        side.wait_stream(torch.cuda.current_stream(eng.device))
    frames, flip = image.tensor, image.flip

    def job():
        torch.cuda.set_device(eng.device)
        with torch.cuda.stream(side):
            pts = _fit_points(eng, frames, flip)
        n, ih = frames.shape
        return pts, _derived_geometry(pts, (ih, n))      # the caller only has to log it
    return _fit_pool.submit(job)


def _derived_geometry(fit_points, shape):
    """What correct_image / ellipse_to_circle derive from the fitted ellipse (reference :100-122, 310-314): the
    correction matrix, the disk's centre and radius after the warp, and the bounding box of the limb points."""
    center, height, phi, ratio, kept, raw, outline = fit_points
    mat, mat3, _, _, theta = geometry.warp_plan(shape, phi, ratio)
    new_center, new_radius = geometry.moved_circle(center, height, phi, ratio, shape)
    pts = np.ones((kept.shape[0], 3))
    pts[:, 0], pts[:, 1] = kept[:, 1], kept[:, 0]                             # (row, col) -> (x, y)
    moved = (np.linalg.inv(mat3) @ pts.T).T
    borders = [np.min(moved[:, 0]), np.min(moved[:, 1]), np.max(moved[:, 0]), np.max(moved[:, 1])]
    return dict(mat=mat, mat3=mat3, theta=theta, new_center=new_center, new_radius=new_radius, borders=borders)


def fit_geometry(image, options, basefich=None):
    """The fit half of ellipse_to_circle: limb search + ellipse fit + the geometry
    of the correction, without warping anything.  Returns a dict with
    phi, ratio, circle (cx, cy, radius after the warp), borders, mat3 and the
    points used (for the diagnostic plot)."""
    eng = get_engine()
    frames, flip = _as_frames(eng, image)
    n, ih = frames.shape
    early = getattr(image, 'fit_future', None)
    if early is not None:
        image.fit_future = None
        fit_points, d = early.result()                                       # started under the reconstruction
    else:
        fit_points = _fit_points(eng, frames, flip)
        d = _derived_geometry(fit_points, (ih, n))
    center, height, phi, ratio, kept, raw, outline = fit_points
    _log_geometry(options, d['mat'], d['theta'], phi, ratio, d['new_center'], d['new_radius'], height)
    borders = d['borders']
    print('sun borders found:' + str(borders))
    return dict(phi=phi, ratio=ratio, circle=(d['new_center'][0], d['new_center'][1], d['new_radius']), borders=borders,
                mat3=d['mat3'], center=center, height=height, kept=kept, raw=raw, outline=outline, frames=frames,
                flip=flip)


def ellipse_to_circle(image, options, basefich):
    """Fit an ellipse to the solar limb and circularise
    (reference ellipse_to_circle.py:294-342).
    Returns (image, (cx, cy, radius), ratio, phi, borders)."""
    eng = get_engine()
    fit = fit_geometry(image, options, basefich)
    src = image if isinstance(image, DeviceImage) else DeviceImage(eng, fit['frames'], 'frames', fit['flip'])
    fixed, circle, mat3 = correct_image(src, fit['phi'], fit['ratio'], fit['center'], fit['height'], options)
    if not options['clahe_only'] and not options['protus_only']:
        _plot_fit(image, fixed, fit['raw'], fit['kept'], fit['outline'], fit['borders'],
                  output_path(basefich + '_ellipse_fit.png', options))
    if not isinstance(image, DeviceImage):
        fixed = np.asarray(fixed)
    return fixed, circle, fit['ratio'], fit['phi'], fit['borders']


def _plot_fit(image, fixed, raw, kept, outline, borders, path):
    """`_ellipse_fit.png` (ellipse_to_circle.py:316-341): edges found, edges kept + fitted ellipse, corrected image."""
    from . import miniplot
    image, fixed = np.asarray(image), np.asarray(fixed)
    outline = np.asarray(outline)
    if not miniplot.have_matplotlib():
        import os
        import cv2
        tmp = [path + '.%d.png' % i for i in range(3)]
        miniplot.image_overlay(tmp[0], image, np.asarray(raw)[:, ::-1], title='edge detection')
        miniplot.image_overlay(tmp[1], image, np.asarray(kept)[:, ::-1], outline[:, ::-1], title='filtered edges / ellipse fit')
        miniplot.image_overlay(tmp[2], fixed, hlines=(borders[1], borders[3]), vlines=(borders[0], borders[2]),
                               title='geometrically corrected image')
        miniplot.panels(path, [cv2.imread(t) for t in tmp])
        for t in tmp:
            os.remove(t)
        return
    import matplotlib.figure
    import matplotlib.pyplot
    fig = matplotlib.figure.Figure()
    ax = [[fig.add_subplot(2, 2, 1), fig.add_subplot(2, 2, 2)], [fig.add_subplot(2, 2, 3), fig.add_subplot(2, 2, 4)]]
    fig.tight_layout()
    gray = matplotlib.pyplot.cm.gray
    ax[0][0].imshow(image, cmap=gray)
    ax[0][0].set_title('uncorrected image', fontsize=11)
    ax[0][1].imshow(image, cmap=gray)
    ax[0][1].plot(raw[:, 1], raw[:, 0], 'ro', label='edge detection')
    ax[0][1].legend(prop={'size': 6})
    ax[1][1].plot(kept[:, 1], kept[:, 0], 'ro', label='filtered edges')
    ax[1][1].plot(outline[:, 1], outline[:, 0], color='b', label='ellipse fit')
    ax[1][1].set_ylim([image.shape[0], 0])
    ax[1][1].legend(prop={'size': 6})
    ax[1][0].imshow(fixed, cmap=gray)
    for y in (borders[1], borders[3]):
        ax[1][0].axhline(y=y)
    for x in (borders[0], borders[2]):
        ax[1][0].axvline(x=x)
    ax[1][0].set_title('geometrically corrected image', fontsize=11)
    for a in (ax[0][0], ax[0][1], ax[1][0], ax[1][1]):
        a.set_aspect('equal')
    fig.savefig(path, dpi=300)
