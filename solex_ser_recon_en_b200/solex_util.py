"""Drop-in for the reference's solex_util module (/root/reference/solex_util.py):
same function names, arguments, return values, log lines and `options` side
effects for everything on the reconstruction path, with the pixel work done by
libshg.so on the GPU.

  compute_mean_max / compute_mean_return_fit   solex_util.py:174-274
  read_video_improved                          solex_util.py:93-144
  correct_transversalium2 / reject_outliers    solex_util.py:76-86, 383-516
  image_process / rescale_brightness           solex_util.py:519-588 (host tail: CLAHE + PNG, outside the hot path)
  logme / clearlog / write_complete / output_path / make_header   solex_util.py:29-63, 147-161

Scans stay resident in HBM between compute_mean_return_fit and
read_video_improved (the reference opens the file three times,
Solex_recon.py:56-63; here it crosses PCIe once).
"""
from __future__ import annotations

import datetime
import math
import os
import sys
import traceback

import cv2
import numpy as np
from scipy.signal import savgol_filter

from . import fits_min as fits
from .device_image import DeviceImage
from .engine import DeviceStack, ScanGeometry, get_engine
from .video_reader import *          # noqa: F401,F403  (the reference re-exports it too)
from .video_reader import device_scan, memory_scan, video_reader


# ------------------------------------------------------------------- logging
def output_path(path, options):
    if options['output_dir'].strip() == '':
        return path
    return os.path.join(options['output_dir'], os.path.basename(path))


def _append(path, options, text, mode):
    try:
        with open(output_path(path, options), mode) as f:
            f.write(text)
    except Exception:
        traceback.print_exc()
        print('ERROR: failed to log file: ' + path)


def clearlog(path, options):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    _append(path, options, 'start time: ' + str(datetime.datetime.now()) + '\n', 'w')


def write_complete(path, options):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    _append(path, options, 'end time: ' + str(datetime.datetime.now()) + '\n', 'a')


def logme(path, options, s):
    if '_nolog' in options or int(os.environ.get('RANK', '0')) != 0:     # one log per scan: rank 0 writes it
        return
    _append(path, options, s + '\n', 'a')


def resource_path(relative_path):
    base = getattr(sys, '_MEIPASS', os.path.abspath('.'))
    return os.path.join(base, relative_path)


def make_header(rdr):
    hdr = fits.Header()
    for key, val in (('SIMPLE', 'T'), ('BITPIX', 32), ('NAXIS', 2), ('NAXIS1', rdr.iw), ('NAXIS2', rdr.ih),
                     ('BZERO', 0), ('BSCALE', 1), ('BIN1', 1), ('BIN2', 1), ('EXPTIME', 0)):
        hdr[key] = val
    return hdr


# ------------------------------------------------------ resident scan cache
class _Resident:
    """The scan most recently ingested by this process (one at a time: a
    config-5 stack is 84 GB of the 180 GB of HBM)."""
    key = None
    stack = None
    stats = None
    owner = None              # the reader object itself for readers without a file identity: keeps it alive, so
    #                           that id() cannot be reused by another reader while its stack is cached

    @classmethod
    def drop(cls):
        cls.key = cls.stack = cls.stats = cls.owner = None


def _reader_key(rdr):
    path = getattr(rdr, 'path', None)
    if isinstance(path, str) and os.path.exists(path):
        st = os.stat(path)
        return (os.path.realpath(path), st.st_size, st.st_mtime_ns)
    return ('object', id(rdr))


def resident_stack(rdr, accumulate=True) -> DeviceStack:
    """Device-resident raw stack for a reader, ingesting it if needed.  Accepts
    this package's video_reader (streamed from the file through the pinned ring)
    or any reader with the reference's interface (all_video_reader's `frames`
    array, or frame-by-frame next_frame())."""
    eng = get_engine()
    from .video_reader import all_video_reader
    if isinstance(rdr, all_video_reader):                         # resident since it was opened; summed during ingest
        if accumulate and not rdr.stack.accumulated:
            eng.accumulate(rdr.stack, reset=True)
        return rdr.stack
    if isinstance(rdr, device_scan):                              # already in HBM: pass 1 just re-reads it
        if accumulate:
            with eng.stage('accumulate'):
                eng.accumulate(rdr.stack, reset=True)
        return rdr.stack
    key = _reader_key(rdr)
    if _Resident.key == key and _Resident.stack is not None and (_Resident.stack.accumulated or not accumulate) \
            and (key[0] != 'object' or _Resident.owner is rdr):
        return _Resident.stack
    _Resident.drop()
    from . import parallel
    stats = None
    if isinstance(rdr, memory_scan):
        geom = rdr.geometry
        k0, k1 = parallel.frame_range(geom.n_frames)
        reuse = getattr(rdr, 'device_stack', None)                # callers may hand in the HBM buffer to refill
        if reuse is not None:
            reuse.sum.zero_()
            reuse.max.zero_()
        with eng.stage('ingest_h2d+accumulate'):
            stack, stats = eng.ingest_host(rdr.host_ptr + k0 * geom.frame_bytes, geom, k1 - k0, k0=k0,
                                           accumulate=accumulate, stack=reuse)
    elif isinstance(rdr, video_reader):
        geom = rdr.geometry
        k0, k1 = parallel.frame_range(geom.n_frames)              # this rank's frames (all of them on one GPU)
        if rdr.streamable:
            stack, stats = eng.ingest_file(rdr.path, geom, rdr.payload_offset, rdr.frame_stride,
                                           k0=k0, n=k1 - k0, accumulate=accumulate)
        else:                                                     # compressed AVI: host decode, then upload
            stack = DeviceStack(geom, k0, k1 - k0, eng.device)
            step = max(1, (64 << 20) // geom.frame_bytes)
            for a in range(k0, k1, step):
                b = min(k1, a + step)
                blk = np.ascontiguousarray(rdr.raw_frames(a, b))
                sub = eng.ingest_array(blk, n_total=geom.n_frames, k0=a, accumulate=False)
                stack.frames[(a - k0) * geom.frame_bytes:(b - k0) * geom.frame_bytes].copy_(sub.frames[:blk.nbytes])
            if accumulate:
                eng.accumulate(stack)
    else:
        if hasattr(rdr, 'frames'):                                # all_video_reader: oriented uint16 frames in RAM
            frames = np.asarray(rdr.frames)
        else:                                                     # foreign streaming reader
            got = []
            while rdr.has_frames():
                got.append(np.array(rdr.next_frame()))
            frames = np.stack(got)
        k0, k1 = parallel.frame_range(frames.shape[0])
        stack = eng.ingest_array(frames[k0:k1], n_total=frames.shape[0], k0=k0, accumulate=accumulate)
    _Resident.key, _Resident.stack, _Resident.stats = key, stack, stats
    _Resident.owner = rdr if key[0] == 'object' else None
    return stack


def release_resident():
    _Resident.drop()


def _oriented_geometry(rdr, stack):
    """(ih, iw) the caller sees.  Oriented host frames are stored unrotated."""
    return stack.geom.ih, stack.geom.iw


# ----------------------------------------------------------- pass 1 + the fit
def reject_outliers(data, m=2):
    med = np.median(data)
    d = np.abs(data - med)
    mdev = np.median(d)
    s = d / mdev if mdev else np.zeros(len(d))
    return data[s < m]


def detect_bord(img, axis):
    """Kept for API compatibility (host, tiny): first / last row or column whose
    blurred mean exceeds a fifth of the median (solex_util.py:165-172)."""
    ymean = np.mean(cv2.blur(img, ksize=(5, 5)), axis)
    where_sun = ymean > np.median(ymean) / 5
    lb = np.argmax(where_sun)
    ub = img.shape[int(not axis)] - 1 - np.argmax(np.flip(where_sun))
    return lb, ub


def _mean_max_device(rdr, options, basefich0):
    logme(basefich0 + '_log.txt', options, 'Width, Height : ' + str(rdr.Width) + ' ' + str(rdr.Height))
    logme(basefich0 + '_log.txt', options, 'Number of frames : ' + str(rdr.FrameCount))
    eng = get_engine()
    stack = resident_stack(rdr, accumulate=True)
    from . import parallel
    with eng.stage('allreduce+finalize'):
        total_sum, total_max, n_total = parallel.combine_stats(stack)
        mean_img, max_img = eng.finalize_mean_max(total_sum, total_max, n_total, stack.geom)
    return eng, stack, mean_img, max_img


def compute_mean_max(rdr, options, basefich0):
    _, _, mean_img, max_img = _mean_max_device(rdr, options, basefich0)
    return mean_img.cpu().numpy(), max_img.cpu().numpy()


def compute_mean_return_fit(vid_rdr, options, hdr, iw, ih, basefich0):
    """Mean frame, slit extent, spectral-line minima and the cubic fit.
    Returns (mean_img uint16 (ih, iw), fit float64 (ih, 4), y1, y2)."""
    eng, stack, mean_dev, max_dev = _mean_max_device(vid_rdr, options, basefich0)
    mean_early = eng.download_early(mean_dev)         # copied to the host underneath the line detection
    mean_img = mean_early.result() if options['save_fit'] or options['flag_display'] else None
    from . import parallel
    writer = parallel.world()[0] == 0                  # several ranks compute the same mean frame: one of them writes it
    if options['save_fit'] and writer:
        fits.PrimaryHDU(mean_img, header=hdr).writeto(output_path(basefich0 + '_mean.fits', options), overwrite='True')
    if options['flag_display'] and writer:
        cv2.namedWindow('Ser mean', cv2.WINDOW_NORMAL)
        cv2.imshow('Ser mean', mean_img)
        if cv2.waitKey(2000) == 27:
            cv2.destroyAllWindows()
            sys.exit()
        cv2.destroyAllWindows()
    with eng.stage('detect+fit'):
        det = eng.detect_line(mean_dev, max_dev)
        y1, y2 = det['y1'], det['y2']
        lf = eng.fit_line(det, stack.geom.ih)
    if mean_img is None:
        mean_img = mean_early.result()
    logme(basefich0 + '_log.txt', options, 'Vertical limits y1, y2 : ' + str(y1) + ' ' + str(y2))
    p = lf['p3']
    logme(basefich0 + '_log.txt', options, 'Spectral line polynomial fit: ' + str(p))
    fit = lf['fit']
    if not options['clahe_only'] and not options['protus_only'] and writer:
        _plot_line_fit(mean_img, det, lf, y1, y2, output_path(basefich0 + '_spectral_line_data.png', options))
    return mean_img, fit, y1, y2


def _plot_line_fit(mean_img, det, lf, y1, y2, path):
    """`_spectral_line_data.png` (solex_util.py:263-273): the mean frame with the detected line minima and the fit."""
    from . import miniplot
    sharp = det['min_sharp'].cpu().numpy()[y1:y2]
    good = lf['mask_good'].cpu().numpy().astype(bool)
    s = (y2 - y1) // 20 + 1
    if not miniplot.have_matplotlib():
        pts = np.stack([sharp[good][::s], np.arange(y1, y2)[good][::s]], axis=1)
        curve = np.stack([lf['fit'][:, 3], np.arange(mean_img.shape[0])], axis=1)
        miniplot.image_overlay(path, mean_img, pts, curve, title='line detection (red) / polynomial fit (blue)')
        return
    import matplotlib.figure
    import matplotlib.pyplot
    fig = matplotlib.figure.Figure()
    ax = fig.add_subplot(1, 1, 1)
    ax.imshow(mean_img, cmap=matplotlib.pyplot.cm.gray)
    ax.plot(sharp[good][::s], np.arange(y1, y2)[good][::s], 'rx', label='line detection')
    ax.plot(lf['fit'][:, 3], np.arange(mean_img.shape[0]), label='polynomial fit')
    ax.legend(loc='center left', bbox_to_anchor=(1, 0.5))
    ax.set_aspect(0.1)
    fig.tight_layout()
    fig.savefig(path, dpi=400)


# ------------------------------------------------------------------- pass 2
def read_video_improved(rdr, fit, options):
    """One (ih, N) uint16 image per entry of options['shift'], sampled along the
    fitted line.  Returns (disk_list, ih, iw, FrameCount); the images are
    DeviceImage objects (array-likes that stay in HBM until their pixels are
    asked for)."""
    eng = get_engine()
    stack = resident_stack(rdr, accumulate=False)
    shifts = [int(s) for s in options['shift']]
    from . import parallel
    # When an ellipse fit will follow (Solex_recon.solex_process, no fixed ratio / tilt), the image of the
    # ellipse-fit shift is reconstructed first and the limb search starts on it in a helper thread and on
    # a side stream, underneath the reconstruction of the other shifts.
    prefetch = {}
    first_done = None
    wants_fit = options.get('_prefetch_fit') and options.get('ratio_fixe') is None and options.get('slant_fix') is None
    if wants_fit:
        def first_done(image0, ready):
            from .ellipse_to_circle import start_fit
            prefetch['fit'] = start_fit(DeviceImage(eng, image0, 'frames', bool(options.get('flip_x'))), ready)
    # several ranks: complete disk images on their owners ('by_shift'), or -- when nothing downstream needs the
    # pixels of a complete disk on one GPU -- every rank keeps its frame rows and the (4-5 x smaller)
    # circularised images are exchanged instead ('post_warp', see parallel.py)
    wait_previous = options.get('_before_recon')
    if wait_previous is not None:                                  # batch mode: the previous scan's owners are done
        wait_previous()                                            # with the buffers this scan reconstructs into
    _, size = parallel.world()
    post_warp = (size > 1 and parallel.exchange_mode() == 'post_warp' and wants_fit and not options['save_fit']
                 and not options['flag_display'] and (options['clahe_only'] or options['protus_only']))
    options['_exchange'] = 'post_warp' if post_warp else None
    if post_warp:
        with eng.stage('recon+gather'):
            disk_list = parallel.reconstruct_partial(stack, np.asarray(fit, dtype=np.float64), shifts, first_done)
        if 'fit' in prefetch and disk_list[0].full is not None:
            disk_list[0].full.fit_future = prefetch['fit']
    else:
        with eng.stage('recon+gather'):
            disk, mins, known = parallel.reconstruct(stack, np.asarray(fit, dtype=np.float64), shifts, first_done)
        # one image per shift; under several ranks only the images this rank owns (None elsewhere)
        disk_list = [None if d is None else DeviceImage(eng, d, 'frames') for d in disk]
        for j, im in enumerate(disk_list):
            if im is not None and known[j]:
                im.min_ref = (mins, j)                           # the kernel tracked this image's minimum
        if 'fit' in prefetch and disk_list[0] is not None:
            disk_list[0].fit_future = prefetch['fit']
    if options['flag_display'] and disk_list[1] is not None:
        cv2.namedWindow('disk', cv2.WINDOW_NORMAL)
        cv2.imshow('disk', np.asarray(disk_list[1]))             # disk_list[1] is always shift = 0
        if cv2.waitKey(1) == 27:
            cv2.destroyAllWindows()
            sys.exit()
    if not options.get('_keep_stack'):
        release_resident()
    ih, iw = _oriented_geometry(rdr, stack)
    return disk_list, ih, iw, rdr.FrameCount


# ----------------------------------------------------------- transversalium
def tukey_taper(n, a=0.05):
    """Tukey window with taper fraction a, sampled at 0..n-1 (solex_util.py:456-472)."""
    x = np.arange(n, dtype=np.float64)
    x = np.where(x > n / 2, n - x, x)                           # t(N - x) for the right half
    rising = 0.5 * (1 - np.cos(2 * np.pi * x / (a * n))) if n else x
    return np.where(x < a * n / 2, rising, 1.0)


def transversalium_gain(y_ratios_r, y1, y2, n_rows, strength):
    """Per-row gain from the per-row robust log-ratios (solex_util.py:400-404, 456-479)."""
    y_ratios_r = np.asarray(y_ratios_r, dtype=np.float64)
    trend = savgol_filter(y_ratios_r, min(strength, len(y_ratios_r) // 2 * 2 - 1), 3)
    detrended = y_ratios_r - trend
    detrended -= np.mean(detrended)
    correction = np.exp(-np.cumsum(detrended))
    n = correction.shape[0]
    tapered = np.ones(n) + (correction - np.ones(n)) * tukey_taper(n)
    c = np.ones(n_rows)
    c[y1:y2] = tapered
    return c


def savgol_window(n, strength):
    """Window of the trend filter for n rows (solex_util.py:400)."""
    return min(strength, n // 2 * 2 - 1)


def savgol_cubic_rows(x, window):
    """scipy.signal.savgol_filter(x, window, 3, axis=-1) (mode 'interp') for a
    (S, n) batch in O(S*n): the smoothing weights of a local cubic fit are
    a + b*k^2 in the offset k, so the filter is a combination of windowed sums of
    x, j*x and j^2*x taken from prefix sums; the first / last window//2 samples
    come from a cubic fitted to the first / last window, as scipy does.  Agrees
    with scipy to ~1e-12 of the signal scale (the gain tolerance is 1e-5)."""
    x = np.asarray(x, dtype=np.float64)
    s, n = x.shape
    m = window // 2
    if window < 5 or window > n:
        return savgol_filter(x, window, 3, axis=-1)
    den = (2 * m + 3) * (2 * m + 1) * (2 * m - 1)
    a = 3.0 * (3 * m * m + 3 * m - 1) / den
    b = -15.0 / den
    j = np.arange(n, dtype=np.float64) - n / 2.0
    p0 = np.concatenate([np.zeros((s, 1)), np.cumsum(x, axis=1)], axis=1)
    p1 = np.concatenate([np.zeros((s, 1)), np.cumsum(x * j, axis=1)], axis=1)
    p2 = np.concatenate([np.zeros((s, 1)), np.cumsum(x * j * j, axis=1)], axis=1)
    k = n - window + 1                                                    # interior outputs i = m .. n-m-1
    s0 = p0[:, window:window + k] - p0[:, :k]
    s1 = p1[:, window:window + k] - p1[:, :k]
    s2 = p2[:, window:window + k] - p2[:, :k]
    ji = j[m:n - m]
    out = np.empty_like(x)
    out[:, m:n - m] = a * s0 + b * (s2 - 2.0 * ji * s1 + ji * ji * s0)
    t = np.arange(window, dtype=np.float64)
    head = np.polyfit(t, x[:, :window].T, 3)                              # (4, S), highest power first
    tail = np.polyfit(t, x[:, n - window:].T, 3)
    th = np.arange(0, m, dtype=np.float64)[:, None]
    tt = np.arange(window - m, window, dtype=np.float64)[:, None]
    out[:, :m] = (((head[0] * th + head[1]) * th + head[2]) * th + head[3]).T
    out[:, n - m:] = (((tail[0] * tt + tail[1]) * tt + tail[2]) * tt + tail[3]).T
    return out


def transversalium_gains(stats, y1, y2, n_rows, strength):
    """transversalium_gain for a (S, n) batch of per-row statistics in one
    vectorised pass (savgol_filter, cumsum and exp work along the last axis)."""
    stats = np.asarray(stats, dtype=np.float64)
    s = stats.shape[0]
    ratios = np.concatenate([np.zeros((s, 1)), stats], axis=1)           # y_ratios_r[0] = 0 (solex_util.py:386)
    n = ratios.shape[1]
    trend = savgol_cubic_rows(ratios, savgol_window(n, strength))
    detrended = ratios - trend
    detrended -= np.mean(detrended, axis=1, keepdims=True)
    correction = np.exp(-np.cumsum(detrended, axis=1))
    tapered = np.ones(n) + (correction - np.ones(n)) * tukey_taper(n)
    c = np.ones((s, n_rows))
    c[:, y1:y2] = tapered
    return c


def correct_transversalium2(img, circle, borders, options, reqFlag, basefich):
    """Remove horizontal line defects: per-row gain from robust row-to-row
    log-ratios inside the disk.  `img` may be a DeviceImage or an ndarray;
    returns the same kind."""
    if options.get('stubborn_transversalium'):
        raise Exception('stubborn transversalium is a GUI-only option of the reference and is not part of this path')
    eng = get_engine()
    on_device = isinstance(img, DeviceImage)
    dev = img.rows_tensor() if on_device else _upload_u16(eng, img)
    y1, y2, rows, xa, xb = eng.transversalium_chords(circle, borders)
    stats = eng.transversalium_row_stats(dev, rows, xa, xb)
    y_ratios_r = np.concatenate([[0.0], stats])
    c = transversalium_gain(y_ratios_r, y1, y2, dev.shape[0], options['trans_strength'])
    options['_transversalium_cache'] = c
    if (not reqFlag) and (not options['clahe_only'] and not options['protus_only']):
        _plot_gain(c, output_path(basefich + '_transversalium_correction.png', options))
    out = eng.row_scale(dev, c)
    return DeviceImage(eng, out) if on_device else out.cpu().numpy()


def _plot_gain(c, path):
    """`_transversalium_correction.png` (solex_util.py:482-488): the per-row correction factor."""
    from . import miniplot
    if not miniplot.have_matplotlib():
        miniplot.line_plot(path, c, 'y', 'transversalium correction factor')
        return
    import matplotlib.figure
    fig = matplotlib.figure.Figure()
    ax = fig.add_subplot(1, 1, 1)
    ax.plot(c)
    ax.set_xlabel('y')
    ax.set_ylabel('transversalium correction factor')
    fig.savefig(path, dpi=300)


def _upload_u16(eng, img):
    import torch
    a = np.ascontiguousarray(np.asarray(img), dtype=np.uint16)
    return torch.from_numpy(a).to(eng.device)


# ---------------------------------------------------- host tail (not hot path)
def rescale_brightness(img, lo, hi, alpha=1.0):
    sat = np.iinfo(img.dtype).max
    assert sat >= hi > lo
    out = float(sat) * alpha * (img - lo) / (hi - lo)
    np.clip(out, 0, sat, out=out)
    return out.astype(img.dtype)


def image_process(frame, cercle, options, header, basefich):
    """CLAHE, brightness rescales, protuberance disk, rotation and the PNG / FITS
    writers (solex_util.py:527-588).  A DeviceImage is processed on the GPU (CLAHE and the rescales: the
    image_process_device path, bit-identical); arrays take the host path with OpenCV like the reference."""
    if isinstance(frame, DeviceImage) and not os.environ.get('SHG_HOST_TAIL'):
        return image_process_device(frame, cercle, options, header, basefich)
    frame = np.asarray(frame).astype(np.uint16)
    cl1 = cv2.createCLAHE(clipLimit=0.8, tileGridSize=(2, 2)).apply(frame)
    bright = np.percentile(frame, 99.9999)
    frame_hc = rescale_brightness(frame, bright * 0.25, bright)
    frame_protus = rescale_brightness(frame, 0, bright * 0.18)
    cc = rescale_brightness(cl1, np.percentile(cl1, 10), np.max(cl1))
    return _finish_images(frame, frame_hc, frame_protus, cc, cl1, cercle, options, header, basefich)


def _percentile_from_hist(hist, q):
    """np.percentile(values, q) (method 'linear') of the uint16 values whose 65536-bin histogram is `hist`."""
    from .ellipse_fit import percentile_from_pair
    n = int(hist.sum())
    cum = np.cumsum(hist)
    virtual = (n - 1) * np.true_divide(q, 100)
    k = int(np.floor(virtual))
    a = int(np.searchsorted(cum, k + 1, side='left'))                 # value of rank k (0-based)
    b = int(np.searchsorted(cum, min(k + 2, n), side='left'))
    return percentile_from_pair(float(a), float(b), n, q)


def image_process_device(image, cercle, options, header, basefich, _pool=None):
    """image_process with CLAHE and the brightness rescales on the GPU (csrc/tail.cu: OpenCV's CLAHE algorithm and
    NumPy's arithmetic step for step, so the images equal the host path's); only the images that are written or
    returned are copied to the host.  With `_pool` the host part (disc, rotation, PNG / FITS encoding) is submitted
    there and its future returned.  SURVEY.md 8f#2."""
    import ctypes as C
    import torch
    from ._lib import call, lib
    eng = get_engine()
    src = image.rows_tensor()
    assert src.is_contiguous() and src.dtype == torch.uint16
    rows, cols = src.shape
    n = rows * cols
    tiles = 2
    hists = torch.empty((tiles * tiles + 2, 65536), dtype=torch.int32, device=eng.device)   # tiles, frame, cl1
    lut = torch.empty((tiles * tiles, 65536), dtype=torch.uint16, device=eng.device)
    cl1 = torch.empty_like(src)
    st = eng.stream
    with eng.stage('tail:clahe'):
        call('shg_tile_hist_u16', src.data_ptr(), rows, cols, tiles, tiles, hists.data_ptr(), hists[tiles * tiles].data_ptr(), st)
        call('shg_clahe_lut', hists.data_ptr(), tiles * tiles, int(lib.shg_clahe_tile_area(rows, cols, tiles, tiles)), 0.8,
             lut.data_ptr(), st)
        call('shg_clahe_apply', src.data_ptr(), rows, cols, tiles, tiles, lut.data_ptr(), cl1.data_ptr(),
             hists[tiles * tiles + 1].data_ptr(), st)
        eng.n_launches += 3
    h = hists[tiles * tiles:].cpu().numpy().view(np.uint32).astype(np.int64)
    bright = _percentile_from_hist(h[0], 99.9999)
    dark_clahe = _percentile_from_hist(h[1], 10)
    bright_clahe = float(np.flatnonzero(h[1])[-1])

    def rescaled(t, lo, hi):
        assert 65535 >= hi > lo                                     # rescale_brightness's own assertion
        out = torch.empty_like(t)
        call('shg_rescale_u16', t.data_ptr(), n, float(lo), float(hi), out.data_ptr(), st)
        eng.n_launches += 1
        return out

    def host(t):
        return DeviceImage(eng, t).numpy()

    plots = not options['clahe_only'] and not options['protus_only']
    want_protus = True                                              # returned to the caller like the reference does
    with eng.stage('tail:rescale'):
        cc = host(rescaled(cl1, dark_clahe, bright_clahe))
        frame_protus = host(rescaled(src, 0, bright * 0.18)) if want_protus else None
        need_hc = plots or options['flag_display']
        frame_hc = host(rescaled(src, bright * 0.25, bright)) if need_hc else None
    frame = image.numpy() if plots else None
    cl1_h = host(cl1) if options['save_fit'] else None
    if _pool is not None:
        return _pool.submit(_finish_images, frame, frame_hc, frame_protus, cc, cl1_h, cercle, dict(options), header, basefich)
    return _finish_images(frame, frame_hc, frame_protus, cc, cl1_h, cercle, options, header, basefich)


def _finish_images(frame, frame_hc, frame_protus, cc, cl1, cercle, options, header, basefich):
    """Protuberance disc, rotation and the writers of image_process (solex_util.py:547-588) -- host side."""
    if not cercle == (-1, -1, -1) and options['disk_display'] and frame_protus is not None:
        r = int(cercle[2]) + options['delta_radius']
        if r > 0:
            frame_protus = cv2.circle(frame_protus, (int(cercle[0]), int(cercle[1])), r, 80, -1)
    turns = options['img_rotate'] // 90
    frame_raw, frame_hc, frame_protus, cc = (None if a is None else np.rot90(a, turns, axes=(0, 1))
                                             for a in (frame, frame_hc, frame_protus, cc))
    png = [cv2.IMWRITE_PNG_COMPRESSION, 0]
    if '_nolog' not in options:
        if options['clahe_only'] or not options['protus_only']:
            print('saving image to:' + basefich + '_clahe.png')
            cv2.imwrite(output_path(basefich + '_clahe.png', options), cc, png)
        if options['protus_only'] or not options['clahe_only']:
            cv2.imwrite(output_path(basefich + '_protus.png', options), frame_protus, png)
        if not options['clahe_only'] and not options['protus_only']:
            cv2.imwrite(output_path(basefich + '_uncontrasted.png', options), frame_raw, png)
            cv2.imwrite(output_path(basefich + '_high_contrast.png', options), frame_hc, png)
    if options['flag_display']:
        cv2.namedWindow('Sun images', cv2.WINDOW_NORMAL)
        cv2.imshow('Sun images', cv2.hconcat([cc, frame_hc, frame_protus]))
        cv2.waitKey(options.get('tempo', 1000))
        cv2.destroyAllWindows()
    if options['save_fit']:
        fits.PrimaryHDU(cl1, header).writeto(output_path(basefich + '_clahe.fits', options), overwrite='True')
    return cc, frame_protus


def removeVignette(frame_circularized, cercle0):
    raise Exception('de-vignette is a GUI-only option of the reference and is not part of this path')
