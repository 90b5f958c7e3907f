"""Drop-in for the reference's Solex_recon module (/root/reference/Solex_recon.py):
solex_do_work, solex_read, solex_process, single_image_process with the same
arguments, return values, `options` mutations and log lines.

Differences that follow from running on a GPU:
 * the reference post-processes in a forked multiprocessing.Pool(4)
   (Solex_recon.py:30-42); a forked child cannot use the parent's CUDA context,
   so the GPU stages of solex_process run in the calling process and only the
   host tail (CLAHE, PNG / FITS writers) is handed to worker threads, which
   keeps the reference's overlap of "read file i+1 while file i is written";
 * disk images are DeviceImage array-likes: they stay in HBM until a writer or
   a caller asks for the pixels.
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import fits_min as fits
from . import parallel, postprocess
from .device_image import DeviceImage, PartialImage
from .ellipse_to_circle import correct_image, ellipse_to_circle, fit_geometry
from .engine import get_engine
from .solex_util import (clearlog, compute_mean_return_fit, correct_transversalium2, image_process, logme,
                         make_header, output_path, read_video_improved, write_complete)
from .video_reader import video_reader


def solex_do_work(tasks, flag_command_line=False):
    """Process a list of (file, options) tasks; returns None, raises on the first bad file
    like the reference (the caller's try/except reports it).

    The reference overlaps "read file i+1" with "post-process file i" through a process pool
    (Solex_recon.py:30-42).  Here file i+1 is ingested (pinned ring -> H2D on the ingest streams,
    the call releases the GIL) while one worker thread runs the transversalium correction and the
    device -> host copies of file i and four more run the host tails (CLAHE + PNG / FITS)."""
    def opened():
        for file, options in tasks:
            print('file %s is processing' % file)
            yield video_reader(file), options, os.path.splitext(file)[0]
    solex_do_work_readers(opened())


def solex_do_work_readers(tasks):
    """solex_do_work on already opened readers: `tasks` yields (reader, options, basefich0).

    Pipeline (any number of GPUs): the calling thread runs solex_read of scan i+1 -- and the part of
    solex_process that contains collectives (ellipse geometry, circularisation) -- while a worker thread finishes
    scan i (transversalium, device -> host copies, hand-over to the host tails).  The worker issues no collective,
    so the ranks' collective order is the calling threads' order; before scan i+1 reconstructs into buffers that
    scan i's owners may still be reading, the caller waits for the worker to have finished scan i
    (options['_before_recon'])."""
    def on_engine_device():
        # CUDA's current device is per host thread: the post-processing thread must use the engine's GPU
        # (SHG_DEVICE / LOCAL_RANK), not device 0
        import torch
        torch.cuda.set_device(get_engine().device)

    with ThreadPoolExecutor(max_workers=1, initializer=on_engine_device) as gpu_post, \
            ThreadPoolExecutor(max_workers=4) as tails:
        posts = []
        for rdr, options, basefich0 in tasks:
            previous = posts[-1] if posts else None
            options['_before_recon'] = (lambda p=previous: p.result()) if previous is not None else None
            disk_list, backup_bounds, hdr = solex_read_reader(rdr, options, basefich0)
            options.pop('_before_recon', None)
            posts.append(solex_process(options, disk_list, backup_bounds, hdr, tails, _defer=gpu_post))
            del disk_list
        for post in posts:
            for fut in post.result():
                fut.result()


def solex_read(file, options):
    """Mean frame + line fit + reconstruction of one scan.
    Returns (disk_list, (y1, y2), hdr); mutates options['basefich0'],
    ['shift_requested'] and ['shift'] exactly as the reference (Solex_recon.py:49-83)."""
    return solex_read_reader(video_reader(file), options, os.path.splitext(file)[0])


def solex_read_reader(rdr, options, basefich0):
    """solex_read on an already opened reader (a video_reader, or a
    video_reader.memory_scan wrapping a payload that sits in host memory)."""
    options['basefich0'] = basefich0
    log = basefich0 + '_log.txt'
    clearlog(log, options)
    logme(log, options, 'Pixel shift : ' + str(options['shift']))
    options['shift_requested'] = options['shift']
    options['shift'] = list(dict.fromkeys([options['ellipse_fit_shift'], 0] + options['shift']))
    hdr = make_header(rdr)
    mean_img, fit, backup_y1, backup_y2 = compute_mean_return_fit(rdr, options, hdr, rdr.iw, rdr.ih, basefich0)
    if not os.environ.get('SHG_NO_EARLY_FIT'):   # solex_process follows: start the limb search under the reconstruction
        options['_prefetch_fit'] = True
    try:
        disk_list, ih, iw, _ = read_video_improved(rdr, fit, options)
    finally:
        options.pop('_prefetch_fit', None)
    hdr['NAXIS1'] = iw
    observer = options.get('_observer')                        # tests / diagnostics: see the seams of a CLI run
    for i in range(len(disk_list)):
        if disk_list[i] is None:                              # image owned by another rank
            continue
        if options['flip_x']:
            disk_list[i] = disk_list[i].flipped() if isinstance(disk_list[i], (DeviceImage, PartialImage)) \
                else np.flip(disk_list[i], axis=1)
        if options['save_fit'] and options['shift'][i] in options['shift_requested']:
            basefich = basefich0 + '_shift=' + str(options['shift'][i])
            fits.PrimaryHDU(np.asarray(disk_list[i]), header=hdr).writeto(
                output_path(basefich + '_raw.fits', options), overwrite='True')
    if observer is not None:
        observer('disks', basefich0, disk_list)
    return disk_list, (backup_y1, backup_y2), hdr


def solex_process(options, disk_list, backup_bounds, hdr, _pool=None, _defer=None):
    """Circularise, de-transversalium, crop and write every requested shift
    (Solex_recon.py:93-133).  With `_pool` the host tail of each image is
    submitted to it and the futures are returned; otherwise it runs inline and
    the return value is None like the reference's.  With `_defer` (an executor) everything after the
    circularisation -- which holds the only collectives -- runs there and a future of the tail futures is
    returned at once."""
    basefich0 = options['basefich0']
    log = basefich0 + '_log.txt'
    if options['transversalium']:
        logme(log, options, 'Transversalium correction : ' + str(options['trans_strength']))
    else:
        logme(log, options, 'Transversalium disabled')
    logme(log, options, 'Mirror X : ' + str(options['flip_x']))
    logme(log, options, 'Post-rotation : ' + str(options['img_rotate']) + ' degrees')
    logme(log, options, f'Protus adjustment : {options["delta_radius"]}')
    logme(log, options, f'de-vignette : {options["de-vignette"]}')
    if options['de-vignette']:
        raise Exception('de-vignette is a GUI-only option of the reference and is not part of this path')
    shifts = options['shift']
    if options.get('_exchange') == 'post_warp':
        # frames of every image on every rank (parallel.reconstruct_partial): rank 0 fits the ellipse on the
        # gathered first image, every rank circularises its own frames, owners receive the result
        requested, circular, cercle0, borders = _circularise_post_warp(options, disk_list, shifts, basefich0)
    else:
        requested, circular, cercle0, borders = _circularise_owned(options, disk_list, shifts, basefich0)
    if _defer is not None:
        import torch
        launched = torch.cuda.Event()
        launched.record()                                     # the worker's kernels follow the circularisation

        def finish():
            torch.cuda.current_stream().wait_event(launched)
            return _finish_process(options, hdr, backup_bounds, shifts, requested, circular, cercle0, borders, _pool)
        return _defer.submit(finish)
    return _finish_process(options, hdr, backup_bounds, shifts, requested, circular, cercle0, borders, _pool)


def _finish_process(options, hdr, backup_bounds, shifts, requested, circular, cercle0, borders, _pool):
    """Part 3 of solex_process: transversalium for all requested shifts (batched), then the host tail per
    image.  No collectives: with several ranks each one works on the images it owns."""
    basefich0 = options['basefich0']
    log = basefich0 + '_log.txt'
    images = [circular[i] for i in requested]
    plot_jobs = []
    if options['transversalium'] and images and all(isinstance(im, DeviceImage) for im in images) \
            and not options['save_fit']:
        if not cercle0 == (-1, -1, -1):
            circle, bord = cercle0, borders
        else:
            circle = (0, 0, 99999)
            bord = [0, backup_bounds[0] + 20, images[0].shape[1] - 1, backup_bounds[1] - 20]
        detrans, gains = postprocess.detransversalium_many(images, circle, bord, options['trans_strength'])
        options['_transversalium_cache'] = gains[-1]
        options['_transversalium_gains'] = {shifts[i]: gains[j] for j, i in enumerate(requested)}
        if not options['clahe_only'] and not options['protus_only']:
            # the reference plots the correction of every requested shift (solex_util.py:482-488)
            from .solex_util import _plot_gain
            for j, i in enumerate(requested):
                target = output_path(basefich0 + '_shift=' + str(shifts[i]) + '_transversalium_correction.png', options)
                if _pool is not None:
                    plot_jobs.append(_pool.submit(_plot_gain, gains[j].copy(), target))
                else:
                    _plot_gain(gains[j], target)
    else:
        detrans = None
    futures = []
    observer = options.get('_observer')
    for j, i in enumerate(requested):
        basefich = basefich0 + '_shift=' + str(shifts[i])
        if observer is not None:
            observer('circular', basefich, images[j])
            if detrans is not None:
                observer('detrans', basefich, detrans[j])
        res = single_image_process(images[j], hdr, options, cercle0, borders, basefich, backup_bounds, _pool=_pool,
                                   _detrans=None if detrans is None else detrans[j])
        if _pool is not None and hasattr(res, 'result'):      # (a result sink returns nothing to wait for)
            futures.append(res)
        write_complete(log, options)
    return futures + plot_jobs if _pool is not None else None


def _circularise_owned(options, disk_list, shifts, basefich0):
    """Geometry + circularisation of the requested shifts whose complete disk images this rank holds
    (all of them on one GPU).  Returns (requested indices, {index: circularised image}, cercle0, borders)."""
    borders = [0, 0, 0, 0]
    cercle0 = (-1, -1, -1)
    # with several ranks each one post-processes the shifts whose images it owns
    requested = [i for i in range(len(disk_list))
                 if shifts[i] in options['shift_requested'] and disk_list[i] is not None]
    circular = {}
    # the min / max of the disks (the warp's clip range) only depends on the disks: start it now, on a
    # side stream, so it runs underneath the ellipse fit
    plots_wanted = not options['clahe_only'] and not options['protus_only']
    early = [i for i in requested if not (i == 0 and (plots_wanted or options['ratio_fixe'] is not None
                                                     or options['slant_fix'] is not None))]
    prepared = postprocess.start_minmax([disk_list[i] for i in early])
    # 1. geometry: disk_list[0] is the ellipse-fit shift; the fit is made once (by its owner) and reused
    if options['ratio_fixe'] is None and options['slant_fix'] is None:
        geom, failure = None, None
        if disk_list[0] is not None:
            basefich = basefich0 + '_shift=' + str(shifts[0])
            plots = not options['clahe_only'] and not options['protus_only']
            try:
                if plots:
                    # diagnostic figure wanted (the reference draws `_ellipse_fit.png` for the ellipse-fit shift
                    # whether or not that shift was requested, ellipse_to_circle.py:316-341): the one-image path
                    fixed, cercle0, ratio_fit, phi, borders = ellipse_to_circle(disk_list[0], options, basefich)
                    if 0 in requested:
                        circular[0] = fixed
                else:
                    fit = fit_geometry(disk_list[0], options, basefich)
                    cercle0, ratio_fit, phi, borders = fit['circle'], fit['ratio'], fit['phi'], fit['borders']
                geom = (tuple(float(v) for v in cercle0), float(ratio_fit), float(phi), [float(b) for b in borders])
            except Exception as e:                            # the other ranks must not wait for a result
                failure = e
        with get_engine().stage('geometry_bcast'):
            cercle0, options['ratio_fixe'], phi, borders = parallel.broadcast_geometry(geom, 0, failure)
        options['slant_fix'] = math.degrees(phi)
        todo = [i for i in requested if i not in circular]
    else:
        todo = list(requested)
        if todo and todo[0] == 0:                             # the reference logs the matrix for i == 0 only
            circular[0] = correct_image(disk_list[0], math.radians(options['slant_fix'] or 0.0),
                                        options['ratio_fixe'] if options['ratio_fixe'] is not None else 1.0,
                                        np.array([-1.0, -1.0]), -1.0, options, print_log=True)[0]
            todo = todo[1:]
    ratio = options['ratio_fixe'] if options['ratio_fixe'] is not None else 1.0
    phi = math.radians(options['slant_fix']) if options['slant_fix'] is not None else 0.0
    # 2. circularise every requested shift with that geometry: one min/max + one warp launch for the set
    if todo:
        warped, _, _, _ = postprocess.circularise_many([disk_list[i] for i in todo], phi, ratio,
                                                       prepared if todo == early else None)
        circular.update(zip(todo, warped))
    return requested, circular, cercle0, borders


def _circularise_post_warp(options, disk_list, shifts, basefich0):
    """The same for frame-sharded images (exchange mode 'post_warp'): the list of requested shifts is the
    same on every rank; the circularised images land on their owners (by position in that list)."""
    req_all = [i for i in range(len(disk_list)) if shifts[i] in options['shift_requested']]
    # everything of the circularisation that does not depend on the geometry (index lists, clip ranges, fill
    # constants: a few small uploads and index kernels) is prepared BEFORE waiting for the fit / the broadcast
    prepared = postprocess.prepare_partial([disk_list[i] for i in req_all])
    geom, failure = None, None
    full0 = disk_list[0].full
    if full0 is not None:                                     # rank 0: the gathered ellipse-fit image
        try:
            fit = fit_geometry(full0, options, basefich0 + '_shift=' + str(shifts[0]))
            geom = (tuple(float(v) for v in fit['circle']), float(fit['ratio']), float(fit['phi']),
                    [float(b) for b in fit['borders']])
        except Exception as e:
            failure = e
    with get_engine().stage('geometry_bcast'):
        cercle0, options['ratio_fixe'], phi, borders = parallel.broadcast_geometry(geom, 0, failure)
    options['slant_fix'] = math.degrees(phi)
    phi = math.radians(options['slant_fix'])                  # the same degrees round trip as the single-GPU path
    warped = postprocess.circularise_partial([disk_list[i] for i in req_all], phi, options['ratio_fixe'], prepared)
    requested = [i for q, i in enumerate(req_all) if warped[q] is not None]
    circular = {i: warped[q] for q, i in enumerate(req_all) if warped[q] is not None}
    return requested, circular, cercle0, borders


def _crop(img, cercle, options):
    """Square / fixed-width crop centred on the disk (Solex_recon.py:155-171)."""
    h, w = img.shape
    nw = h if options['fixed_width'] is None else options['fixed_width']
    half = nw // 2
    cx = w // 2 if cercle == (-1, -1, -1) else int(cercle[0])
    shift = half - cx
    out = np.full((h, nw), img[0, 0], dtype=img.dtype)
    a, b = max(0, cx - half), min(cx + half, w)
    out[:, :b - a] = img[:, a:b]
    if shift > 0:
        out = np.roll(out, shift, axis=1)
        out[:, :shift] = img[0, 0]
    if not cercle == (-1, -1, -1):
        cercle = (half, cercle[1], cercle[2])
    return out, cercle


def single_image_process(frame_circularized, hdr, options, cercle0, borders, basefich, backup_bounds, _pool=None,
                         _detrans=None):
    """Transversalium correction on the GPU, then the host tail
    (Solex_recon.py:136-174).  Returns image_process's (cc, frame_protus), or a
    future of it when a pool is given.  `_detrans` carries the already corrected
    image when solex_process batched that stage."""
    if options['save_fit']:
        fits.PrimaryHDU(np.asarray(frame_circularized), header=hdr).writeto(
            output_path(basefich + '_circular.fits', options), overwrite='True')
    if _detrans is not None:
        detrans = _detrans
    elif options['transversalium']:
        if not cercle0 == (-1, -1, -1):
            detrans = correct_transversalium2(frame_circularized, cercle0, borders, options, 0, basefich)
        else:
            detrans = correct_transversalium2(
                frame_circularized, (0, 0, 99999),
                [0, backup_bounds[0] + 20, frame_circularized.shape[1] - 1, backup_bounds[1] - 20], options, 0, basefich)
    else:
        detrans = frame_circularized
    if _detrans is None and options.get('_observer') is not None:
        options['_observer']('detrans', basefich, detrans)
    sink = options.get('_result_sink')
    if sink is not None:                                      # callers that want the hot-path result itself
        return sink(basefich, detrans, cercle0)
    crop = options['fixed_width'] is not None or options['crop_width_square']
    if isinstance(detrans, DeviceImage) and not crop and not (options['save_fit'] and options['transversalium']) \
            and not os.environ.get('SHG_HOST_TAIL'):
        # CLAHE + brightness rescales on the device; only the images that get written come back to the host
        from .solex_util import image_process_device
        return image_process_device(detrans, cercle0, options, hdr, basefich, _pool=_pool)
    host = np.asarray(detrans)                                # the one device -> host copy of this image
    if options['save_fit'] and options['transversalium']:
        fits.PrimaryHDU(host, header=hdr).writeto(output_path(basefich + '_detransversaliumed.fits', options),
                                                  overwrite='True')
    cercle = cercle0
    if options['fixed_width'] is not None or options['crop_width_square']:
        host, cercle = _crop(host, cercle, options)
    if _pool is not None:
        return _pool.submit(image_process, host, cercle, dict(options), hdr, basefich)
    return image_process(host, cercle, options, hdr, basefich)
