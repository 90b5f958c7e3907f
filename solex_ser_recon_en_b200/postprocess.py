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


def _common_base(pairs):
    """If every image is a slice of one (S, N, ih) disk tensor with the same flip,
    return (base tensor view, indices); else None."""
    first = pairs[0][0]
    flip = pairs[0][1]
    stride = first.shape[0] * first.shape[1]
    base_ptr = min(p[0].data_ptr() for p in pairs)
    idx = []
    for fr, fl in pairs:
        off = (fr.data_ptr() - base_ptr) // 2
        if fl != flip or fr.shape != first.shape or not fr.is_contiguous() or off % stride:
            return None
        idx.append(off // stride)
    return base_ptr, idx, stride


def start_minmax(images):
    """Launch the per-image min / max of the disks on a side stream.  It depends
    only on the disks, so solex_process starts it before the (host-heavy) ellipse
    fit and the two overlap.  Returns a handle for circularise_many, or None."""
    eng = get_engine()
    if len(images) < 2 or not all(isinstance(im, DeviceImage) and im.layout == 'frames' for im in images):
        return None
    pairs = [(im.tensor, im.flip) for im in images]
    common = _common_base(pairs)
    if common is None:
        return None
    base_ptr, idx, stride = common
    n, ih = pairs[0][0].shape
    first = min(pairs, key=lambda p: p[0].data_ptr())[0]
    base = torch.as_strided(first, (max(idx) + 1, n, ih), (stride, ih, 1))
    if not hasattr(eng, '_side_stream'):
        eng._side_stream = torch.cuda.Stream(device=eng.device)
    side = eng._side_stream
    side.wait_stream(torch.cuda.current_stream(eng.device))
    with torch.cuda.stream(side):
        with eng.stage('minmax'):
            mm = _minmax_with_tracked(eng, images, base, idx)
    done = torch.cuda.Event()
    done.record(side)
    return dict(ptrs=[p[0].data_ptr() for p in pairs], base=base, idx=idx, mm=mm, done=done)


def _minmax_with_tracked(eng, images, base, idx):
    """(n, 2) int32 clip range of every image.  The reconstruction kernel tracks the minimum of the
    images it writes (DeviceImage.min_ref) and only the LOWER clip can change a truncated pixel (see
    csrc/warp.cu), so tracked images get (min, 65535) without touching their pixels; the others (the
    ellipse-fit shift, reconstructed by the direct-load kernel) go through the min / max kernel."""
    refs = [getattr(im, 'min_ref', None) for im in images]
    unknown = [n for n, r in enumerate(refs) if r is None]
    if len(unknown) == len(images):
        return eng.minmax_device(base, idx)
    mm = torch.empty((len(images), 2), dtype=torch.int32, device=eng.device)
    mm[:, 1] = 65535
    by_tensor = {}
    for n, r in enumerate(refs):
        if r is not None:
            by_tensor.setdefault(id(r[0]), (r[0], [], []))
            by_tensor[id(r[0])][1].append(n)
            by_tensor[id(r[0])][2].append(r[1])
    for mins, pos, src in by_tensor.values():
        mm[eng.upload(np.asarray(pos, dtype=np.int64)), 0] = mins[eng.upload(np.asarray(src, dtype=np.int64))]
    if unknown:
        mm[eng.upload(np.asarray(unknown, dtype=np.int64))] = eng.minmax_device(base, [idx[n] for n in unknown])
    return mm


def circularise_many(images, phi, ratio, prepared=None):
    """Warp every image with the same (phi, ratio): one min/max launch and one
    warp launch for the whole set when the images are slices of one disk tensor
    (the normal case: read_video_improved's output).  Returns
    (list of row-major DeviceImage, mat 2x2, mat3, theta)."""
    eng = get_engine()
    if not images:
        return [], None, None, None
    pairs = [frames_of(eng, im) for im in images]
    n, ih = pairs[0][0].shape
    mat, mat3, out_shape, _, theta = geometry.warp_plan((ih, n), phi, ratio)
    common = _common_base(pairs)
    if common is not None and len(pairs) > 1:
        base_ptr, idx, stride = common
        first = min(pairs, key=lambda p: p[0].data_ptr())[0]
        span = max(idx) + 1
        # a (span, N, ih) view over the shared storage that starts at the lowest slice
        base = torch.as_strided(first, (span, n, ih), (stride, ih, 1))
        if prepared is not None and prepared['ptrs'] == [p[0].data_ptr() for p in pairs]:
            torch.cuda.current_stream(eng.device).wait_event(prepared['done'])
            mm = prepared['mm']
        else:
            with eng.stage('minmax'):
                mm = eng.minmax_device(base, idx)
        with eng.stage('warp'):
            out = eng.warp_batch(base, idx, pairs[0][1], mat3, out_shape, mm)
        return [DeviceImage(eng, out[i]) for i in range(len(pairs))], mat, mat3, theta
    res = []
    for fr, flip in pairs:
        with eng.stage('minmax'):
            mm = eng.minmax_device(fr)
        with eng.stage('warp'):
            res.append(DeviceImage(eng, eng.warp_batch(fr, None, flip, mat3, out_shape, mm)[0]))
    return res, mat, mat3, theta


def prepare_partial(parts):
    """The geometry-independent half of circularise_partial: which slice of the local buffer each image is, the
    launch order, clip ranges and fill constants on the device.  Called before the ellipse geometry is known (while
    rank 0 fits it), so that nothing but the warp plan stands between the broadcast and the launch."""
    from . import parallel
    eng = get_engine()
    rank, size = parallel.world()
    with eng.stage('circ_prep'):
        p0 = parts[0]
        flip = p0.flip
        n_ext, ih = p0.tensor.shape
        stride = n_ext * ih
        base_ptr = min(p.tensor.data_ptr() for p in parts)
        first = min(parts, key=lambda p: p.tensor.data_ptr()).tensor
        idx = []
        for p in parts:
            off = (p.tensor.data_ptr() - base_ptr) // 2
            assert off % stride == 0 and p.tensor.is_contiguous() and p.flip == flip and p.tensor.shape == p0.tensor.shape
            idx.append(off // stride)
        base = torch.as_strided(first, (max(idx) + 1, n_ext, ih), (stride, ih, 1))
        owner = parallel.shift_owner(len(parts), size, 'by_shift')
        # Launch order of the images (blockIdx.z): every rank starts with the images of the NEXT rank and goes
        # round, so that at any moment the ranks store into different owners.  In list order all ranks wrote into
        # rank 0's memory first, then all into rank 1's, ...: the owner's NVLink ingress was the bottleneck and
        # everybody else's egress idled (measured at four GPUs: 3.6 - 4.1 ms for this step instead of 2.3).
        order = sorted(range(len(parts)), key=lambda q: ((owner[q] - rank - 1) % size, q))
        mins, red = p0.min_ref[0], p0.cval_ref[0]
        assert all(p.min_ref[0].data_ptr() == mins.data_ptr() and p.cval_ref[0].data_ptr() == red.data_ptr() for p in parts)
        src = eng.upload(np.asarray([parts[q].min_ref[1] for q in order], dtype=np.int64))
        mm = torch.empty((len(parts), 2), dtype=torch.int32, device=eng.device)
        mm[:, 0] = mins[src]
        mm[:, 1] = 65535
        cvals = red[2 if flip else 1][src].contiguous()              # image[0][0]: last / first physical frame
        sel = eng.upload(np.asarray([idx[q] for q in order], dtype=np.int32))
    return dict(base=base, order=order, sel=sel, mm=mm, cvals=cvals, owner=owner)


def circularise_partial(parts, phi, ratio, prepared=None):
    """Exchange mode 'post_warp' (parallel.py): every rank warps ITS frames of every image in `parts`
    (device_image.PartialImage, all slices of one local buffer) into a local full-width image and copies the
    pixels it produced into the rank that owns the image (peer memory).  Returns a list with a row-major
    DeviceImage for the images this rank owns and None for the others; owners are assigned by position in the list."""
    from . import parallel
    eng = get_engine()
    rank, size = parallel.world()
    prep = prepared if prepared is not None else prepare_partial(parts)
    p0 = parts[0]
    ih, n_frames = p0.shape
    flip = p0.flip
    mat, mat3, out_shape, _, theta = geometry.warp_plan((ih, n_frames), phi, ratio)
    oh, ow = int(out_shape[0]), int(out_shape[1])
    ex = parallel.circ_exchange(len(parts), oh, ow)
    if getattr(ex, 'ptrs_rot', None) is None:
        ex.ptrs_rot = eng.upload(ex.ptrs.view(np.int64)[prep['order']]).view(torch.int64)
    own_lo, own_hi = parallel.owned_logical_frames(n_frames, rank, size, flip)
    parallel.device_barrier()              # the owners are done reading the previous scan's circularised images
    with eng.stage('warp'):
        eng.warp_batch(prep['base'], prep['sel'], flip, mat3, (oh, ow), prep['mm'], n_frames=n_frames,
                       frame_origin=p0.k0 - p0.halo, cvals=prep['cvals'], window=(own_lo, own_hi), out_ptrs=ex.ptrs_rot)
    parallel.device_barrier()              # every rank's pixels have landed
    out = [None] * len(parts)
    for n, q in enumerate(ex.mine):
        out[q] = DeviceImage(eng, ex.images[n])
    return out


def _stack_rows(eng, images):
    """(S, h, w) tensor of row-major device images (a view when they already are
    consecutive slices of one tensor, else a copy)."""
    tens = [im.rows_tensor() if isinstance(im, DeviceImage) else
            torch.from_numpy(np.ascontiguousarray(np.asarray(im), dtype=np.uint16)).to(eng.device) for im in images]
    h, w = tens[0].shape
    step = h * w * 2
    if all(t.is_contiguous() and t.shape == (h, w) and t.data_ptr() == tens[0].data_ptr() + i * step
           for i, t in enumerate(tens)) and tens[0].untyped_storage().nbytes() - \
            (tens[0].data_ptr() - tens[0].untyped_storage().data_ptr()) >= len(tens) * step:
        return torch.as_strided(tens[0], (len(tens), h, w), (h * w, w, 1))
    return torch.stack(tens)


def detransversalium_many(images, circle, borders, strength):
    """Transversalium correction of several row-major device images that share
    the disk geometry.  Returns (list of DeviceImage, gains (S, rows))."""
    eng = get_engine()
    if not images:
        return [], np.zeros((0, 0))
    with eng.stage('transv_prep'):
        batch = _stack_rows(eng, images)
        y1, y2, rows, xa, xb = eng.transversalium_chords(circle, borders)
    with eng.stage('transv_stats'):
        stats = eng.transversalium_row_stats(batch, rows, xa, xb, device=True)
    h = batch.shape[1]
    with eng.stage('transv_gain'):
        gains_d = eng.transversalium_gains(stats, y1, y2, h, strength)
    early = eng.download_early(gains_d)
    with eng.stage('row_scale'):
        out = eng.row_scale(batch, gains_d)
    gains = early.result()                           # options['_transversalium_cache'] is a host array upstream;
    #                                                  copied while the row scaling runs, not behind it
    return [DeviceImage(eng, out[i]) for i in range(len(images))], gains
