"""A uint16 image that lives in HBM but behaves like the NumPy array the
reference's callers expect (Solex_recon.solex_read returns a list of (ih, N)
uint16 arrays, /root/reference/Solex_recon.py:49-83).

At config 5 the 101 reconstructed disks are 16.5 GB: they stay on the device
and are only copied to the host when somebody actually asks for the pixels
(np.asarray(img), indexing, arithmetic, FITS/PNG writers).
"""
from __future__ import annotations

import numpy as np


class DeviceImage:
    """Logical shape (rows, cols), dtype uint16.

    layout 'rows'   : tensor is (rows, cols) row-major -- the reference layout;
    layout 'frames' : tensor is (cols, rows), i.e. a frame-major disk (N, ih);
                      ``flip`` means the logical image is np.flip(axis=1) of it
                      (the reference's flip_x, Solex_recon.py:75-76).
    """
    __array_priority__ = 100.0

    def __init__(self, engine, tensor, layout='rows', flip=False):
        assert layout in ('rows', 'frames')
        self.engine, self.tensor, self.layout, self.flip = engine, tensor, layout, bool(flip)
        self._host = None
        self._rows = None
        self.fit_future = None        # limb search started early on this image (solex_util.read_video_improved)
        self.min_ref = None           # (int32 device tensor, index): minimum pixel, tracked by the reconstruction

    # ---- array protocol -----------------------------------------------------
    @property
    def shape(self):
        a, b = self.tensor.shape
        return (int(a), int(b)) if self.layout == 'rows' else (int(b), int(a))

    dtype = np.dtype(np.uint16)
    ndim = 2

    @property
    def size(self):
        return self.shape[0] * self.shape[1]

    def rows_tensor(self):
        """Row-major (rows, cols) device tensor of the logical image."""
        if self.layout == 'rows':
            return self.tensor
        if self._rows is None:
            self._rows = self.engine.to_reference_layout(self.tensor, flip=self.flip)
        return self._rows

    def numpy(self):
        if self._host is None:
            import torch
            src = self.rows_tensor()
            pinned = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)      # cached pinned pool: full PCIe rate
            pinned.copy_(src, non_blocking=True)
            torch.cuda.current_stream(src.device).synchronize()
            self._host = pinned.numpy()
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __getitem__(self, idx):
        return self.numpy()[idx]

    def __len__(self):
        return self.shape[0]

    def flipped(self):
        """np.flip(self, axis=1) without touching the pixels."""
        if self.layout == 'frames':
            out = DeviceImage(self.engine, self.tensor, 'frames', not self.flip)
            out.fit_future = self.fit_future      # the early fit was started on the image as it will be used
            out.min_ref = self.min_ref
            return out
        return np.flip(self.numpy(), axis=1)

    def __getattr__(self, name):
        # any other ndarray attribute (astype, T, mean, ...) is answered by the host copy; a name ndarray does not
        # have (a typo, a hasattr probe) must not cost a 164 MB device -> host copy before it fails
        if name.startswith('__') or not hasattr(np.ndarray, name):
            raise AttributeError('%s object has no attribute %r' % (type(self).__name__, name))
        return getattr(self.numpy(), name)

    def _binary(self, other, op):
        return op(self.numpy(), np.asarray(other) if isinstance(other, DeviceImage) else other)

    def __truediv__(self, o):
        return self._binary(o, np.true_divide)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    __rmul__ = __mul__

    def __add__(self, o):
        return self._binary(o, np.add)

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __eq__(self, o):
        return self._binary(o, np.equal)

    __hash__ = None

    def __repr__(self):
        return 'DeviceImage(shape=%s, layout=%s, flip=%s)' % (self.shape, self.layout, self.flip)


class PartialImage:
    """This rank's frame rows of an image whose frames are spread over the GPUs of the box
    (parallel.reconstruct, exchange mode 'post_warp'): logical shape (ih, N) like every disk image,
    but only frames [k0, k1) -- plus `halo` frames of each neighbour -- are here.
    tensor: (halo + (k1 - k0) + halo, ih) frame-major; row r is frame k0 - halo + r of the scan.
    The pixels of the whole image never exist on one GPU: the circularisation works on the parts and
    the ranks exchange the (4-5 x smaller) circularised column blocks instead."""
    dtype = np.dtype(np.uint16)
    ndim = 2

    def __init__(self, engine, tensor, k0, k1, halo, n_frames, flip=False):
        self.engine, self.tensor = engine, tensor
        self.k0, self.k1, self.halo, self.n_frames, self.flip = int(k0), int(k1), int(halo), int(n_frames), bool(flip)
        self.min_ref = None           # (int32 device tensor, index): minimum pixel of the WHOLE image
        self.cval_ref = None          # (int32 device tensor, index): pixel [0][0] of the whole image
        self.full = None              # the complete image, on the rank that also holds it (the ellipse-fit shift)
        self.fit_future = None

    @property
    def shape(self):
        return (int(self.tensor.shape[1]), self.n_frames)

    def flipped(self):
        out = PartialImage(self.engine, self.tensor, self.k0, self.k1, self.halo, self.n_frames, not self.flip)
        out.min_ref, out.cval_ref = self.min_ref, self.cval_ref
        out.full = None if self.full is None else self.full.flipped()
        return out

    def __array__(self, dtype=None, copy=None):
        raise TypeError('this image is spread over the GPUs (exchange mode post_warp); its pixels exist only after '
                        'the circularisation.  Set SHG_EXCHANGE=by_shift to have complete disk images on their owners')

    def __repr__(self):
        return 'PartialImage(shape=%s, frames=[%d, %d), halo=%d, flip=%s)' % (self.shape, self.k0, self.k1, self.halo,
                                                                            self.flip)
