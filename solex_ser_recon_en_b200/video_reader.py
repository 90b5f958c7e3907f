"""SER / AVI frame source with the reference's interface
(/root/reference/video_reader.py:10-158) plus what the device path needs.

Same attributes (FrameCount, Width, Height, ih, iw, flag_rotate, FrameIndex,
count, infilebytes, infiledatatype, PixelDepthPerPlane, SER_flag, AVI_flag) and
methods (next_frame, has_frames; reset on all_video_reader), so existing callers
keep working.  What is new: the payload is described as (offset, frame stride)
so that the ingest ring can stream it into HBM without per-frame Python; the
frame-by-frame next_frame() is kept for callers that still want host frames
(display, the spectral analyser) and is served from a memory map.
"""
from __future__ import annotations

import os
import struct

import numpy as np

SER_HEADER_BYTES = 178
_RAW8_FOURCCS = (b'Y800', b'GREY', b'Y8  ', b'Y8\x00\x00')


def _parse_avi_raw8(path):
    """Locate the frames of an uncompressed 8-bit grey AVI (top-down Y800/GREY
    payload, one chunk per frame).  Returns (width, height, offsets) or None when
    the file is anything else (compressed, palettised DIB, ...) and has to be
    decoded by OpenCV/FFmpeg like the reference does (video_reader.py:68-80)."""
    size = os.path.getsize(path)
    with open(path, 'rb') as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b'RIFF' or head[8:12] != b'AVI ':
            return None
        width = height = None
        ok_format = False
        offsets = []
        pos = 12
        movi = None
        while pos + 8 <= size:
            f.seek(pos)
            cc, sz = struct.unpack('<4sI', f.read(8))
            if cc == b'LIST':
                kind = f.read(4)
                if kind == b'movi':
                    movi = (pos + 12, pos + 8 + sz)
                    pos += 8 + sz + (sz & 1)
                    continue
                pos += 12                       # descend into hdrl / strl
                continue
            if cc == b'strf' and width is None:
                bih = f.read(min(sz, 40))
                if len(bih) >= 40:
                    _, w, h, _, bits, comp = struct.unpack('<IiiHH4s', bih[:20])
                    if bits == 8 and comp in _RAW8_FOURCCS and w > 0 and h != 0:
                        width, height, ok_format = w, abs(h), True
            pos += 8 + sz + (sz & 1)
        if not ok_format or movi is None:
            return None
        pos, end = movi
        while pos + 8 <= min(end, size):
            f.seek(pos)
            cc, sz = struct.unpack('<4sI', f.read(8))
            if cc == b'LIST':                   # 'rec ' groups
                pos += 12
                continue
            if cc[2:4] in (b'dc', b'db') and cc[:2] == b'00':
                if sz != width * height:
                    return None
                offsets.append(pos + 8)
            pos += 8 + sz + (sz & 1)
    if not offsets:
        return None
    return width, height, np.asarray(offsets, dtype=np.int64)


class video_reader:

    def __init__(self, file, buffer_size=25):
        self.file = file
        self.path = file
        self.buffer_size = buffer_size
        self.buffer_remaining = 0
        upper = file.upper()
        if upper.endswith('.SER'):
            self.SER_flag, self.AVI_flag = True, False
        elif upper.endswith('.AVI'):
            self.SER_flag, self.AVI_flag = False, True
            self.infiledatatype = 'uint8'
        else:
            raise Exception('error input file ' + file + 'neither is SER nor AVI')

        self._map = None
        self._cap = None
        self._offsets = None
        if self.SER_flag:
            with open(file, 'rb') as f:
                hdr = f.read(SER_HEADER_BYTES)
            if len(hdr) < 42:
                raise Exception('error input file ' + file + ' is too short to be a SER file')
            self.FileID = np.frombuffer(hdr, dtype='int8', count=14)
            fields = np.frombuffer(hdr, dtype='<u4', count=7, offset=14)
            # the reference keeps these as 1-element arrays / numpy scalars
            self.LuID, self.ColorID, self.littleEndian = fields[0:1], fields[1:2], fields[2:3]
            self.Width, self.Height = fields[3], fields[4]
            self.PixelDepthPerPlane = fields[5]
            self.FrameCount = fields[6]
            self.count = self.Width * self.Height
            if self.PixelDepthPerPlane == 8:
                self.infiledatatype, self.infilebytes = 'uint8', 1
            else:
                self.infiledatatype, self.infilebytes = 'uint16', 2
            self.offset = SER_HEADER_BYTES
            self.fileoffset = SER_HEADER_BYTES
            self.payload_offset = SER_HEADER_BYTES
            self.frame_stride = int(self.count) * self.infilebytes
        else:
            raw = _parse_avi_raw8(file)
            if raw is not None:
                self.Width, self.Height, self._offsets = int(raw[0]), int(raw[1]), raw[2]
                self.FrameCount = int(len(self._offsets))
                steps = np.diff(self._offsets)
                if len(steps) == 0 or np.all(steps == steps[0]):
                    self.payload_offset = int(self._offsets[0])
                    self.frame_stride = int(steps[0]) if len(steps) else self.Width * self.Height
                else:
                    self.payload_offset = self.frame_stride = None       # irregular: host gather
            else:
                import cv2
                self._cap = cv2.VideoCapture(file)
                self.file = self._cap
                self.Width = int(self._cap.get(cv2.CAP_PROP_FRAME_WIDTH))
                self.Height = int(self._cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
                self.FrameCount = int(self._cap.get(cv2.CAP_PROP_FRAME_COUNT))
                self.payload_offset = self.frame_stride = None
            self.PixelDepthPerPlane = 1 * 8
            self.count = self.Width * self.Height
            self.infilebytes = 1
            self.offset = 0
            self.fileoffset = 0
        self.FrameIndex = -1

        if self.Width > self.Height:
            self.flag_rotate = True
            self.ih, self.iw = self.Width, self.Height
        else:
            self.flag_rotate = False
            self.iw, self.ih = self.Width, self.Height

    # ---- what the device path uses -----------------------------------------
    @property
    def geometry(self):
        from .engine import ScanGeometry
        return ScanGeometry(int(self.Width), int(self.Height), int(self.infilebytes), int(self.FrameCount))

    @property
    def streamable(self):
        """True when the frames sit in the file at a constant stride and can be
        pread straight into the pinned ring."""
        return self.payload_offset is not None

    def raw_frames(self, k0, k1):
        """Raw (k1-k0, H, W) frames on the host, file dtype, no rotation / scaling."""
        W, H = int(self.Width), int(self.Height)
        dt = np.dtype(self.infiledatatype)
        if self.streamable:
            if self._map is None:
                self._map = np.memmap(self.path, dtype=np.uint8, mode='r')
            fb = W * H * dt.itemsize
            if self.frame_stride == fb:
                a = self._map[self.payload_offset + k0 * fb:self.payload_offset + k1 * fb]
                return a.view(dt.newbyteorder('<')).reshape(k1 - k0, H, W)
            out = np.empty((k1 - k0, H, W), dtype=dt)
            for j, k in enumerate(range(k0, k1)):
                o = self.payload_offset + k * self.frame_stride
                out[j] = self._map[o:o + fb].view(dt).reshape(H, W)
            return out
        if self._offsets is not None:
            if self._map is None:
                self._map = np.memmap(self.path, dtype=np.uint8, mode='r')
            out = np.empty((k1 - k0, H, W), dtype=np.uint8)
            for j, k in enumerate(range(k0, k1)):
                o = int(self._offsets[k])
                out[j] = self._map[o:o + W * H].reshape(H, W)
            return out
        # compressed AVI: sequential decode, as the reference (video_reader.py:111-113)
        import cv2
        out = np.empty((k1 - k0, H, W), dtype=np.uint8)
        if int(self._cap.get(cv2.CAP_PROP_POS_FRAMES)) != k0:
            self._cap.set(cv2.CAP_PROP_POS_FRAMES, k0)
        for j in range(k1 - k0):
            ret, img = self._cap.read()
            if not ret:
                raise Exception('error reading frame %d of the AVI file' % (k0 + j))
            out[j] = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if img.ndim == 3 else img
        return out

    # ---- the reference's frame interface ------------------------------------
    def next_frame(self):
        self.FrameIndex += 1
        self.offset = self.fileoffset + self.FrameIndex * int(self.count) * self.infilebytes
        img = self.raw_frames(self.FrameIndex, self.FrameIndex + 1)[0]
        if self.flag_rotate:
            img = np.rot90(img)
        if self.infiledatatype == 'uint8':
            img = np.asarray(img, dtype='uint16') * 256
        return img

    def has_frames(self):
        return self.FrameIndex + 1 < self.FrameCount


class _DeviceFrames:
    """all_video_reader.frames for a scan that lives in HBM: indexes like the (N, ih, iw) uint16 array of the
    reference, copying only the frames that are asked for to the host (oriented and scaled like next_frame)."""
    dtype = np.dtype(np.uint16)
    ndim = 3

    def __init__(self, owner):
        self._o = owner
        self.shape = (int(owner.FrameCount), owner.ih, owner.iw)

    def __len__(self):
        return self.shape[0]

    def _block(self, k0, k1):
        o = self._o
        blk = o.stack.host_frames(k0, k1)
        if o.flag_rotate:
            blk = np.rot90(blk, axes=(1, 2))
        blk = blk.astype(np.uint16)
        return blk * 256 if o.infiledatatype == 'uint8' else blk

    def __getitem__(self, idx):
        first = idx[0] if isinstance(idx, tuple) else idx
        rest = idx[1:] if isinstance(idx, tuple) else ()
        n = self.shape[0]
        if isinstance(first, slice):
            a, b, step = first.indices(n)
            if step == 1:
                blk = self._block(a, max(a, b))
                return blk[(slice(None),) + tuple(rest)] if rest else blk
        elif isinstance(first, (int, np.integer)):
            k = int(first) + (n if first < 0 else 0)
            fr = self._block(k, k + 1)[0]
            return fr[tuple(rest)] if rest else fr
        return np.asarray(self)[idx]

    def __array__(self, dtype=None, copy=None):
        a = self._block(0, self.shape[0])
        return a if dtype is None else a.astype(dtype)


class all_video_reader:
    """Whole scan resident in HBM behind the interface of the reference's in-RAM reader
    (/root/reference/video_reader.py:129-158; used by the spectral analyser,
    spectralAnalyserUI.py:155-175, 345-359: one compute_mean_return_fit, then read_video_improved again and
    again at new shifts after reset()).  The file crosses PCIe ONCE here; `frames` indexes like the reference's
    (N, ih, iw) uint16 array but copies only what is asked for to the host; `means` are the per-frame means.
    solex_util.compute_mean_return_fit / read_video_improved recognise `.stack` and work on it in place."""

    def __init__(self, file, buffer_size=25):
        from . import parallel
        from .engine import get_engine
        if parallel.world()[1] > 1:
            raise Exception('all_video_reader keeps the WHOLE scan on one GPU (the interactive tools\' reader); '
                            'under torchrun use video_reader, whose frames are sharded across the ranks')
        rdr = video_reader(file, buffer_size)
        self.file = self.path = file
        self.ih, self.iw = rdr.ih, rdr.iw
        self.Width, self.Height = rdr.Width, rdr.Height
        self.FrameCount = rdr.FrameCount
        self.count = rdr.count
        self.flag_rotate, self.infiledatatype = rdr.flag_rotate, rdr.infiledatatype
        self.FrameIndex = -1
        eng = get_engine()
        g = rdr.geometry
        if rdr.streamable:
            self.stack, _ = eng.ingest_file(rdr.path, g, rdr.payload_offset, rdr.frame_stride, accumulate=True)
        else:                                           # compressed AVI: host decode
            self.stack = eng.ingest_array(np.ascontiguousarray(rdr.raw_frames(0, int(rdr.FrameCount))))
        self.means = eng.frame_means(self.stack)
        self.frames = _DeviceFrames(self)

    @property
    def geometry(self):
        return self.stack.geom

    def has_frames(self):
        return self.FrameIndex + 1 < self.FrameCount

    def next_frame(self):
        self.FrameIndex += 1
        return self.frames[self.FrameIndex]

    def reset(self):
        self.FrameIndex = -1


class memory_scan:
    """A scan whose raw payload already sits in host memory (a pinned buffer, an
    mmap, an ndarray): same attributes as video_reader, consumed by the device
    path through the ingest ring (or copied straight from where it lies when
    the memory is pinned)."""

    def __init__(self, host_ptr, width, height, depth_bits, n_frames, name='memory', keepalive=None):
        self.file = self.path = name
        self.host_ptr = int(host_ptr)
        self.keepalive = keepalive
        self.SER_flag, self.AVI_flag = True, False
        self.Width, self.Height = np.uint32(width), np.uint32(height)
        self.PixelDepthPerPlane = np.uint32(depth_bits)
        self.FrameCount = np.uint32(n_frames)
        self.count = self.Width * self.Height
        self.infilebytes = 1 if depth_bits == 8 else 2
        self.infiledatatype = 'uint8' if depth_bits == 8 else 'uint16'
        self.FrameIndex = -1
        self.flag_rotate = bool(width > height)
        self.ih, self.iw = (int(width), int(height)) if self.flag_rotate else (int(height), int(width))

    @property
    def geometry(self):
        from .engine import ScanGeometry
        return ScanGeometry(int(self.Width), int(self.Height), int(self.infilebytes), int(self.FrameCount))

    def raw_frames(self, k0, k1):
        import ctypes
        g = self.geometry
        n = (k1 - k0) * g.frame_bytes
        buf = (ctypes.c_uint8 * n).from_address(self.host_ptr + k0 * g.frame_bytes)
        return np.frombuffer(buf, dtype=self.infiledatatype).reshape(k1 - k0, g.height, g.width)

    def next_frame(self):
        self.FrameIndex += 1
        img = self.raw_frames(self.FrameIndex, self.FrameIndex + 1)[0]
        if self.flag_rotate:
            img = np.rot90(img)
        if self.infiledatatype == 'uint8':
            img = np.asarray(img, dtype='uint16') * 256
        return img

    def has_frames(self):
        return self.FrameIndex + 1 < self.FrameCount


class device_scan(memory_scan):
    """A scan (or this rank's frame range of it) that is already resident in HBM
    as an engine.DeviceStack: the interactive / benchmark case where the stack is
    reused (the spectral analyser re-reconstructs the same scan at new shifts,
    /root/reference/spectralAnalyserUI.py:345-346)."""

    def __init__(self, stack, name='device'):
        g = stack.geom
        super().__init__(0, g.width, g.height, 8 if g.bytes_per_px == 1 else 16, g.n_frames, name=name)
        self.stack = stack

    def raw_frames(self, k0, k1):
        return self.stack.host_frames(k0 - self.stack.k0, k1 - self.stack.k0)
