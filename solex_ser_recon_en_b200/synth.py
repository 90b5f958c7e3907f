"""Deterministic synthetic spectroheliograph scans (SER / AVI) for tests and benches.

There is no sample data in the reference tree, so every
parity pin is produced from scans made here.  The recipe follows SURVEY.md
section 8(d): every raw frame is a narrow image of a spectrum (``H`` rows of
dispersion, ``W`` columns of slit) with one dark absorption line whose centre
bends quadratically along the slit, multiplied by the brightness of the solar
disk at the slit position for that frame, plus a pedestal and Gaussian noise.
Noise is required: the reference divides by ``std(delta)`` and needs at least
three distinct residual bins (/root/reference/solex_util.py:236-246).

The file formats are the ones the reference reader parses
(/root/reference/video_reader.py:30-80): a 178-byte SER header followed by the
little-endian payload, or an uncompressed 8-bit Y800 AVI.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field

import numpy as np

SER_HEADER_BYTES = 178


@dataclass
class ScanSpec:
    """Geometry and line parameters of one synthetic scan."""
    n_frames: int
    width: int            # raw frame width  (slit axis when width > height)
    height: int           # raw frame height (dispersion axis when width > height)
    depth_bits: int = 16  # 8 or 16
    sigma: float = 3.0    # line width (px)
    line_depth: float = 0.75
    amplitude: float = 30000.0
    pedestal: float = 300.0
    noise: float = 50.0
    seed: int = 1
    dust: tuple = ()      # slit positions with a 3 % dust shadow (transversalium)
    chunk: int = 50       # frames per rng stream
    ellipse_x: float = 0.42   # disk semi-axis along the scan, fraction of n_frames
    ellipse_y: float = 0.40   # disk semi-axis along the slit, fraction of width
    extra: dict = field(default_factory=dict)

    @property
    def dtype(self):
        return np.uint8 if self.depth_bits == 8 else np.uint16

    @property
    def frame_bytes(self):
        return self.width * self.height * (1 if self.depth_bits == 8 else 2)


def halpha(n_frames, width, height, seed=1, **kw):
    return ScanSpec(n_frames, width, height, 16, 3.0, 0.75, 30000.0, 300.0, 50.0, seed, **kw)


def ca_k_8bit(n_frames, width, height, seed=2, **kw):
    return ScanSpec(n_frames, width, height, 8, 8.0, 0.9, 110.0, 4.0, 1.2, seed, **kw)


def line_centre(spec: ScanSpec):
    """Line centre (dispersion coordinate) for every slit position."""
    if spec.width > spec.height:
        n_slit, n_disp = spec.width, spec.height
    else:
        n_slit, n_disp = spec.height, spec.width
    u = (np.arange(n_slit) - n_slit / 2) / (n_slit / 2)
    return n_disp / 2 + 6.0 * u * u + 1.5 * u


def frames(spec: ScanSpec, k0: int, k1: int) -> np.ndarray:
    """Raw frames ``k0 <= k < k1`` as an array ``(k1-k0, height, width)``.

    The rng stream is keyed on ``(seed, chunk index)`` so that any frame range
    can be produced independently (used to write shards per rank).
    """
    W, H, N = spec.width, spec.height, spec.n_frames
    rotated = W > H
    n_slit, n_disp = (W, H) if rotated else (H, W)
    slit = np.arange(n_slit, dtype=np.float64)
    disp = np.arange(n_disp, dtype=np.float64)
    centre = line_centre(spec)
    # profile[disp, slit]
    profile = 1.0 - spec.line_depth * np.exp(-0.5 * ((disp[:, None] - centre[None, :]) / spec.sigma) ** 2)
    dust = np.ones(n_slit)
    for x in spec.dust:
        dust[int(x)] = 0.97
    out = np.empty((k1 - k0, H, W), dtype=spec.dtype)
    top = 255.0 if spec.depth_bits == 8 else 65535.0
    k = k0
    while k < k1:
        c = k // spec.chunk
        ke = min(k1, (c + 1) * spec.chunk)
        rng = np.random.default_rng([spec.seed, c])
        noise_all = rng.normal(0.0, spec.noise, size=(spec.chunk, n_disp, n_slit))
        for kk in range(k, ke):
            rho2 = ((kk - N / 2) / (spec.ellipse_x * N)) ** 2 + ((slit - n_slit / 2) / (spec.ellipse_y * n_slit)) ** 2
            disk = np.where(rho2 < 1.0, np.sqrt(np.maximum(1.0 - 0.6 * rho2, 0.0)), 0.02)
            tex = 1.0 + 0.05 * np.sin(0.37 * kk) * np.cos(0.11 * slit)
            sig = spec.amplitude * (disk * tex * dust)[None, :] * profile + spec.pedestal
            sig = sig + noise_all[kk - c * spec.chunk]
            sig = np.clip(np.rint(sig), 0, top)
            if rotated:
                out[kk - k0] = sig.astype(spec.dtype)           # (H, W) = (disp, slit)
            else:
                out[kk - k0] = sig.T.astype(spec.dtype)         # (H, W) = (slit, disp)
        k = ke
    return out


def ser_header(width: int, height: int, depth_bits: int, n_frames: int) -> bytes:
    """178-byte SER header with the fields the reference reads
    (/root/reference/video_reader.py:31-54: LuID@14 ... FrameCount@38)."""
    hdr = bytearray(SER_HEADER_BYTES)
    hdr[0:14] = b'LUCAM-RECORDER'
    struct.pack_into('<7I', hdr, 14, 0, 0, 1, width, height, depth_bits, n_frames)
    return bytes(hdr)


def write_ser(path: str, spec: ScanSpec, k0: int = 0, k1: int | None = None, batch: int = 100) -> str:
    """Write frames ``[k0, k1)`` of the scan as a SER file (FrameCount = k1-k0)."""
    k1 = spec.n_frames if k1 is None else k1
    with open(path, 'wb') as f:
        f.write(ser_header(spec.width, spec.height, spec.depth_bits, k1 - k0))
        for a in range(k0, k1, batch):
            b = min(k1, a + batch)
            f.write(frames(spec, a, b).astype('<u2' if spec.depth_bits != 8 else 'u1').tobytes())
    return path


def write_ser_from_array(path: str, stack: np.ndarray) -> str:
    """Write an ``(N, H, W)`` uint8/uint16 array as a SER file."""
    n, h, w = stack.shape
    bits = 8 if stack.dtype == np.uint8 else 16
    with open(path, 'wb') as f:
        f.write(ser_header(w, h, bits, n))
        f.write(np.ascontiguousarray(stack).tobytes())
    return path


def write_avi(path: str, spec: ScanSpec, batch: int = 100) -> str:
    """Uncompressed 8-bit grey AVI (fourcc 0 => Y800), which round-trips
    losslessly through cv2.VideoCapture, the reference's AVI reader
    (/root/reference/video_reader.py:68-80,111-113)."""
    import cv2
    assert spec.depth_bits == 8
    vw = cv2.VideoWriter(path, 0, 25.0, (spec.width, spec.height), isColor=False)
    if not vw.isOpened():
        raise RuntimeError('cannot open AVI writer for ' + path)
    for a in range(0, spec.n_frames, batch):
        b = min(spec.n_frames, a + batch)
        for fr in frames(spec, a, b):
            vw.write(fr)
    vw.release()
    return path


# The five BASELINE.json configurations (geometry only; the big ones are never
# materialised on the host -- see bench.py which synthesises them on the device).
CONFIGS = {
    1: dict(spec=lambda: halpha(1000, 1280, 200, seed=1, dust=(400, 401, 700)), flags=''),
    2: dict(spec=lambda: ca_k_8bit(2000, 1920, 256, seed=2), flags='-ms'),
    3: dict(spec=lambda: halpha(4000, 2048, 300, seed=3), flags='-w-10:10:1'),
    4: dict(spec=lambda: halpha(3000, 2048, 256, seed=4), flags='', files=16),
    5: dict(spec=lambda: halpha(20000, 4096, 512, seed=5), flags='-w-50:50:1'),
}


def shifts_from_flag(flag: str):
    """Shift list for a ``-w`` flag spec the way the reference CLI parses it
    (/root/reference/CLI_handler.py:50-73) -- used only to label configs."""
    if not flag.startswith('-w'):
        return [0]
    s = flag[2:].split(':')
    if len(s) == 1:
        return [int(x) for x in s[0].split(',')]
    if len(s) == 2:
        return list(range(int(s[0]), int(s[1]) + 1))
    return list(range(int(s[0]), int(s[1]) + 1, int(s[2])))
