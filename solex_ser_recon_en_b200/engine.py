"""Host side of the B200 reconstruction path: device buffers (torch tensors used
purely as memory), streams, and one method per stage of the reference's path,
each a thin call into libshg.so through ctypes.

Stage map (reference file:line -> method):
  video_reader.py:94-123 + solex_util.py:174-188   ingest_file / ingest_array / accumulate
  solex_util.py:188                                 finalize_mean_max
  solex_util.py:165-172,223-231,242                 detect_line
  solex_util.py:233-259                             fit_line
  solex_util.py:93-144                              recon
  ellipse_to_circle.py:94-145                       warp
  solex_util.py:383-516                             transversalium

There is no CPU fallback anywhere in this module: it needs libshg.so and a
CUDA device, and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import ShgError, call, lib


@dataclass(frozen=True)
class ScanGeometry:
    """Raw-file geometry of a scan (video_reader.py:31-91)."""
    width: int
    height: int
    bytes_per_px: int          # 1 (8-bit, scaled by 256 on use) or 2
    n_frames: int              # frames of the whole scan

    @property
    def rotated(self):
        return self.width > self.height

    @property
    def ih(self):
        return self.width if self.rotated else self.height

    @property
    def iw(self):
        return self.height if self.rotated else self.width

    @property
    def frame_px(self):
        return self.width * self.height

    @property
    def frame_bytes(self):
        return self.frame_px * self.bytes_per_px


class DeviceStack:
    """Frames [k0, k0+n) of a scan, resident in HBM in raw file layout, with
    the running integer sum / max of those frames."""

    def __init__(self, geom: ScanGeometry, k0: int, n: int, device):
        self.geom, self.k0, self.n = geom, int(k0), int(n)
        self.frames = torch.empty(max(1, self.n * geom.frame_bytes), dtype=torch.uint8, device=device)
        self.sum = torch.zeros(geom.frame_px, dtype=torch.int64, device=device)      # uint64 bit pattern
        self.max = torch.zeros(geom.frame_px, dtype=torch.int32, device=device)      # uint32 bit pattern
        self.accumulated = False

    def frame_ptr(self, k_local: int) -> int:
        return self.frames.data_ptr() + k_local * self.geom.frame_bytes

    def host_frames(self, k_local0: int, k_local1: int) -> np.ndarray:
        """Raw frames copied back to the host as (n, H, W) -- tests only."""
        g = self.geom
        a = self.frames[k_local0 * g.frame_bytes:k_local1 * g.frame_bytes].cpu().numpy()
        return a.view(np.uint8 if g.bytes_per_px == 1 else np.uint16).reshape(-1, g.height, g.width)


def _bind_to_gpu_numa_node(index: int):
    """Pin this process (one rank per GPU) to the CPUs NVML reports as local to
    the GPU, so that pinned buffers are first-touched on the GPU's own NUMA node
    and every rank's H2D stream stays on its own socket / PCIe root.  Returns
    the CPU list, or None when NVML is unavailable or SHG_NO_AFFINITY is set."""
    if os.environ.get('SHG_NO_AFFINITY') or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(visible.split(',')[index]) if visible and visible.split(',')[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


class _Stage:
    """One stage of the path: an NVTX range (always; free without a profiler attached) and, while
    eng.profile_stages is set, a pair of CUDA events on the current stream plus the host times at which the
    block was entered and left (stage timers / trace ranges)."""

    def __init__(self, eng, name):
        self.eng, self.name = eng, name

    def __enter__(self):
        torch.cuda.nvtx.range_push('shg:' + self.name)
        if getattr(self.eng, 'profile_stages', False):
            import time
            self.h0 = time.perf_counter()
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.eng.device))
        return self

    def __exit__(self, *exc):
        if getattr(self.eng, 'profile_stages', False) and hasattr(self, 'a'):
            import time
            b = torch.cuda.Event(enable_timing=True)
            b.record(torch.cuda.current_stream(self.eng.device))
            if not hasattr(self.eng, '_stage_events'):
                self.eng._stage_events = []
            self.eng._stage_events.append((self.name, self.a, b, self.h0, time.perf_counter()))
        torch.cuda.nvtx.range_pop()
        return False


class _EarlyDownload:
    def __init__(self, eng, t, ev):
        self.eng, self.t, self.ev = eng, t, ev

    def result(self) -> np.ndarray:
        up = self.eng._upload_stream
        up.wait_event(self.ev)
        with torch.cuda.stream(up):
            host = torch.empty(self.t.shape, dtype=self.t.dtype, pin_memory=True)
            host.copy_(self.t, non_blocking=True)
        self.t.record_stream(up)
        up.synchronize()
        return host.numpy()


class Engine:
    """One per process / GPU."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise ShgError('no CUDA device: the reconstruction path has no CPU fallback')
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', os.environ.get('SHG_DEVICE', '0')))
        self.index = int(device)
        self.device = torch.device('cuda', self.index)
        torch.cuda.set_device(self.device)
        self.cpu_affinity = _bind_to_gpu_numa_node(self.index)
        info = (C.c_int64 * 6)()
        call('shg_device_info', self.index, info)
        self.sm_count, self.cc = int(info[0]), (int(info[1]), int(info[2]))
        self.hbm_bytes, self.l2_bytes, self.smem_optin = int(info[3]), int(info[4]), int(info[5])
        self._logtab = None
        self._ingest = None
        self._ingest_cfg = None
        self.n_launches = 0            # kernels launched through this engine (bench's gpu_launches)
        self.ring_creates = 0          # pinned ingest rings allocated / reused without reallocation
        self.ring_reuses = 0
        self.ingest_log = []           # (kind, (seconds, seconds waiting for reads, bytes, chunks)) of every ingest

    # ------------------------------------------------------------------ util
    def stage(self, name):
        """Context manager: when self.stage_times is a dict, brackets the block
        with CUDA events on the current stream and accumulates its milliseconds
        (read after a synchronize through stage_report())."""
        return _Stage(self, name)

    def stage_report(self):
        torch.cuda.synchronize(self.device)
        out = {}
        for name, a, b, _, _ in getattr(self, '_stage_events', []):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        self._stage_events = []
        return out

    def stage_timeline(self):
        """[(name, gpu start_ms, gpu end_ms, host enter_ms, host exit_ms)] of the recorded stages relative to
        the first one (diagnostics: which stages overlap, where the device idles, where the host is late).
        Clears the record like stage_report()."""
        torch.cuda.synchronize(self.device)
        ev = getattr(self, '_stage_events', [])
        self._stage_events = []
        if not ev:
            return []
        base, h = ev[0][1], ev[0][3]
        return [(name, base.elapsed_time(a), base.elapsed_time(b), (h0 - h) * 1e3, (h1 - h) * 1e3)
                for name, a, b, h0, h1 in ev]

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def upload(self, arr) -> torch.Tensor:
        """Small host array -> device tensor without waiting for the work queued on the
        current stream.  torch.tensor(..., device=...) / .to(device) from pageable memory
        synchronise the stream they copy on; index lists and chord tables uploaded that
        way held the host until the kernel in front of them had finished, and every
        launch behind them started late.  Here the copy runs on a stream of its own."""
        if not hasattr(self, '_upload_stream'):
            self._upload_stream = torch.cuda.Stream(device=self.device)
        host = torch.from_numpy(np.ascontiguousarray(arr))
        with torch.cuda.stream(self._upload_stream):
            t = host.to(self.device)                    # synchronises the upload stream only
        t.record_stream(torch.cuda.current_stream(self.device))
        return t

    def sync(self):
        torch.cuda.synchronize(self.device)

    # ---------------------------------------------------------------- ingest
    def download_early(self, t: torch.Tensor) -> np.ndarray:
        """Host copy of a device tensor that is complete on the current stream NOW, without
        waiting for kernels queued after this call: the copy runs on the upload stream
        behind an event recorded here.  Call it right after the producer and before
        launching the next stage; the result is read with .result()."""
        if not hasattr(self, '_upload_stream'):
            self._upload_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return _EarlyDownload(self, t, ev)

    def pinned_alloc(self, nbytes: int) -> int:
        """Exact-size pinned host buffer (freed with pinned_free)."""
        p = C.c_void_p()
        call('shg_host_alloc', int(nbytes), C.byref(p))
        return int(p.value)

    @staticmethod
    def pinned_free(ptr: int):
        call('shg_host_free', int(ptr))

    def copy(self, dst_ptr: int, src_ptr: int, nbytes: int, kind: str):
        call('shg_memcpy_async', int(dst_ptr), int(src_ptr), int(nbytes), {'h2d': 1, 'd2h': 2, 'd2d': 3}[kind],
             self.stream)

    def _ring(self, slot_bytes, n_slots, n_threads):
        cfg = (int(slot_bytes), int(n_slots), int(n_threads))
        if self._ingest is not None and self._ingest_cfg == cfg:
            self.ring_reuses += 1                     # pinned slots reused by the next scan (batch mode)
            return self._ingest
        self.close_ring()
        self.ring_creates += 1
        h = C.c_void_p()
        call('shg_ingest_create', self.index, cfg[0], cfg[1], cfg[2], C.byref(h))
        self._ingest, self._ingest_cfg = h, cfg
        return h

    def close_ring(self):
        if self._ingest is not None:
            lib.shg_ingest_destroy(self._ingest)
            self._ingest = None

    def __del__(self):
        try:
            self.close_ring()
        except Exception:
            pass

    @staticmethod
    def _slot_bytes(geom, slot_mb):
        per = max(1, (slot_mb << 20) // geom.frame_bytes)
        return per * geom.frame_bytes

    def ingest_file(self, path: str, geom: ScanGeometry, payload_offset: int, frame_stride: int | None = None,
                    k0: int = 0, n: int | None = None, accumulate: bool = True, slot_mb: int | None = None,
                    n_slots: int = 6, n_threads: int | None = None, stack: DeviceStack | None = None):
        """File -> pinned ring -> HBM, mean/max accumulation overlapped.
        Returns (stack, stats) with stats = (seconds, seconds reading, bytes, chunks).
        Defaults (16 reader threads, 6 slots of 256 MB for big scans) measured 43 GB/s from a
        page-cached file on the B200 box (1 thread: 5.8, 8 threads: 31 GB/s)."""
        n = geom.n_frames - k0 if n is None else n
        if n_threads is None:
            n_threads = max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else 8))
        if slot_mb is None:
            slot_mb = 256 if n * geom.frame_bytes >= (4 << 30) else 32
        if stack is None:
            stack = DeviceStack(geom, k0, n, self.device)
        ring = self._ring(self._slot_bytes(geom, slot_mb), n_slots, n_threads)
        stats = (C.c_double * 4)()
        torch.cuda.current_stream(self.device).synchronize()        # zero-fill of sum/max must be done
        call('shg_ingest_file', ring, os.fsencode(path), int(payload_offset), geom.frame_bytes,
             int(frame_stride or geom.frame_bytes), int(k0), int(n), stack.frames.data_ptr(), geom.bytes_per_px,
             _ptr(stack.sum) if accumulate else 0, _ptr(stack.max) if accumulate else 0, stats)
        stack.accumulated = bool(accumulate)
        self.n_launches += int(stats[3]) if accumulate else 0
        self.ingest_log.append(('file', tuple(stats)))
        return stack, tuple(stats)

    def ingest_host(self, host_ptr: int, geom: ScanGeometry, n: int, k0: int = 0, frame_stride: int | None = None,
                    accumulate: bool = True, slot_mb: int = 256, n_slots: int = 4, n_threads: int = 8,
                    stack: DeviceStack | None = None):
        """Host memory image of the payload (mmap, pinned buffer, ndarray) -> HBM."""
        if stack is None:
            stack = DeviceStack(geom, k0, n, self.device)
        ring = self._ring(self._slot_bytes(geom, slot_mb), n_slots, n_threads)
        stats = (C.c_double * 4)()
        torch.cuda.current_stream(self.device).synchronize()
        call('shg_ingest_memory', ring, int(host_ptr), geom.frame_bytes, int(frame_stride or geom.frame_bytes),
             int(n), stack.frames.data_ptr(), geom.bytes_per_px,
             _ptr(stack.sum) if accumulate else 0, _ptr(stack.max) if accumulate else 0, stats)
        stack.accumulated = bool(accumulate)
        self.n_launches += int(stats[3]) if accumulate else 0
        self.ingest_log.append(('memory', tuple(stats)))
        return stack, tuple(stats)

    def ingest_array(self, frames: np.ndarray, n_total: int | None = None, k0: int = 0, accumulate: bool = True):
        """(N, H, W) uint8/uint16 ndarray -> stack."""
        frames = np.ascontiguousarray(frames)
        n, h, w = frames.shape
        geom = ScanGeometry(w, h, frames.dtype.itemsize, n if n_total is None else n_total)
        stack, _ = self.ingest_host(frames.ctypes.data, geom, n, k0=k0, accumulate=accumulate, slot_mb=16)
        return stack

    def synth_stack(self, geom: ScanGeometry, k0: int = 0, n: int | None = None, seed: int = 1) -> DeviceStack:
        n = geom.n_frames - k0 if n is None else n
        stack = DeviceStack(geom, k0, n, self.device)
        call('shg_synth_fill', stack.frames.data_ptr(), geom.bytes_per_px, k0, n, geom.n_frames, geom.width,
             geom.height, seed, self.stream)
        return stack

    # ------------------------------------------------------- pass 1: mean/max
    def accumulate(self, stack: DeviceStack, reset: bool = True):
        if reset:
            stack.sum.zero_()
            stack.max.zero_()
        g = stack.geom
        call('shg_accumulate', stack.frames.data_ptr(), g.bytes_per_px, stack.n, g.frame_px,
             stack.sum.data_ptr(), stack.max.data_ptr(), self.stream)
        self.n_launches += 1
        stack.accumulated = True

    def frame_means(self, stack: DeviceStack) -> np.ndarray:
        """np.mean of every oriented frame of the stack (all_video_reader.means, video_reader.py:143-147)."""
        g = stack.geom
        sums = self.empty((stack.n,), torch.int64)
        call('shg_frame_sums', stack.frames.data_ptr(), g.bytes_per_px, stack.n, g.frame_px, sums.data_ptr(), self.stream)
        self.n_launches += 1
        scale = 256 if g.bytes_per_px == 1 else 1
        return sums.cpu().numpy().astype(np.float64) * scale / g.frame_px

    def finalize_mean_max(self, sum_t, max_t, n_total: int, geom: ScanGeometry):
        mean_img = self.empty((geom.ih, geom.iw), torch.uint16)
        max_img = self.empty((geom.ih, geom.iw), torch.uint16)
        call('shg_finalize_mean_max', sum_t.data_ptr(), max_t.data_ptr(), int(n_total), geom.width, geom.height,
             1 if geom.bytes_per_px == 1 else 0, mean_img.data_ptr(), max_img.data_ptr(), self.stream)
        self.n_launches += 1
        return mean_img, max_img

    # ------------------------------------------------- line detection and fit
    def box_blur(self, img, kw: int, kh: int):
        rows, cols = img.shape
        out = self.empty((rows, cols), torch.uint16)
        tmp = self.empty((rows, cols), torch.int32)
        call('shg_box_blur_u16', img.data_ptr(), rows, cols, int(kw), int(kh), out.data_ptr(), tmp.data_ptr(),
             self.stream)
        self.n_launches += 2
        return out

    def detect_line(self, mean_img, max_img):
        """Slit extent from the max frame and the blurred / sharp per-row minima of
        the mean frame (solex_util.py:165-172, 223-231, 242)."""
        ih, iw = mean_img.shape
        blur5 = self.box_blur(max_img, 5, 5)
        sums = self.empty((ih,), torch.int64)
        call('shg_row_sums_u16', blur5.data_ptr(), ih, iw, sums.data_ptr(), self.stream)
        self.n_launches += 1
        ymean = sums.cpu().numpy().astype(np.float64) / iw                # np.mean(blur, axis=1): exact sums
        where_sun = ymean > np.median(ymean) / 5
        lb = int(np.argmax(where_sun))
        ub = int(ih - 1 - np.argmax(where_sun[::-1]))
        clip = int((ub - lb) * 0.05)
        y1 = min(ih - 1, lb + clip)
        y2 = max(0, ub - clip)
        bwy = int((y2 - y1) * 0.01)
        blur = self.box_blur(mean_img, 25, bwy)                           # raises for bwy < 1 like cv2.blur
        mi = self.empty((ih,), torch.int32)
        ms = self.empty((ih,), torch.int32)
        if iw - 13 <= 12:
            raise ShgError('frames are too narrow along the dispersion axis (%d px) for the 25 px blur window' % iw)
        call('shg_row_argmin_u16', blur.data_ptr(), ih, iw, 12, iw - 13, mi.data_ptr(), self.stream)
        call('shg_row_argmin_u16', mean_img.data_ptr(), ih, iw, 0, iw, ms.data_ptr(), self.stream)
        self.n_launches += 2
        return dict(y1=y1, y2=y2, min_intensity=mi, min_sharp=ms)

    def fit_line(self, det, ih: int):
        """The three cubic fits with their outlier logic (solex_util.py:233-259).
        Moment sums, residuals and masks on the device; the O(ih) mode-of-tenths
        bookkeeping calls NumPy exactly as the reference does."""
        y1, y2 = det['y1'], det['y2']
        n = y2 - y1
        if n < 4:
            raise ShgError('spectral line fit needs at least 4 slit rows, got %d (y1=%d, y2=%d)' % (n, y1, y2))
        mi, ms = det['min_intensity'][y1:y2], det['min_sharp'][y1:y2]
        packed = self.empty((12 + 4 * ih,), torch.float64)          # coefficients and fit table: one download
        coef = packed[:12].view(3, 4)
        resid = self.empty((n,), torch.float64)
        keep = self.empty((n,), torch.uint8)
        good = self.empty((n,), torch.uint8)
        st = self.stream
        call('shg_polyfit3', mi.data_ptr(), 0, y1, n, coef[0].data_ptr(), mi.data_ptr(), resid.data_ptr(), st)
        call('shg_sigma_mask', resid.data_ptr(), n, 3.0, keep.data_ptr(), st)
        call('shg_polyfit3', mi.data_ptr(), keep.data_ptr(), y1, n, coef[1].data_ptr(), ms.data_ptr(),
             resid.data_ptr(), st)
        self.n_launches += 3
        delta_sharp = resid.cpu().numpy()
        values, counts = np.unique(np.around(delta_sharp, 1), return_counts=True)
        ind = np.argpartition(-counts, kth=2)[:2]                         # ValueError for < 3 bins, as upstream
        shift = float(values[ind[0]])
        call('shg_window_mask', resid.data_ptr(), n, shift, 5.0, good.data_ptr(), st)
        call('shg_polyfit3', ms.data_ptr(), good.data_ptr(), y1, n, coef[2].data_ptr(), 0, 0, st)
        fit = packed[12:].view(ih, 4)
        call('shg_fit_table', coef[2].data_ptr(), ih, fit.data_ptr(), st)
        self.n_launches += 3
        packed_h = packed.cpu().numpy()
        coef_h = packed_h[:12].reshape(3, 4)
        return dict(p1=coef_h[0], p2=coef_h[1], p3=coef_h[2], shift=shift, fit=packed_h[12:].reshape(ih, 4),
                    keep=keep, mask_good=good)

    # ------------------------------------------------------- pass 2: recon
    def alloc_disk(self, n_shifts: int, n_frames: int, ih: int):
        return self.empty((n_shifts, n_frames, ih), torch.uint16)

    def recon(self, stack: DeviceStack, fit: np.ndarray, shifts, disk=None, k0_out: int | None = None,
              impl: int = 0, out_ptrs=None, mins=None):
        """disk[s, k, i] (frame-major) for the frames of `stack`; with `disk`
        given the rows land at frame offset k0_out (default: the stack's k0).
        `out_ptrs` (one device address per shift, possibly on peer GPUs) replaces
        `disk`: each shift's rows are written into that image at offset k0_out.
        `mins` (int32 device tensor, one per shift, pre-filled with 65535 or a partial
        minimum): the kernel folds the minimum of what it writes into it;
        self.recon_min_done tells whether the variant that ran does so."""
        g = stack.geom
        shifts = np.ascontiguousarray(shifts, dtype=np.int32)
        fit = np.ascontiguousarray(fit, dtype=np.float64)
        assert fit.shape == (g.ih, 4)
        assert mins is None or (mins.numel() == len(shifts) and mins.dtype == torch.int32 and mins.is_contiguous())
        done = C.c_int(0)
        wb = int(lib.shg_recon_workspace_bytes(g.ih, len(shifts)))
        if getattr(self, '_recon_ws', None) is None or self._recon_ws.numel() < wb:
            self._recon_ws = self.empty((wb,), torch.uint8)
        if out_ptrs is not None:
            ptrs = np.ascontiguousarray(out_ptrs, dtype=np.uint64)
            assert len(ptrs) == len(shifts)
            call('shg_recon', stack.frames.data_ptr(), g.bytes_per_px, stack.n, g.width, g.height,
                 fit.ctypes.data, shifts.ctypes.data, len(shifts), 0, 0, ptrs.ctypes.data,
                 int(stack.k0 if k0_out is None else k0_out), int(impl), self._recon_ws.data_ptr(), wb,
                 _ptr(mins), C.byref(done), self.stream)
            self.n_launches += 1
            self.recon_min_done = bool(done.value)
            return None
        if disk is None:
            disk = self.alloc_disk(len(shifts), stack.n, g.ih)
            k0_out = 0
        elif k0_out is None:
            k0_out = stack.k0
        assert disk.shape[0] == len(shifts) and disk.shape[2] == g.ih and disk.is_contiguous()
        call('shg_recon', stack.frames.data_ptr(), g.bytes_per_px, stack.n, g.width, g.height,
             fit.ctypes.data, shifts.ctypes.data, len(shifts), disk.data_ptr(), disk.stride(0), 0, int(k0_out),
             int(impl), self._recon_ws.data_ptr(), wb, _ptr(mins), C.byref(done), self.stream)
        self.n_launches += 1
        self.recon_min_done = bool(done.value)
        return disk

    def to_reference_layout(self, disk_s, flip: bool = False):
        """(N, ih) frame-major -> the reference's (ih, N) image; flip = np.flip(axis=1)."""
        n, ih = disk_s.shape
        out = self.empty((ih, n), torch.uint16)
        call('shg_transpose_u16', disk_s.data_ptr(), n, ih, out.data_ptr(), 1 if flip else 0, self.stream)
        self.n_launches += 1
        return out

    def minmax_device(self, imgs, sel=None):
        """Per-image (min, max) of a (S, ...) batch (or one 2-D image), left on the
        device as an int32 (n, 2) tensor: the warp reads its clip range from it."""
        if imgs.dim() == 2:
            imgs = imgs.unsqueeze(0)
        assert imgs[0].is_contiguous()                  # images may be slices of a larger buffer (stride(0) > numel)
        n_imgs = imgs.shape[0] if sel is None else len(sel)
        sel_t = None if sel is None else self.upload(np.asarray(list(sel), dtype=np.int32))
        mm = self.empty((n_imgs, 2), torch.int32)
        call('shg_minmax_u16', imgs.data_ptr(), imgs[0].numel(), imgs.stride(0), _ptr(sel_t), n_imgs,
             mm.data_ptr(), self.stream)
        self.n_launches += 2
        return mm

    def checksum(self, img) -> int:
        """Position-sensitive 64-bit checksum of a contiguous uint16 device tensor (shg_checksum_u16)."""
        assert img.is_contiguous() and img.dtype == torch.uint16
        out = self.empty((1,), torch.int64)
        call('shg_checksum_u16', img.data_ptr(), img.numel(), out.data_ptr(), self.stream)
        self.n_launches += 1
        return int(out.cpu().item()) & 0xFFFFFFFFFFFFFFFF

    def minmax(self, img):
        lo, hi = self.minmax_device(img)[0].cpu().tolist()
        return int(lo), int(hi)

    # -------------------------------------------------- circularisation warp
    def warp_batch(self, disks, sel, flip: bool, mat3: np.ndarray, out_shape, minmax_dev=None, out=None,
                   n_frames: int | None = None, frame_origin: int = 0, cvals=None, window=None, out_ptrs=None):
        """correct_image's pixel work (ellipse_to_circle.py:112-118) on frame-major
        disks (S, N, ih): images `sel` (None = all) -> (n, rows, cols) uint16, one launch.

        Frame-sharded scans (parallel.py, exchange after the warp): `disks` holds PHYSICAL frames
        [frame_origin, frame_origin + disks.shape[1]) of a scan of `n_frames`; only the output
        pixels whose left tap lies in the logical frames window = (own_lo, own_hi) are produced,
        `cvals` (int32 device tensor, one per image) stands in for each image's pixel [0][0], and
        `out_ptrs` (int64 device tensor, one address per image, possibly on peer GPUs) says where
        each image goes."""
        if disks.dim() == 2:
            disks = disks.unsqueeze(0)
        assert disks.is_contiguous() or disks.stride(1) == disks.shape[2]
        _, n, ih = disks.shape
        if not (mat3[1, 0] == 0 and mat3[1, 1] == 1 and mat3[1, 2] == 0 and
                mat3[2, 0] == 0 and mat3[2, 1] == 0 and mat3[2, 2] == 1):
            raise ShgError('warp matrix is not the row-preserving form get_correction_matrix produces')
        n_imgs = disks.shape[0] if sel is None else len(sel)
        if isinstance(sel, torch.Tensor):                   # already on the device (prepared ahead of time)
            sel_t = sel
        else:
            sel_t = None if sel is None else self.upload(np.asarray(list(sel), dtype=np.int32))
        if minmax_dev is None:
            assert n_frames is None, 'a frame-sharded warp needs the clip range of the whole images'
            minmax_dev = self.minmax_device(disks, sel)
        oh, ow = int(out_shape[0]), int(out_shape[1])
        sharded = n_frames is not None
        m00, m01, m02 = float(mat3[0, 0]), float(mat3[0, 1]), float(mat3[0, 2])
        if not sharded:
            if out is None:
                out = self.empty((n_imgs, oh, ow), torch.uint16)
            if lib.shg_warp_rows_tma_ok(disks.data_ptr(), disks.stride(0), ih, out.data_ptr(), out.stride(0), None):
                call('shg_warp_rows_tma', disks.data_ptr(), disks.stride(0), disks.shape[0], n, 0, _ptr(sel_t), n_imgs,
                     n, ih, 1 if flip else 0, m00, m01, m02, minmax_dev.data_ptr(), out.data_ptr(), out.stride(0),
                     oh, ow, None, -2 ** 31, 2 ** 31 - 1, None, self.stream)
            else:                          # odd geometries (ih not a multiple of 8, unaligned views): direct-load kernel
                call('shg_warp_rows', disks.data_ptr(), disks.stride(0), _ptr(sel_t), n_imgs, n, ih, 1 if flip else 0,
                     m00, m01, m02, minmax_dev.data_ptr(), out.data_ptr(), out.stride(0), oh, ow, self.stream)
            self.n_launches += 1
            return out
        assert cvals is not None and window is not None and (out is not None or out_ptrs is not None)
        if lib.shg_warp_rows_tma_ok(disks.data_ptr(), disks.stride(0), ih, _ptr(out), 0 if out is None else out.stride(0),
                                    None) and (out is not None or (oh * ow) % 8 == 0):
            # this rank's pixels into a local full-width image (16-byte stores), then one copy kernel moves every
            # row interval into the image's owner in 512-byte runs (peer stores of 16 or 64 bytes crawl over NVLink)
            local = out if out is not None else self.empty((n_imgs, oh, ow), torch.uint16)
            call('shg_warp_rows_tma', disks.data_ptr(), disks.stride(0), disks.shape[0], disks.shape[1],
                 int(frame_origin), _ptr(sel_t), n_imgs, int(n_frames), ih, 1 if flip else 0, m00, m01, m02,
                 minmax_dev.data_ptr(), local.data_ptr(), local.stride(0), oh, ow, cvals.data_ptr(), int(window[0]),
                 int(window[1]), None, self.stream)
            self.n_launches += 1
            if out_ptrs is not None:
                call('shg_exchange_rows', local.data_ptr(), local.stride(0), n_imgs, oh, ow, m00, m01, m02,
                     int(window[0]), int(window[1]), out_ptrs.data_ptr(), self.stream)
                self.n_launches += 1
                local.record_stream(torch.cuda.current_stream(self.device))
            return out
        base = disks.data_ptr() - int(frame_origin) * ih * 2          # frame 0 of the whole scan (never dereferenced there)
        call('shg_warp_rows_window', base, disks.stride(0), _ptr(sel_t), n_imgs, int(n_frames), ih, 1 if flip else 0,
             m00, m01, m02, minmax_dev.data_ptr(), _ptr(out),
             0 if out is None else out.stride(0), oh, ow, cvals.data_ptr(), int(window[0]), int(window[1]),
             _ptr(out_ptrs), self.stream)
        self.n_launches += 1
        return out

    def warp(self, disk_s, flip: bool, mat3: np.ndarray, out_shape, cval=None, lo=None, hi=None, out=None):
        """One image.  cval is image[0][0] and is read on the device; lo / hi
        default to the image's min / max (what skimage's warp clips to)."""
        mm = None
        if lo is not None and hi is not None:
            mm = self.upload(np.array([[int(lo), int(hi)]], dtype=np.int32))
        res = self.warp_batch(disk_s, None, flip, mat3, out_shape, mm,
                              None if out is None else out.unsqueeze(0))
        return res[0]

    def downscale4(self, disk_s, flip: bool):
        """4x4 block sums of the (ih, N) image (ellipse_to_circle.py:301)."""
        n, ih = disk_s.shape
        out = self.empty(((ih + 3) // 4, (n + 3) // 4), torch.int32)
        call('shg_downscale4_sum', disk_s.data_ptr(), n, ih, 1 if flip else 0, out.data_ptr(), out.shape[0],
             out.shape[1], self.stream)
        self.n_launches += 1
        return out

    # ------------------------------------------- limb detection (ellipse fit)
    def box_sum_u32(self, img, kw: int, kh: int):
        rows, cols = img.shape
        out = self.empty((rows, cols), torch.int32)
        tmp = self.empty((rows, cols), torch.int32)
        call('shg_box_sum_u32', img.data_ptr(), rows, cols, int(kw), int(kh), out.data_ptr(), tmp.data_ptr(),
             self.stream)
        self.n_launches += 2
        return out

    def sum_u32(self, img) -> int:
        out = self.empty((1,), torch.int64)
        call('shg_sum_u32', img.data_ptr(), img.numel(), out.data_ptr(), self.stream)
        self.n_launches += 1
        return int(out.cpu().item())

    def select_u32(self, vals, ranks):
        """Exact order statistics (0-based ranks) of a uint32 device array."""
        ranks = [int(r) for r in ranks]
        r = (C.c_int64 * len(ranks))(*ranks)
        out = (C.c_uint32 * len(ranks))()
        work = self.empty((2112,), torch.int32)
        call('shg_select_u32', vals.data_ptr(), vals.numel(), r, len(ranks), out, work.data_ptr(), self.stream)
        self.n_launches += 8
        return [int(v) for v in out]

    def blur_range(self, box, scale: float, ceiling: float):
        out = self.empty((2,), torch.int32)
        call('shg_blur_range', box.data_ptr(), box.numel(), float(scale), float(ceiling), out.data_ptr(), self.stream)
        self.n_launches += 1
        lo, hi = out.cpu().numpy().view(np.uint32).tolist()
        return int(lo), int(hi)

    def blur_hist(self, box, scale: float, ceiling: float, edges: np.ndarray):
        edges = np.ascontiguousarray(edges, dtype=np.float64)
        n_bins = len(edges) - 1
        counts = self.empty((32,), torch.int64)
        call('shg_blur_hist', box.data_ptr(), box.numel(), float(scale), float(ceiling),
             edges.ctypes.data_as(C.POINTER(C.c_double)), n_bins, counts.data_ptr(), self.stream)
        self.n_launches += 1
        return counts.cpu().numpy()[:n_bins]

    def canny_candidates(self, box, scale: float, level: float, weights: np.ndarray, low: float):
        """Thin-edge pixels of canny(flood image): (flat indices sorted ascending, magnitudes)."""
        rows, cols = box.shape
        n = rows * cols
        w = np.ascontiguousarray(weights, dtype=np.float64)
        buf = self.empty((6, n), torch.float64)          # smoothed, tmp0, tmp1, gi, gj, mag
        call('shg_flood_smooth', box.data_ptr(), rows, cols, float(scale), float(level),
             w.ctypes.data_as(C.POINTER(C.c_double)), len(w) - 1, float(np.finfo(np.float64).eps),
             buf[0].data_ptr(), buf[1].data_ptr(), self.stream)
        call('shg_sobel_mag', buf[0].data_ptr(), rows, cols, buf[3].data_ptr(), buf[4].data_ptr(), buf[5].data_ptr(),
             self.stream)
        self.n_launches += 6
        cap = 1 << 20
        while True:
            count = self.empty((1,), torch.int32)
            idx = self.empty((cap,), torch.int32)
            mag = self.empty((cap,), torch.float64)
            call('shg_nms_candidates', buf[3].data_ptr(), buf[4].data_ptr(), buf[5].data_ptr(), rows, cols,
                 float(low), count.data_ptr(), cap, idx.data_ptr(), mag.data_ptr(), self.stream)
            self.n_launches += 1
            m = int(count.cpu().item())
            if m <= cap:
                break
            cap = m
        idx_h = idx[:m].cpu().numpy().view(np.uint32).astype(np.int64)
        mag_h = mag[:m].cpu().numpy()
        order = np.argsort(idx_h, kind='stable')
        return idx_h[order], mag_h[order]

    def _limb_buffers(self, rows, cols):
        """Scratch of the limb search for one image shape, allocated once (the search is on the critical
        path of every multi-GPU step: no allocator calls, no pinned-memory registration per scan)."""
        key = (rows, cols)
        if getattr(self, '_limb_key', None) != key:
            n = rows * cols
            cap = 1 << 17
            self._limb = dict(
                u32=self.empty((3, rows, cols), torch.int32),            # box, box5, tmp
                f64=self.empty((6, n), torch.float64),                   # smoothed, tmp0, tmp1, gi, gj, mag
                state=self.empty((int(lib.shg_limb_state_bytes()),), torch.uint8),
                count=self.empty((1,), torch.int32), cap=cap,
                idx=self.empty((cap,), torch.int32), mag=self.empty((cap,), torch.float64),
                h_idx=torch.empty((cap,), dtype=torch.int32, pin_memory=True),
                h_mag=torch.empty((cap,), dtype=torch.float64, pin_memory=True),
                h_cnt=torch.empty((1,), dtype=torch.int32, pin_memory=True))
            self._limb_key = key
        return self._limb

    def limb_front(self, sums, bw: int, ranks4, gamma: float, n_bins: int = 20):
        """The threshold search of the limb detection as one queue of kernels and ONE read-back
        (shg_limb_front).  Returns (box, dict of exact host scalars)."""
        rows, cols = sums.shape
        L = self._limb_buffers(rows, cols)
        bufs = L['u32']
        r = (C.c_int64 * 4)(*[int(v) for v in ranks4])
        out = (C.c_double * 80)()
        call('shg_limb_front', sums.data_ptr(), rows, cols, int(bw), r, float(gamma), int(n_bins),
             bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), L['state'].data_ptr(), out, self.stream)
        self.n_launches += 17
        res = dict(total=int(out[0]), stats=[int(out[1 + q]) for q in range(4)], ceiling=float(out[5]),
                   range=(int(out[6]), int(out[7])), edges=np.array(out[8:9 + n_bins], dtype=np.float64),
                   counts=np.array(out[41:41 + n_bins], dtype=np.float64).astype(np.int64))
        return bufs[0], res

    def limb_canny(self, box, scale: float, level: float, weights: np.ndarray, low: float):
        """canny_candidates() with one round trip: (flat indices ascending, magnitudes)."""
        rows, cols = box.shape
        w = np.ascontiguousarray(weights, dtype=np.float64)
        L = self._limb_buffers(rows, cols)
        while True:
            cap = L['cap']
            call('shg_limb_canny', box.data_ptr(), rows, cols, float(scale), float(level),
                 w.ctypes.data_as(C.POINTER(C.c_double)), len(w) - 1, float(np.finfo(np.float64).eps), float(low),
                 L['f64'].data_ptr(), L['count'].data_ptr(), cap, L['idx'].data_ptr(), L['mag'].data_ptr(),
                 min(cap, 1 << 15), L['h_cnt'].data_ptr(), L['h_idx'].data_ptr(), L['h_mag'].data_ptr(), self.stream)
            self.n_launches += 7
            m = int(L['h_cnt'].numpy().view(np.uint32)[0])
            if m <= cap:
                break
            cap = 1 << int(m - 1).bit_length()                             # rare: a very long edge list
            L.update(cap=cap, idx=self.empty((cap,), torch.int32), mag=self.empty((cap,), torch.float64),
                     h_idx=torch.empty((cap,), dtype=torch.int32, pin_memory=True),
                     h_mag=torch.empty((cap,), dtype=torch.float64, pin_memory=True))
        idx_h = L['h_idx'].numpy()[:m].view(np.uint32).astype(np.int64)
        order = np.argsort(idx_h)                                          # pixel indices are distinct
        return idx_h[order], L['h_mag'].numpy()[:m][order]

    # ------------------------------------------------------- transversalium
    @property
    def logtab(self):
        """L(v) of the row statistics tabulated for v in [0, 65536) (tests / diagnostics only)."""
        if self._logtab is None:
            self._logtab = self.empty((65536,), torch.float64)
            call('shg_log_table', self._logtab.data_ptr(), self.stream)
        return self._logtab

    @staticmethod
    def transversalium_chords(circle, borders):
        """Row range and per-row chord of correct_transversalium2 (solex_util.py:384-391)."""
        cx, cy, rad = (float(v) for v in circle)
        b0, b1, b2, b3 = (float(v) for v in borders)
        y1 = math.ceil(max(cy - rad, b1))
        y2 = math.floor(min(cy + rad, b3))
        rows = np.arange(y1 + 1, max(y2, y1 + 1), dtype=np.int64)
        # dx = floor((r^2 - (y - cy)^2) ** 0.5): vectorised with sqrt (the per-row Python loop cost 2-5 ms of host
        # time, exposed whenever the GPU waits for the geometry); pow(x, 0.5) and sqrt(x) can differ in the last
        # bit, which matters only within an ulp of an integer: those rows (there are normally none) are redone
        # with the reference's own expression
        arg = rad ** 2 - (rows.astype(np.float64) - cy) ** 2
        with np.errstate(invalid='ignore'):
            root = np.sqrt(arg)
        near = ~(np.abs(root - np.rint(root)) > 1e-9 * np.maximum(1.0, np.abs(root)))     # also catches nan
        for i in np.flatnonzero(near):
            root[i] = (rad ** 2 - (float(rows[i]) - cy) ** 2) ** 0.5        # raises / complex like the reference
        dx = np.floor(root)
        xa = np.ceil(np.maximum(cx - dx, b0))
        xb = np.floor(np.minimum(cx + dx, b2))
        return y1, y2, rows.astype(np.int32), xa.astype(np.int32), xb.astype(np.int32)

    def transversalium_row_stats(self, imgs, rows, xa, xb, device: bool = False):
        """Per-row robust mean of log(img[y]/img[y-1]) over [xa, xb) (solex_util.py:392-395)
        for one (h, w) image -> (n,), or a (S, h, w) batch sharing the chords -> (S, n).
        One launch, one device -> host copy."""
        single = imgs.dim() == 2
        if single:
            imgs = imgs.unsqueeze(0)
        n = len(rows)
        if n == 0:
            return np.zeros(0) if single else np.zeros((imgs.shape[0], 0))
        n_imgs, h, w = imgs.shape
        assert imgs.stride(1) == w and imgs.stride(2) == 1
        if rows.min() < 1 or rows.max() >= h or xa.min() < 0 or xb.max() > w:
            raise IndexError('transversalium chord outside the image')     # the reference would raise / wrap too
        idx = self.upload(np.stack([rows, xa, xb]).astype(np.int32))
        out = self.empty((n_imgs, n), torch.float64)
        max_len = int(max(0, (xb - xa).max()))
        wb = int(lib.shg_transv_workspace_bytes(n, max_len, n_imgs))
        work = self.empty((wb,), torch.uint8) if wb > 0 else None
        call('shg_transv_row_stats', imgs.data_ptr(), h, w, n_imgs, imgs.stride(0), idx[0].data_ptr(),
             idx[1].data_ptr(), idx[2].data_ptr(), n, max_len, out.data_ptr(),
             _ptr(work), wb, self.stream)
        # the register-resident kernel, then the classic kernel over the rows it handed back (csrc/transv.cu)
        classic_only = os.environ.get('SHG_TRANSV_HIST', '1') != '1' or os.environ.get('SHG_TRANSV_REG', '1') == '0'
        self.n_launches += 1 if classic_only else 2
        self._transv_work = work           # first word: how many rows went to the classic kernel (diagnostics)
        if device:
            return out
        res = out.cpu().numpy()
        return res[0] if single else res

    def transversalium_gains(self, stats_dev, y1: int, y2: int, n_rows: int, strength: int):
        """(S, n_rows) device tensor of per-row gains from the (S, n-1) device row
        statistics (solex_util.py:400-404, 456-479)."""
        from .solex_util import savgol_window, tukey_taper
        n = y2 - y1
        window = savgol_window(n, strength)
        if window < 5:                     # scipy.signal.savgol_filter: polyorder (3) must be less than window_length
            raise ValueError('polyorder must be less than window_length (only %d disk rows)' % n)
        from scipy.signal import savgol_coeffs
        key = (n, window)
        if getattr(self, '_gain_tab_key', None) != key:
            tab = np.concatenate([savgol_coeffs(window, 3), tukey_taper(n)])
            self._gain_tab = self.upload(tab)
            self._gain_tab_key = key
        n_imgs = stats_dev.shape[0]
        assert stats_dev.shape[1] == n - 1 and stats_dev.is_contiguous()
        gains = self.empty((n_imgs, n_rows), torch.float64)
        call('shg_transv_gain', stats_dev.data_ptr(), n_imgs, n, window, self._gain_tab.data_ptr(),
             self._gain_tab[window:].data_ptr(), int(y1), int(n_rows), gains.data_ptr(), self.stream)
        self.n_launches += 1
        return gains

    def row_scale(self, imgs, gain, out=None):
        """(img.T * c).T, clip 65535, truncate (solex_util.py:489,515-516) for one
        image with gain (h,), or a (S, h, w) batch with gains (S, h)."""
        single = imgs.dim() == 2
        if single:
            imgs = imgs.unsqueeze(0)
        n_imgs, h, w = imgs.shape
        assert imgs.stride(1) == w and imgs.stride(2) == 1
        g = gain if isinstance(gain, torch.Tensor) else \
            self.upload(np.ascontiguousarray(gain, dtype=np.float64))
        assert g.numel() == n_imgs * h
        if out is None:
            out = self.empty((n_imgs, h, w), torch.uint16)
        elif out.dim() == 2:
            out = out.unsqueeze(0)
        assert out.stride(0) == imgs.stride(0) or n_imgs == 1
        call('shg_row_scale_u16', imgs.data_ptr(), h, w, n_imgs, imgs.stride(0), g.data_ptr(), out.data_ptr(),
             self.stream)
        self.n_launches += 1
        return out[0] if single else out


_engines = {}


def get_engine(device: int | None = None) -> Engine:
    """Process-wide engine for a device (created on first use; never in a forked child)."""
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', os.environ.get('SHG_DEVICE', '0')))
    key = (os.getpid(), int(device))
    if key not in _engines:
        _engines[key] = Engine(device)
    return _engines[key]
