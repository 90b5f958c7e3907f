#!/usr/bin/env python
"""Benchmark of the frame-stack reconstruction hot path (BASELINE.json metric:
frames/s of a full reconstruction, with HBM / PCIe roofline fractions).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the oracle port on host cores

Workload (config.workload): BASELINE.json configs[4] -- a synthetic 16-bit scan
of 20 000 frames x 4096x512, H-alpha line, -w-50:50:1 (101 shift images), ellipse
fit + circularisation + transversalium on.  It fits one B200 (84 GB stack +
16.5 GB of disk images), so it is the N=1 workload; at N>1 the same 20 000
frames are sharded by frame range across the ranks (strong scaling).

One step = one pass of the whole path over the scan:
  value : stack already resident in HBM (pass 1 re-reads it) -> mean/max -> all-reduce -> line detection + fit
          -> reconstruction at 101 shifts, each rank writing its frame rows into the owner rank's image over NVLink
          -> ellipse fit (rank 0, broadcast) -> warp + transversalium of the shifts each rank owns.
  e2e   : K scans back to back through the public batch entry point (Solex_recon.solex_do_work_readers, what the CLI
          runs for a list of files), each starting from the payload in pinned HOST memory (H2D inside the timed
          region, overlapped with pass 1) and ending with the final images copied back to pinned host memory (D2H
          inside the timed region); scan i+1's ingest overlaps the tail of scan i, as in the reference's batch mode.
Timed with CUDA events bracketed by barrier + synchronize, max over ranks.  The
84 GB input is far larger than the 126 MB L2, so no explicit flush is needed.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = dict(frames=20000, width=4096, height=512, bits=16, shift_lo=-50, shift_hi=50)
METRIC = 'frames/sec, full reconstruction (mean + line fit + 101-shift recon + circularise + transversalium)'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    # development overrides (the driver never passes these; a run with them is labelled in config)
    ap.add_argument('--frames', type=int, default=WORKLOAD['frames'])
    ap.add_argument('--width', type=int, default=WORKLOAD['width'])
    ap.add_argument('--height', type=int, default=WORKLOAD['height'])
    ap.add_argument('--sample-frames', type=int, default=0, help='CPU arm: frames per step (0 = auto)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--timeline', action='store_true', help='print the stage timeline of one extra resident pass (stderr)')
    ap.add_argument('--e2e-steps', type=int, default=0, help='0 = same as --steps')
    ap.add_argument('--no-configs', action='store_true', help='skip the small-configuration lines (config_lines)')
    ap.add_argument('--record-crc', action='store_true',
                    help='store this run\'s outputs_crc in profiles/outputs_crc.json as the expected value (run at N=1)')
    return ap.parse_args()


def workload_config(a, extra=None):
    shifts = list(range(WORKLOAD['shift_lo'], WORKLOAD['shift_hi'] + 1))
    cfg = {
        'workload': 'BASELINE configs[4]: synthetic 16-bit SER, %d frames x %dx%d, H-alpha, -w%d:%d:1 (%d shifts), '
                    'ellipse fit + circularisation + transversalium' % (a.frames, a.width, a.height, shifts[0],
                                                                        shifts[-1], len(shifts)),
        'frames': a.frames, 'width': a.width, 'height': a.height, 'bits': 16, 'n_shifts': len(shifts),
        'stack_bytes': a.frames * a.width * a.height * 2,
        'l2_policy': 'inputs (84 GB stack) far larger than the 126 MB L2; no flush needed',
    }
    if (a.frames, a.width, a.height) != (WORKLOAD['frames'], WORKLOAD['width'], WORKLOAD['height']):
        cfg['reduced'] = True
    if extra:
        cfg.update(extra)
    return cfg, shifts


def default_options(shifts):
    """The reference's default options (SHG_MAIN.py:41-68) for a -c style run with no files written."""
    from solex_ser_recon_en_b200 import SHG_MAIN
    o = dict(SHG_MAIN.options)
    o.update(shift=list(shifts), clahe_only=True, _nolog=True)
    return o


# =============================================================== clocks sampler
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, smax, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.3:
                continue
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax or None,
                'reasons': sorted(reasons), 'samples': len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ================================================================== CPU arm
def _cpu_worker(args):
    """One host process: frames [k0, k1) of the sample.  phase 'sum' -> partial sum/max; phase 'recon' -> disks."""
    from oracle import shg_oracle as O
    phase, frames, fit, shifts = args
    if phase == 'sum':
        return O.raw_sum_max(frames)
    return O.recon(frames, fit, shifts)


def cpu_reference_step(sample, shifts, pool, n_workers):
    """The reference's solex_read work (mean frame, line detection + fit, reconstruction at every shift) on a
    frame sample, restated by the oracle and spread over the host cores by frame range."""
    from oracle import shg_oracle as O
    n = sample.shape[0]
    cuts = [n * i // n_workers for i in range(n_workers + 1)]
    parts = [sample[cuts[i]:cuts[i + 1]] for i in range(n_workers) if cuts[i + 1] > cuts[i]]
    t0 = time.perf_counter()
    res = list(pool.map(_cpu_worker, [('sum', p, None, None) for p in parts]))
    s = sum(r[0] for r in res)
    m = res[0][1]
    for r in res[1:]:
        m = np.maximum(m, r[1])
    mean_img, max_img = O.finalize_mean_max(s, m, n, False)
    y1, y2 = O.slit_extent(max_img)
    mi, ms = O.line_minima(mean_img, y1, y2)
    lf = O.line_fit(mi, ms, y1, y2, mean_img.shape[0])
    all_shifts = O.shift_list(shifts)
    list(pool.map(_cpu_worker, [('recon', p, lf['fit'], all_shifts) for p in parts]))
    return time.perf_counter() - t0


def make_cpu_sample(a, n_sample):
    """Frames from the middle of the scan (so the Sun is in the slit), same recipe family as the device synth."""
    from solex_ser_recon_en_b200 import synth
    spec = synth.halpha(a.frames, a.width, a.height, seed=5)
    spec.chunk = 8
    k0 = (a.frames - n_sample) // 2
    return synth.frames(spec, k0, k0 + n_sample)


def synthetic_disk_image(a, seed=5):
    """(ih, N) uint16 image with the statistics of a reconstructed config-5 disk (elliptical Sun, limb
    darkening, texture, dust rows, pedestal, noise): the input of the post-processing half of the CPU arm."""
    ih, n = a.width, a.frames
    rng = np.random.default_rng(seed)
    img = np.empty((ih, n), dtype=np.uint16)
    k = np.arange(n, dtype=np.float64)
    ck = ((k - n / 2) / (0.42 * n)) ** 2
    dust = np.ones(ih)
    dust[rng.integers(int(0.2 * ih), int(0.8 * ih), 12)] = 0.97
    for r0 in range(0, ih, 256):
        r = np.arange(r0, min(ih, r0 + 256), dtype=np.float64)
        rho2 = ck[None, :] + (((r - ih / 2) / (0.40 * ih)) ** 2)[:, None]
        disk = np.where(rho2 < 1.0, np.sqrt(np.maximum(1.0 - 0.6 * rho2, 0.0)), 0.02)
        tex = 1.0 + 0.05 * np.sin(0.37 * k)[None, :] * np.cos(0.11 * r)[:, None]
        sig = 7500.0 * disk * tex * dust[r0:r0 + len(r), None] + 300.0 + rng.normal(0.0, 50.0, size=rho2.shape)
        img[r0:r0 + len(r)] = np.clip(np.rint(sig), 0, 65535).astype(np.uint16)
    return img


def run_unmodified_reference(a):
    """CPU arm, kind "reference": the UNMODIFIED reference modules (oracle/_ref or /root/reference, loaded by
    oracle/ref_shim) in ONE process with NumPy / OpenCV default threading, as the reference itself runs.

    A full config-5 pass takes the reference ~20 min, so every step times a bounded sample of the same work
    through the reference's own entry points and extrapolates (SURVEY 8d / BASELINE.md section 3):
      read half    Solex_recon.solex_read on an n-frame excerpt of the scan (4096x512 frames, all 101 shifts):
                   it scales with the frame count -> t_read * N / n;
      process half Solex_recon.solex_process on one full-size (4096 x 20000) disk image: the ellipse fit runs
                   once per scan (t_fit), circularisation + transversalium once per requested shift (t_image)
                   -> t_fit + 101 * t_image.
    image_process (CLAHE / PNG tail) is outside the metric and replaced by a no-op on this arm."""
    import shutil
    import tempfile

    import cv2
    from oracle import ref_shim
    from solex_ser_recon_en_b200 import synth
    ref = ref_shim.load()
    cfg, shifts = workload_config(a)
    n_shifts = len(shifts)
    work = tempfile.mkdtemp(prefix='shg_ref_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    spec = synth.halpha(a.frames, a.width, a.height, seed=5)
    spec.chunk = 8
    disk = synthetic_disk_image(a)
    timers = {}

    def timed(fn, key):
        def wrapper(*args, **kw):
            t = time.perf_counter()
            try:
                return fn(*args, **kw)
            finally:
                timers[key] = timers.get(key, 0.0) + time.perf_counter() - t
        return wrapper

    real_e2c, real_ip = ref.Solex_recon.ellipse_to_circle, ref.Solex_recon.image_process
    ref.Solex_recon.ellipse_to_circle = timed(real_e2c, 'fit')
    ref.Solex_recon.image_process = lambda frame, cercle, options, header, basefich: (None, None)
    state = {'n': 0, 'path': None}

    def sample_file(n):
        if state['n'] != n:
            k0 = (a.frames - n) // 2
            state['path'] = synth.write_ser_from_array(os.path.join(work, 'sample.SER'), synth.frames(spec, k0, k0 + n))
            state['n'] = n
        return state['path']

    def one_step(n):
        timers.clear()
        opt = ref_shim.default_options(shift=list(shifts), output_dir=work)
        t0 = time.perf_counter()
        disk_list, bounds, hdr = ref.Solex_recon.solex_read(sample_file(n), opt)
        t_read = time.perf_counter() - t0
        del disk_list
        # post-processing of ONE requested shift on a full-size image (index 0 = the ellipse-fit shift, index 1 =
        # shift 0, exactly the list solex_read builds: Solex_recon.py:55)
        opt2 = ref_shim.default_options(shift=[0], output_dir=work)
        opt2.update(basefich0=os.path.join(work, 'sample'), shift_requested=[0], shift=[10, 0])
        t1 = time.perf_counter()
        ref.Solex_recon.solex_process(opt2, [disk, disk], (int(0.1 * a.width), int(0.9 * a.width)), hdr)
        t_proc = time.perf_counter() - t1
        t_fit = timers.get('fit', 0.0)
        return t_read, t_fit, t_proc - t_fit

    try:
        budget = max(5.0, min(14.0, 270.0 / max(1, a.steps + a.warmup)))
        t_read, t_fit, t_img = one_step(16)                        # calibration (untimed)
        n = a.sample_frames or int(max(8, min(512, 16 * max(1.0, budget - t_fit - t_img) / max(t_read, 1e-3))))
        wall, full = [], []
        for it in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            t_read, t_fit, t_img = one_step(n)
            if it >= a.warmup:
                wall.append(time.perf_counter() - t0)
                full.append(a.frames / n * t_read + t_fit + n_shifts * t_img)
    finally:
        ref.Solex_recon.ellipse_to_circle, ref.Solex_recon.image_process = real_e2c, real_ip
        shutil.rmtree(work, ignore_errors=True)
    t_full = float(np.mean(full))
    value = a.frames / t_full
    sample_desc = ('UNMODIFIED reference modules (%s), one process: per step solex_read on a %d-frame excerpt (%dx%d, %d '
                   'shifts; last step %.2f s) + solex_process of one full-size %dx%d disk image (ellipse fit %.2f s, '
                   'circularise + transversalium %.2f s per image); full scan = %d/%d x read + fit + %d x image = %.0f s; '
                   'image_process (CLAHE / PNG, outside the metric) is a no-op; scikit-image / lsq-ellipse are not '
                   'installable here: oracle/thirdparty.py NumPy restatements stand in for warp / canny / LsqEllipse' %
                   ('oracle/_ref archive' if ref_shim.source() == 'staged' else '/root/reference', n, a.width, a.height,
                    n_shifts, t_read, a.width, a.frames, t_fit, t_img, a.frames, n, n_shifts, t_full))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': 1e3 * float(np.mean(wall)), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'u16+f64', 'data': 'synthetic', 'config': cfg,
        'extrapolated_full_scan_s': t_full,
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': 1, 'kind': 'reference', 'sample': sample_desc,
                         'host_cpus': os.cpu_count(), 'cv2_threads': cv2.getNumThreads()},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def run_reference_arm(a):
    import concurrent.futures as cf
    import multiprocessing as mp
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from oracle import ref_shim
    if ref_shim.available() and not os.environ.get('SHG_REF_PORT'):
        return run_unmodified_reference(a)
    # fallback (kind "port"): the oracle's restatement of the read half on every host core
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    n_sample = a.sample_frames or 16 * workers
    cfg, shifts = workload_config(a)
    sample = make_cpu_sample(a, n_sample)
    ctx = mp.get_context('fork')
    times = []
    with cf.ProcessPoolExecutor(max_workers=workers, mp_context=ctx) as pool:
        for it in range(a.warmup + a.steps):
            t = cpu_reference_step(sample, shifts, pool, workers)
            if it >= a.warmup:
                times.append(t)
    ms = 1e3 * float(np.mean(times))
    value = n_sample / (ms / 1e3)
    sample_desc = ('oracle/_ref absent: %d frames of %dx%d at %d shifts per step through the oracle PORT of solex_read '
                   '(mean/max, line detection + cubic fit, reconstruction); circularisation + transversalium (O(1) in '
                   'frame count) not included, which favours the CPU arm' % (n_sample, a.width, a.height, len(shifts)))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'u16+f64', 'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': workers, 'kind': 'port', 'sample': sample_desc},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ===================================================== the other BASELINE configurations
def config_lines(eng, max_files=16):
    """BASELINE.json configs[0..3] at full size through the command line front end (SHG_MAIN.main, scan files on
    tmpfs, `-c`: CLAHE + one PNG per image written): wall time of the whole invocation, frames/s, the file ->
    pinned ring -> HBM ingest rate and how often the pinned ring was reused instead of reallocated.  These are
    the parity-test configurations (tests/test_gpu_cli_configs.py checks them against the oracle); the lines put
    their timings into the driver's record."""
    import shutil
    import tempfile

    import cv2
    from solex_ser_recon_en_b200 import SHG_MAIN, synth
    from solex_ser_recon_en_b200.engine import ScanGeometry
    base = '/dev/shm' if os.path.isdir('/dev/shm') else tempfile.gettempdir()
    free = shutil.disk_usage(base).free
    work = tempfile.mkdtemp(prefix='shg_cfg_', dir=base)
    os.environ['SHG_NO_CONFIG'] = '1'
    out = []

    def device_payload(n, w, h, bpp, seed):
        st = eng.synth_stack(ScanGeometry(w, h, bpp, n), seed=seed)
        host = st.frames.cpu().numpy()
        del st
        return host

    def ser(name, n, w, h, seed):
        p = os.path.join(work, name)
        with open(p, 'wb') as f:
            f.write(synth.ser_header(w, h, 16, n))
            f.write(device_payload(n, w, h, 2, seed).tobytes())
        return p

    def run(label, flags, files, frames, payload_bytes, reps=2):
        best = None
        for _ in range(reps):
            SHG_MAIN.options.update(shift=[0], ratio_fixe=None, slant_fix=None, flip_x=False, crop_width_square=False,
                                    clahe_only=False, save_fit=False, fixed_width=None, output_dir=work)
            n_log, c0, r0 = len(eng.ingest_log), eng.ring_creates, eng.ring_reuses
            t0 = time.perf_counter()
            rc = SHG_MAIN.main(flags + files)
            wall = time.perf_counter() - t0
            assert rc == 0
            ing = [st for _, st in eng.ingest_log[n_log:]]
            rec = {'config': label, 'flags': ' '.join(flags), 'files': len(files), 'frames': frames, 'wall_s': round(wall, 4),
                   'frames_per_s': round(frames / wall, 1), 'payload_GB': round(payload_bytes / 1e9, 3),
                   'file_to_hbm_GBps': round(sum(s[2] for s in ing) / max(1e-9, sum(s[0] for s in ing)) / 1e9, 2)
                   if ing else None,
                   'pinned_ring_allocations': eng.ring_creates - c0, 'pinned_ring_reuses': eng.ring_reuses - r0}
            if best is None or rec['wall_s'] < best['wall_s']:
                best = rec
        out.append(best)

    try:
        f1 = ser('cfg1.SER', 1000, 1280, 200, 1)
        run('configs[0]: 16-bit SER 1000 x 1280x200, shift 0', ['-c'], [f1], 1000, 1000 * 1280 * 200 * 2)
        # f3: the spectral analyser's pattern on the same file -- all_video_reader keeps the scan in
        # HBM; reset() + read_video_improved at a NEW shift is one kernel over the resident stack
        from solex_ser_recon_en_b200 import solex_util, video_reader
        rdr = video_reader.all_video_reader(f1)
        opt = dict(SHG_MAIN.options, shift=[0], _nolog=True, clahe_only=True, output_dir=work)
        _, fit, _, _ = solex_util.compute_mean_return_fit(rdr, opt, {}, rdr.iw, rdr.ih, '')
        lat, lat_host = [], []
        for sh in list(range(-12, 13)):
            opt['shift'] = [sh]
            rdr.reset()
            eng.sync()
            t0 = time.perf_counter()
            disks, _, _, _ = solex_util.read_video_improved(rdr, fit, opt)
            eng.sync()
            t1 = time.perf_counter()
            host = np.asarray(disks[0])
            t2 = time.perf_counter()
            lat.append((t1 - t0) * 1e3)
            lat_host.append((t2 - t0) * 1e3)
        out[-1]['resident_rereconstruction'] = {
            'what': 'all_video_reader (scan resident in HBM): reset() + read_video_improved at one new shift, 25 shifts',
            'ms_per_shift_median': round(float(np.median(lat)), 4), 'ms_per_shift_max': round(float(np.max(lat[2:])), 4),
            'ms_per_shift_with_host_copy_median': round(float(np.median(lat_host)), 4),
            'image_shape': list(host.shape)}
        del rdr, disks
        os.remove(f1)
        f2 = os.path.join(work, 'cfg2.avi')
        pay = device_payload(2000, 1920, 256, 1, 2).reshape(2000, 256, 1920)
        vw = cv2.VideoWriter(f2, 0, 25.0, (1920, 256), isColor=False)
        for fr in pay:
            vw.write(fr)
        vw.release()
        del pay
        run('configs[1]: 8-bit AVI 2000 x 1920x256, mirror X + crop square', ['-cms'], [f2], 2000, 2000 * 1920 * 256)
        os.remove(f2)
        f3 = ser('cfg3.SER', 4000, 2048, 300, 3)
        run('configs[2]: 16-bit SER 4000 x 2048x300, -w-10:10:1 (21 shifts)', ['-cw-10:10:1'], [f3], 4000,
            4000 * 2048 * 300 * 2)
        os.remove(f3)
        per_file = 3000 * 2048 * 256 * 2
        n_files = int(max(2, min(max_files, (free * 0.7 - 2e9) // per_file)))
        f4 = [ser('cfg4_%02d.SER' % i, 3000, 2048, 256, 40 + i) for i in range(n_files)]
        run('configs[3]: %d x 16-bit SER 3000 x 2048x256 back to back (one invocation)' % n_files, ['-c'], f4,
            3000 * n_files, per_file * n_files, reps=1 if n_files > 4 else 2)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return out


# ============================================================ output checksum
CRC_FILE = os.path.join(ROOT, 'profiles', 'outputs_crc.json')


def source_sha():
    """sha256 of everything that decides the output bits (kernels, host logic, the synthetic scan recipe):
    a stored checksum is only binding for the build it was recorded with."""
    import glob
    import hashlib
    h = hashlib.sha256()
    pkg = os.path.join(ROOT, 'solex_ser_recon_en_b200')
    files = sorted(glob.glob(os.path.join(pkg, 'csrc', '*.cu*')) + glob.glob(os.path.join(pkg, '*.py')) +
                   [os.path.join(ROOT, 'include', 'shg.h')])
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, 'rb').read())
    return h.hexdigest()[:16]


def outputs_crc(eng, results, shifts, world):
    """CRC32 over the 64-bit checksums (shg_checksum_u16) of every final image in shift order.  Each image
    lives on exactly one rank; the per-image values are combined with one all-reduce."""
    import zlib
    import torch
    import torch.distributed as dist
    vec = torch.zeros(len(shifts), dtype=torch.int64)
    for j, sh in enumerate(shifts):
        im = results.get('bench_shift=%d' % sh)
        if im is not None:
            v = eng.checksum(im.rows_tensor().contiguous())
            vec[j] = v - (1 << 64) if v >= (1 << 63) else v
    if world > 1:
        dev = vec.to(eng.device)
        dist.all_reduce(dev, op=dist.ReduceOp.SUM)          # one non-zero contribution per image (mod 2^64)
        vec = dev.cpu()
    return '%08x' % (zlib.crc32(vec.numpy().astype('<i8').tobytes()) & 0xffffffff)


def crc_key(a, shifts):
    return '%dx%dx%d_s%d' % (a.frames, a.width, a.height, len(shifts))


def check_crc(a, shifts, crc, world):
    """(expected, match): match is True / False when a value recorded by THIS build exists, else None."""
    try:
        book = json.load(open(CRC_FILE))
    except Exception:
        book = {}
    rec = book.get(crc_key(a, shifts))
    if a.record_crc and world == 1:
        book[crc_key(a, shifts)] = {'crc': crc, 'source_sha': source_sha(), 'n_gpus': 1}
        os.makedirs(os.path.dirname(CRC_FILE), exist_ok=True)
        with open(CRC_FILE, 'w') as f:
            json.dump(book, f, indent=1, sort_keys=True)
        return crc, True
    if rec is None:
        return None, None
    if rec.get('source_sha') != source_sha():
        return rec['crc'], (True if rec['crc'] == crc else None)     # other build: informative only
    return rec['crc'], rec['crc'] == crc


# ================================================================== GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    from solex_ser_recon_en_b200 import Solex_recon, parallel
    from solex_ser_recon_en_b200.engine import DeviceStack, ScanGeometry, get_engine
    from solex_ser_recon_en_b200.solex_util import release_resident
    from solex_ser_recon_en_b200.video_reader import device_scan, memory_scan

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    eng = get_engine(local)
    cfg, shifts = workload_config(a)
    geom = ScanGeometry(a.width, a.height, 2, a.frames)
    k0, k1 = parallel.frame_range(a.frames, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the scan, synthesised straight into HBM (this rank's frame range)
    stack = eng.synth_stack(geom, k0=k0, n=k1 - k0, seed=5)
    torch.cuda.synchronize()
    results = {}

    def sink(basefich, image, cercle):
        results[basefich] = image                       # DeviceImage: final hot-path output of one shift
        return None

    def one_pass(reader, to_host):
        opt = default_options(shifts)
        opt['_result_sink'] = sink
        results.clear()
        disk_list, bounds, hdr = Solex_recon.solex_read_reader(reader, opt, 'bench')
        # every rank circularises / corrects the shifts whose images it owns (all of them at N=1)
        Solex_recon.solex_process(opt, disk_list, bounds, hdr)
        nbytes = 0
        if to_host:
            for im in results.values():
                nbytes += im.numpy().nbytes             # D2H into pinned memory, through this rank's own PCIe link
        return nbytes

    sampler = ClockSampler(local) if rank == 0 else None

    def timed_loop(reader_factory, steps, warmup, to_host, label):
        for _ in range(warmup):
            one_pass(reader_factory(), to_host)
        eng.profile_stages = True
        eng.stage_report()
        launches0 = eng.n_launches
        barrier()
        t_wall0 = time.time()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        d2h = 0
        for _ in range(steps):
            d2h = one_pass(reader_factory(), to_host)
        e1.record()
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        stages = eng.stage_report()
        eng.profile_stages = False
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return dict(ms_per_step=ms / steps, stages={k: v / steps for k, v in stages.items()},
                    launches=(eng.n_launches - launches0) // max(1, steps), d2h=d2h, wall=(t_wall0, t_wall1))

    def timed_batch(reader_factory, steps, warmup):
        """End to end through the public batch entry point (Solex_recon.solex_do_work_readers: what SHG_MAIN runs
        for a list of files): `steps` scans back to back, each from pinned host memory to final images in pinned
        host memory.  The driver pipelines them like the reference's reader / worker-pool split: scan i+1 is
        ingested while a worker thread finishes scan i (transversalium, device -> host copies).  Every step's
        H2D and D2H lie inside the timed region; the first ingest and the last tail are not overlapped."""
        d2h_log = []

        def sink_host(basefich, image, cercle):
            results[basefich] = image
            d2h_log.append(image.numpy().nbytes)        # D2H into pinned memory, through this rank's own PCIe link
            return None

        def tasks(n):
            for _ in range(n):
                opt = default_options(shifts)
                opt['_result_sink'] = sink_host
                yield reader_factory(), opt, 'bench'

        if warmup:
            Solex_recon.solex_do_work_readers(tasks(warmup))
        eng.profile_stages = True
        eng.stage_report()
        launches0 = eng.n_launches
        del d2h_log[:]
        barrier()
        t_wall0 = time.time()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        Solex_recon.solex_do_work_readers(tasks(steps))
        e1.record()
        barrier()
        t_wall1 = time.time()
        ms = e0.elapsed_time(e1)
        stages = eng.stage_report()
        eng.profile_stages = False
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return dict(ms_per_step=ms / steps, stages={k: v / steps for k, v in stages.items()},
                    launches=(eng.n_launches - launches0) // max(1, steps), d2h=sum(d2h_log) // max(1, steps),
                    wall=(t_wall0, t_wall1))

    # ---- value: stack resident in HBM
    dev = timed_loop(lambda: device_scan(stack), a.steps, a.warmup, False, 'device')
    dev['crc'] = outputs_crc(eng, results, shifts, world)          # images of the LAST timed step (outside the timing)

    if a.timeline:                                   # every rank runs the pass (it has collectives); rank 0 prints
        eng.profile_stages = True
        eng.stage_report()
        barrier()
        t0 = time.perf_counter()
        one_pass(device_scan(stack), False)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        tl = sorted(eng.stage_timeline(), key=lambda t: t[1])
        lines = ['timeline r%d %-28s gpu %8.3f -> %8.3f (%.3f ms)   host %8.3f -> %8.3f' %
                 (rank, name, s0, s1, s1 - s0, h0, h1) for name, s0, s1, h0, h1 in tl]
        lines.append('timeline r%d wall %.3f ms' % (rank, wall))
        if rank == 0:
            print('\n'.join(lines), file=sys.stderr)
        try:
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(ROOT, 'gpurun_out', 'timeline_n%d_rank%d.txt' % (world, rank)), 'w') as f:
                f.write('\n'.join(lines) + '\n')
        except Exception:
            pass
        eng.profile_stages = False

    # ---- e2e: payload in pinned host memory -> H2D -> ... -> D2H
    e2e = None
    if not a.no_e2e:
        nbytes = (k1 - k0) * geom.frame_bytes
        host_ptr = eng.pinned_alloc(nbytes)               # exact size: 84 GB at N=1
        eng.copy(host_ptr, stack.frames.data_ptr(), nbytes, 'd2h')
        torch.cuda.synchronize()

        def host_reader():
            release_resident()
            r = memory_scan(host_ptr - k0 * geom.frame_bytes, a.width, a.height, 16, a.frames)
            r.device_stack = stack                      # refill the same HBM buffer (pinned-buffer / HBM reuse)
            return r
        # PCIe roofline denominator, measured in this run: pinned H2D copies back to back for ~1.5 s on every rank at
        # the same time (the CONCURRENT, SUSTAINED per-GPU rate: a best-of-3 single copy overstated it)
        probe = min(nbytes, 4 << 30)
        barrier()
        eng.copy(stack.frames.data_ptr(), host_ptr, probe, 'h2d')      # warm-up
        torch.cuda.synchronize()
        barrier()
        p0 = torch.cuda.Event(enable_timing=True)
        p1 = torch.cuda.Event(enable_timing=True)
        t_probe = time.perf_counter()
        n_probe = 0
        p0.record()
        while time.perf_counter() - t_probe < 1.5:
            eng.copy(stack.frames.data_ptr(), host_ptr, probe, 'h2d')
            n_probe += 1
            if n_probe % 2 == 0:
                torch.cuda.synchronize()
        p1.record()
        torch.cuda.synchronize()
        pcie_peak = n_probe * probe / (p0.elapsed_time(p1) * 1e-3) / 1e9
        if world > 1:                                   # the slowest rank's rate bounds a step that waits for all
            t = torch.tensor([pcie_peak], dtype=torch.float64, device='cuda')
            tmin = t.clone()
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            pcie_peak_min, pcie_peak = float(tmin.item()), float(t.item()) / world
        else:
            pcie_peak_min = pcie_peak
        e_steps = a.e2e_steps or a.steps
        e2e = timed_batch(host_reader, e_steps, min(a.warmup, 3))
        e2e['crc'] = outputs_crc(eng, results, shifts, world)
        e2e['h2d'] = a.frames * geom.frame_bytes
        e2e['pcie_peak'] = pcie_peak
        e2e['pcie_peak_min'] = pcie_peak_min
        if world > 1:
            t = torch.tensor([float(e2e['d2h'])], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            e2e['d2h'] = int(t.item())
        eng.pinned_free(host_ptr)
    if world > 1:
        parallel.release_exchange()

    # ---- roofline of the dominant kernel (pass-1 accumulation), timed live on its stream
    acc_ms = []
    for _ in range(3):
        eng.accumulate(stack)
    for _ in range(max(3, a.steps)):
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        eng.accumulate(stack)
        s1.record()
        torch.cuda.synchronize()
        acc_ms.append(s0.elapsed_time(s1))
    acc_bytes = (k1 - k0) * geom.frame_bytes
    clocks = sampler.window(*dev['wall']) if sampler else None
    if sampler:
        sampler.stop()
    config_results = None
    if world == 1 and not a.no_configs and not cfg.get('reduced'):
        del stack
        results.clear()
        release_resident()
        torch.cuda.empty_cache()
        import contextlib
        import io
        try:
            with contextlib.redirect_stdout(io.StringIO()):          # the CLI prints per-file progress lines
                config_results = config_lines(eng)
        except Exception as e:                                       # never lose the main line to a side measurement
            config_results = [{'error': repr(e)}]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'MEASURED_PEAKS.json hbm_gbs (measured copy)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'accumulate_traffic.json'))).get('dram_bytes_per_launch')
        except Exception:
            pass
        in_step = dev['stages'].get('accumulate')
        t_acc = float(np.mean(acc_ms))
        achieved = acc_bytes / (t_acc * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': a.frames / (dev['ms_per_step'] * 1e-3), 'unit': 'frames/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': dev['ms_per_step'], 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'u16+f64', 'data': 'synthetic',
            'config': dict(cfg, parallelism='frames sharded over %d rank(s); images owned by shift block' % world),
            'clocks': clocks, 'gpu_launches': dev['launches'], 'outputs_crc': dev['crc'],
            'stages_ms': {k: round(v, 3) for k, v in sorted(dev['stages'].items())},
            'roofline': {'bound': 'hbm', 'kernel': 'accumulate_u16_kernel (pass 1: integer sum + max of the stack)',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': acc_bytes, 'ms_per_launch': t_acc,
                         'ms_per_launch_inside_step': in_step, 'frac_of_8TBps_nominal': achieved / 8000.0},
        }
        # secondary roofline: pass 2 (reconstruction), algorithmic bytes of SURVEY 8(d) over its stage time
        nb = shifts[-1] - shifts[0] + 2
        recon_bytes = (k1 - k0) * (geom.ih * nb * 2 + len(shifts) * geom.ih * 2)
        t_rec = dev['stages'].get('recon+gather')
        if t_rec:
            line['roofline_recon'] = {'bound': 'hbm' if world == 1 else 'hbm+nvlink', 'kernel': 'recon_tma_pair_kernel (pass 2; also tracks each image minimum)',
                                      'achieved': recon_bytes / (t_rec * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                                      'frac': recon_bytes / (t_rec * 1e-3) / 1e9 / peak,
                                      'algorithmic_bytes_per_launch': recon_bytes, 'ms_in_step': t_rec}
        if e2e is not None:
            line['e2e'] = {'value': a.frames / (e2e['ms_per_step'] * 1e-3), 'unit': 'frames/s',
                           'h2d_bytes_per_step': e2e['h2d'], 'd2h_bytes_per_step': e2e['d2h'],
                           'ms_per_step': e2e['ms_per_step'],
                           'h2d_GBps': e2e['h2d'] / (e2e['ms_per_step'] * 1e-3) / 1e9,
                           'pcie_h2d_peak_GBps_per_gpu': e2e['pcie_peak'],
                           'pcie_h2d_peak_GBps_slowest_gpu': e2e['pcie_peak_min'],
                           'pcie_peak_method': 'pinned 4 GiB H2D copies back to back for 1.5 s on all ranks at once; mean over ranks',
                           'pcie_frac': e2e['h2d'] / (e2e['ms_per_step'] * 1e-3) / 1e9 / (e2e['pcie_peak'] * world),
                           'stages_ms': {k: round(v, 3) for k, v in sorted(e2e['stages'].items())}}
        expected, match = check_crc(a, shifts, dev['crc'], world)
        line['outputs_crc_expected'] = expected
        line['outputs_crc_match'] = match
        line['outputs_crc_note'] = ('CRC32 over the 64-bit position-sensitive checksums of the %d final images in shift '
                                    'order (last timed step); expected = the value recorded at N=1 by '
                                    'bench.py --record-crc (profiles/outputs_crc.json), binding when recorded with '
                                    'this build (source sha %s)' % (len(shifts), source_sha()))
        if e2e is not None:
            line['e2e']['outputs_crc'] = e2e['crc']
        if world == 1 and not a.no_cpu:
            line['cpu_baseline'] = cpu_baseline_subprocess(a)
        if config_results is not None:
            line['config_lines'] = config_results
        print(json.dumps(line))
        sys.stdout.flush()
        if match is False or (e2e is not None and e2e['crc'] != dev['crc']):
            raise AssertionError('outputs_crc mismatch: resident %s, e2e %s, expected %s (N=1, this build)' %
                                 (dev['crc'], e2e['crc'] if e2e else None, expected))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline_subprocess(a):
    """The CPU arm on a bounded sample in a clean child process (no CUDA context to fork)."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '2', '--warmup', '1',
           '--frames', str(a.frames), '--width', str(a.width), '--height', str(a.height)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK='0'))
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)['cpu_baseline']
        return {'error': (out.stderr or 'no output')[-300:]}
    except Exception as e:
        return {'error': repr(e)}


def main():
    a = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    try:
        if a.impl == 'reference':
            return run_reference_arm(a)
        return run_b200(a)
    except BaseException as e:
        if isinstance(e, SystemExit) and not e.code:
            raise
        # torchrun's own summary pushes a rank's traceback out of the captured tail: say who failed, where, and
        # leave a per-rank file behind; then take the whole job down so the other ranks do not wait in NCCL
        import traceback
        text = 'RANK %d: %s' % (rank, traceback.format_exc())
        sys.stderr.write(text + '\n')
        sys.stderr.flush()
        try:
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(ROOT, 'gpurun_out', 'bench_rank%d.err' % rank), 'w') as f:
                f.write(text)
        except Exception:
            pass
        sys.stdout.flush()
        os._exit(1)


if __name__ == '__main__':
    sys.exit(main())
