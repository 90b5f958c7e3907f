"""GPU: BASELINE.json configs[1..3] at their FULL sizes through the command line front end
(`python -m solex_ser_recon_en_b200.SHG_MAIN <flags> files...`, here SHG_MAIN.main(argv) in-process):

  configs[1]  2000-frame 8-bit AVI 1920x256, -cms (mirror X + crop square)
  configs[2]  4000-frame 16-bit SER 2048x300, -cw-10:10:1 (21 shift images)
  configs[3]  16 SER files x 3000 frames x 2048x256 in ONE invocation (batch mode, pinned-ring reuse)

checked against the CPU oracle on the same files: integer mean frame and line fit (the oracle reads the whole
file), disks bit-exact on sampled frames at every shift, and one complete image per config pushed through the
oracle's ellipse fit + warp + transversalium (<= 1 DN).  The seams of the CLI run are observed through
options['_observer'] (a callback; it changes nothing in the run: the PNG tail still executes).
"""
import os
import shutil

import numpy as np
import pytest

from oracle import shg_oracle as O

pytestmark = pytest.mark.gpu
os.environ['SHG_NO_CONFIG'] = '1'


def _scratch(tmp_path, need_bytes):
    """A directory with room for the scan files: tmpfs when it is large enough (the page cache then IS the file)."""
    for base in ('/dev/shm', str(tmp_path)):
        try:
            if shutil.disk_usage(base).free > need_bytes * 1.2:
                d = os.path.join(base, 'shg_cli_cfg_%d' % os.getpid())
                os.makedirs(d, exist_ok=True)
                return d
        except Exception:
            continue
    pytest.skip('no scratch space for %.1f GB of scan files' % (need_bytes / 1e9))


def _write_ser_from_device(path, n, w, h, seed):
    from solex_ser_recon_en_b200 import synth
    from solex_ser_recon_en_b200.engine import ScanGeometry, get_engine
    eng = get_engine(0)
    st = eng.synth_stack(ScanGeometry(w, h, 2, n), seed=seed)
    host = st.frames.cpu().numpy()
    with open(path, 'wb') as f:
        f.write(synth.ser_header(w, h, 16, n))
        f.write(host.tobytes())
    return host.view(np.uint16).reshape(n, h, w)


class Seams:
    """Collects what the run produced at its seams (host copies of sampled frames / chosen images only)."""

    def __init__(self, sample_frames, keep_full):
        self.sample_frames, self.keep_full = sample_frames, keep_full
        self.disks, self.full, self.final, self.shapes = {}, {}, {}, {}

    def __call__(self, stage, name, obj):
        import torch
        base = os.path.basename(name)
        if stage == 'disks':
            for i, d in enumerate(obj):
                t = d.tensor                                       # frame-major (N, ih)
                rows = t.view(torch.int16)[self.sample_frames].cpu().numpy().view(np.uint16)
                self.disks[(base, i)] = (rows, d.flip)
                if i in self.keep_full.get(base, ()):
                    self.full[(base, i)] = np.asarray(d).copy()
        elif stage == 'detrans':
            self.final[base] = np.asarray(obj).copy()
            self.shapes[base] = tuple(obj.shape)


def _run_cli(flags, files, out_dir, observer):
    from solex_ser_recon_en_b200 import SHG_MAIN
    from solex_ser_recon_en_b200.SHG_MAIN import options
    options.update(shift=[0], ratio_fixe=None, slant_fix=None, flip_x=False, crop_width_square=False, clahe_only=False,
                   save_fit=False, fixed_width=None, output_dir=out_dir, _observer=observer)
    try:
        assert SHG_MAIN.main(flags + files) == 0
    finally:
        options.pop('_observer', None)


def _oracle_fit(stack, eight_bit):
    """mean frame + line fit of a whole file by the oracle, in chunks (the stack may be several GB)."""
    n = stack.shape[0]
    s = np.zeros(stack.shape[1:], np.uint64)
    m = np.zeros(stack.shape[1:], stack.dtype)
    for a in range(0, n, 256):
        ps, pm = O.raw_sum_max(stack, a, min(n, a + 256))
        s += ps
        np.maximum(m, pm, out=m)
    mean_img, max_img = O.finalize_mean_max(s, m, n, eight_bit)
    y1, y2 = O.slit_extent(max_img)
    mi, ms = O.line_minima(mean_img, y1, y2)
    return O.line_fit(mi, ms, y1, y2, mean_img.shape[0]), (y1, y2)


def _check_sampled_disks(seams, base, stack, fit, shifts, ks, flip):
    ref = O.recon(np.stack([stack[k] for k in ks]), fit, shifts)
    for i, sh in enumerate(shifts):
        rows, flipped = seams.disks[(base, i)]
        assert flipped == flip
        assert np.array_equal(rows.T, ref[i]), 'disk of shift %d differs from the oracle on frames %s' % (sh, ks)


def _check_full_image(seams, base, shifts, bounds, final_name, flip):
    """One complete image through the oracle's post-processing, from the run's own disks."""
    disks = [seams.full[(base, 0)], seams.full[(base, 1)]]        # ellipse-fit shift, shift 0 (np.asarray applies flip)
    want, geom = O.solex_process(disks, shifts[:2], [0], bounds)
    det = seams.final[final_name]
    ref = want[0][1]
    assert det.shape == ref.shape
    d = np.abs(det.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) < 1e-3, (int(d.max()), float(np.mean(d != 0)))
    return geom


def test_config2_full_size_avi_mirror_crop(tmp_path):
    import cv2
    from solex_ser_recon_en_b200 import synth
    spec = synth.ca_k_8bit(2000, 1920, 256, seed=2)
    work = _scratch(tmp_path, 3e9)
    try:
        stack = synth.frames(spec, 0, spec.n_frames)
        path = os.path.join(work, 'cfg2.avi')
        vw = cv2.VideoWriter(path, 0, 25.0, (spec.width, spec.height), isColor=False)      # uncompressed Y800
        assert vw.isOpened()
        for fr in stack:
            vw.write(fr)
        vw.release()
        ks = [0, 1, 999, 1999]
        seams = Seams(ks, {'cfg2': (0, 1)})
        _run_cli(['-cms'], [path], work, seams)
        lf, bounds = _oracle_fit(stack, True)
        shifts = [10, 0]
        _check_sampled_disks(seams, 'cfg2', stack, lf['fit'], shifts, ks, True)
        _check_full_image(seams, 'cfg2', shifts, bounds, 'cfg2_shift=0', True)
        png = cv2.imread(os.path.join(work, 'cfg2_shift=0_clahe.png'), cv2.IMREAD_UNCHANGED)
        assert png is not None and png.dtype == np.uint16 and png.shape == (1920, 1920)       # -s: cropped square
        assert not os.path.exists(os.path.join(work, 'cfg2_shift=0_protus.png'))             # -c: clahe only
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_config3_full_size_21_shifts(tmp_path):
    import cv2
    n, w, h = 4000, 2048, 300
    work = _scratch(tmp_path, 6e9)
    try:
        path = os.path.join(work, 'cfg3.SER')
        stack = _write_ser_from_device(path, n, w, h, seed=3)
        ks = [0, 1, n // 2, n - 1]
        seams = Seams(ks, {'cfg3': (0, 1)})
        _run_cli(['-cw-10:10:1'], [path], work, seams)
        lf, bounds = _oracle_fit(stack, False)
        shifts = O.shift_list(list(range(-10, 11)))
        assert len(shifts) == 21 and shifts[:2] == [10, 0]
        _check_sampled_disks(seams, 'cfg3', stack, lf['fit'], shifts, ks, False)
        _check_full_image(seams, 'cfg3', shifts, bounds, 'cfg3_shift=0', False)
        for sh in range(-10, 11):                                                            # one image per shift
            assert ('cfg3_shift=%d' % sh) in seams.final
            assert os.path.exists(os.path.join(work, 'cfg3_shift=%d_clahe.png' % sh))
        assert len({seams.shapes['cfg3_shift=%d' % sh] for sh in range(-10, 11)}) == 1
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_config4_sixteen_files_back_to_back(tmp_path):
    """One CLI invocation over 16 files: every file's result equals the result of running that file alone (the
    pipelining of file i+1's ingest under file i's post-processing and the reuse of the pinned ring and of the
    HBM buffers change nothing), file 0 is checked against the oracle, and the pinned ring is allocated once."""
    from solex_ser_recon_en_b200.engine import get_engine
    n, w, h, n_files = 3000, 2048, 256, 16
    work = _scratch(tmp_path, n_files * n * w * h * 2 + 2e9)
    try:
        paths, first = [], None
        for i in range(n_files):
            p = os.path.join(work, 'cfg4_%02d.SER' % i)
            st = _write_ser_from_device(p, n, w, h, seed=40 + i)
            if i == 0:
                first = st
            paths.append(p)
        eng = get_engine(0)
        ks = [0, 1, n // 2, n - 1]
        seams = Seams(ks, {'cfg4_00': (0, 1)})
        creates0, reuses0 = eng.ring_creates, eng.ring_reuses
        _run_cli(['-c'], paths, work, seams)
        assert eng.ring_creates - creates0 <= 1 and eng.ring_reuses - reuses0 >= n_files - 1
        assert sorted(seams.final) == ['cfg4_%02d_shift=0' % i for i in range(n_files)]
        lf, bounds = _oracle_fit(first, False)
        _check_sampled_disks(seams, 'cfg4_00', first, lf['fit'], [10, 0], ks, False)
        _check_full_image(seams, 'cfg4_00', [10, 0], bounds, 'cfg4_00_shift=0', False)
        # files run alone give the same images as inside the batch
        for i in (3, 15):
            alone = Seams(ks, {})
            _run_cli(['-c'], [paths[i]], work, alone)
            name = 'cfg4_%02d_shift=0' % i
            assert np.array_equal(alone.final[name], seams.final[name]), name
        assert len({seams.final['cfg4_%02d_shift=0' % i].tobytes()[:4096] for i in range(n_files)}) > 1   # distinct scans
    finally:
        shutil.rmtree(work, ignore_errors=True)
