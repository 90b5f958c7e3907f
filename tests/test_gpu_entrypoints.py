"""GPU: the drop-in entry points (video_reader -> Solex_recon.solex_read ->
solex_process) against the arrays recorded at the same seams of the unmodified
reference (tests/golden, made by oracle/make_golden.py), including the
`options` side effects and the log file."""
import math
import os

import numpy as np
import pytest

from oracle import ref_shim
from oracle.make_golden import CASES
from helpers import ALL_CASES, case_file, golden

pytestmark = pytest.mark.gpu


def options_for(name, tmp_path, **over):
    case = CASES[name]
    o = ref_shim.default_options(shift=list(case['shift']), flip_x=case['flip_x'], ratio_fixe=case.get('ratio_fixe'))
    o.pop('_nolog')
    o['output_dir'] = str(tmp_path)
    o.update(over)
    return o


@pytest.mark.parametrize('name', ALL_CASES)
def test_solex_read_matches_reference(name, tmp_path):
    from solex_ser_recon_en_b200 import Solex_recon
    g = golden(name)
    path = case_file(name, tmp_path)
    opt = options_for(name, tmp_path)
    disk_list, bounds, hdr = Solex_recon.solex_read(path, opt)
    assert [int(s) for s in opt['shift']] == [int(s) for s in g['shift']]
    assert opt['shift_requested'] == list(CASES[name]['shift'])
    assert opt['basefich0'] == os.path.splitext(path)[0]
    assert (int(bounds[0]), int(bounds[1])) == (int(g['y1']), int(g['y2']))
    assert hdr['NAXIS1'] == int(g['iw']) and hdr['NAXIS2'] == int(g['ih'])
    assert len(disk_list) == len(opt['shift'])
    for i, d in enumerate(disk_list):
        a = np.asarray(d)
        assert a.dtype == np.uint16 and a.shape == (int(g['ih']), int(g['n']))
        assert np.array_equal(a, g[f'disk{i}']), f'shift {opt["shift"][i]}'
    log = open(os.path.join(str(tmp_path), os.path.basename(opt['basefich0']) + '_log.txt')).read()
    assert 'Pixel shift : ' + str(list(CASES[name]['shift'])) in log
    assert 'Number of frames : %d' % int(g['n']) in log
    assert 'Vertical limits y1, y2 : %d %d' % (int(g['y1']), int(g['y2'])) in log
    assert 'Spectral line polynomial fit: [' in log


@pytest.mark.parametrize('name', ALL_CASES)
def test_solex_process_matches_reference(name, tmp_path):
    """Circularised and detransversaliumed images per requested shift, the
    ellipse geometry written back into `options`, and the gain vector."""
    from solex_ser_recon_en_b200 import Solex_recon
    g = golden(name)
    path = case_file(name, tmp_path)
    opt = options_for(name, tmp_path, save_fit=True)
    disk_list, bounds, hdr = Solex_recon.solex_read(path, opt)
    seen = {}
    real = Solex_recon.single_image_process

    def spy(frame_circularized, hdr_, options, cercle0, borders, basefich, backup_bounds, **kw):
        sh = int(basefich.rsplit('_shift=', 1)[1])
        out = real(frame_circularized, hdr_, options, cercle0, borders, basefich, backup_bounds, **kw)
        seen[sh] = (np.asarray(frame_circularized), np.array(options['_transversalium_cache']),
                    np.array(cercle0, dtype='d'), np.array(borders, dtype='d'))
        return out

    Solex_recon.single_image_process = spy
    try:
        assert Solex_recon.solex_process(opt, disk_list, bounds, hdr) is None
    finally:
        Solex_recon.single_image_process = real
    assert sorted(seen) == sorted(int(s) for s in g['shift_requested'])
    np.testing.assert_allclose(float(opt['ratio_fixe']), float(g['ratio']), rtol=1e-9)
    if not math.isnan(float(g['slant'])):
        np.testing.assert_allclose(float(opt['slant_fix']), float(g['slant']), rtol=1e-6, atol=1e-9)
    for sh, (circ, gain, cercle, borders) in seen.items():
        want = g[f'circ_{sh}']
        assert circ.shape == want.shape and circ.dtype == np.uint16
        assert np.abs(circ.astype(np.int32) - want.astype(np.int32)).max() <= 1          # north_star: <= 1 DN
        assert np.mean(circ != want) < 1e-3
        np.testing.assert_allclose(gain, g[f'gain_{sh}'], rtol=1e-5)                      # north_star: 1e-5 relative
        np.testing.assert_allclose(cercle, g['cercle'], rtol=1e-9)
        np.testing.assert_allclose(borders, g['borders'], rtol=1e-9, atol=1e-9)
        # -f output: the detransversaliumed FITS holds the final hot-path image
        from solex_ser_recon_en_b200 import fits_min
        det_path = os.path.join(str(tmp_path), os.path.basename(opt['basefich0']) + f'_shift={sh}_detransversaliumed.fits')
        det = read_fits_u16(det_path)
        d = np.abs(det.astype(np.int32) - g[f'det_{sh}'].astype(np.int32))
        assert d.max() <= 1 and np.mean(d != 0) < 1e-3
        assert os.path.exists(os.path.join(str(tmp_path), os.path.basename(opt['basefich0']) + f'_shift={sh}_clahe.png'))
    log = open(os.path.join(str(tmp_path), os.path.basename(opt['basefich0']) + '_log.txt')).read()
    assert 'Transversalium correction : 301' in log and 'Mirror X : ' + str(CASES[name]['flip_x']) in log
    if CASES[name].get('ratio_fixe') is None:       # with a fixed ratio the reference only logs it when shift 10 is requested
        assert 'Y/X ratio : ' + '{:.3f}'.format(float(g['ratio'])) in log
    assert 'end time: ' in log


def read_fits_u16(path):
    raw = open(path, 'rb').read()
    head_end = raw.index(b'END' + b' ' * 77) + 80
    head_len = (head_end + 2879) // 2880 * 2880
    cards = {raw[i:i + 8].strip().decode(): raw[i + 10:i + 30].strip().decode() for i in range(0, head_end, 80)}
    w, h = int(cards['NAXIS1']), int(cards['NAXIS2'])
    data = np.frombuffer(raw, dtype='>i2', count=w * h, offset=head_len).reshape(h, w)
    return (data.astype(np.int32) + int(float(cards.get('BZERO', '0')))).astype(np.uint16)


def test_solex_do_work_batch_and_flags(tmp_path):
    """Two files back to back through the CLI entry (config-4 style), -c flag:
    only the clahe PNGs appear; the resident stack of file 1 is released."""
    from solex_ser_recon_en_b200 import SHG_MAIN
    os.environ['SHG_NO_CONFIG'] = '1'
    a = case_file('ser16_rot', tmp_path)
    b = case_file('ser8_rot_flip', tmp_path)
    SHG_MAIN.options['output_dir'] = str(tmp_path)
    try:
        assert SHG_MAIN.main(['-cw0,2', a, b]) == 0
    finally:
        SHG_MAIN.options['output_dir'] = ''
    names = set(os.listdir(str(tmp_path)))
    for base in ('ser16_rot', 'ser8_rot_flip'):
        for sh in (0, 2):
            assert f'{base}_shift={sh}_clahe.png' in names
            assert f'{base}_shift={sh}_protus.png' not in names
        assert f'{base}_log.txt' in names


def test_bad_file_raises_like_reference(tmp_path):
    from solex_ser_recon_en_b200 import Solex_recon
    bad = os.path.join(str(tmp_path), 'scan.txt')
    open(bad, 'w').write('x')
    with pytest.raises(Exception, match='neither is SER nor AVI'):
        Solex_recon.solex_read(bad, options_for('ser16_rot', tmp_path))


def test_reader_on_host_frames(tmp_path):
    """compute_mean_return_fit / read_video_improved accept the reference-style
    all_video_reader (oriented frames in RAM) as the spectral analyser uses them."""
    from solex_ser_recon_en_b200 import solex_util, video_reader
    g = golden('ser16_rot')
    path = case_file('ser16_rot', tmp_path)
    rdr = video_reader.all_video_reader(path)
    opt = options_for('ser16_rot', tmp_path, _nolog=True)
    opt['shift'] = [int(s) for s in g['shift']]
    mean_img, fit, y1, y2 = solex_util.compute_mean_return_fit(rdr, opt, {}, rdr.iw, rdr.ih, os.path.splitext(path)[0])
    assert np.array_equal(mean_img, g['mean_img']) and (y1, y2) == (int(g['y1']), int(g['y2']))
    np.testing.assert_allclose(fit, g['fit'], rtol=0, atol=1e-7)
    rdr.reset()
    disks, ih, iw, n = solex_util.read_video_improved(rdr, fit, opt)
    assert (ih, iw, n) == (int(g['ih']), int(g['iw']), int(g['n']))
    for i, d in enumerate(disks):
        assert np.array_equal(np.asarray(d), g[f'disk{i}'])


def test_solex_process_batched_path_matches_per_image_path(tmp_path):
    """Without -f the warp and transversalium stages run batched over all shifts;
    results must be identical to the per-image path, and identical to the reference."""
    from solex_ser_recon_en_b200 import Solex_recon
    name = 'ser16_rot'
    g = golden(name)
    path = case_file(name, tmp_path)
    got = {}

    def sink(basefich, image, cercle):
        got[int(basefich.rsplit('_shift=', 1)[1])] = np.asarray(image).copy()

    opt = options_for(name, tmp_path, _result_sink=sink)
    disk_list, bounds, hdr = Solex_recon.solex_read(path, opt)
    Solex_recon.solex_process(opt, disk_list, bounds, hdr)
    assert sorted(got) == sorted(int(s) for s in g['shift_requested'])
    for sh, det in got.items():
        d = np.abs(det.astype(np.int32) - g[f'det_{sh}'].astype(np.int32))
        assert d.max() <= 1 and np.mean(d != 0) < 1e-3
        np.testing.assert_allclose(opt['_transversalium_gains'][sh], g[f'gain_{sh}'], rtol=1e-5)


def test_resident_scan_is_reconstructed_again_without_reingest(tmp_path):
    """The spectral-analyser pattern (reference spectralAnalyserUI.py:345-346): the same scan
    reconstructed again at new shifts.  With options['_keep_stack'] the stack stays in HBM and the
    second read_video_improved costs one kernel, not another pass over the file."""
    from oracle import shg_oracle as O
    from solex_ser_recon_en_b200 import solex_util, video_reader
    from solex_ser_recon_en_b200.engine import get_engine
    from helpers import case_stack
    g = golden('ser16_rot')
    _, stack = case_stack('ser16_rot')
    path = case_file('ser16_rot', tmp_path)
    opt = options_for('ser16_rot', tmp_path, _nolog=True, _keep_stack=True)
    opt['shift'] = [10, 0]
    rdr = video_reader.video_reader(path)
    mean_img, fit, y1, y2 = solex_util.compute_mean_return_fit(rdr, opt, {}, rdr.iw, rdr.ih, os.path.splitext(path)[0])
    assert np.array_equal(mean_img, g['mean_img'])
    eng = get_engine()
    real_ingest = eng.ingest_file
    calls = []
    eng.ingest_file = lambda *a, **k: (calls.append(1), real_ingest(*a, **k))[1]
    try:
        for shifts in ([10, 0], [-7, 3, 12], [1]):
            opt['shift'] = shifts
            disks, ih, iw, n = solex_util.read_video_improved(video_reader.video_reader(path), fit, opt)
            ref = O.recon(stack, fit, shifts)
            for i in range(len(shifts)):
                assert np.array_equal(np.asarray(disks[i]), ref[i]), shifts[i]
    finally:
        eng.ingest_file = real_ingest
        solex_util.release_resident()
    assert calls == []


@pytest.mark.parametrize('name', ['ser16_rot', 'ser8_rot_flip', 'ser16_norot'])
def test_all_video_reader_is_device_resident_and_matches_the_reference(name, tmp_path):
    """f3: all_video_reader (reference video_reader.py:129-158, the spectral analyser's reader) keeps the scan in
    HBM: same attributes, `frames` slices and `means` as the reference's in-RAM reader; compute_mean_return_fit
    and repeated reset() + read_video_improved at new shifts (spectralAnalyserUI.py:155-175, 345-346) give the
    reference's results without the file crossing PCIe again."""
    import time
    from oracle import ref_shim
    from oracle import shg_oracle as O
    from solex_ser_recon_en_b200 import solex_util, video_reader
    from solex_ser_recon_en_b200.engine import get_engine
    from helpers import case_stack
    g = golden(name)
    _, stack = case_stack(name)
    path = case_file(name, tmp_path)
    eng = get_engine()
    n_ingests = len(eng.ingest_log)
    rdr = video_reader.all_video_reader(path)
    assert len(eng.ingest_log) == n_ingests + 1
    want_frames = np.stack([O.orient(f) for f in stack])
    assert rdr.frames.shape == want_frames.shape and (rdr.ih, rdr.iw) == want_frames.shape[1:]
    assert np.array_equal(rdr.frames[3:9, :, :], want_frames[3:9])
    assert np.array_equal(rdr.frames[-1], want_frames[-1])
    assert np.array_equal(np.asarray(rdr.frames), want_frames)
    np.testing.assert_array_equal(rdr.means, want_frames.reshape(len(want_frames), -1).mean(axis=1))
    if ref_shim.available():                                    # the reference's own reader on the same file
        ref = ref_shim.load().video_reader.all_video_reader(path)
        assert np.array_equal(ref.frames, want_frames) and np.array_equal(ref.means, rdr.means)
        assert (ref.ih, ref.iw, int(ref.FrameCount)) == (rdr.ih, rdr.iw, int(rdr.FrameCount))
    assert rdr.has_frames() and np.array_equal(rdr.next_frame(), want_frames[0]) and rdr.FrameIndex == 0
    opt = options_for(name, tmp_path, _nolog=True)
    mean_img, fit, y1, y2 = solex_util.compute_mean_return_fit(rdr, opt, {}, rdr.iw, rdr.ih, '')
    assert np.array_equal(mean_img, g['mean_img']) and (y1, y2) == (int(g['y1']), int(g['y2']))
    lat = []
    for shifts in ([10], [0], [-7, 3, 12], [1]):
        opt['shift'] = shifts
        rdr.reset()
        eng.sync()
        t0 = time.perf_counter()
        disks, ih, iw, n = solex_util.read_video_improved(rdr, fit, opt)
        eng.sync()
        lat.append((time.perf_counter() - t0) * 1e3 / len(shifts))
        ref_disks = O.recon(stack, fit, shifts)
        for i in range(len(shifts)):
            assert np.array_equal(np.asarray(disks[i]), ref_disks[i]), shifts[i]
    assert len(eng.ingest_log) == n_ingests + 1                 # the scan crossed PCIe once
    assert min(lat) < 5.0, lat                                  # ms per shift on these tiny scans (launch-bound)


@pytest.mark.parametrize('shape', [(200, 300), (201, 303), (416, 531), (1024, 1221)])
def test_device_tail_equals_opencv_clahe_and_numpy_rescales(shape, tmp_path):
    """f2: image_process on a DeviceImage (CLAHE + percentile rescales on the GPU, csrc/tail.cu) against the host
    path that calls cv2.createCLAHE / np.percentile like the reference (solex_util.py:527-546): identical images,
    for even and odd sizes (OpenCV extends odd images by reflection before cutting tiles)."""
    import cv2
    import torch
    from solex_ser_recon_en_b200 import solex_util
    from solex_ser_recon_en_b200.device_image import DeviceImage
    from solex_ser_recon_en_b200.engine import get_engine
    eng = get_engine()
    rng = np.random.default_rng(shape[0])
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    rho2 = ((yy - shape[0] / 2) / (0.4 * shape[0])) ** 2 + ((xx - shape[1] / 2) / (0.42 * shape[1])) ** 2
    img = np.where(rho2 < 1, 30000 * np.sqrt(np.clip(1 - 0.6 * rho2, 0, 1)), 600.0) + rng.normal(0, 50, shape) + 300
    img = np.clip(img, 0, 65535).astype(np.uint16)
    img[5, 7] = 65535                                         # a hot pixel: the 99.9999 percentile interpolates to it
    opt = options_for('ser16_rot', tmp_path, clahe_only=False)
    opt.pop('_nolog', None)
    cercle = (shape[1] / 2.0, shape[0] / 2.0, 0.4 * shape[0])
    out_h, out_d = os.path.join(str(tmp_path), 'host'), os.path.join(str(tmp_path), 'dev')
    os.makedirs(out_h), os.makedirs(out_d)
    cc_h, pr_h = solex_util.image_process(img, cercle, dict(opt, output_dir=out_h), {}, 'tail_shift=0')
    t = torch.from_numpy(img.view(np.int16)).to(eng.device).view(torch.uint16)
    cc_d, pr_d = solex_util.image_process(DeviceImage(eng, t), cercle, dict(opt, output_dir=out_d), {}, 'tail_shift=0')
    assert np.array_equal(cc_d, cc_h) and np.array_equal(pr_d, pr_h)
    assert np.array_equal(cc_h, solex_util.rescale_brightness(
        cv2.createCLAHE(clipLimit=0.8, tileGridSize=(2, 2)).apply(img), np.percentile(
            cv2.createCLAHE(clipLimit=0.8, tileGridSize=(2, 2)).apply(img), 10),
        np.max(cv2.createCLAHE(clipLimit=0.8, tileGridSize=(2, 2)).apply(img))))
    for name in ('clahe', 'protus', 'uncontrasted', 'high_contrast'):
        a = cv2.imread(os.path.join(out_h, 'tail_shift=0_%s.png' % name), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(os.path.join(out_d, 'tail_shift=0_%s.png' % name), cv2.IMREAD_UNCHANGED)
        assert a is not None and b is not None and np.array_equal(a, b), name
