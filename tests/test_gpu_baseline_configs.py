"""GPU: the BASELINE.json configurations at (or near) their real sizes.

Config 1 and config 2 run at full size through the drop-in entry points and
are compared with the oracle on the same file (seconds of NumPy); configs 3-5
are too big for the CPU oracle, so they are checked through size-independent
properties of a device-synthesised scan: linearity of the integer sums over
frame ranges, bit-exact disks on sampled frames, and shard-invariance."""
import math
import os

import numpy as np
import pytest

from oracle import ref_shim
from oracle import shg_oracle as O

pytestmark = pytest.mark.gpu


def _options(tmp_path, **over):
    o = ref_shim.default_options(**over)
    o['output_dir'] = str(tmp_path)
    return o


@pytest.mark.parametrize('cfg', [1, 2])
def test_config_full_size_against_oracle(cfg, tmp_path):
    """configs[0]: 16-bit SER 1000 x 1280x200, shift 0, ellipse fit + transversalium;
    configs[1]: 8-bit AVI 2000 x 1920x256 (reduced to 600 frames to keep the CPU oracle in seconds), -ms."""
    from solex_ser_recon_en_b200 import Solex_recon, synth
    got = {}

    def sink(basefich, image, cercle):
        got[int(basefich.rsplit('_shift=', 1)[1])] = np.asarray(image).copy()

    if cfg == 1:
        spec = synth.halpha(1000, 1280, 200, seed=1, dust=(400, 401, 700))
        path = synth.write_ser(os.path.join(str(tmp_path), 'cfg1.SER'), spec)
        opt = _options(tmp_path, shift=[0], _result_sink=sink)
    else:
        spec = synth.ca_k_8bit(600, 1920, 256, seed=2)
        path = synth.write_avi(os.path.join(str(tmp_path), 'cfg2.avi'), spec)
        opt = _options(tmp_path, shift=[0], flip_x=True, crop_width_square=True, _result_sink=sink)
    disk_list, bounds, hdr = Solex_recon.solex_read(path, opt)
    stack = synth.frames(spec, 0, spec.n_frames)
    disks, shifts, lf = O.solex_read(stack, [0], flip_x=opt['flip_x'])
    assert shifts == opt['shift'] and (lf['y1'], lf['y2']) == tuple(int(b) for b in bounds)
    for i in range(len(shifts)):
        assert np.array_equal(np.asarray(disk_list[i]), disks[i]), shifts[i]          # bit-exact
    Solex_recon.solex_process(opt, disk_list, bounds, hdr)
    want, geom = O.solex_process(disks, shifts, [0], bounds)
    np.testing.assert_allclose(float(opt['ratio_fixe']), geom['ratio'], rtol=1e-9)
    det = got[0]
    ref = want[0][1]
    assert det.shape == ref.shape
    d = np.abs(det.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) < 1e-3                                      # north_star: <= 1 DN


@pytest.mark.parametrize('geom,n_shift', [((2048, 300, 2, 1200), 10), ((4096, 512, 2, 600), 50), ((2048, 256, 2, 900), 0)])
def test_big_geometry_properties(geom, n_shift):
    """Geometries of configs 3 / 5 / 4 (frame counts reduced so the readback stays small):
    sums of frame ranges add up exactly; disks of sampled frames are bit-exact against
    the oracle; reconstructing in two frame shards equals reconstructing at once."""
    import torch
    from solex_ser_recon_en_b200.engine import ScanGeometry, get_engine
    eng = get_engine(0)
    W, H, bpp, N = geom
    g = ScanGeometry(W, H, bpp, N)
    st = eng.synth_stack(g, seed=5)
    eng.accumulate(st)
    total = st.sum.clone()
    # linearity: two half-range stacks accumulate to the same integers
    a = eng.synth_stack(g, k0=0, n=N // 3, seed=5)
    b = eng.synth_stack(g, k0=N // 3, n=N - N // 3, seed=5)
    eng.accumulate(a)
    eng.accumulate(b)
    assert torch.equal(a.sum + b.sum, total) and torch.equal(torch.maximum(a.max, b.max), st.max)
    mean_img, max_img = eng.finalize_mean_max(st.sum, st.max, N, g)
    det = eng.detect_line(mean_img, max_img)
    fit = eng.fit_line(det, g.ih)
    shifts = O.shift_list(list(range(-n_shift, n_shift + 1)))
    disk = eng.recon(st, fit['fit'], shifts)
    # sampled frames against the oracle (bit-exact)
    ks = sorted({0, 1, N // 2, N - 1, N // 3 - 1, N // 3})
    frames = np.stack([st.host_frames(k, k + 1)[0] for k in ks])
    ref = O.recon(frames, fit['fit'], shifts)
    got = disk.view(torch.int16)[:, ks, :].cpu().numpy().view(np.uint16)
    for i in range(len(shifts)):
        assert np.array_equal(got[i].T, ref[i]), shifts[i]
    # shard invariance (what ranks do)
    disk2 = eng.alloc_disk(len(shifts), N, g.ih)
    eng.recon(a, fit['fit'], shifts, disk=disk2)
    eng.recon(b, fit['fit'], shifts, disk=disk2)
    assert torch.equal(disk2.view(torch.int16), disk.view(torch.int16))
    # the mean frame of the device scan agrees with the oracle on a readback of the sums
    mo, xo = O.finalize_mean_max(st.sum.cpu().numpy().view(np.uint64).reshape(H, W),
                                 st.max.cpu().numpy().reshape(H, W).astype(np.uint16), N, False)
    assert np.array_equal(mean_img.cpu().numpy(), mo) and np.array_equal(max_img.cpu().numpy(), xo)


def test_config5_full_size_against_oracle_on_one_shift(tmp_path):
    """BASELINE configs[4] at its real size (20 000 frames x 4096x512, 84 GB in HBM), through the drop-in entry
    points.  The CPU oracle cannot read 84 GB, so: the integer mean frame is checked by linearity against four
    frame-range shards; the disks are checked bit-exactly on sampled frames; and ONE complete disk image is copied
    back and pushed through the oracle's ellipse fit, warp and transversalium for a full-size comparison of the
    post-processing (<= 1 DN, gains 1e-5)."""
    import torch
    from solex_ser_recon_en_b200 import Solex_recon
    from solex_ser_recon_en_b200.engine import ScanGeometry, get_engine
    from solex_ser_recon_en_b200.video_reader import device_scan
    eng = get_engine(0)
    free, _ = torch.cuda.mem_get_info()
    N = 20000 if free > 120e9 else 4000
    g = ScanGeometry(4096, 512, 2, N)
    st = eng.synth_stack(g, seed=5)
    got = {}

    def sink(basefich, image, cercle):
        got[int(basefich.rsplit('_shift=', 1)[1])] = image

    opt = _options(tmp_path, shift=[0, -50, 50], _result_sink=sink)
    disk_list, bounds, hdr = Solex_recon.solex_read_reader(device_scan(st), opt, os.path.join(str(tmp_path), 'cfg5'))
    # pass 1: four shards add up to the same integers
    total = st.sum.clone()
    acc = torch.zeros_like(total)
    for q in range(4):
        part = eng.synth_stack(g, k0=q * N // 4, n=N // 4, seed=5)
        eng.accumulate(part)
        acc += part.sum
        del part
    assert torch.equal(acc, total)
    # pass 2: sampled frames of every disk, bit-exact
    lf_fit = None
    ks = [0, 1, N // 2, N - 1]
    frames = np.stack([st.host_frames(k, k + 1)[0] for k in ks])
    mean_img, max_img = eng.finalize_mean_max(st.sum, st.max, N, g)
    det = eng.detect_line(mean_img, max_img)
    lf_fit = eng.fit_line(det, g.ih)['fit']
    ref = O.recon(frames, lf_fit, opt['shift'])
    for i in range(len(opt['shift'])):
        img = disk_list[i].tensor.view(torch.int16)[ks].cpu().numpy().view(np.uint16)
        assert np.array_equal(img.T, ref[i]), opt['shift'][i]
    # post-processing of the whole set on the GPU ...
    Solex_recon.solex_process(opt, disk_list, bounds, hdr)
    # ... and of shift 0 on the CPU oracle, from the GPU's own disk images
    disk10 = np.asarray(disk_list[0])
    disk0 = np.asarray(disk_list[1])
    _, circle, ratio, phi, borders = O.ellipse_to_circle(disk10)
    np.testing.assert_allclose(float(opt['ratio_fixe']), ratio, rtol=1e-9)
    np.testing.assert_allclose(math.radians(float(opt['slant_fix'])), phi, rtol=1e-6, atol=1e-12)
    circ, _ = O.warp_rows(disk0, phi, ratio)
    det_ref, gain_ref = O.correct_transversalium(circ, circle, borders)
    det_gpu = np.asarray(got[0])
    assert det_gpu.shape == det_ref.shape
    d = np.abs(det_gpu.astype(np.int32) - det_ref.astype(np.int32))
    assert d.max() <= 1 and np.mean(d != 0) < 1e-3
    np.testing.assert_allclose(opt['_transversalium_gains'][0], gain_ref, rtol=1e-5)
