"""CPU tests: the oracle restatement against the fixtures recorded from the
unmodified reference (oracle/make_golden.py), and its cv2.blur rounding model
against cv2 itself."""
import math

import cv2
import numpy as np
import pytest

from oracle import shg_oracle as O
from oracle.make_golden import CASES
from helpers import ALL_CASES, ELLIPSE_CASES, case_stack, golden


@pytest.mark.parametrize('name', ALL_CASES)
def test_mean_max_bit_exact(name):
    g = golden(name)
    _, stack = case_stack(name)
    mean_img, max_img = O.mean_max(stack)
    assert mean_img.dtype == np.uint16 and mean_img.shape == (int(g['ih']), int(g['iw']))
    assert np.array_equal(mean_img, g['mean_img'])
    assert np.array_equal(max_img, g['max_img'])


@pytest.mark.parametrize('name', ALL_CASES)
def test_detection_and_fit(name):
    g = golden(name)
    _, stack = case_stack(name)
    lf = O.mean_and_fit(stack)
    assert (lf['y1'], lf['y2']) == (int(g['y1']), int(g['y2']))
    y1, y2 = lf['y1'], lf['y2']
    # integer line indices: bit-exact
    assert np.array_equal(lf['min_intensity'][y1:y2], g['polyfit0_y'])
    assert np.array_equal(lf['min_sharp'], g['min_sharp'])
    assert np.array_equal(lf['min_intensity'][y1:y2][lf['keep']], g['polyfit1_y'])
    assert np.array_equal(lf['min_sharp'][y1:y2][lf['mask_good']], g['polyfit2_y'])
    for k, key in enumerate(('p1', 'p2', 'p3')):
        np.testing.assert_allclose(lf[key], g[f'polyfit{k}_p'], rtol=1e-6, atol=0)
    assert np.array_equal(lf['fit'][:, 0], g['fit'][:, 0])
    np.testing.assert_allclose(lf['fit'], g['fit'], rtol=0, atol=1e-9)


@pytest.mark.parametrize('name', ALL_CASES)
def test_recon_bit_exact(name):
    g = golden(name)
    _, stack = case_stack(name)
    shifts = [int(s) for s in g['shift']]
    assert shifts == O.shift_list([int(s) for s in g['shift_requested']])
    disks = O.recon(stack, g['fit'], shifts)
    for i, d in enumerate(disks):
        if CASES[name]['flip_x']:
            d = np.flip(d, axis=1)
        assert np.array_equal(d, g[f'disk{i}']), f'shift {shifts[i]}'


@pytest.mark.parametrize('name', ALL_CASES)
def test_warp_rows_matches_2d_warp(name):
    g = golden(name)
    phi = 0.0 if math.isnan(float(g['slant'])) else math.radians(float(g['slant']))
    ratio = float(g['ratio'])
    shifts = [int(s) for s in g['shift']]
    for sh in (int(s) for s in g['shift_requested']):
        disk = g[f'disk{shifts.index(sh)}']
        out, _ = O.warp_rows(disk, phi, ratio)
        assert out.shape == g[f'circ_{sh}'].shape
        assert np.array_equal(out, g[f'circ_{sh}'])


@pytest.mark.parametrize('name', ALL_CASES)
def test_transversalium(name):
    g = golden(name)
    cercle = tuple(float(v) for v in g['cercle'])
    for sh in (int(s) for s in g['shift_requested']):
        circ = g[f'circ_{sh}']
        if cercle == (-1.0, -1.0, -1.0):
            circle, borders = (0, 0, 99999), [0, int(g['y1']) + 20, circ.shape[1] - 1, int(g['y2']) - 20]
        else:
            circle, borders = cercle, list(g['borders'])
        det, gain = O.correct_transversalium(circ, circle, borders)
        np.testing.assert_allclose(gain, g[f'gain_{sh}'], rtol=1e-12)
        assert np.array_equal(det, g[f'det_{sh}'])


@pytest.mark.parametrize('name', ELLIPSE_CASES)
def test_ellipse_fit(name):
    g = golden(name)
    fix, circle, ratio, phi, borders = O.ellipse_to_circle(g['disk0'])
    np.testing.assert_allclose(ratio, float(g['ratio']), rtol=1e-9)
    np.testing.assert_allclose(math.degrees(phi), float(g['slant']), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(circle, g['cercle'], rtol=1e-9)
    np.testing.assert_allclose(borders, g['borders'], rtol=1e-9)


def test_ellipse_comes_out_circular():
    """Acceptance check for the unpinned LsqEllipse / warp restatements: the
    synthetic Sun (semi-axes 0.42 N frames x 0.40 W pixels) becomes a circle."""
    g = golden('ser16_rot')
    spec = CASES['ser16_rot']['spec']
    truth = 0.40 * spec['width'] / (0.42 * spec['n_frames'])
    assert abs(float(g['ratio']) / truth - 1) < 0.03
    cx, cy, rad = g['cercle']
    assert abs(rad / (0.40 * spec['width']) - 1) < 0.03
    circ = g['circ_0']
    assert circ.shape[0] == spec['width']
    # measured extent of the bright disk along both axes agrees
    bright = circ > 0.25 * circ.max()
    ext_y = np.ptp(np.nonzero(bright.any(axis=1))[0])
    ext_x = np.ptp(np.nonzero(bright.any(axis=0))[0])
    assert abs(ext_x / ext_y - 1) < 0.04


def test_blur_model_matches_cv2():
    rng = np.random.default_rng(0)
    for trial in range(60):
        H = int(rng.integers(60, 300)); W = int(rng.integers(40, 260))
        kh = int(rng.integers(1, 40)); kw = 25 if trial % 3 else int(rng.integers(1, 30))
        if trial % 5 == 0:
            kw, kh = 5, 5
        hi = [65535, 4000, 300][trial % 3]
        img = rng.integers(0, hi + 1, size=(H, W)).astype(np.uint16)
        assert np.array_equal(O.box_blur_u16(img, kw, kh), cv2.blur(img, ksize=(kw, kh))), (H, W, kw, kh)


def test_mean_floor_division_claim():
    """(sum / N).astype(uint16) == sum // N for the N range of interest."""
    rng = np.random.default_rng(1)
    for n in (1, 3, 7, 999, 1000, 20000, 65537, 10 ** 6):
        q = rng.integers(0, 65536, size=2000, dtype=np.uint64)
        r = rng.integers(0, n, size=2000, dtype=np.uint64)
        r[:50] = n - 1
        s = q * np.uint64(n) + r
        assert np.array_equal((s / np.uint32(n)).astype('uint16'), (s // np.uint64(n)).astype('uint16'))


def test_partial_sums_combine_exactly():
    _, stack = case_stack('ser16_rot')
    s_all, m_all = O.raw_sum_max(stack)
    s0, m0 = O.raw_sum_max(stack, 0, 77)
    s1, m1 = O.raw_sum_max(stack, 77, stack.shape[0])
    assert np.array_equal(s0 + s1, s_all) and np.array_equal(np.maximum(m0, m1), m_all)
