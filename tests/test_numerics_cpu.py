"""CPU models of the in-kernel logarithms of csrc/transv.cu (log_u16, log_ratio_series) -- the error bounds quoted in
the kernel comments and in DESIGN.md, checked in extended precision -- and of its two selection schemes, without a GPU.  The GPU tests
(test_log_u16_matches_numpy, test_row_stats_single_pixel_chords_give_log_ratio) check the kernels themselves."""
import numpy as np

L = np.longdouble


def _fma(a, b, c):
    """a*b + c rounded once to fp64 (80-bit intermediate: exact enough for these magnitudes)."""
    return (a.astype(L) * b.astype(L) + c).astype(np.float64)


def test_log_u16_model_is_within_one_ulp_of_log_65535():
    v = np.arange(1, 65536, dtype=np.uint32)
    e = np.floor(np.log2(v.astype(np.float64))).astype(np.int64)
    m16 = (v << (15 - e).astype(np.uint32)).astype(np.uint32)
    assert m16.min() >= 32768 and m16.max() < 65536
    h = np.arange(128)
    c = 1.0 / (1.0 + (h + 0.5) / 128.0)                       # the 128-segment table: {c / 2^15, -log(c)}
    t = (-np.log(c.astype(L))).astype(np.float64)
    idx = (m16 >> 8) & 127
    r = _fma(m16.astype(np.float64), (c / 32768.0)[idx], L(-1.0))
    assert np.abs(r).max() <= 2.0 ** -8                       # the degree-6 log1p remainder is then < 2e-18
    p = _fma(r, np.full_like(r, -1.0 / 6.0), L(0.2))
    p = _fma(r, p, L(-0.25))
    p = _fma(r, p, L(1.0 / 3.0))
    p = _fma(r, p, L(-0.5))
    ed = e.astype(np.float64)
    hi = _fma(ed, np.full_like(r, 6.93147180369123816490e-01), t[idx].astype(L))
    lo = _fma(ed, np.full_like(r, 1.90821492927058770002e-10), _fma(r * r, p, r.astype(L)).astype(L))
    got = hi + lo
    err = np.abs(got.astype(L) - np.log(v.astype(L))).astype(np.float64)
    assert err.max() <= 2.0e-15                               # ~1 ulp at log(65535) = 11.09


def test_log_ratio_series_model_is_a_few_ulp_of_the_result():
    rng = np.random.default_rng(3)
    a = rng.integers(1, 65536, 200000).astype(np.int64)
    b = np.clip(a + rng.integers(-2000, 2001, a.size), 1, 65535)
    small = (np.abs(a - b) << 6) <= (a + b)                   # the kernel's test for |z| <= 2^-6
    a, b = a[small], b[small]
    assert a.size > 50000
    z = ((a - b).astype(L) / (a + b).astype(L)).astype(np.float64)    # (the kernel's Newton division is <= 1 ulp)
    z2 = z * z
    p = _fma(z2, np.full_like(z, 1.0 / 9.0), L(1.0 / 7.0))
    p = _fma(z2, p, L(0.2))
    p = _fma(z2, p, L(1.0 / 3.0))
    got = 2.0 * _fma(z * z2, p, z.astype(L))
    true = np.log1p((a - b).astype(L) / b.astype(L))
    err = np.abs(got.astype(L) - true).astype(np.float64)
    nz = got != 0
    assert np.all(err[nz] <= 3 * np.spacing(np.abs(got[nz])))
    assert np.all(got[~nz] == 0) and np.all(true[~nz] == 0)
    # the reference's own fp64 log(a/b) carries the rounding of the quotient: ~1e-16 absolute
    ref = np.log(a.astype(np.float64) / b.astype(np.float64))
    assert np.abs(got - ref).max() <= 2.3e-16


def _counting_select_model(x, t0, t1, qlo, qhi, kind, med=0.0, bins=1024):
    """NumPy model of hist_select (csrc/transv.cu): one-level counting select with arbitrary bin edges."""
    keys = x if kind == 0 else np.abs(x - med)
    iqr = qhi - qlo
    lo = qlo - 4.0 * iqr if kind == 0 else 0.0
    scale = (bins - 2) / (9.0 * iqr) if kind == 0 else (bins - 1) / (4.0 * iqr)
    b = np.floor((keys - lo) * scale).astype(np.int64) + (1 if kind == 0 else 0)
    b = np.clip(b, 0, bins - 1)
    counts = np.bincount(b, minlength=bins)
    cum = np.concatenate([[0], np.cumsum(counts)])
    b0 = int(np.searchsorted(cum, t0, side='right') - 1)
    b1 = int(np.searchsorted(cum, t1, side='right') - 1)
    cand = np.sort(keys[(b >= b0) & (b <= b1)])               # the kernel ranks these exactly in fp64
    below = cum[b0]
    return cand[t0 - below], cand[t1 - below], cand.size


def test_counting_select_model_is_exact_for_any_bin_edges():
    """Monotone binning + exact ranking inside the target bins gives the exact order statistics whatever
    the sample quartiles were -- good edges only keep the candidate list short."""
    rng = np.random.default_rng(8)
    for trial in range(60):
        n = int(rng.integers(257, 4000))
        x = rng.normal(0.001, 0.01, n)
        x[rng.integers(0, n, 30)] = rng.uniform(-2.5, 2.5, 30)            # limb pixels
        if trial % 3 == 0:
            x = np.round(x, 3)                                             # heavy ties
        if trial % 2:                                                      # quartiles of a 32-element sample ...
            s = np.sort(x[((2 * np.arange(32) + 1) * n) >> 6])
            qlo, qhi = s[8], s[23]
        else:                                                              # ... or nonsense edges
            qlo, qhi = sorted(rng.uniform(-1, 1, 2))
        if not qhi - qlo >= 1e-5:
            continue
        t0, t1 = (n - 1) // 2, n // 2
        srt = np.sort(x)
        a0, a1, m = _counting_select_model(x, t0, t1, qlo, qhi, 0)
        assert (a0, a1) == (srt[t0], srt[t1])
        med = (a0 + a1) / 2.0
        d = np.sort(np.abs(x - med))
        b0, b1, _ = _counting_select_model(x, t0, t1, qlo, qhi, 1, med)
        assert (b0, b1) == (d[t0], d[t1])
        if trial % 2 and trial % 3:
            assert m <= 256                                                # representative edges: short candidate list


def _window_select_model(keys, lo, hi, t0, t1, cap=128):
    """NumPy model of window_select (transv_row_stats_reg_kernel): count the keys below [lo, hi), list those
    inside, verify that both middle ranks are in the list, rank the list exactly.  None = row handed back."""
    below = int((keys < lo).sum())
    cand = np.sort(keys[(keys >= lo) & (keys < hi)])
    if cand.size > cap or t0 < below or t1 >= below + cand.size:
        return None
    return cand[t0 - below], cand[t1 - below], cand.size


def _reg_kernel_model(x, qlo, qhi, bins=2048):
    """Median and MAD as the register-resident row-statistics kernel finds them: histogram of the values only,
    value window of the middle bins for the median, distance window for the MAD from the SAME histogram's prefix
    table (no second histogram).  Returns (med, mad, median list length, MAD list length) or None."""
    n = x.size
    t0, t1 = (n - 1) // 2, n // 2
    iqr = qhi - qlo
    lo, scale, w = qlo - 3.0 * iqr, (bins - 2) / (7.0 * iqr), iqr * (7.0 / (bins - 2))
    b = np.clip(np.floor((x - lo) * scale + 1.0).astype(np.int64), 0, bins - 1)
    P = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=bins))])      # P[i] = #values in bins < i
    b0 = int(np.searchsorted(P, t0, side='right') - 1)
    b1 = int(np.searchsorted(P, t1, side='right') - 1)
    elo = -np.inf if b0 <= 0 else lo + (b0 - 1) * w
    ehi = np.inf if b1 >= bins - 1 else lo + b1 * w
    got = _window_select_model(x, elo, ehi, t0, t1)
    if got is None:
        return None
    med = got[0] if t0 == t1 else (got[0] + got[1]) / 2.0
    fl = int(np.floor(min(max((med - lo) * scale + 1.0, -4096.0), 8192.0)))

    def Pc(i):
        return int(P[min(max(i, 0), bins)])

    def upper(k):                      # every |r - med| < k w lies in bins fl-k .. fl+k
        return 0 if k == 0 else Pc(fl + k + 1) - Pc(fl - k)

    def lower(k):                      # all of the regular bins fl-k+1 .. fl+k-1 are that close
        hi_, lo_ = min(max(fl + k, 1), bins - 1), min(max(fl - k + 1, 1), bins - 1)
        return Pc(hi_) - Pc(lo_) if hi_ > lo_ else 0

    d = np.abs(x - med)
    for k in (0, 1, 2, 5, 17, 100, 700, 2047):                                # the brackets hold for every k
        true = int((d < k * w).sum())
        assert lower(k) <= true <= upper(k), (k, lower(k), true, upper(k))
    ks = np.arange(bins)
    k_lo = int(ks[[upper(int(k)) <= t0 for k in ks]].max())
    ok = [lower(int(k)) > t1 for k in ks]
    if not any(ok):
        return None
    k_hi = int(ks[ok].min())
    assert k_hi > k_lo
    got2 = _window_select_model(d, k_lo * w, k_hi * w, t0, t1)
    if got2 is None:
        return None
    mad = got2[0] if t0 == t1 else (got2[0] + got2[1]) / 2.0
    return med, mad, got[2], got2[2]


def test_window_select_with_histogram_brackets_is_exact():
    """The median's histogram brackets #{|r - med| < k bins} from both sides for every k, so a distance window
    [k_lo w, k_hi w) that holds both middle ranks of |r - med| follows from it without a second counting pass;
    counts taken from the data make the result exact whatever the bins are (or the row is handed back)."""
    rng = np.random.default_rng(12)
    taken = 0
    for trial in range(80):
        n = int(rng.integers(257, 4000))
        x = rng.normal(0.001, 0.01, n)
        x[rng.integers(0, n, 30)] = rng.uniform(-2.5, 2.5, 30)            # limb pixels
        if trial % 4 == 0:
            x = np.round(x, 3)                                             # heavy ties: lists overflow
        if trial % 5 == 4:
            qlo, qhi = sorted(rng.uniform(-0.05, 0.05, 2))                 # a sample that misjudges the row
        else:
            s = np.sort(x[((2 * np.arange(32) + 1) * n) >> 6])
            qlo, qhi = s[8], s[23]
        if not qhi - qlo >= 1e-5:
            continue
        got = _reg_kernel_model(x, qlo, qhi)
        if got is None:                                                    # handed back to the classic kernel
            continue
        taken += 1
        med, mad, m1, m2 = got
        assert med == np.median(x)
        assert mad == np.median(np.abs(x - np.median(x)))
        if trial % 4 and trial % 5 != 4:
            assert m1 <= 32 and m2 <= 96                                   # representative bins: short lists
    assert taken >= 50


def test_limb_selection_on_the_sparse_list_equals_the_image_pipeline():
    """ellipse_fit.select_limb_pixels (hysteresis, two largest regions, hull test, crop -- on the sparse list
    of thin-edge pixels the GPU returns) gives the same points as limb_points() working on whole images
    with scipy.ndimage.label, for a tilted ellipse, a Sun cut by the frame, and two separate arcs."""
    import cv2
    from solex_ser_recon_en_b200 import ellipse_fit as E
    rng = np.random.default_rng(4)
    for case in range(4):
        rows, cols = 160 + 30 * case, 420 - 40 * case
        yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float64)
        cy, cx = rows * (0.5 if case != 1 else 0.30), cols * 0.5
        t = 0.1 * case
        u = (yy - cy) * np.cos(t) + (xx - cx) * np.sin(t)
        v = -(yy - cy) * np.sin(t) + (xx - cx) * np.cos(t)
        rho2 = (u / (0.40 * rows)) ** 2 + (v / (0.42 * cols)) ** 2
        img = np.where(rho2 < 1, 30000 * np.sqrt(np.clip(1 - 0.6 * rho2, 0, 1)), 500.0)
        if case == 3:
            img[:, cols // 2 - 6:cols // 2 + 6] = 500.0             # a dark lane: the limb breaks into two arcs
        img = (img + 300 + rng.normal(0, 15, img.shape)) / 65536
        want_pts, want_raw = E.limb_points(img)
        flood = E.flood_image(img)
        low = np.median(cv2.blur(img, ksize=(5, 5))) / 10
        flat, mag = E.thin_edges(flood, 2.0, low)
        got_pts, got_raw = E.select_limb_pixels(flat, mag, low * 1.5, cols)
        assert np.array_equal(got_raw, want_raw), case
        assert np.array_equal(got_pts, want_pts), case
        assert len(got_pts) > 50


def test_vectorised_chords_equal_the_reference_loop():
    """Engine.transversalium_chords (vectorised) against the reference's per-row expressions
    (solex_util.py:384-391), including circles whose chords end exactly on integers."""
    import math
    from solex_ser_recon_en_b200.engine import Engine

    def loop(circle, borders):
        y1 = math.ceil(max(circle[1] - circle[2], borders[1]))
        y2 = math.floor(min(circle[1] + circle[2], borders[3]))
        out = []
        for y in range(y1 + 1, y2):
            dx = math.floor((circle[2] ** 2 - (y - circle[1]) ** 2) ** 0.5)
            out.append((y, math.ceil(max(circle[0] - dx, borders[0])), math.floor(min(circle[0] + dx, borders[2]))))
        return y1, y2, np.asarray(out, dtype=np.int64).reshape(-1, 3)

    rng = np.random.default_rng(11)
    cases = [((1951.93, 2047.5, 1638.1), [313.02, 460.0, 3590.84, 3624.0]),
             ((500.0, 400.0, 250.0), [0.0, 0.0, 1000.0, 800.0]),               # integer geometry: 3-4-5 rows hit integers
             ((0, 0, 99999), [0, 120, 1220, 1100]),                            # the no-ellipse fallback
             ((100.0, 100.0, 5.0), [0, 300, 200, 310])]                        # empty row range
    for _ in range(40):
        r = rng.uniform(20, 3000)
        c = (rng.uniform(0, 4000), rng.uniform(0, 4000), r)
        b = [c[0] - r * rng.uniform(0.5, 1.2), c[1] - r * rng.uniform(0.5, 1.2), c[0] + r * rng.uniform(0.5, 1.2),
             c[1] + r * rng.uniform(0.5, 1.2)]
        cases.append((c, b))
    for circle, borders in cases:
        y1, y2, want = loop(circle, borders)
        g1, g2, rows, xa, xb = Engine.transversalium_chords(circle, borders)
        assert (g1, g2) == (y1, y2)
        assert np.array_equal(rows, want[:, 0]) and np.array_equal(xa, want[:, 1]) and np.array_equal(xb, want[:, 2])
