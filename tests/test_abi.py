"""CPU: the C-ABI library loads and exports every symbol include/shg.h declares
(no compute calls -- those need a GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'shg.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(shg_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from solex_ser_recon_en_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(_lib.lib, n), n
    assert sorted(_lib.EXPORTS) == names, 'ctypes prototypes and shg.h disagree'
    assert _lib.lib.shg_version() == 1
    assert _lib.lib.shg_last_error() == b''


def test_every_declaration_cites_the_reference():
    src = open(os.path.join(ROOT, 'include', 'shg.h')).read()
    for ref in ('solex_util.py:174-188', 'solex_util.py:93-144', 'ellipse_to_circle.py:94-118',
                'video_reader.py:94-123', 'solex_util.py:233-259', 'solex_util.py:76-86'):
        assert ref in src, ref


def test_engine_fails_loudly_without_a_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from solex_ser_recon_en_b200.engine import Engine
    from solex_ser_recon_en_b200._lib import ShgError
    with pytest.raises(ShgError):
        Engine(0)


def test_label_points_matches_scipy_label():
    """The host-side labelling helper (no GPU needed) against scipy.ndimage.label."""
    import numpy as np
    from scipy import ndimage as ndi
    from solex_ser_recon_en_b200 import ellipse_fit as E
    rng = np.random.default_rng(0)
    for t in range(100):
        rows, cols = int(rng.integers(1, 40)), int(rng.integers(1, 50))
        m = rng.random((rows, cols)) < rng.uniform(0.05, 0.7)
        lab, cnt = ndi.label(m, np.ones((3, 3)))
        flat = np.flatnonzero(m)
        c2, l2 = E._components(flat, cols)
        assert c2 == cnt and np.array_equal(l2, lab.ravel()[flat]), t


def test_hull_vertices_match_qhull():
    """The host-side convex hull helper against scipy.spatial.ConvexHull on integer point sets: same vertex
    SET (collinear points on an edge are not vertices in either), including limb-like thin rings."""
    import numpy as np
    from scipy.spatial import ConvexHull
    from solex_ser_recon_en_b200 import ellipse_fit as E
    rng = np.random.default_rng(1)
    for t in range(120):
        n = int(rng.integers(3, 400))
        if t % 3 == 0:                                   # ring of pixels around an ellipse
            ang = rng.uniform(0, 2 * np.pi, n)
            pts = np.stack([200 + np.rint(150 * np.cos(ang)), 300 + np.rint(260 * np.sin(ang))], axis=1)
        elif t % 3 == 1:                                 # small grid: many collinear and duplicate points
            pts = rng.integers(0, 7, size=(n, 2)).astype(float)
        else:
            pts = rng.integers(-1000, 1000, size=(n, 2)).astype(float)
        pts = pts.astype(np.int64)
        try:
            want = set(map(tuple, pts[ConvexHull(pts).vertices]))
        except Exception:
            continue                                     # Qhull refuses degenerate input
        got = E._hull_vertices(pts[:, 0], pts[:, 1])
        assert set(map(tuple, pts[got])) == want, t
        assert len(got) == len(want), t
