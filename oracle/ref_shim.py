"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference modules from
/root/reference in this container, with the non-numeric packages it imports
(matplotlib, astropy, tkinter, FreeSimpleGUI) replaced by inert stubs and the
two absent numeric packages (scikit-image, lsq-ellipse) replaced by the
restatements in oracle/thirdparty.py (parity unpinned for those call sites).

/root/reference does not exist on the GPU box.  There the loader falls back to
oracle/_ref/reference_modules.zip -- the five modules, byte for byte, archived by
oracle/stage_ref.py in a git-ignored directory that travels with the snapshot --
so that bench.py's reference arm can time the unmodified reference on the box's
host cores.  Used by oracle/make_golden.py (fixture generation, run here), by
bench.py --impl reference, and by tests that are skipped when neither tree exists.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'reference_modules.zip')
REFERENCE_DIR = os.environ.get('SHG_REFERENCE_DIR', '/root/reference')
if not os.path.isfile(os.path.join(REFERENCE_DIR, 'solex_util.py')) and os.path.isfile(_STAGED):
    REFERENCE_DIR = _STAGED                      # a zip archive on sys.path: Python imports the modules from it


def available() -> bool:
    return os.path.isfile(REFERENCE_DIR) or os.path.isfile(os.path.join(REFERENCE_DIR, 'solex_util.py'))


def source() -> str:
    """'tree' (/root/reference itself), 'staged' (the oracle/_ref archive) or 'absent'."""
    if not available():
        return 'absent'
    return 'staged' if os.path.isfile(REFERENCE_DIR) else 'tree'


class _Anything:
    """Callable, attribute-able, context-manageable do-nothing."""
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _Anything()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything


class _Header(dict):
    pass


class _PrimaryHDU:
    written = []

    def __init__(self, data=None, header=None):
        self.data = data
        self.header = header

    def writeto(self, path, overwrite=False):
        _PrimaryHDU.written.append(path)


def _install_stubs():
    from . import thirdparty
    def stub(name, **attrs):
        if name in sys.modules and not isinstance(sys.modules[name], _StubModule):
            return sys.modules[name]
        m = _StubModule(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    for name in ('matplotlib', 'matplotlib.figure', 'matplotlib.pyplot', 'matplotlib.patches',
                 'tkinter', 'FreeSimpleGUI', 'PIL', 'PIL.Image', 'PIL.ImageTk'):
        try:
            importlib.import_module(name)
        except Exception:
            stub(name)
    mpl = sys.modules['matplotlib']
    for sub in ('figure', 'pyplot', 'patches'):
        if isinstance(mpl, _StubModule):
            setattr(mpl, sub, sys.modules['matplotlib.' + sub])
    try:
        importlib.import_module('astropy.io.fits')
    except Exception:
        fits = stub('astropy.io.fits', Header=_Header, PrimaryHDU=_PrimaryHDU)
        io = stub('astropy.io', fits=fits)
        stub('astropy', io=io)
    try:
        importlib.import_module('skimage')
    except Exception:
        tr = stub('skimage.transform', warp=thirdparty.warp,
                  ProjectiveTransform=thirdparty.ProjectiveTransform,
                  downscale_local_mean=thirdparty.downscale_local_mean)
        ft = stub('skimage.feature', canny=thirdparty.canny)
        fl = stub('skimage.filters')
        fe = stub('skimage.data._fetchers')
        da = stub('skimage.data', _fetchers=fe)
        stub('skimage', transform=tr, feature=ft, filters=fl, data=da)
    try:
        importlib.import_module('ellipse')
    except Exception:
        stub('ellipse', LsqEllipse=thirdparty.LsqEllipse)


_loaded = {}


def load():
    """Return a namespace with the reference modules
    (video_reader, solex_util, ellipse_to_circle, Solex_recon, CLI_handler)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_DIR)
    _install_stubs()
    # The product ships drop-in modules with the same names; make sure the
    # reference's own files win while we import them, then restore.
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in
                  ('video_reader', 'solex_util', 'ellipse_to_circle', 'Solex_recon', 'CLI_handler')
                  if k in sys.modules}
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import scipy.ndimage
        if not hasattr(scipy.ndimage, 'measurements'):
            scipy.ndimage.measurements = scipy.ndimage
        for name in ('video_reader', 'solex_util', 'ellipse_to_circle', 'Solex_recon', 'CLI_handler'):
            _loaded[name] = importlib.import_module(name)
    finally:
        sys.path[:] = saved_path
        for k in ('video_reader', 'solex_util', 'ellipse_to_circle', 'Solex_recon', 'CLI_handler'):
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)
    return types.SimpleNamespace(**_loaded)


def default_options(**over):
    """The reference's default options dict (/root/reference/SHG_MAIN.py:41-68)
    with plotting and logging silenced."""
    o = {
        'language': 'English', 'shift': [0], 'flag_display': False, 'ratio_fixe': None,
        'slant_fix': None, 'save_fit': False, 'clahe_only': True, 'protus_only': False,
        'disk_display': True, 'delta_radius': 0, 'crop_width_square': False,
        'transversalium': True, 'stubborn_transversalium': False, 'trans_strength': 301,
        'img_rotate': 0, 'flip_x': False, 'workDir': '', 'fixed_width': None,
        'output_dir': '', 'input_dir': '', 'specDir': '', 'selected_mode': 'File input mode',
        'continuous_detect_mode': False, 'dispersion': 0.05, 'ellipse_fit_shift': 10,
        'de-vignette': False, '_nolog': True,
    }
    o.update(over)
    return o
