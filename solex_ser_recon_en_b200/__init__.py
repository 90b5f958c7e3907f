"""B200-native frame-stack reconstruction for Sol'Ex spectroheliograph scans:
a drop-in for the reconstruction path of thelondonsmiths/Solex_ser_recon_EN.

The pixel work lives in libshg.so (hand-written CUDA for sm_100a, C ABI in
include/shg.h); the modules here keep the reference's Python entry points:

    video_reader, solex_util, Solex_recon, ellipse_to_circle, CLI_handler, SHG_MAIN

install_aliases() registers them under the reference's top-level module names
so that existing code (`import Solex_recon`) picks up this implementation.
"""
from __future__ import annotations

import importlib
import sys

__version__ = '0.1.0'

REFERENCE_MODULES = ('video_reader', 'solex_util', 'ellipse_to_circle', 'Solex_recon', 'CLI_handler', 'SHG_MAIN')


def install_aliases():
    """Make `import Solex_recon` (etc.) resolve to this package's modules."""
    for name in REFERENCE_MODULES:
        sys.modules[name] = importlib.import_module(__name__ + '.' + name)
    return [sys.modules[n] for n in REFERENCE_MODULES]
