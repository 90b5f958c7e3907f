"""Shared test helpers: golden fixtures and the scans they were made from."""
import hashlib
import os

import numpy as np

from oracle.make_golden import CASES, make_spec
from solex_ser_recon_en_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_cache = {}


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))


def case_stack(name):
    """Raw (N, H, W) payload of a golden case, regenerated from its recipe and
    checked against the sha256 stored with the fixture."""
    if name not in _cache:
        spec = make_spec(CASES[name]['spec'])
        stack = synth.frames(spec, 0, spec.n_frames)
        digest = hashlib.sha256(stack.tobytes()).hexdigest()
        assert digest == str(golden(name)['sha256']), 'synthetic scan differs from the one the fixture was made from'
        _cache[name] = (spec, stack)
    return _cache[name]


def case_file(name, tmpdir):
    """Write the case as a SER/AVI file and return its path."""
    spec, stack = case_stack(name)
    if CASES[name]['kind'] == 'avi':
        return synth.write_avi(os.path.join(str(tmpdir), name + '.avi'), spec)
    return synth.write_ser_from_array(os.path.join(str(tmpdir), name + '.SER'), stack)


ALL_CASES = list(CASES)
ELLIPSE_CASES = [c for c in CASES if CASES[c].get('ratio_fixe') is None]
