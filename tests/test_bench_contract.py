"""The bench.py output contract: one JSON line with the keys the driver reads.
CPU: the reference arm on a tiny geometry.  GPU: the CUDA arm on a reduced scan."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'cpu_baseline'}


def _last_json(stdout):
    for ln in reversed(stdout.strip().splitlines()):
        if ln.startswith('{'):
            return json.loads(ln)
    raise AssertionError('no JSON line in:\n' + stdout[-2000:])


def _reference_arm(env_extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup',
                        '0', '--frames', '400', '--width', '512', '--height', '128', '--sample-frames', '32'],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, **env_extra))
    assert r.returncode == 0, r.stderr[-2000:]
    d = _last_json(r.stdout)
    assert BASE_KEYS <= set(d)
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['value'] > 0 and d['higher_is_better'] is True
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['vs_baseline'] is None and d['gpu_launches'] == 0
    return d


def test_reference_arm_line_port_fallback():
    """Without the reference modules (neither /root/reference nor the staged oracle/_ref archive) the arm times the
    oracle's port and says so."""
    d = _reference_arm({'SHG_REF_PORT': '1'})
    assert d['cpu_baseline']['kind'] == 'port'


def test_reference_arm_line_unmodified_reference():
    """With the reference modules present the arm runs the UNMODIFIED solex_read + solex_process in one process."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference modules not present (run python -m oracle.stage_ref where /root/reference exists)')
    d = _reference_arm({})
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] == 1 and 'UNMODIFIED' in cb['sample']
    assert cb['host_cpus'] >= 1 and cb['cv2_threads'] >= 1 and d['extrapolated_full_scan_s'] > 0


def test_staged_reference_archive_imports_without_the_tree(tmp_path):
    """oracle/stage_ref.py archives the reference modules byte for byte; ref_shim imports them from the archive when
    /root/reference is absent (the GPU box)."""
    from oracle import stage_ref
    if not stage_ref.stage():
        pytest.skip('no reference tree and no staged archive')
    code = ('import sys; sys.path.insert(0, %r); from oracle import ref_shim; assert ref_shim.source() == "staged"; '
            'ref = ref_shim.load(); assert ref.solex_util.__file__.endswith("reference_modules.zip/solex_util.py"); '
            'print(ref.Solex_recon.solex_read.__name__)' % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, SHG_REFERENCE_DIR=str(tmp_path / 'absent')))
    assert r.returncode == 0 and r.stdout.strip() == 'solex_read', r.stderr[-1500:]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.gpu
def test_cuda_arm_line_reduced_scan():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '2', '--warmup', '1', '--frames', '1500',
                        '--no-cpu'], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _last_json(r.stdout)
    assert (BASE_KEYS - {'cpu_baseline'}) <= set(d)
    assert d['config'].get('reduced') is True and d['n_gpus'] == 1 and d['gpu_launches'] > 20
    roof = d['roofline']
    assert roof['bound'] == 'hbm' and roof['unit'] == 'GB/s' and 0 < roof['frac'] < 1.5
    assert abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-9
    e = d['e2e']
    assert e['h2d_bytes_per_step'] == 1500 * 4096 * 512 * 2 and e['d2h_bytes_per_step'] > 0
    assert e['value'] < d['value']                      # the copies are inside the end-to-end number
    assert set(d['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
