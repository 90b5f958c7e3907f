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


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup',
                        '0', '--frames', '400', '--width', '512', '--height', '128', '--sample-frames', '32'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _last_json(r.stdout)
    assert BASE_KEYS <= set(d)
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['value'] > 0 and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['vs_baseline'] is None and d['gpu_launches'] == 0


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
