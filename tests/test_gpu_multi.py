"""GPU, >= 2 devices: frames sharded across ranks (NCCL all-reduce of the integer
mean / max frame, reconstruction rows written into the owner rank's images over
NVLink peer memory) must reproduce the single-GPU / reference results exactly.
Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize('mode', ['by_shift', 'gather0', 'post_warp'])
@pytest.mark.parametrize('name', ['ser16_rot', 'ser8_rot_flip'])
def test_two_ranks_match_reference(name, mode, tmp_path):
    if _n_gpus() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29631', os.path.join(ROOT, 'tests', 'mgpu_worker.py'),
           name, str(tmp_path), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(os.path.join(str(tmp_path), 'ok0')) and os.path.exists(os.path.join(str(tmp_path), 'ok1'))
