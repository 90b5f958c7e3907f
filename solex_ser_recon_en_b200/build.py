"""Build libshg.so (the sm_100a kernels + C ABI) in-tree with nvcc.

    python -m solex_ser_recon_en_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so
it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
BUILD_DIR = os.path.join(ROOT, 'build', 'libshg')
LIB_PATH = os.path.join(PKG_DIR, 'libshg.so')

SOURCES = ['api.cu', 'mean_max.cu', 'detect.cu', 'recon.cu', 'layout.cu', 'warp.cu', 'transv.cu', 'limb.cu',
           'synth.cu', 'ingest.cu', 'tail.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC,-pthread', '-I', INCLUDE, '-I', CSRC,
              '-DSHG_BUILDING=1']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libshg.so cannot be built')
    return exe


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(INCLUDE, 'shg.h'), os.path.join(CSRC, 'common.cuh'), os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD_DIR, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed on ' + cmd[-3])
    if jobs or force or _newer(LIB_PATH, objs):
        cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-Xcompiler', '-pthread', '-ldl']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link of libshg.so failed')
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
