"""Launch-shape sweep of the TMA warp kernel on config-5 sized disks (development aid)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solex_ser_recon_en_b200 import geometry as G                        # noqa: E402
from solex_ser_recon_en_b200.engine import get_engine                    # noqa: E402

eng = get_engine(0)
n_img, n, ih = int(os.environ.get('SWEEP_IMGS', '101')), 20000, 4096
disk = torch.randint(300, 30000, (n_img, n, ih), dtype=torch.int16, device=eng.device).view(torch.uint16)
mm = eng.minmax_device(disk)
res = {}
for phi in (-0.0004, 0.02):
    mat3, (oh, ow) = G.warp_plan((ih, n), phi, 0.195)[1:3]
    out = eng.empty((n_img, oh, ow), torch.uint16)
    nbytes = n_img * (n * ih * 2 + oh * ow * 2)
    ref = None
    shapes = ['old', '64,256,1,0', '64,512,1,0', '128,512,1,0', '64,256,2,2', '64,256,3,1', '64,512,3,1', '64,512,4,1',
              '64,256,4,1', '64,256,2,1', '128,512,2,1', '64,256,1,4', '64,256,1,8']
    for shape in shapes:
        os.environ.pop('SHG_WARP_OLD', None)
        os.environ.pop('SHG_WARP_SHAPE', None)
        if shape == 'old':
            os.environ['SHG_WARP_OLD'] = '1'
        else:
            os.environ['SHG_WARP_SHAPE'] = shape
        try:
            for _ in range(2):
                eng.warp_batch(disk, None, False, mat3, (oh, ow), mm, out=out)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            for a, b in ev:
                a.record()
                eng.warp_batch(disk, None, False, mat3, (oh, ow), mm, out=out)
                b.record()
            torch.cuda.synchronize()
            ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
            chk = eng.checksum(out[n_img // 2].contiguous())
            ref = chk if ref is None else ref
            res['phi=%g %s' % (phi, shape)] = dict(ms=round(ms, 3), GBps=round(nbytes / ms / 1e6), same=chk == ref)
        except Exception as e:
            res['phi=%g %s' % (phi, shape)] = repr(e)[:120]
        print(list(res.items())[-1], flush=True)
json.dump(res, open('gpurun_out/warp_sweep.json', 'w'), indent=1)
