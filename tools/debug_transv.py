import sys, os, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solex_ser_recon_en_b200.engine import get_engine
from solex_ser_recon_en_b200._lib import call, lib
from oracle import shg_oracle as O
eng = get_engine(0)
def ref(a, b):
    with np.errstate(divide='ignore', invalid='ignore'):
        return O.reject_outliers_mean(np.log(a / b))
for n in (300, 1000, 5000, 26000, 28000, 30000, 40000):
    for force_global in (False, True):
        rng = np.random.default_rng(n)
        img = np.empty((5, n + 3), np.uint16)
        img[0] = rng.integers(20000, 20400, n + 3); img[1] = rng.integers(20000, 20400, n + 3)
        img[2] = rng.integers(78, 82, n + 3) * 256; img[3] = rng.integers(78, 82, n + 3) * 256; img[4] = img[3]
        d = torch.from_numpy(img).to(eng.device)
        rows = np.array([1, 2, 3, 4, 3, 3], np.int32); xa = np.array([1, 0, 2, 1, 2, 0], np.int32); xb = xa + n
        idx = torch.from_numpy(np.stack([rows, xa, xb])).to(eng.device)
        out = eng.empty((len(rows),), torch.float64)
        max_len = 40000 if force_global else n
        wb = int(lib.shg_transv_workspace_bytes(len(rows), max_len))
        work = eng.empty((max(wb, 1),), torch.uint8)
        call('shg_transv_row_stats', d.data_ptr(), 5, n + 3, idx[0].data_ptr(), idx[1].data_ptr(), idx[2].data_ptr(),
             len(rows), max_len, eng.logtab.data_ptr(), out.data_ptr(), work.data_ptr(), wb, eng.stream)
        got = out.cpu().numpy()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            want = np.array([ref(img[y, a:b], img[y - 1, a:b]) for y, a, b in zip(rows, xa, xb)])
        print(n, 'global' if wb else 'smem', np.abs(got - want) < 1e-12, got[2], want[2])
