"""Wall-clock of the drop-in CLI on the small BASELINE configurations (files on tmpfs).

    python tools/config_bench.py            # config 1 (1 file), config 3 (21 shifts), config 4 (4 files back to back)

Scans are synthesised on the device and written as SER files; the timed call is
solex_ser_recon_en_b200.SHG_MAIN.main([...flags, files]) -- CLI parsing, ingest from
the file, the whole GPU path, CLAHE + PNG writers (-c: only the clahe PNG)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['SHG_NO_CONFIG'] = '1'

from solex_ser_recon_en_b200 import SHG_MAIN, synth                       # noqa: E402
from solex_ser_recon_en_b200.engine import ScanGeometry, get_engine      # noqa: E402


def write_scan(path, n, w, h, seed):
    eng = get_engine(0)
    st = eng.synth_stack(ScanGeometry(w, h, 2, n), seed=seed)
    host = st.frames.cpu().numpy()
    with open(path, 'wb') as f:
        f.write(synth.ser_header(w, h, 16, n))
        f.write(host.tobytes())
    return path


def run(flags, files, reps=3):
    times = []
    for _ in range(reps):
        SHG_MAIN.options.update(shift=[0], ratio_fixe=None, slant_fix=None, flip_x=False, crop_width_square=False,
                                clahe_only=False, output_dir=OUT)
        t0 = time.perf_counter()
        assert SHG_MAIN.main(flags + files) == 0
        times.append(time.perf_counter() - t0)
    return min(times)


OUT = '/dev/shm/shg_cfg_out'
os.makedirs(OUT, exist_ok=True)
res = {}
f1 = write_scan('/dev/shm/shg_cfg1.SER', 1000, 1280, 200, 1)
t = run(['-c'], [f1])
res['config1: 1000 x 1280x200, shift 0, -c'] = dict(seconds=t, frames_per_s=1000 / t)
f3 = write_scan('/dev/shm/shg_cfg3.SER', 4000, 2048, 300, 3)
t = run(['-cw-10:10:1'], [f3])
res['config3: 4000 x 2048x300, 21 shifts, -c'] = dict(seconds=t, frames_per_s=4000 / t)
f4 = [write_scan('/dev/shm/shg_cfg4_%d.SER' % i, 3000, 2048, 256, 40 + i) for i in range(4)]
t = run(['-c'], f4)
res['config4: 4 files x 3000 x 2048x256, shift 0, -c'] = dict(seconds=t, frames_per_s=12000 / t, files=4)
for p in [f1, f3] + f4:
    os.remove(p)
print(json.dumps(res, indent=1))
