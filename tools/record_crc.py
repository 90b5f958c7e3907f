"""Record the N=1 outputs_crc of a bench line as the expected value for THIS build (profiles/outputs_crc.json).

    python tools/record_crc.py gpurun_out/bench_n1.log [note]

`bench.py --record-crc` does the same on the machine that ran the bench; a gpurun box only returns gpurun_out/, so
the value is taken from the returned bench line here.  The record is binding for runs of the same sources only
(bench.source_sha: kernels, host logic, the synthetic scan recipe)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
assert line['n_gpus'] == 1 and line.get('impl', 'b200') != 'reference', 'record from the 1-GPU line of the product arm'
cfg = line['config']
key = '%dx%dx%d_s%d' % (cfg['frames'], cfg['width'], cfg['height'], cfg['n_shifts'])
if 'e2e' in line and 'outputs_crc' in line['e2e']:
    assert line['e2e']['outputs_crc'] == line['outputs_crc'], 'resident and end-to-end images differ'
try:
    book = json.load(open(bench.CRC_FILE))
except Exception:
    book = {}
book[key] = {'crc': line['outputs_crc'], 'n_gpus': 1, 'source_sha': bench.source_sha()}
if len(sys.argv) > 2:                      # e.g. what changed in the sources since the run the line comes from
    book[key]['note'] = sys.argv[2]
with open(bench.CRC_FILE, 'w') as f:
    json.dump(book, f, indent=1, sort_keys=True)
print(key, book[key])
