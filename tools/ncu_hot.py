"""Print the hot SASS instructions of an ncu report (source page): share of executed
warp instructions and of stall samples per instruction.  usage: ncu_hot.py report.ncu-rep [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ix['Instructions Executed']]), int(r[ix['# Samples']]), r[ix['Source']].strip()))
    except Exception:
        pass
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print('kernel:', rows[0][1][:100])
print('warp instructions executed: %d   stall samples: %d' % (tot, tots))
for n, (e, s, src) in enumerate(data):
    if e > tot * min_pct / 100 or s > tots * min_pct / 100:
        print('%5d %6.2f%% inst %6.2f%% stall  %s' % (n, 100 * e / tot, 100 * s / max(tots, 1), src[:100]))
