"""Distil gpurun_out/ ncu artefacts into small tracked files under profiles/:
  profiles/<tag>_launches.md       per-kernel share of the bench step (ncu launch list of the bench command)
  profiles/<tag>_<kernel>.txt      headline metrics of each `ncu --set full` capture
  profiles/accumulate_traffic.json DRAM bytes per launch of the dominant kernel (read by bench.py)
usage: python tools/ncu_digest.py r02"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
os.makedirs(PROF, exist_ok=True)

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
        'smsp__pcsamp_warps_issue_stalled_branch_resolving']


def to_bytes(v, unit):
    f = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)
    return float(v) * f


def digest_full(path, name, pick=None, note=None):
    """Headline metrics of one kernel of an ncu report -> profiles/<tag>_<name>.txt.  `pick`: substring of the
    kernel name (a report may hold several kernels / launches); the LAST matching launch is taken (in
    tools/kernel_bench.py that is the timed repetition, after the warm-up launches)."""
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr) and (pick is None or pick in r[ix['Kernel Name']])]
    if not body:
        return {}
    vals = body[-1]
    lines = [note or 'ncu --set full --clock-control none, the production launch inside tools/kernel_bench.py (config-5 '
             'geometry: whole stack / 101-image batch)',
             'kernel: ' + vals[ix['Kernel Name']][:160],
             'launches of this kernel in the report: %d (this is the last)' % len(body)]
    got = {}
    for w in WANT:
        if w in ix:
            lines.append('%-62s %s %s' % (w, vals[ix[w]], units[ix[w]]))
            got[w] = (vals[ix[w]], units[ix[w]])
    sass = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'solex_ser_recon_en_b200', 'libshg.so')],
                          capture_output=True, text=True).stdout
    if name in ('recon_tma', 'warp_tma'):
        lines.append('SASS evidence of TMA in libshg.so (all kernels): UTMALDG x%d, SYNCS.ARRIVE.TRANS64 x%d' %
                     (sass.count('UTMALDG'), sass.count('SYNCS.ARRIVE.TRANS64')))
    open(os.path.join(PROF, '%s_%s.txt' % (tag, name)), 'w').write('\n'.join(lines) + '\n')
    return got


def digest_launches(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[-70:]
        v, u = float(r[ix['Metric Value']]), r[ix['Metric Unit']]
        us = v / 1000 if u.startswith('n') else (v if u.startswith('u') else v * 1000)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    setup = {k: v for k, v in agg.items() if 'synth_kernel' in k or 'log_table' in k}
    step = {k: v for k, v in agg.items() if k not in setup}
    tot = sum(v[1] for v in step.values())
    lines = ['# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs` (config 5, 1 x B200)', '',
             '`ncu --metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and serialised,',
             'so the SHARE of the step is what is comparable with the CUDA-event stage times in the bench line.',
             '3 passes of the step (1 warm-up + 2 timed) + the roofline loop of the accumulate kernel are in the list.', '',
             '| kernel | launches | total us | share of listed step kernels |', '|---|---|---|---|']
    for k, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
        lines.append('| `%s` | %d | %.1f | %.1f %% |' % (k, n, t, 100 * t / tot))
    lines += ['', 'one-off set-up kernels (not part of the step): ' +
              ', '.join('`%s` %.0f us' % (k, v[1]) for k, v in setup.items())]
    open(os.path.join(PROF, '%s_launches.md' % tag), 'w').write('\n'.join(lines) + '\n')


if os.path.exists(os.path.join(OUT, 'launches.csv')):
    digest_launches(os.path.join(OUT, 'launches.csv'))
JOBS = [('accumulate_u16', 'prof_accumulate_u16.ncu-rep', 'accumulate_u16'),
        ('recon_tma', 'prof_recon_tma.ncu-rep', 'recon_tma_pair'),
        ('warp_tma', 'prof_batch.ncu-rep', 'warp_tma'),
        ('row_scale', 'prof_batch.ncu-rep', 'row_scale_kernel'),
        ('transv_row_stats_reg', 'prof_transv_reg.ncu-rep', 'transv_row_stats_reg'),
        ('warp_rows', 'prof_warp_rows.ncu-rep', None), ('transv_row_stats', 'prof_transv_row_stats.ncu-rep', None),
        ('minmax_u16', 'prof_minmax_u16.ncu-rep', None)]
for name, rep, pick in JOBS:
    p = os.path.join(OUT, rep)
    if tag != 'r01' and name in ('warp_rows', 'transv_row_stats', 'minmax_u16'):
        continue                                   # round-1 captures (kernels since replaced / no longer in the step)
    if os.path.exists(p):
        got = digest_full(p, name, pick)
        if name == 'accumulate_u16' and 'dram__bytes_read.sum' in got:
            rd = to_bytes(*got['dram__bytes_read.sum'])
            wr = to_bytes(*got['dram__bytes_write.sum'])
            json.dump({'kernel': 'accumulate_u16_kernel', 'dram_bytes_per_launch': rd + wr, 'read': rd, 'write': wr,
                       'frames': 20000, 'algorithmic_bytes': 20000 * 4096 * 512 * 2,
                       'source': 'ncu --set full, profiles/%s_accumulate_u16.txt' % tag},
                      open(os.path.join(PROF, 'accumulate_traffic.json'), 'w'), indent=1)
print(sorted(os.listdir(PROF)))
