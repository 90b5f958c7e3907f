# scratch script for one short development call on a GPU box (edit freely):
#   gpurun --timeout 900 -- 'bash tools/dev_session.sh'
mkdir -p gpurun_out
true
timeout 800 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo rc=$?
tail -1 gpurun_out/bench_n1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['outputs_crc'], d['outputs_crc_match'], d['e2e']['ms_per_step'], d['e2e']['pcie_frac'], d['e2e']['pcie_h2d_peak_GBps_per_gpu'], d['e2e']['outputs_crc'])
for c in d.get('config_lines', []): print(c)"
tail -3 gpurun_out/bench_n1.err

