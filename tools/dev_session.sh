# scratch script for one short development call on a GPU box (edit freely):
#   gpurun --timeout 900 -- 'bash tools/dev_session.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests_gpu.log 2>&1; tail -5 gpurun_out/tests_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 800 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo rc=$?
tail -1 gpurun_out/bench_n1.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['outputs_crc'], d['outputs_crc_match'], d['e2e']['ms_per_step'], d['e2e']['pcie_frac'])
print(json.dumps(d['cpu_baseline'])[:900])
for c in d.get('config_lines', []): print(c)"
tail -3 gpurun_out/bench_n1.err
