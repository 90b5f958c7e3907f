# scratch script for one short development call on a GPU box (edit freely):
#   gpurun --timeout 900 -- 'bash tools/dev_session.sh'
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py --only limb --reps 7 > gpurun_out/kb_limb.log 2>&1; tail -12 gpurun_out/kb_limb.log
for h in 1 5 9; do
SHG_TRANSV_HIST=$h timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_h$h.log 2>/dev/null
tail -1 gpurun_out/bench_dev_h$h.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('hist=$h', d['ms_per_step'], d['stages_ms']['transv_stats'])"
done
