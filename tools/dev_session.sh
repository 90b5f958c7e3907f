# scratch script for one short development call on a GPU box (edit freely):
#   gpurun --timeout 900 -- 'bash tools/dev_session.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --timeline > gpurun_out/bench_dev.log 2> gpurun_out/timeline.txt
tail -1 gpurun_out/bench_dev.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"
grep timeline gpurun_out/timeline.txt
