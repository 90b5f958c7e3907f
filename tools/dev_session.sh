timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_dev.log 2>&1; tail -1 gpurun_out/bench_dev.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms']); print(d['e2e']['ms_per_step'], d['e2e']['stages_ms'])"
