mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/tests_multi.log 2>&1; tail -2 gpurun_out/tests_multi.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_entrypoints.py -m gpu -q -x -k "warp or circular or correct_image or process or entry or cli" > gpurun_out/t_warp.log 2>&1; tail -2 gpurun_out/t_warp.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev.log 2>&1; tail -1 gpurun_out/bench_dev.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms']['warp'], d['stages_ms'])"
