mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests_gpu.log 2>&1; tail -4 gpurun_out/tests_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-configs --timeline > gpurun_out/bench_dev.log 2> gpurun_out/timeline.txt
tail -1 gpurun_out/bench_dev.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['outputs_crc'], d['outputs_crc_match'], d['stages_ms'])"
grep timeline gpurun_out/timeline.txt | cut -c1-110
