mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "transv or row_stats or log_u16" > gpurun_out/t_transv.log 2>&1; tail -3 gpurun_out/t_transv.log
for h in 1 5 9; do echo "== bench HIST=$h"; SHG_TRANSV_HIST=$h timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_h$h.log 2>&1; tail -1 gpurun_out/bench_dev_h$h.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms']['transv_stats'])"; done
