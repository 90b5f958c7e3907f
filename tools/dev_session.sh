mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
echo "== bench + timeline"; timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --timeline > gpurun_out/bench_dev.log 2> gpurun_out/timeline.txt; tail -1 gpurun_out/bench_dev.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"; grep timeline gpurun_out/timeline.txt
echo "== bench NO_MIN"; SHG_RECON_NO_MIN=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_nomin.log 2>&1; tail -1 gpurun_out/bench_dev_nomin.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"
for m in 0 1; do echo "== recon kernel NO_MIN=$m"; if [ $m = 1 ]; then export SHG_RECON_NO_MIN=1; fi; timeout 120 python tools/kernel_bench.py --only recon 2>&1 | grep -E '"ms"' | head -1; done
