timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
for cfg in "128 4" "256 4" "256 2" "128 2" "64 4"; do set -- $cfg; echo "== recon TX=$1 G=$2"; SHG_RECON_TX=$1 SHG_RECON_G=$2 timeout 120 python tools/kernel_bench.py --only recon 2>&1 | grep -E '"ms"' | head -1; done
timeout 120 python tools/kernel_bench.py --only warp,transv 2>&1 | grep -E '"(warp|transv_stats)"|"ms"'
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev.log 2>&1; tail -1 gpurun_out/bench_dev.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"
