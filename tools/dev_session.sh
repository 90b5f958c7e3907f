mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "recon" > gpurun_out/t_recon.log 2>&1; tail -3 gpurun_out/t_recon.log
for cfg in "1 4" "1 2" "0 2"; do set -- $cfg; echo "== recon PAIR=$1 G=$2"; SHG_RECON_PAIR=$1 SHG_RECON_G=$2 timeout 120 python tools/kernel_bench.py --only recon 2>&1 | grep -E '"ms"' | head -1; done
echo "== bench"; timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev.log 2>&1; tail -1 gpurun_out/bench_dev.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"
