timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
for t in 128 256; do echo "== transv T=$t"; SHG_TRANSV_T=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'])"; done
