timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
SHG_TRANSV_T=64 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_entrypoints.py -m gpu -q -k "row_stats or transversalium or solex_process" 2>&1 | tail -3
for t in 128 64; do echo "== transv T=$t"; SHG_TRANSV_T=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms']['transv_stats'])"; done
timeout 300 python tools/kernel_bench.py --only ingest 2>&1 | grep -E "ingest_file|GBps"
