mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "transv or row_stats or log_u16" > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
timeout 300 python tools/kernel_bench.py --only tbatch --reps 7 > gpurun_out/kb_tbatch.log 2>&1; tail -40 gpurun_out/kb_tbatch.log
