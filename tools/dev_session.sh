mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "limb or ellipse or process" > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
timeout 300 python tools/kernel_bench.py --only limb --reps 9 > gpurun_out/kb_limb.log 2>&1; tail -13 gpurun_out/kb_limb.log
SHG_LIMB_UNFUSED=1 timeout 300 python tools/kernel_bench.py --only limb --reps 9 2>&1 | grep -A3 limb_chained
