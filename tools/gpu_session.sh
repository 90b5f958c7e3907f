#!/bin/bash
# One 1-GPU session (round 2): the bench line, the reference arm, the ncu launch list of the bench command, ONE
# `ncu --set full` run that captures the production launches of the top kernels (the 101-image batches of
# tools/kernel_bench.py; distilled into profiles/ by tools/ncu_digest.py) and compute-sanitizer over the new kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tests]'
mkdir -p gpurun_out
if [ "$1" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 --timeline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.log | cut -c1-1500
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:accumulate_u16|recon_tma_pair|warp_tma|transv_row_stats_reg|row_scale_kernel' -c 40 -o gpurun_out/prof_all -f python tools/kernel_bench.py --reps 1 --only accumulate,recon,batch > gpurun_out/ncu_all.log 2>&1
ls -la gpurun_out/prof_all.ncu-rep 2>/dev/null | awk '{print $5, $9}'
# compute-sanitizer over the row-statistics kernels (new this round) and the limb kernels (fused this round)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k 'row_stats or transv or limb' > gpurun_out/sanitizer_memcheck.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | tail -2
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -k 'row_stats or transv or limb' > gpurun_out/sanitizer_racecheck.log 2>&1; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck.log | tail -2
