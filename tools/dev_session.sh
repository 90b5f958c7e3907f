#!/bin/bash
# One 1-GPU session: parity tests, the bench line, the ncu launch list of the bench command and
# full captures of the top kernels (distilled into profiles/ by tools/ncu_digest.py), plus the
# small-config CLI timings and the file-ingest sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-3200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
for k in accumulate_u16 recon_tma warp_rows transv_row_stats minmax_u16; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k -f python tools/kernel_bench.py --reps 1 --only accumulate,recon,warp,transv,minmax > gpurun_out/ncu_$k.log 2>&1
  ls -la gpurun_out/prof_$k.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
timeout 600 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; tail -16 gpurun_out/config_bench.log
timeout 300 python tools/kernel_bench.py --only ingest > gpurun_out/ingest_bench.log 2>&1; grep -E "ingest_file|GBps" gpurun_out/ingest_bench.log
# compute-sanitizer over the kernel / entry-point parity tests (memcheck) and the shared-memory-heavy kernels (racecheck)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_entrypoints.py -m gpu -q > gpurun_out/sanitizer_memcheck.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | tail -2
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -k 'row_stats or recon or warp or gain_kernel or transpose or limb' > gpurun_out/sanitizer_racecheck.log 2>&1; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck.log | tail -2
