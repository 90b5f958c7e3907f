#!/bin/bash
# One 1-GPU session (round 2): the bench line, the ncu launch list of the bench command and `ncu --set full` captures
# of the production launches of the top kernels (tools/kernel_bench.py: whole stack / 101-image batches), distilled
# into profiles/ by tools/ncu_digest.py.  gpurun copies back at most 64 MiB: no --import-source here (the source-level
# captures used during development are taken one kernel at a time), and the reports go first if the cap is near.
#   gpurun --timeout 1200 -- 'bash tools/gpu_session.sh [tests] [ref] [sanitizer]'
mkdir -p gpurun_out
for a in "$@"; do eval "do_$a=1"; done
if [ -n "$do_tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 --timeline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.log | cut -c1-300
if [ -n "$do_ref" ]; then
  timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log | cut -c1-300
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:accumulate_u16 -s 2 -c 1 -o gpurun_out/prof_accumulate_u16 -f python tools/kernel_bench.py --reps 1 --only accumulate > gpurun_out/ncu_accumulate_u16.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:recon_tma_pair -s 2 -c 1 -o gpurun_out/prof_recon_tma -f python tools/kernel_bench.py --reps 1 --only recon > gpurun_out/ncu_recon_tma.log 2>&1
timeout 300 ncu --set full --clock-control none -k 'regex:warp_tma|row_scale_kernel' -c 8 -o gpurun_out/prof_batch -f python tools/kernel_bench.py --reps 1 --only batch > gpurun_out/ncu_batch.log 2>&1
if [ -n "$do_sanitizer" ]; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k 'row_stats or transv or limb' > gpurun_out/sanitizer_memcheck.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | tail -2
  timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -k 'row_stats or transv or limb' > gpurun_out/sanitizer_racecheck.log 2>&1; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck.log | tail -2
fi
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/prof_batch.ncu-rep; echo "dropped prof_batch (size cap)"; fi
du -sm gpurun_out
