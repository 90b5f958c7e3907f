#!/bin/bash
# One GPU session: tests, recon tuning sweep, ncu launch list + full captures of the top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t5.log 2>&1; tail -5 gpurun_out/t5.log
for cfg in "128 4" "256 4" "128 8" "64 8" "128 2" "256 2"; do
  set -- $cfg
  echo "== recon TX=$1 G=$2"
  SHG_RECON_TX=$1 SHG_RECON_G=$2 timeout 120 python tools/kernel_bench.py --only recon 2>&1 | grep -E '"ms"|GBps' | head -2
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_full4.log 2>&1; tail -1 gpurun_out/bench_full4.log | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
for k in accumulate_u16 recon_tma warp_rows transv_row_stats; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k -f python tools/kernel_bench.py --reps 1 --only accumulate,recon,warp,transv > gpurun_out/ncu_$k.log 2>&1
  ls -la gpurun_out/prof_$k.ncu-rep 2>/dev/null
done
