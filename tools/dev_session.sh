mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
for k in recon_tma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k -f python tools/kernel_bench.py --reps 1 --only accumulate,recon,warp,transv,minmax > gpurun_out/ncu_$k.log 2>&1
  ls -la gpurun_out/prof_$k.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
