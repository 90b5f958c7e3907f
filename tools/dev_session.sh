mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-300
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
