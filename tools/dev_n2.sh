mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/tests_multi.log 2>&1; tail -3 gpurun_out/tests_multi.log
for mode in by_shift post_warp; do
  SHG_EXCHANGE=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/bench_n2_$mode.log 2> gpurun_out/bench_n2_$mode.err
  echo "$mode rc=$?"
  tail -1 gpurun_out/bench_n2_$mode.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['outputs_crc'], d['outputs_crc_match'], d['e2e']['ms_per_step'], d['e2e']['pcie_frac'], d['e2e']['pcie_h2d_peak_GBps_per_gpu'], d['e2e']['outputs_crc'])"
  grep "RANK" -A12 gpurun_out/bench_n2_$mode.err | head -30
done
