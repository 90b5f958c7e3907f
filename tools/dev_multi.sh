# the bench on N GPUs of one box as the driver launches it (+ per-rank timelines); at 8 GPUs also the H2D map
#   gpurun --gpus N --timeout 900 -- 'bash tools/dev_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout ${T:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 --timeline > gpurun_out/bench_n${N}.log 2> gpurun_out/bench_n${N}.err
echo "rc=$?"
tail -1 gpurun_out/bench_n${N}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['outputs_crc'], d['outputs_crc_match'], d['stages_ms'], d.get('e2e',{}).get('ms_per_step'), d.get('e2e',{}).get('pcie_frac'))"
grep "RANK" -A12 gpurun_out/bench_n${N}.err | head -30
if [ "$N" = "8" ] || [ -n "$H2D" ]; then
  timeout 200 python tools/h2d_probe.py --seconds 1.0 --gb 2 --subsets '0,2;0,4' --out gpurun_out/h2d_matrix.json > gpurun_out/h2d_probe.log 2>&1; grep aggregate gpurun_out/h2d_probe.log | cut -c1-200
fi
if [ -n "$BOTH" ]; then     # the other exchange mode, resident step only
  SHG_EXCHANGE=post_warp timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs --timeline > gpurun_out/bench_n${N}_post_warp.log 2> gpurun_out/bench_n${N}_post_warp.err
  tail -1 gpurun_out/bench_n${N}_post_warp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('post_warp', d['ms_per_step'], d['outputs_crc'], d['outputs_crc_match'], d['stages_ms'])"
fi
