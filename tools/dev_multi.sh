# development call on N GPUs: the 2-rank parity tests, then bench.py in both exchange modes
#   gpurun --gpus N --timeout 900 -- 'bash tools/dev_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ] && [ -n "$TESTS" ]; then
  timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/tests_multi.log 2>&1; tail -3 gpurun_out/tests_multi.log
fi
for mode in by_shift post_warp; do
  SHG_EXCHANGE=$mode timeout ${T:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ${2:---no-e2e} --timeline > gpurun_out/bench_n${N}_$mode.log 2> gpurun_out/bench_n${N}_$mode.err
  echo "$mode rc=$?"
  tail -1 gpurun_out/bench_n${N}_$mode.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['outputs_crc'], d['outputs_crc_match'], d['stages_ms'], d.get('e2e',{}).get('ms_per_step'))"
  grep "timeline\|RANK" gpurun_out/bench_n${N}_$mode.err | head -40
done
