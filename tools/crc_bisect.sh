mkdir -p gpurun_out
for env in "X=1" "SHG_WARP_OLD=1" "SHG_LIMB_UNFUSED=1"; do
  env $env timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$env', d['outputs_crc'], d['ms_per_step'])"
done
