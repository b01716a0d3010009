# Round 2, call 12 (2 GPUs): effect of the acquire load (instead of fence.sys) after the exchange and of the uniform carve-out, boundary profile
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561"
for carve in 1 0; do
for cfg in "--solver cg --global 4096 --max-iters 1500" "--solver cg --global 2048 --max-iters 1500" "--solver cheby --global 4096 --max-iters 2600"; do
  TEALEAF_B200_CARVEOUT=$carve timeout 300 $TR tools/config_bench.py $cfg --comm fused --reps 1 --prof 2>> gpurun_out/r02s12.err | grep "^{" | sed "s/^{/{\"carveout\": $carve, /" >> gpurun_out/r02s12_boundary_profile_n2.jsonl
done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02s12_boundary_profile_n2.jsonl'):
    d=json.loads(l); p=d['boundary_profile_us_per_kernel']['max_over_ranks']
    print(d['carveout'], d['solver'], d['global_cells'][0], 'us/sweep %.2f'%d['us_per_sweep'], {k: round(v,2) for k,v in p.items() if k!='solve_ms_with_stamps'})
PY
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "tiled_solvers and default" 2>&1 | tail -3
