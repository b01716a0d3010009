# Round 2, call 18 (2 GPUs): split exchange v2 (fast path at kernel entry) vs blocking exchange
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29581"
for d in 0 1 0 1; do
for cfg in "--solver cg --global 4096 --max-iters 1500" "--solver cg --global 2048 --max-iters 1500" "--solver cheby --global 4096 --max-iters 2600" "--solver ppcg --global 8192 --max-iters 2600 --ppcg-halo-depth 0"; do
  timeout 300 $TR tools/config_bench.py $cfg --comm fused --reps 1 --prof --opt xchg_deferred=$d 2>> gpurun_out/r02s18.err | grep "^{" >> gpurun_out/r02s18_split_exchange_ab_n2.jsonl
done
done
python - <<'PY'
import json
for l in open('gpurun_out/r02s18_split_exchange_ab_n2.jsonl'):
    d=json.loads(l); p=d['boundary_profile_us_per_kernel']['max_over_ranks']
    print(d['options'], d['solver'], d['global_cells'][0], 'us/sweep %.2f'%d['us_per_sweep'], 'iters', d['iters'], 'err %r'%d['error'], {k: round(v,2) for k,v in p.items() if k not in('solve_ms_with_stamps','kernels')})
PY
( TEALEAF_B200_OPTS=xchg_deferred=1 timeout 300 $TR tests/mgpu_check.py ) > gpurun_out/r02s18_mgpu_parity_split_n2.log 2>&1
grep -c "OK" gpurun_out/r02s18_mgpu_parity_split_n2.log; grep -E "FAIL|Error" gpurun_out/r02s18_mgpu_parity_split_n2.log | head -5
