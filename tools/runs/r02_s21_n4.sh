# Round 2, 4-GPU call: parity on 2x2 tiles + bench --gpus 4 (the driver's SCALE run includes N = 4)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29601"
( time timeout 420 $TR tests/mgpu_check.py ) > gpurun_out/r02s21_mgpu_parity_n4.log 2>&1
grep -c " OK" gpurun_out/r02s21_mgpu_parity_n4.log; grep -E "FAIL|real" gpurun_out/r02s21_mgpu_parity_n4.log | cut -c1-200
timeout 500 $TR bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02s21_bench_n4.json 2> gpurun_out/r02s21_bench_n4.err
tail -2 gpurun_out/r02s21_bench_n4.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s21_bench_n4.json').read())
print('value', d['value'], 'solo', d.get('same_tile_single_gpu',{}).get('value'), 'e2e', d['e2e']['value'])
for k,v in d['other_configs'].items():
    if isinstance(v,dict):
        print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('solve_ms','cg_phase_ms','us_per_cheby_iteration','us_per_sweep','us_per_iteration','halo_depth_k','skipped','error')})
PY
