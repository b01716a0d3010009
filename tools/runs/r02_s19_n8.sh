# Round 2, final 8-GPU call: HEAD (acquire-load exchange, kernel-zeroed control block, diagnostics) -- parity on 2x4 tiles, one context over 8 GPUs, bench --gpus 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29591"
( time timeout 420 $TR tests/mgpu_check.py ) > gpurun_out/r02s19_mgpu_parity_n8.log 2>&1
grep -c " OK" gpurun_out/r02s19_mgpu_parity_n8.log; grep -E "FAIL|real" gpurun_out/r02s19_mgpu_parity_n8.log | cut -c1-200
( timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "one_context or two_devices" ) > gpurun_out/r02s19_pytest_multi_ctx_n8.log 2>&1
tail -2 gpurun_out/r02s19_pytest_multi_ctx_n8.log | cut -c1-300
timeout 500 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02s19_bench_n8.json 2> gpurun_out/r02s19_bench_n8.err
tail -2 gpurun_out/r02s19_bench_n8.err | cut -c1-300; cut -c1-300 gpurun_out/r02s19_bench_n8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s19_bench_n8.json').read())
print('value', d['value'], 'solo', d.get('same_tile_single_gpu',{}).get('value'), 'e2e', d['e2e']['value'])
for k,v in d['other_configs'].items():
    if isinstance(v,dict):
        print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('solve_ms','cg_phase_ms','us_per_cheby_iteration','us_per_sweep','us_per_iteration','halo_depth_k','skipped','error')})
PY
timeout 200 $TR tools/config_bench.py --solver cg --global 4096 --max-iters 1500 --comm fused --reps 1 --prof 2>> gpurun_out/r02s19.err | grep "^{" > gpurun_out/r02s19_boundary_profile_cg_n8.jsonl
cut -c1-900 gpurun_out/r02s19_boundary_profile_cg_n8.jsonl
