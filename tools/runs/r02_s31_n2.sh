# Round 2, call 31 (2 GPUs): HEAD with the lazy u update -- parity of every solver across 2 GPUs (mgpu_check), the multi-GPU tests, bench N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29581"
( timeout 400 $TR tests/mgpu_check.py ) > gpurun_out/r02s31_mgpu_parity_n2.log 2>&1
grep -c "OK" gpurun_out/r02s31_mgpu_parity_n2.log; grep -E "FAIL|Error" gpurun_out/r02s31_mgpu_parity_n2.log | head -5
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q ) > gpurun_out/r02s31_pytest_multi_gpu.log 2>&1
tail -3 gpurun_out/r02s31_pytest_multi_gpu.log | cut -c1-300
timeout 600 $TR bench.py --gpus 2 > gpurun_out/r02s31_bench_n2.json 2> gpurun_out/r02s31_bench_n2.err
cut -c1-300 gpurun_out/r02s31_bench_n2.json
for cfg in "--solver cg --global 4096 --max-iters 1500" "--solver cg --global 4096 --max-iters 1500 --opt cg_lazy_u=0"; do
  timeout 300 $TR tools/config_bench.py $cfg --comm fused --reps 2 2>> gpurun_out/r02s31.err | grep "^{" >> gpurun_out/r02s31_cg4096_lazy_ab_n2.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02s31_cg4096_lazy_ab_n2.jsonl'):
    d=json.loads(l); print(d['options'], d['solver'], d['global_cells'][0], 'us/sweep %.2f'%d['us_per_sweep'], 'iters', d['iters'], 'err %r'%d['error'])
PY
