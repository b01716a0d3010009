# Round 2, 8-GPU call: parity on 2x4 tiles (both schedules), one context over 8 GPUs, bench --gpus 8 with the config legs, boundary profile
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29551"
( time timeout 420 $TR tests/mgpu_check.py ) > gpurun_out/r02s11_mgpu_parity_n8.log 2>&1
grep -E "mgpu|real" gpurun_out/r02s11_mgpu_parity_n8.log | cut -c1-260
( timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "one_context or two_devices" ) > gpurun_out/r02s11_pytest_multi_ctx_n8.log 2>&1
tail -4 gpurun_out/r02s11_pytest_multi_ctx_n8.log | cut -c1-300
timeout 500 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02s11_bench_n8.json 2> gpurun_out/r02s11_bench_n8.err
tail -2 gpurun_out/r02s11_bench_n8.err | cut -c1-300; cut -c1-600 gpurun_out/r02s11_bench_n8.json
for cfg in "--solver cg --global 4096 --max-iters 1500" "--solver cheby --global 4096 --max-iters 2600" "--solver ppcg --global 8192 --max-iters 2600 --ppcg-halo-depth 0"; do
  timeout 200 $TR tools/config_bench.py $cfg --comm fused --reps 1 --prof >> gpurun_out/r02s11_boundary_profile_n8.jsonl 2>> gpurun_out/r02s11_boundary_profile_n8.err
done
cut -c1-1300 gpurun_out/r02s11_boundary_profile_n8.jsonl
