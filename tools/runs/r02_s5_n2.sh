# Round 2, call 5 (2 GPUs): real-NVLink validation of the tiled pair kernels and of tl_create_multi; bench --gpus 2; boundary micro-profile
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541"
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --durations=6 ) > gpurun_out/r02s5_pytest_n2.log 2>&1
tail -25 gpurun_out/r02s5_pytest_n2.log | cut -c1-400
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02s5_bench_n2.json 2> gpurun_out/r02s5_bench_n2.err
tail -3 gpurun_out/r02s5_bench_n2.err | cut -c1-300; cut -c1-1500 gpurun_out/r02s5_bench_n2.json
for cfg in "--solver cg --global 4096 --max-iters 1500" "--solver cheby --global 4096 --max-iters 2600" "--solver cheby --global 4096 --max-iters 2600 --opt pair_tiled=0" "--solver ppcg --global 8192 --max-iters 2600 --ppcg-halo-depth 0" "--solver ppcg --global 8192 --max-iters 2600 --ppcg-halo-depth 1 --opt ppcg_pair=0"; do
  timeout 300 $TR tools/config_bench.py $cfg --comm fused --reps 1 --prof >> gpurun_out/r02s5_config_bench_n2.jsonl 2>> gpurun_out/r02s5_config_bench_n2.err
done
cut -c1-1200 gpurun_out/r02s5_config_bench_n2.jsonl
