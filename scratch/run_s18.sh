# N=2: multi-GPU parity at HEAD + depth-k (matrix-powers) PPCG timing A/B + strong-scaling references
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29641"
( time timeout 300 python -m pytest tests/test_multi_gpu.py tests/test_tiled_one_gpu.py -m gpu -x -q ) > gpurun_out/s18_pytest_n2.log 2>&1
tail -6 gpurun_out/s18_pytest_n2.log
for N in 8192 4096; do
for k in 1 2 4; do
  timeout 120 $TR tools/config_bench.py --solver ppcg --global $N --max-iters 130 --inner 10 --halo-depth 4 --ppcg-halo-depth $k --comm fused --reps 2 2>/dev/null | grep '^{' | tee -a gpurun_out/s18_depthk_n2.jsonl | cut -c1-400
done
done
timeout 120 $TR tools/config_bench.py --solver ppcg --global 8192 --max-iters 130 --inner 20 --halo-depth 4 --ppcg-halo-depth 1 --comm fused --reps 2 2>/dev/null | grep '^{' | tee -a gpurun_out/s18_depthk_n2.jsonl | cut -c1-400
timeout 120 $TR tools/config_bench.py --solver ppcg --global 8192 --max-iters 130 --inner 20 --halo-depth 4 --ppcg-halo-depth 4 --comm fused --reps 2 2>/dev/null | grep '^{' | tee -a gpurun_out/s18_depthk_n2.jsonl | cut -c1-400
timeout 120 $TR tools/config_bench.py --solver cheby --global 4096 --max-iters 2000 --comm fused --reps 2 2>/dev/null | grep '^{' | tee -a gpurun_out/s18_strong_n2.jsonl | cut -c1-400
timeout 120 $TR tools/config_bench.py --solver cg --global 4096 --max-iters 1000 --comm fused --reps 2 2>/dev/null | grep '^{' | tee -a gpurun_out/s18_strong_n2.jsonl | cut -c1-400
