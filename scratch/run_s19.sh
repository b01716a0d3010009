# N=2: depth-k (matrix-powers) PPCG timing A/B; max-iters chosen so that PPCG really runs (CG presteps end at 1018 / 2278 iterations)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29641"
timeout 100 $TR tools/config_bench.py --solver ppcg --global 4096 --max-iters 1600 --inner 10 --halo-depth 4 --ppcg-halo-depth 1,2,4,1,4 --comm fused --reps 1 2>/dev/null | grep '^{' | tee -a gpurun_out/s19_depthk_n2.jsonl | cut -c1-100,250-420
timeout 100 $TR tools/config_bench.py --solver ppcg --global 8192 --max-iters 2600 --inner 10 --halo-depth 4 --ppcg-halo-depth 1,4 --comm fused --reps 1 2>/dev/null | grep '^{' | tee -a gpurun_out/s19_depthk_n2.jsonl | cut -c1-100,250-420
