# First GPU calls of the next round (what this round's budget no longer covered).  Run pieces with gpurun.
mkdir -p gpurun_out
# 1. the EXPERIMENTAL tiled Chebyshev pairs (never run on a GPU): correctness on one GPU through the thread harness
TL_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_tiled_one_gpu.py -m gpu -x -q -k tiled_cheby_pairs 2>&1 | tail -15 | tee gpurun_out/n1_pair_tiled_pytest.log
# 2. ncu of the pair kernels (launch list + full capture): python scratch/pair_ab.py drives k_cheby_pair_ring; scratch/ppcg_pair_ab.py k_ppcg_pair_ring
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_cheby_pair_ring|k_ppcg_pair_ring" -s 4 -c 2 -o gpurun_out/n1_pair_full -f python scratch/pair_ab.py > gpurun_out/n1_ncu_pair.log 2>&1
# 3. (gpurun --gpus 8) config 3 strong scaling with and without the tiled pairs:
#    python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 tools/config_bench.py --solver cheby --global 4096 --max-iters 2000 --comm fused
#    ... --opt pair_tiled=1
#    and depth-k PPCG: --solver ppcg --global 8192 --max-iters 4000 --halo-depth 4 --ppcg-halo-depth 1,2,4 --comm fused --reps 1
