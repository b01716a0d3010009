# Round 2, call 1 (1 GPU): the never-run tiled Chebyshev pairs + ncu --set full of the kernels behind the round-1 claims
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02s1_smi.log 2>&1
( TL_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_tiled_one_gpu.py -m gpu -q -k tiled_cheby_pairs ) > gpurun_out/r02s1_pair_tiled_pytest.log 2>&1
tail -30 gpurun_out/r02s1_pair_tiled_pytest.log
# one capture per kernel: the 4th launch (3 warm-up launches inside tl_time_kernel)
for kk in cheby_pair:k_cheby_pair_ring ppcg_pair:k_ppcg_pair_ring ppcg_inner:k_ppcg_inner_ring jacobi_fused:k_jacobi_fused_ring cheby_fused:k_cheby_fused_ring; do
  name=${kk%%:*}; kern=${kk##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -o gpurun_out/r02s1_${name}_4096_full -f python tools/ncu_targets.py --n 4096 --reps 1 --kernels $name > gpurun_out/r02s1_ncu_${name}.log 2>&1
  tail -2 gpurun_out/r02s1_ncu_${name}.log
done
for kk in cg_fused_w:k_cg_fused_w_ring cg_fused_r:k_cg_fused_r; do
  name=${kk%%:*}; kern=${kk##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -o gpurun_out/r02s1_${name}_8192_full -f python tools/ncu_targets.py --n 8192 --reps 1 --kernels $name > gpurun_out/r02s1_ncu_${name}_8192.log 2>&1
  tail -2 gpurun_out/r02s1_ncu_${name}_8192.log
done
timeout 120 python tools/ncu_targets.py --n 4096 --reps 30 --kernels cheby_pair,ppcg_pair,ppcg_inner,jacobi_fused,cheby_fused,cg_fused_w,cg_fused_r > gpurun_out/r02s1_kernel_times_4096.log 2>&1
cat gpurun_out/r02s1_kernel_times_4096.log
ls -la gpurun_out | tail -12
