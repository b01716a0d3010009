# Round 2, call 24 (1 GPU): DRAM bytes of the CG kernels at the 16384^2 tile of the weak-scaling runs (ncu, 3 metrics), final suite + bench at HEAD
mkdir -p gpurun_out
for kk in cg_fused_w:k_cg_fused_w_ring cg_fused_r:k_cg_fused_r; do
  name=${kk%%:*}; kern=${kk##*:}
  timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$kern -s 3 -c 1 --csv --log-file gpurun_out/r02s24_${name}_16384.csv python tools/ncu_targets.py --n 16384 --reps 1 --kernels $name > gpurun_out/r02s24_ncu_${name}_16384.log 2>&1
  tail -1 gpurun_out/r02s24_ncu_${name}_16384.log | cut -c1-200; tail -4 gpurun_out/r02s24_${name}_16384.csv | cut -c1-300
done
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02s24_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02s24_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02s24_bench_n1.json 2> gpurun_out/r02s24_bench_n1.err
cut -c1-400 gpurun_out/r02s24_bench_n1.json
