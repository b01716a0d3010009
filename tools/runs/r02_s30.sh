# Round 2, call 30 (1 GPU): lazy u update on by default -- DRAM bytes of its two kinds of launch (ncu, 3 metrics) at 4096^2 and 16384^2,
# full GPU suite, bench, one more A/B of the u-updating launch at 2 vs 3 CTAs per SM
mkdir -p gpurun_out
for n in 4096 16384; do
  for name in cg_fused_w cg_fused_w_odd; do
    timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_cg_fused_w_ring -s 3 -c 1 --csv --log-file gpurun_out/r02s30_${name}_${n}.csv python tools/ncu_targets.py --n $n --reps 1 --kernels $name > gpurun_out/r02s30_ncu_${name}_${n}.log 2>&1
    tail -1 gpurun_out/r02s30_ncu_${name}_${n}.log | cut -c1-200; tail -3 gpurun_out/r02s30_${name}_${n}.csv | cut -c1-300
  done
done
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02s30_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02s30_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02s30_bench_n1.json 2> gpurun_out/r02s30_bench_n1.err
cut -c1-300 gpurun_out/r02s30_bench_n1.json
( timeout 300 python tools/ab/lazy_u_ab.py 4096:1500; timeout 300 python tools/ab/lazy_u_ab.py 4096:1500 ) > gpurun_out/r02s30_lazy_heavy_ctas_ab.log 2>&1
grep "lazy_u': 1" gpurun_out/r02s30_lazy_heavy_ctas_ab.log | cut -c1-330
