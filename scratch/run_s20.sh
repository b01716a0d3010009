mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent or fused_cg or kernels_bitwise or determinism" ) > gpurun_out/s20_pytest.log 2>&1
tail -15 gpurun_out/s20_pytest.log
timeout 240 python scratch/persist_ab.py 2>&1 | tee gpurun_out/s20_persist_ab.log | cut -c1-260
TEALEAF_B200_OPTS=cg_persist=1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s20_bench_persist.json 2>gpurun_out/s20_bench_persist.err
cut -c1-200 gpurun_out/s20_bench_persist.json; tail -3 gpurun_out/s20_bench_persist.err
