mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ppcg_pair" ) > gpurun_out/s29_pytest.log 2>&1
tail -4 gpurun_out/s29_pytest.log | cut -c1-200
timeout 40 python scratch/ppcg_pair_ab.py 2>&1 | grep "^\[ppair\]\|rror" | tee gpurun_out/s29_ppcg_pair_ab.log
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s29_bench_n1.json 2>gpurun_out/s29_bench.err; cut -c1-160 gpurun_out/s29_bench_n1.json
