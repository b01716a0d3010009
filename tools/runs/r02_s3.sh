# Round 2, call 3 (1 GPU): full -m gpu suite (tiled pair kernels default on, PPCG pairs on tiles, odd inner steps), bench N=1 with the new legs
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r02s3_pytest_gpu.log 2>&1
tail -45 gpurun_out/r02s3_pytest_gpu.log | cut -c1-600
timeout 600 python bench.py > gpurun_out/r02s3_bench_n1.json 2> gpurun_out/r02s3_bench_n1.err
tail -3 gpurun_out/r02s3_bench_n1.err; cut -c1-3000 gpurun_out/r02s3_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02s3_bench_ref.json 2> gpurun_out/r02s3_bench_ref.err
tail -3 gpurun_out/r02s3_bench_ref.err; cat gpurun_out/r02s3_bench_ref.json
