# Round 2, call 6 (1 GPU): TMA flavour of kernel A (bit-identity + A/B), faster last-block partial sums (A/B via bench), property tests
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_measured_paths.py tests/test_multi_context.py -m gpu -q -k "a_tma or full_size or multi_context_solves or run_to_run or fused_cg" --durations=6 ) > gpurun_out/r02s6_pytest.log 2>&1
tail -30 gpurun_out/r02s6_pytest.log | cut -c1-700
timeout 600 python tools/ab/tma_ab.py > gpurun_out/r02s6_tma_ab.log 2>&1
cat gpurun_out/r02s6_tma_ab.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-legs > gpurun_out/r02s6_bench_n1_short.json 2> gpurun_out/r02s6_bench.err
cut -c1-1600 gpurun_out/r02s6_bench_n1_short.json
