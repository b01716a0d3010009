# Round 2, call 4 (1 GPU): tl_create_multi tests (tiles sharing the GPU), the 8192^2 property tests, pair_rows sweep
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_multi_context.py tests/test_gpu_measured_paths.py -m gpu -q -k "multi_context or c_host_with_one or full_size_8192" --durations=8 ) > gpurun_out/r02s4_pytest.log 2>&1
tail -40 gpurun_out/r02s4_pytest.log | cut -c1-1200
timeout 600 python tools/ab/pair_rows_sweep.py > gpurun_out/r02s4_pair_rows_sweep.log 2>&1
cat gpurun_out/r02s4_pair_rows_sweep.log
