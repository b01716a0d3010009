# Round 2, call 17 (1 GPU): split exchange v2 (fast path) -- bit-identity on tiles sharing the GPU, whole tiled + multi-context suites with the option on
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_tiled_one_gpu.py -m gpu -q -k "split_exchange" ) > gpurun_out/r02s17_split_pytest.log 2>&1
tail -6 gpurun_out/r02s17_split_pytest.log | cut -c1-1500
( TEALEAF_B200_OPTS=xchg_deferred=1 timeout 900 python -m pytest tests/test_tiled_one_gpu.py tests/test_multi_context.py -m gpu -q ) > gpurun_out/r02s17_split_env_pytest.log 2>&1
tail -6 gpurun_out/r02s17_split_env_pytest.log | cut -c1-1500
