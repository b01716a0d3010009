# Round 2, call 13 (1 GPU): split exchange -- bit-identity on tiles sharing the GPU, multi-context with the option on
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_tiled_one_gpu.py -m gpu -q -k "split_exchange" ) > gpurun_out/r02s13_split_pytest.log 2>&1
tail -15 gpurun_out/r02s13_split_pytest.log | cut -c1-1500
( TEALEAF_B200_OPTS=xchg_deferred=1 timeout 900 python -m pytest tests/test_tiled_one_gpu.py tests/test_multi_context.py -m gpu -q -x ) > gpurun_out/r02s13_split_env_pytest.log 2>&1
tail -8 gpurun_out/r02s13_split_env_pytest.log | cut -c1-1500
