mkdir -p gpurun_out
( timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cheby_pair or fused_cheby" ) > gpurun_out/s27_pytest.log 2>&1
tail -12 gpurun_out/s27_pytest.log
timeout 150 python scratch/pair_ab.py 2>&1 | grep "^\[pair\]\|Error\|error" | tee gpurun_out/s27_pair_ab.log
