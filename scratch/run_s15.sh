python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cg or kernels_bitwise or golden or determinism or full_size" 2>&1 | tail -3 | tee gpurun_out/s15_pytest.log
python scratch/ab_kernels.py fastpath 2>&1 | grep "^\[" | tee gpurun_out/s15_ab.log
python scratch/ab_kernels.py fastpath 2>&1 | grep "^\[" | tee -a gpurun_out/s15_ab.log
