python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/s8_pytest.log
tail -30 gpurun_out/s8_pytest.log
python scratch/sweep6.py 2>&1 | tee gpurun_out/s8_sweep6.log
