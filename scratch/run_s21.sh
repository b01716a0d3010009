mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent or b_ring or fused_cg or determinism" ) > gpurun_out/s21_pytest.log 2>&1
tail -8 gpurun_out/s21_pytest.log
timeout 300 python scratch/bring_ab.py 2>&1 | tee gpurun_out/s21_bring_ab.log | cut -c1-200
