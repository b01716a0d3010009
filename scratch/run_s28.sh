mkdir -p gpurun_out
( time timeout 170 python -m pytest tests -m gpu -q ) > gpurun_out/s28_pytest.log 2>&1
tail -30 gpurun_out/s28_pytest.log | cut -c1-220
timeout 50 python scratch/ppcg_pair_ab.py 2>&1 | grep "^\[ppair\]\|rror" | tee gpurun_out/s28_ppcg_pair_ab.log
