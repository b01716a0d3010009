# Round 2, call 2 (1 GPU): diagnostics for the two failing tiled-pair cases + the new measured-path parity tests
mkdir -p gpurun_out
( TL_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_tiled_one_gpu.py -m gpu -q -k "tiled_cheby_pairs and (2x2-129x67 or 3x3)" 2>&1 | grep -E "TeaLeafError|RuntimeError|passed|failed" ) > gpurun_out/r02s2_pair_tiled_diag.log 2>&1
cat gpurun_out/r02s2_pair_tiled_diag.log | cut -c1-1500
( time timeout 1500 python -m pytest tests/test_gpu_measured_paths.py -m gpu -q --durations=15 ) > gpurun_out/r02s2_measured_paths_pytest.log 2>&1
tail -40 gpurun_out/r02s2_measured_paths_pytest.log
