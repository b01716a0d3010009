# Round 2, call 34 (1 GPU): rows per warp task with the lazy-u CG loop at 4096^2
mkdir -p gpurun_out
timeout 300 python tools/ab/chunk_rows_sweep.py > gpurun_out/r02s34_chunk_rows_sweep.log 2>&1
cat gpurun_out/r02s34_chunk_rows_sweep.log | cut -c1-300
