mkdir -p gpurun_out
timeout 900 python tools/ab/multi_stress.py 80 > gpurun_out/r02s9_stress.log 2>&1; grep -E "FAILED|failures" gpurun_out/r02s9_stress.log | cut -c1-900
