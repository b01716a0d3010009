mkdir -p gpurun_out
TL_COMM_TABLE_LEGACY=1 timeout 600 python tools/ab/multi_stress.py 24 > gpurun_out/r02s8_stress_legacy.log 2>&1; tail -30 gpurun_out/r02s8_stress_legacy.log | cut -c1-900
timeout 600 python tools/ab/multi_stress.py 24 > gpurun_out/r02s8_stress_fixed.log 2>&1; tail -30 gpurun_out/r02s8_stress_fixed.log | cut -c1-900
