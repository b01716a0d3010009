mkdir -p gpurun_out
timeout 900 python tools/ab/multi_stress.py 120 > gpurun_out/r02s14_stress.log 2>&1; grep -E "FAILED|failures" gpurun_out/r02s14_stress.log | cut -c1-1800
TEALEAF_B200_OPTS=xchg_deferred=1 timeout 900 python tools/ab/multi_stress.py 120 > gpurun_out/r02s14_stress_deferred.log 2>&1; grep -E "FAILED|failures" gpurun_out/r02s14_stress_deferred.log | cut -c1-1800
