mkdir -p gpurun_out
TEALEAF_B200_OPTS=xchg_deferred=1 timeout 900 python tools/ab/multi_stress.py 500 > gpurun_out/r02s15_stress_deferred_touch.log 2>&1; grep -E "FAILED|failures" gpurun_out/r02s15_stress_deferred_touch.log | cut -c1-700
TL_NO_TOUCH=1 TEALEAF_B200_OPTS=xchg_deferred=1 timeout 900 python tools/ab/multi_stress.py 500 > gpurun_out/r02s15_stress_deferred_notouch.log 2>&1; grep -E "FAILED|failures" gpurun_out/r02s15_stress_deferred_notouch.log | cut -c1-700
