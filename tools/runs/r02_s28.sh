# Round 2, call 28 (1 GPU): is kernel A sensitive to the L1 / shared-memory split?
mkdir -p gpurun_out
( timeout 200 python tools/ab/carveout_probe.py 4096; TEALEAF_B200_CARVEOUT=1 timeout 200 python tools/ab/carveout_probe.py 4096 ) > gpurun_out/r02s28_carveout_probe.log 2>&1
cat gpurun_out/r02s28_carveout_probe.log | cut -c1-400
