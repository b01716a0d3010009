mkdir -p gpurun_out
timeout 600 python tools/ab/pair_stages_ab.py > gpurun_out/r02s22_pair_stages_ab.log 2>&1; cat gpurun_out/r02s22_pair_stages_ab.log
