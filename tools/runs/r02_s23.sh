mkdir -p gpurun_out
timeout 600 python tools/ab/pair_threads_ab.py > gpurun_out/r02s23_pair_threads_ab.log 2>&1; cat gpurun_out/r02s23_pair_threads_ab.log
