# Round 2, calls 26-29 (1 GPU): lazy u update of CG kernel A -- bit-identity tests, then the A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tiled_one_gpu.py -m gpu -q -k "lazy_u or kernel_a_tma or test_solver_parity or golden" ) > gpurun_out/r02s29_pytest_lazy.log 2>&1
tail -8 gpurun_out/r02s29_pytest_lazy.log | cut -c1-300
timeout 600 python tools/ab/lazy_u_ab.py > gpurun_out/r02s29_lazy_u_ab.log 2>&1
cat gpurun_out/r02s29_lazy_u_ab.log | cut -c1-400
