# Round 2, call 7 (1 GPU): multi-context flakiness hunt (tiles sharing the GPU), PDL A/B + boundary profile
mkdir -p gpurun_out
for rep in 1 2 3; do
  ( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_multi_context.py -m gpu -q -x -k "fused_cg or multi_context" 2>&1 | grep -E "passed|failed|TL_ERR" | cut -c1-2500 ) >> gpurun_out/r02s7_multi_rep.log 2>&1
done
cat gpurun_out/r02s7_multi_rep.log
for rep in 1 2; do
  ( CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_multi_context.py -m gpu -q -x -k "fused_cg or multi_context" 2>&1 | grep -E "passed|failed|TL_ERR" | cut -c1-2500 ) >> gpurun_out/r02s7_multi_rep_conn8.log 2>&1
done
cat gpurun_out/r02s7_multi_rep_conn8.log
timeout 900 python tools/ab/pdl_ab.py > gpurun_out/r02s7_pdl_ab.log 2>&1
cat gpurun_out/r02s7_pdl_ab.log
