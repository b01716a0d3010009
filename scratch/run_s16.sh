python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/s16_pytest.log
for rep in 1 2; do
TL_OPTS=kotf=0 python scratch/ab_kernels.py kotf0 2>&1 | grep "^\[" | tee -a gpurun_out/s16_ab.log
TL_OPTS=kotf=1 python scratch/ab_kernels.py kotf1 2>&1 | grep "^\[" | tee -a gpurun_out/s16_ab.log
done
TL_OPTS=kotf=1,ring_stages=4 python scratch/ab_kernels.py kotf1_ring4 2>&1 | grep "^\[" | tee -a gpurun_out/s16_ab.log
