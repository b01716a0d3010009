for rep in 1 2; do
TEALEAF_B200_LIB=$PWD/scratch/libtealeaf_old.so python scratch/ab_kernels.py old
python scratch/ab_kernels.py new
done 2>&1 | grep "^\[" | tee gpurun_out/s10_ab.log
