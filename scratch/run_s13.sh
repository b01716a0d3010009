python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "halo_depths" 2>&1 | tail -5 | tee gpurun_out/s13_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err; cat gpurun_out/s13_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/s13_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/s13_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cg_fused -s 40 -c 2 -o gpurun_out/s13_cg_full -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/s13_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_cheby_fused_ring|k_ppcg_inner_ring|k_ppcg_ur_sd" -s 30 -c 6 -o gpurun_out/s13_chpp_full -f python scratch/sweep4b.py > gpurun_out/s13_ncu_chpp.log 2>&1
ls -la gpurun_out/*.ncu-rep
