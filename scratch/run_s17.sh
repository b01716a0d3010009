# HEAD verification on one B200: full GPU suite, default bench line, launch list + full capture of the CG kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s17_smi.log 2>&1
( time timeout 480 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/s17_pytest.log 2>&1
tail -25 gpurun_out/s17_pytest.log
timeout 240 python bench.py > gpurun_out/s17_bench_n1.json 2> gpurun_out/s17_bench_n1.err
cat gpurun_out/s17_bench_n1.json | cut -c1-600
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/s17_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/s17_ncu_launch.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_cg_fused -s 40 -c 2 -o gpurun_out/s17_cg_full -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/s17_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
