# Round 2, call 38 (1 GPU): ncu launch list (+ full capture of three launches) of the HEAD build's CG kernels through the bench command
mkdir -p gpurun_out
timeout 75 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02s38_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-legs > gpurun_out/r02s38_ncu_launch.log 2>&1
tail -1 gpurun_out/r02s38_ncu_launch.log | cut -c1-200
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_cg_fused -s 40 -c 3 -o gpurun_out/r02s38_cg_full -f python bench.py --steps 1 --warmup 1 --no-cpu --no-legs > gpurun_out/r02s38_ncu_full.log 2>&1
tail -1 gpurun_out/r02s38_ncu_full.log | cut -c1-200
ls -la gpurun_out | grep r02s38
