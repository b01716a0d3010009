# Round 2, call 20 (1 GPU): ncu launch list + full capture of the round-2 build's CG kernels through the bench command; sanitizer passes
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02s20_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-legs > gpurun_out/r02s20_ncu_launch.log 2>&1
tail -2 gpurun_out/r02s20_ncu_launch.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cg_fused -s 40 -c 2 -o gpurun_out/r02s20_cg_full -f python bench.py --steps 1 --warmup 1 --no-cpu --no-legs > gpurun_out/r02s20_ncu_full.log 2>&1
tail -2 gpurun_out/r02s20_ncu_full.log | cut -c1-200
SANITIZE_TOOLS="memcheck racecheck" SANITIZE_TIMEOUT=240 bash tools/sanitize.sh gpurun_out 2>&1 | tail -4
ls -la gpurun_out | tail -6
