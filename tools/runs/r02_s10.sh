# Round 2, call 10 (1 GPU): full -m gpu suite at HEAD + default bench line + reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/r02s10_pytest_gpu.log 2>&1
tail -22 gpurun_out/r02s10_pytest_gpu.log | cut -c1-400
timeout 600 python bench.py > gpurun_out/r02s10_bench_n1.json 2> gpurun_out/r02s10_bench_n1.err
tail -2 gpurun_out/r02s10_bench_n1.err | cut -c1-300; cut -c1-700 gpurun_out/r02s10_bench_n1.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s10_smoke.log 2>&1; tail -6 gpurun_out/r02s10_smoke.log
