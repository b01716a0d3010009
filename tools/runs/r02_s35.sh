# Round 2, call 35 (1 GPU): stencil chunks of 12 rows by default -- full GPU suite, smoke, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02s35_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02s35_pytest_gpu.log | cut -c1-300
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02s35_smoke.log 2>&1; tail -2 gpurun_out/r02s35_smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r02s35_bench_n1.json 2> gpurun_out/r02s35_bench_n1.err
cut -c1-300 gpurun_out/r02s35_bench_n1.json
