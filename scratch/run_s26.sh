mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/s26_pytest.log 2>&1
tail -16 gpurun_out/s26_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/s26_smoke.log
timeout 240 python bench.py > gpurun_out/s26_bench_n1.json 2> gpurun_out/s26_bench_n1.err
cut -c1-300 gpurun_out/s26_bench_n1.json
