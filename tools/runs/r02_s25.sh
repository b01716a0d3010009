# Round 2, call 25 (1 GPU): HEAD validation -- full GPU suite (new CLI test included), smoke, reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02s25_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02s25_pytest_gpu.log | cut -c1-300
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02s25_smoke.log 2>&1
tail -3 gpurun_out/r02s25_smoke.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02s25_bench_ref.json 2> gpurun_out/r02s25_bench_ref.err
cut -c1-600 gpurun_out/r02s25_bench_ref.json
timeout 300 python -m tealeaf_jl_b200.run -i decks/tea_bm_small.in -s ppcg -x 512 -y 512 --tea-out gpurun_out/r02s25_tea.out > gpurun_out/r02s25_cli.log 2>&1
cat gpurun_out/r02s25_tea.out | head -30
