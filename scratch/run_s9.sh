python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/s9_pytest.log
tail -5 gpurun_out/s9_pytest.log
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/s9_bench1.json 2> gpurun_out/s9_bench1.err; cat gpurun_out/s9_bench1.json; tail -3 gpurun_out/s9_bench1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s9_bench2.json 2> gpurun_out/s9_bench2.err; cat gpurun_out/s9_bench2.json; tail -3 gpurun_out/s9_bench2.err
