#!/bin/bash
# 4 GPUs: parity over 2x2, 4x1, 1x4 and the weak-scaling bench line
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -45 > gpurun_out/s7_pytest4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/s7_bench4.json 2> gpurun_out/s7_bench4.err
cat gpurun_out/s7_pytest4.log | grep -v "^$" | tail -45
cat gpurun_out/s7_bench4.json
tail -3 gpurun_out/s7_bench4.err
