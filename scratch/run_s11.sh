#!/bin/bash
# 8 GPUs: parity on the 2x4 decomposition, the weak-scaling bench line, strong-scaling configs 2 and 3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29714"
nvidia-smi -L | wc -l > gpurun_out/s11_ngpus.txt
timeout 300 $TR tests/mgpu_check.py cg cheby ppcg jacobi 2>&1 | grep "mgpu" > gpurun_out/s11_mgpu8.log
timeout 400 $TR bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/s11_bench8.json 2> gpurun_out/s11_bench8.err
out=gpurun_out/s11_config_bench8.jsonl; : > $out
timeout 200 $TR tools/config_bench.py --solver cheby --global 4096 --max-iters 2000 --comm fused,nccl >> $out 2>gpurun_out/s11_cb.err
timeout 200 $TR tools/config_bench.py --solver ppcg --global 8192 --max-iters 4000 --comm fused >> $out 2>>gpurun_out/s11_cb.err
timeout 200 $TR tools/config_bench.py --solver cg --global 4096 --max-iters 1000 --comm fused >> $out 2>>gpurun_out/s11_cb.err
cat gpurun_out/s11_mgpu8.log; cat gpurun_out/s11_bench8.json; tail -3 gpurun_out/s11_bench8.err; cat $out; tail -3 gpurun_out/s11_cb.err
