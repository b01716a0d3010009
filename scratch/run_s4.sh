#!/bin/bash
# A/B of the multi-GPU data paths + single-GPU references, 2 GPUs
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
out=gpurun_out/s4_config_bench.jsonl
: > $out
python tools/config_bench.py --solver cg --global 4096 --max-iters 1000 >> $out 2>gpurun_out/s4.err
python tools/config_bench.py --solver cheby --global 4096 --max-iters 2000 >> $out 2>>gpurun_out/s4.err
python tools/config_bench.py --solver ppcg --global 8192 --max-iters 200 >> $out 2>>gpurun_out/s4.err
python tools/config_bench.py --solver cg --tile 8192 --max-iters 300 >> $out 2>>gpurun_out/s4.err
$TR tools/config_bench.py --solver cg --global 4096 --max-iters 1000 >> $out 2>>gpurun_out/s4.err
$TR tools/config_bench.py --solver cheby --global 4096 --max-iters 2000 >> $out 2>>gpurun_out/s4.err
$TR tools/config_bench.py --solver ppcg --global 8192 --max-iters 200 >> $out 2>>gpurun_out/s4.err
$TR tools/config_bench.py --solver cg --tile 8192 --max-iters 300 >> $out 2>>gpurun_out/s4.err
$TR tools/config_bench.py --solver cg --global 1024 --max-iters 2000 >> $out 2>>gpurun_out/s4.err
python tools/config_bench.py --solver cg --global 1024 --max-iters 2000 >> $out 2>>gpurun_out/s4.err
cat $out
tail -5 gpurun_out/s4.err
