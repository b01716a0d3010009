# Round 2, call 36 (4 GPUs): bench N=4 at HEAD (2x2 tiles; lazy-u CG loop, 12-row chunks)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29581"
timeout 400 $TR bench.py --gpus 4 > gpurun_out/r02s36_bench_n4.json 2> gpurun_out/r02s36_bench_n4.err
cut -c1-300 gpurun_out/r02s36_bench_n4.json; tail -2 gpurun_out/r02s36_bench_n4.err | cut -c1-300
