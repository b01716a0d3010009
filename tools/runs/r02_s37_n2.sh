# Round 2, call 37 (2 GPUs): bench N=2, solo tile reported for the slowest GPU of the job
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29581"
timeout 200 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-legs --no-cpu > gpurun_out/r02s37_bench_n2.json 2> gpurun_out/r02s37_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s37_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['same_tile_single_gpu'], d['value']/(2*d['same_tile_single_gpu']['value']))
PY
tail -2 gpurun_out/r02s37_bench_n2.err | cut -c1-300
