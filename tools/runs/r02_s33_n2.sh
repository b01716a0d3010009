# Round 2, call 33 (2 GPUs): bench N=2 at HEAD (solo tile measured like for like: same timestep, same iteration cap)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29581"
timeout 600 $TR bench.py --gpus 2 > gpurun_out/r02s33_bench_n2.json 2> gpurun_out/r02s33_bench_n2.err
cut -c1-300 gpurun_out/r02s33_bench_n2.json; tail -3 gpurun_out/r02s33_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s33_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['same_tile_single_gpu'], d['value']/(2*d['same_tile_single_gpu']['value']))
PY
