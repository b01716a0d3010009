# Round 2, call 39 (1 GPU, the last seconds of the budget): kernel-boundary micro-profile of the CG loop at HEAD, 4096^2
mkdir -p gpurun_out
timeout 40 python tools/config_bench.py --solver cg --global 4096 --max-iters 1500 --comm fused --reps 1 --prof 2> gpurun_out/r02s39.err | grep "^{" > gpurun_out/r02s39_boundary_profile_cg4096_n1.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02s39_boundary_profile_cg4096_n1.jsonl'):
    d=json.loads(l); print(d['us_per_sweep'], d['boundary_profile_us_per_kernel']['max_over_ranks'])
PY
