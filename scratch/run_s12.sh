python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s12_pytest.log
python scratch/sweep7.py 2>&1 | tee gpurun_out/s12_sweep7.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715"
for pdl in 0 1; do
$TR tools/config_bench.py --solver cg --global 4096 --max-iters 1000 --comm fused --opt use_pdl=$pdl 2>/dev/null | grep '^{' | tee -a gpurun_out/s12_cb2.jsonl
$TR tools/config_bench.py --solver cheby --global 2048 --max-iters 2000 --comm fused --opt use_pdl=$pdl 2>/dev/null | grep '^{' | tee -a gpurun_out/s12_cb2.jsonl
done
$TR bench.py --gpus 2 --steps 2 --warmup 3 --tile 8192 2>gpurun_out/s12_bench2.err | tee gpurun_out/s12_bench2.json
