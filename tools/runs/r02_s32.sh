# Round 2, call 32 (1 GPU): compute-sanitizer memcheck + racecheck at HEAD (lazy u kernels included)
mkdir -p gpurun_out
SANITIZE_TOOLS="memcheck racecheck" SANITIZE_TIMEOUT=200 bash tools/sanitize.sh gpurun_out
for t in memcheck racecheck; do cp gpurun_out/sanitize_$t.log gpurun_out/r02s32_sanitize_$t.log; tail -4 gpurun_out/sanitize_$t.log | cut -c1-200; done
