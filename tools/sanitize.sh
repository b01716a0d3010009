#!/bin/bash
# compute-sanitizer passes over every kernel family (race detection / sanitizers).
# Usage (on a GPU box): bash tools/sanitize.sh [outdir]      -- logs go to <outdir>/sanitize_<tool>.log
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
rc=0
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
  timeout ${SANITIZE_TIMEOUT:-150} $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py > "$OUT/sanitize_$tool.log" 2>&1
  r=$?
  echo "compute-sanitizer --tool $tool: exit $r; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_$tool.log" | tail -1)"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
