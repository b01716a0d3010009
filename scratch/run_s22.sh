mkdir -p gpurun_out
sed -i 's/for tool in memcheck racecheck synccheck/for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}/' tools/sanitize.sh
SANITIZE_TOOLS="memcheck racecheck" SANITIZE_TIMEOUT=90 bash tools/sanitize.sh gpurun_out
tail -4 gpurun_out/sanitize_memcheck.log; tail -4 gpurun_out/sanitize_racecheck.log
