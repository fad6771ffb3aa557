# compute-sanitizer over the shipped kernels (run under gpurun, one GPU): memcheck, racecheck (shared-memory
# hazards: the TMA ring, the column bins, the vectorscope bins), synccheck (named barriers / mbarriers).
# Logs land in gpurun_out/sanitizer/; summaries are copied to profiles/ by hand.   usage: bash tools/run_sanitizer.sh [TAG]
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
O=gpurun_out/sanitizer; mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > $O/${tool}_$TAG.log 2>&1
  echo "exit $?" >> $O/${tool}_$TAG.log
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver ok|exit ' $O/${tool}_$TAG.log | tr '\n' ' ')"
done
