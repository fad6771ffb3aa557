# Profiles (run under gpurun, single GPU): one `ncu --set full` capture of the dominant kernel and the
# launch list of a short bench run.  Reports stay in gpurun_out/; summaries go to profiles/ by hand
# (read the .ncu-rep here with `ncu -i ... --page raw --csv` / `--page source --csv`).
#   usage: bash tools/run_ncu.sh [TAG] [FRAMES]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r02a}; FRAMES=${2:-64}
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"scope_strip|scope_fused" -s 3 -c 1 \
  -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --frames-per-gpu $FRAMES --no-e2e --no-cpu-baseline --no-config4 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scope_strip|scope_fused|finalize|hist_max" -s 9 -c 24 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-config4 > gpurun_out/launches_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
