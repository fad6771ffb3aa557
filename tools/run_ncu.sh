# one `ncu --set full` capture of the dominant kernel (single GPU) + the launch list of a short
# bench run; reports are left in gpurun_out/ (summaries go to profiles/ by hand)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r01f}
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:scope_strip -s 3 -c 1 \
  -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --frames-per-gpu 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scope_strip|finalize|hist_max" -s 9 -c 24 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_$TAG.csv
