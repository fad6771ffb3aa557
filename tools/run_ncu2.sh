cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-r01g}
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:scope_strip -s 3 -c 1 \
  -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --frames-per-gpu 64 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
