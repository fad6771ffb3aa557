# one `ncu --set full` capture of the dominant kernel (single GPU), report left in gpurun_out/
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:scope_strip -s 3 -c 1 \
  -o gpurun_out/prof_r01e -f python bench.py --steps 1 --warmup 3 --frames-per-gpu 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r01e.log 2>&1
tail -3 gpurun_out/ncu_r01e.log
ls -la gpurun_out/*.ncu-rep
