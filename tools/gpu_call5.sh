cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/c5; O=gpurun_out/c5
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt 2>&1
timeout -s KILL 900 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -4 $O/pytest.full
timeout -s KILL 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config4 --width 1920 --height 1080 > $O/bench_1080p.json 2>/dev/null
bash tools/run_ncu.sh r02b 64
