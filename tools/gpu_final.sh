cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-final}; mkdir -p $O
timeout -s KILL 900 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -3 $O/pytest.full
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout -s KILL 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
B="timeout -s KILL 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
$B --width 1920 --height 1080 > $O/bench_1080p.json 2>/dev/null
$B --width 7680 --height 4320 --frames-per-gpu 16 > $O/bench_8k.json 2>/dev/null
for c in random ramp solid natural ui; do $B --content $c > $O/bench_4k_$c.json 2>/dev/null; done
for sc in hist wave vscope hist,wave; do $B --scopes $sc > $O/bench_4k_scopes_$(echo $sc | tr , _).json 2>/dev/null; done
timeout 200 python bench.py --workload stream-vscope-4k --steps 3 > $O/bench_stream_vscope.json 2>/dev/null
bash tools/run_ncu.sh ${2:-r02d} 64
