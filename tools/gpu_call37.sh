# bench.py after the clock-sampling change: the default line, the lines that had come back without samples, config 4 at N = 1
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-c37}; mkdir -p $O
timeout -s KILL 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 400 $O/bench_default.json
B="timeout -s KILL 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
$B --width 7680 --height 4320 --frames-per-gpu 16 > $O/bench_8k.json 2>$O/bench_8k.err
$B --scopes hist,wave > $O/bench_4k_scopes_hist_wave.json 2>$O/hw.err
$B --width 1920 --height 1080 > $O/bench_1080p.json 2>$O/1080.err
timeout -s KILL 120 python bench.py --workload roi-tiled-8k --graph --reduce peers --bands cols --in-flight 8 --emulate-world 4 --steps 400 > $O/cfg4_e4.json 2>$O/cfg4_e4.err
timeout 200 python bench.py --workload stream-vscope-4k --steps 3 > $O/bench_stream_vscope.json 2>$O/sv.err
for f in $O/*.json; do echo $(basename $f) $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), d.get('roofline',{}).get('frac'), d['clocks'], (d.get('config4') or {}).get('best'))" 2>&1 | tail -1); done
tail -3 $O/*.err
