cd $GRAFT_REPO_ROOT
O=gpurun_out/final; mkdir -p $O
timeout -s KILL 200 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -6 $O/pytest.full > $O/pytest.log
timeout -s KILL 200 python bench.py > $O/bench_n1.json 2>$O/bench_n1.err
B="timeout -s KILL 60 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B --scopes wave > $O/waveonly.json 2>/dev/null
$B --scopes hist > $O/histonly.json 2>/dev/null
timeout -s KILL 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scope_strip|finalize|hist_max" -s 9 -c 24 --csv \
  --log-file $O/launches.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > $O/launches.log 2>&1
$B --scopes hist,wave > $O/histwave.json 2>/dev/null
$B --scopes vscope > $O/vsonly.json 2>/dev/null
$B --width 7680 --height 4320 --frames-per-gpu 16 > $O/8k.json 2>/dev/null
$B --width 1920 --height 1080 > $O/1080p.json 2>/dev/null
timeout -s KILL 100 python bench.py --impl reference --steps 3 --warmup 1 > $O/reference.json 2>/dev/null
echo "== pytest";  cat $O/pytest.log
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), d.get('roofline') and round(d['roofline']['frac'],4), d.get('e2e') and round(d['e2e']['value']), d.get('cpu_baseline') and round(d['cpu_baseline']['value']))" 2>&1 | tail -1); done
