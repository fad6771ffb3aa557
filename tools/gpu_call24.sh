cd $GRAFT_REPO_ROOT
O=gpurun_out/c24; mkdir -p $O
timeout -s KILL 900 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -3 $O/pytest.full
timeout -s KILL 600 python tools/gpu_debug_ui.py > $O/dbg_ui.txt 2>&1; grep -c "hist differ 0 total diff 0 wave differ 0 vscope differ 0" $O/dbg_ui.txt; grep -v "hist differ 0 total diff 0 wave differ 0 vscope differ 0" $O/dbg_ui.txt | head -5
B="timeout -s KILL 100 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
$B > $O/intree_mixed.json 2>/dev/null
for c in random ramp solid natural ui; do $B --content $c > $O/intree_$c.json 2>/dev/null; done
$B --width 1920 --height 1080 > $O/intree_1080p.json 2>/dev/null
$B --width 7680 --height 4320 --frames-per-gpu 16 > $O/intree_8k.json 2>/dev/null
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['parity']['mismatches'])" 2>&1 | tail -1); done
