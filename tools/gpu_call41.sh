# background lanes folded into the one cold block copy: A/B against the HEAD build, then the artefact set of the product build
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-c41}; mkdir -p $O
timeout -s KILL 600 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -3 $O/pytest.full
B="timeout -s KILL 100 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
for c in mixed natural ui solid random; do
  SCOPE_LIB=$PWD/variants_tmp/v3_head.so $B --content $c > $O/head_$c.json 2>/dev/null
  $B --content $c > $O/fold_$c.json 2>/dev/null
done
SCOPE_LIB=$PWD/variants_tmp/v3_head.so $B --width 1920 --height 1080 > $O/head_1080p.json 2>/dev/null
$B --width 1920 --height 1080 > $O/fold_1080p.json 2>/dev/null
$B --width 7680 --height 4320 --frames-per-gpu 16 > $O/fold_8k.json 2>/dev/null
for f in $O/*.json; do echo $(basename $f) $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['samples'], d['parity']['mismatches'])" 2>&1 | tail -1); done
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout -s KILL 500 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 200 $O/bench_default.json
bash tools/run_ncu.sh r02j 64
