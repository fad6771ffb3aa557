cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-c7}; mkdir -p $O
timeout -s KILL 900 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -4 $O/pytest.full
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
$B > $O/intree_mixed.json 2>$O/intree_mixed.err
for c in random ramp solid natural ui; do $B --content $c > $O/intree_$c.json 2>/dev/null; done
$B --width 1920 --height 1080 > $O/intree_1080p.json 2>/dev/null
$B --width 7680 --height 4320 --frames-per-gpu 16 > $O/intree_8k.json 2>/dev/null
for v in variants_tmp/*.so; do
  n=$(basename $v .so)
  SCOPE_LIB=$PWD/$v timeout -s KILL 200 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_headline_parity.py > $O/pytest_$n.full 2>&1; echo "exit $?" >> $O/pytest_$n.full
  SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content random > $O/${n}_random.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --width 1920 --height 1080 > $O/${n}_1080p.json 2>/dev/null
done
for f in $O/pytest_*.full; do echo "$f: $(tail -2 $f | tr '\n' ' ')"; done
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['parity']['mismatches'])" 2>&1 | tail -1); done
