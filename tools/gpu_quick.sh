cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-q}; mkdir -p $O
timeout -s KILL 300 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_headline_parity.py > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -2 $O/pytest.full
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
$B > $O/intree_mixed.json 2>/dev/null
$B --content random > $O/intree_random.json 2>/dev/null
$B --content natural > $O/intree_natural.json 2>/dev/null
$B --width 1920 --height 1080 > $O/intree_1080p.json 2>/dev/null
export SCOPE_BENCH_DIAGNOSTIC=1
for v in variants_tmp/*.so; do
  n=$(basename $v .so)
  SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --width 1920 --height 1080 > $O/${n}_1080p.json 2>/dev/null
done
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['parity']['mismatches'])" 2>&1 | tail -1); done
