# background-lane block A/B: HEAD build / noinline cold block without the skip / with it; mixed, ui, natural, solid, 1080p
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-c40}; mkdir -p $O
timeout -s KILL 600 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_headline_parity.py > $O/pytest.txt 2>&1; tail -2 $O/pytest.txt
B="timeout -s KILL 100 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
for v in variants_tmp/*.so; do
  n=$(basename $v .so)
  SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content ui > $O/${n}_ui.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content natural > $O/${n}_natural.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content solid > $O/${n}_solid.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content random > $O/${n}_random.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --width 1920 --height 1080 > $O/${n}_1080p.json 2>/dev/null
done
for f in $O/*.json; do echo $(basename $f) $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['parity']['mismatches'])" 2>&1 | tail -1); done
