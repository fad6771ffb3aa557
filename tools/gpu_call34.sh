# diagnostic builds of the current v3 kernel (what do the end-of-strip write-out, its barriers, the flush, the loader cost?)
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-c34}; mkdir -p $O
B="timeout -s KILL 100 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
export SCOPE_BENCH_DIAGNOSTIC=1
$B > $O/shipped_mixed.json 2>/dev/null
$B --content random > $O/shipped_random.json 2>/dev/null
$B --content natural > $O/shipped_natural.json 2>/dev/null
for v in variants_tmp/*.so; do
  n=$(basename $v .so)
  SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content random > $O/${n}_random.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content natural > $O/${n}_natural.json 2>/dev/null
done
for f in $O/*.json; do echo $(basename $f) $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['parity']['mismatches'])" 2>&1 | tail -1); done
