# A/B runner used during kernel work: GPU tests on the in-tree library, then bench lines for
# the in-tree library and every variants_tmp/*.so (selected through SCOPE_LIB).
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/ab; mkdir -p $O
( timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest.log 2>&1
B="timeout -s KILL 120 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for c in mixed random natural solid; do
  $B --content $c > $O/new_$c.json 2>$O/err_new_$c.log
done
SCOPE_SPLIT=1 $B --content mixed > $O/newsplit_mixed.json 2>/dev/null
$B --content mixed --scopes vscope > $O/new_vsonly.json 2>/dev/null
for v in $(ls variants_tmp/*.so 2>/dev/null); do
  n=$(basename $v .so)
  for c in mixed random natural; do
    SCOPE_LIB=$PWD/$v $B --content $c > $O/${n}_$c.json 2>/dev/null
  done
  SCOPE_LIB=$PWD/$v $B --content mixed --scopes vscope > $O/${n}_vsonly.json 2>/dev/null
done
set +x
cat $O/pytest.log
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4))" 2>&1 | tail -1); done
