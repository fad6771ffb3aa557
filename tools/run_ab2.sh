# Round-1 closing A/B (short GPU budget): parity of the new default build and of the row-group
# kernel, then bench lines for every build.  Everything is written to gpurun_out/ab2 as it
# finishes, most important first, so a cut-off call still leaves the early results.
cd $GRAFT_REPO_ROOT
O=gpurun_out/ab2; mkdir -p $O
date +%s > $O/t0
PT="timeout -s KILL 200 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py"
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$PT > $O/pytest_tile.full 2>&1; echo "exit $?" >> $O/pytest_tile.full; tail -8 $O/pytest_tile.full > $O/pytest_tile.log
$B > $O/tile_mixed.json 2>$O/tile_mixed.err
SCOPE_LIB=$PWD/variants_tmp/base.so $B > $O/base_mixed.json 2>/dev/null
SCOPE_KERNEL=group $PT > $O/pytest_group.full 2>&1; G=$?; echo "exit $G" >> $O/pytest_group.full; tail -8 $O/pytest_group.full > $O/pytest_group.log
if [ $G -eq 0 ]; then
  SCOPE_KERNEL=group $B > $O/group24_mixed.json 2>$O/group24_mixed.err
  for v in g27 g31 g23; do
    SCOPE_LIB=$PWD/variants_tmp/$v.so $B > $O/${v}_mixed.json 2>/dev/null
  done
fi
SCOPE_LIB=$PWD/variants_tmp/addswz.so $B > $O/addswz_mixed.json 2>/dev/null
for c in random natural solid ramp; do
  $B --content $c > $O/tile_$c.json 2>/dev/null
  [ $G -eq 0 ] && SCOPE_KERNEL=group $B --content $c > $O/group24_$c.json 2>/dev/null
done
( timeout -s KILL 200 python -m pytest -x -q -m gpu tests/test_gpu_shim.py 2>&1 | tail -5 ) > $O/pytest_shim.log 2>&1
date +%s > $O/t1
echo "== pytest tile";  cat $O/pytest_tile.log
echo "== pytest group"; cat $O/pytest_group.log
echo "== shim"; cat $O/pytest_shim.log
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
