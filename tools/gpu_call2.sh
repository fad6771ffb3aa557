cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out

O=gpurun_out/ab3b; mkdir -p $O
PV="timeout -s KILL 120 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py"
SHORT="fused_all_scopes_host or device_batch or saturation_solid or batch_order or tall_and_wide or tiles_add_up"
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
for n in wide15 wide15_straight wide15_dephase w19_x w19_straight_x w23_x w23_nopipe_x w20_immcoef_x; do
  v=variants_tmp/$n.so; [ -f $v ] || continue
  SCOPE_LIB=$PWD/$v $PV -k "$SHORT" > $O/pytest_$n.full 2>&1; rc=$?
  echo "exit $rc" >> $O/pytest_$n.full
  SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content random > $O/${n}_random.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --width 1920 --height 1080 > $O/${n}_1080p.json 2>/dev/null
done
for f in $O/pytest_*.full; do echo "$f: $(tail -2 $f | tr '\n' ' ')"; done
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done

