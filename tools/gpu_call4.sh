cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/ab4; mkdir -p $O
timeout -s KILL 600 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -5 $O/pytest.full
PV="timeout -s KILL 120 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py"
SHORT="fused_all_scopes_host or device_batch or saturation_solid or batch_order or tall_and_wide or tiles_add_up"
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
for n in w20 w20_straight w21 w22 w23 w23_nopipe w18 w20_nodefer; do
  v=variants_tmp/$n.so; [ -f $v ] || continue
  SCOPE_LIB=$PWD/$v $PV -k "$SHORT" > $O/pytest_$n.full 2>&1; rc=$?
  echo "exit $rc" >> $O/pytest_$n.full
  SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --width 1920 --height 1080 > $O/${n}_1080p.json 2>/dev/null
done
for f in $O/pytest_*.full; do echo "$f: $(tail -2 $f | tr '\n' ' ')"; done
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
bash tools/run_ncu.sh r02a 64
