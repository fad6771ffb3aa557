cd $GRAFT_REPO_ROOT
O=gpurun_out/ab3; mkdir -p $O
PT="timeout -s KILL 200 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_shim.py"
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$PT > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -8 $O/pytest.full > $O/pytest.log
$B > $O/new_mixed.json 2>$O/new_mixed.err
for v in nofaddr noemit chunk20 nodefer base; do
  SCOPE_LIB=$PWD/variants_tmp/$v.so $B > $O/${v}_mixed.json 2>/dev/null
done
for c in random natural solid ramp; do
  $B --content $c > $O/new_$c.json 2>/dev/null
done
$B --scopes wave > $O/new_waveonly.json 2>/dev/null
$B --scopes vscope > $O/new_vsonly.json 2>/dev/null
$B --width 1920 --height 1080 > $O/new_1080p.json 2>/dev/null
echo "== pytest";  cat $O/pytest.log
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
