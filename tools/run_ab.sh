# A/B runner for kernel work (run under gpurun, after `make variants` here).  In this order, so that a
# cut-off call keeps the most important results:
#   1. the full GPU test suite and one bench line (mixed content) for the in-tree library;
#   2. for every variants_tmp/*.so (selected through SCOPE_LIB): the parity tests, then - only if they
#      pass - one bench line; a variant whose name ends in _x was built with SCOPE_EXPERIMENT (its
#      two-plane surface-mode ring does not fit), so its surface-mode tests are left out;
#   3. unless "quick": per content, per scope and per frame size for the in-tree library.
# Everything lands in gpurun_out/ab as it finishes.
#   usage: bash tools/run_ab.sh [quick]
cd $GRAFT_REPO_ROOT
O=gpurun_out/ab; mkdir -p $O
PT="timeout -s KILL 300 python -m pytest -x -q -m gpu tests"
PV="timeout -s KILL 240 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py"
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$PT > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -6 $O/pytest.full > $O/pytest.log
$B > $O/new_mixed.json 2>$O/new_mixed.err
for v in $(ls variants_tmp/*.so 2>/dev/null); do
  n=$(basename $v .so); sel=""
  case $n in *_x) sel="not surface and not golden and not shim";; esac
  SCOPE_LIB=$PWD/$v $PV ${sel:+-k "$sel"} > $O/pytest_$n.full 2>&1; rc=$?
  echo "exit $rc" >> $O/pytest_$n.full
  if [ $rc -eq 0 ]; then
    SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
    SCOPE_LIB=$PWD/$v $B --content natural > $O/${n}_natural.json 2>/dev/null
    case $n in *ballot*) SCOPE_LIB=$PWD/$v $B --content ui > $O/${n}_ui.json 2>/dev/null;; esac
  fi
done
for n in deepring w8_deepring; do   # what the deeper rings are for: the passes with room to spare
  v=variants_tmp/$n.so; [ -f $v ] || continue
  SCOPE_LIB=$PWD/$v $B --scopes vscope > $O/${n}_vsonly.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --scopes wave > $O/${n}_waveonly.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --scopes hist,wave > $O/${n}_histwave.json 2>/dev/null
done
SCOPE_KERNEL=group $B > $O/group_mixed.json 2>/dev/null
if [ "$1" != "quick" ]; then
  for c in random natural solid ramp ui; do $B --content $c > $O/new_$c.json 2>/dev/null; done
  $B --scopes wave > $O/new_waveonly.json 2>/dev/null
  $B --scopes hist > $O/new_histonly.json 2>/dev/null
  $B --scopes hist,wave > $O/new_histwave.json 2>/dev/null
  $B --scopes vscope > $O/new_vsonly.json 2>/dev/null
  $B --width 1920 --height 1080 > $O/new_1080p.json 2>/dev/null
  $B --width 7680 --height 4320 --frames-per-gpu 16 > $O/new_8k.json 2>/dev/null
fi
# the open micro-benchmark questions of DESIGN 8.1 (build tools/ubench2 first; about a minute)
[ -x tools/ubench2 ] && bash tools/run_ubench2.sh > /dev/null 2>&1
echo "== pytest (in-tree)"; cat $O/pytest.log
for f in $O/pytest_*.full; do echo "$f: $(tail -2 $f | tr '\n' ' ')"; done
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
