# A/B runner for kernel work (run under gpurun, after `make variants` here).  In this order, so that a
# cut-off call keeps the most important results:
#   1. the full GPU test suite and one bench line (mixed content) for the in-tree library;
#   2. for the variants (variants_tmp/NAME.so, selected through SCOPE_LIB), most promising first: a short
#      GPU parity run - every variant already passed tests/test_kernel_emulation.py on the CPU, this is the
#      check that the hardware agrees - then, only if it passes, one bench line on the mixed batch;
#      names ending in _x were built with SCOPE_EXPERIMENT (their two-plane surface-mode ring does not fit);
#   3. unless "quick": natural / ui content for the variants, then per content, per scope and per frame
#      size for the in-tree library, then the micro-benchmarks of tools/ubench2.cu (if built).
# Everything lands in gpurun_out/ab as it finishes.  About 25 s of GPU time per variant.
#   usage: bash tools/run_ab.sh [quick] [variant names...]
cd $GRAFT_REPO_ROOT
O=gpurun_out/ab; mkdir -p $O
QUICK=0; [ "$1" = "quick" ] && { QUICK=1; shift; }
ORDER="dephase wide_dephase wide_straight_dephase wide wide_straight w16n8_immcoef_r120_x w16n8_r120_x w16n8_straight_immcoef_r120_x w16n8_straight_r120_x immcoef w8_straight w12n8_straight_x w8 w12n8_x straight w16n6_straight_r120_x w12n6_x \
       w16n6_straight_x w16n6_x r120 w8_dephase ballot w8_straight_ballot deepring w8_deepring nopipe rawflat nofaddr nodefer"
[ $# -gt 0 ] && ORDER="$*"
PT="timeout -s KILL 300 python -m pytest -x -q -m gpu tests"
PV="timeout -s KILL 120 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_properties.py"
SHORT="fused_all_scopes_host or device_batch or saturation_solid or batch_order or tall_and_wide or tiles_add_up"
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
$PT > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -6 $O/pytest.full > $O/pytest.log
$B > $O/new_mixed.json 2>$O/new_mixed.err
PASSED=""
for n in $ORDER; do
  v=variants_tmp/$n.so; [ -f $v ] || continue
  SCOPE_LIB=$PWD/$v $PV -k "$SHORT" > $O/pytest_$n.full 2>&1; rc=$?
  echo "exit $rc" >> $O/pytest_$n.full
  if [ $rc -eq 0 ]; then
    PASSED="$PASSED $n"
    SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  fi
done
if [ $QUICK -eq 0 ]; then
  for n in $PASSED; do
    v=variants_tmp/$n.so
    SCOPE_LIB=$PWD/$v $B --content natural > $O/${n}_natural.json 2>/dev/null
    case $n in *ballot*) SCOPE_LIB=$PWD/$v $B --content ui > $O/${n}_ui.json 2>/dev/null;; esac
    case $n in *deepring*)   # what the deeper rings are for: the passes with room to spare
      SCOPE_LIB=$PWD/$v $B --scopes vscope > $O/${n}_vsonly.json 2>/dev/null
      SCOPE_LIB=$PWD/$v $B --scopes wave > $O/${n}_waveonly.json 2>/dev/null
      SCOPE_LIB=$PWD/$v $B --scopes hist,wave > $O/${n}_histwave.json 2>/dev/null;;
    esac
  done
  SCOPE_KERNEL=group $B > $O/group_mixed.json 2>/dev/null
  for c in random natural solid ramp ui; do $B --content $c > $O/new_$c.json 2>/dev/null; done
  $B --scopes wave > $O/new_waveonly.json 2>/dev/null
  $B --scopes hist > $O/new_histonly.json 2>/dev/null
  $B --scopes hist,wave > $O/new_histwave.json 2>/dev/null
  $B --scopes vscope > $O/new_vsonly.json 2>/dev/null
  $B --width 1920 --height 1080 > $O/new_1080p.json 2>/dev/null
  $B --width 7680 --height 4320 --frames-per-gpu 16 > $O/new_8k.json 2>/dev/null
  # the open micro-benchmark questions of DESIGN 8.1 (build tools/ubench2 first; about a minute)
  [ -x tools/ubench2 ] && bash tools/run_ubench2.sh > /dev/null 2>&1
fi
echo "== pytest (in-tree)"; cat $O/pytest.log
for f in $O/pytest_*.full; do echo "$f: $(tail -2 $f | tr '\n' ' ')"; done
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
