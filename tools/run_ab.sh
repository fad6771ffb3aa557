# A/B runner for kernel work (run under gpurun).  GPU parity tests on the in-tree library (and on the
# row-group kernel), then one bench line for the in-tree library, for every variants_tmp/*.so
# (`make variants`; selected through SCOPE_LIB), per content and per scope.  Everything lands in
# gpurun_out/ab as it finishes, most important first, so a cut-off call keeps the early results.
#   usage: bash tools/run_ab.sh [quick]        quick = parity + mixed-content lines only
cd $GRAFT_REPO_ROOT
O=gpurun_out/ab; mkdir -p $O
PT="timeout -s KILL 300 python -m pytest -x -q -m gpu tests"
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$PT > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -6 $O/pytest.full > $O/pytest.log
$B > $O/new_mixed.json 2>$O/new_mixed.err
for v in $(ls variants_tmp/*.so 2>/dev/null); do
  SCOPE_LIB=$PWD/$v $B > $O/$(basename $v .so)_mixed.json 2>/dev/null
done
SCOPE_KERNEL=group $B > $O/group_mixed.json 2>/dev/null
if [ "$1" != "quick" ]; then
  for c in random natural solid ramp; do $B --content $c > $O/new_$c.json 2>/dev/null; done
  $B --scopes wave > $O/new_waveonly.json 2>/dev/null
  $B --scopes hist > $O/new_histonly.json 2>/dev/null
  $B --scopes hist,wave > $O/new_histwave.json 2>/dev/null
  $B --scopes vscope > $O/new_vsonly.json 2>/dev/null
  $B --width 1920 --height 1080 > $O/new_1080p.json 2>/dev/null
  $B --width 7680 --height 4320 --frames-per-gpu 16 > $O/new_8k.json 2>/dev/null
fi
echo "== pytest"; cat $O/pytest.log
for f in $O/*.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
