cd $GRAFT_REPO_ROOT
O=gpurun_out/c19; mkdir -p $O
timeout -s KILL 900 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -4 $O/pytest.full
bash tools/run_config4.sh 1 c19
B="timeout -s KILL 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
$B > $O/mixed.json 2>/dev/null
$B --scopes vscope > $O/vs_only.json 2>/dev/null
$B --scopes wave > $O/wave_only.json 2>/dev/null
$B --scopes hist > $O/hist_only.json 2>/dev/null
timeout 200 python bench.py --workload stream-vscope-4k --steps 3 > $O/stream.json 2>$O/stream.err
for f in $O/mixed.json $O/vs_only.json $O/wave_only.json $O/hist_only.json; do echo $f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['parity']['mismatches'])" 2>&1 | tail -1); done
tail -c 700 $O/stream.json
