# last check of the committed tree: GPU tests, smoke, the default bench line and the reference arm
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-c43}; mkdir -p $O
timeout -s KILL 300 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -3 $O/pytest.full
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout -s KILL 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err; python -c "
import json; d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks'], 'e2e', round(d['e2e']['value']), d['parity']['mismatches'], d['gpu_launches'])"
