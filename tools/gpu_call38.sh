# 2-GPU check of the final code: multi-rank tests, then the default bench line under torchrun (what the driver's scaling run does)
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-mg2b}; mkdir -p $O
timeout -s KILL 400 python -m pytest -x -q -s -m gpu tests/test_gpu_multirank.py > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -4 $O/pytest.full
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 > $O/batch_n2.json 2>$O/batch_n2.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('$O/batch_n2.json').read().strip().splitlines()[-1])
print(round(d['value']), d['roofline']['frac'], d['clocks'], 'e2e', d['e2e']['value'], {k: (round(v['value']), v['frames_in_flight'], v['parity'], v['clocks']['samples']) for k, v in d['config4'].items() if isinstance(v, dict) and 'value' in v and 'bands' in v and 'parity' in v}, d['config4'].get('best'))
"
tail -c 600 $O/batch_n2.err
