# Multi-GPU runner (run under `gpurun --gpus N`, N = 2 / 4 / 8; charged N x the box time, so it is short):
#   1. the N-GPU tests (NCCL tile sharding incl. the all-gather column mode, frame sharding, the peer-memory reduce and
#      the strip kernel's direct column stores over real NVLink mappings);
#   2. BASELINE config 4 (one 7680x4320 frame, luma waveform) at 1 rank and at N ranks: row / column bands x
#      nccl | peers (two-shot) | peers-one-shot | nvls | nvls-one-shot, CUDA-graphed, two frames in flight;
#   3. the headline batch workload at N ranks (frame-sharded, weak scaling; the line carries e2e with the concurrent
#      H2D ceiling and the config4 sub-record).
# Everything lands in gpurun_out/mg$N as it finishes.   usage: bash tools/run_multigpu.sh N [quick]
cd $GRAFT_REPO_ROOT
N=${1:-2}
O=gpurun_out/mg$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29611
timeout -s KILL 300 python -m pytest -x -q -m gpu tests --deselect tests/test_gpu_multirank.py > $O/pytest_1gpu.full 2>&1; tail -3 $O/pytest_1gpu.full
timeout -s KILL 500 python -m pytest -x -q -s -m gpu tests/test_gpu_multirank.py > $O/pytest.full 2>&1
echo "exit $?" >> $O/pytest.full
for bands in rows cols; do
  timeout -s KILL 120 python bench.py --workload roi-tiled-8k --bands $bands --reduce peers --graph --steps 200 --warmup 20 > $O/tiled_${bands}_n1.json 2>$O/tiled_${bands}_n1.err
  for red in peers nccl peers-one-shot nvls nvls-one-shot; do
    [ "$2" = "quick" ] && [ $red != peers ] && [ $red != nccl ] && continue
    $TR --master-port $((PORT++)) bench.py --gpus $N --workload roi-tiled-8k --bands $bands --reduce $red --graph --steps 200 --warmup 20 \
      > $O/tiled_${bands}_${red}_n$N.json 2>$O/tiled_${bands}_${red}_n$N.err
  done
done
$TR --master-port $((PORT++)) bench.py --gpus $N --steps 10 --warmup 3 > $O/batch_n$N.json 2>$O/batch_n$N.err
echo "== pytest"; tail -4 $O/pytest.full
for f in $O/*.json; do echo $f $(python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), d.get('ms_per_step'), d.get('graph'), (d.get('parity') or {}).get('mismatches'))" 2>&1 | tail -1); done
