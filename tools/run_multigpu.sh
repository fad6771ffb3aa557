# Multi-GPU runner (run under `gpurun --gpus N`, N = 2 / 4 / 8; charged N x the box time, so it is short):
#   1. the N-GPU tests (NCCL tile sharding, frame sharding, the peer-memory reduce over real NVLink mappings);
#   2. the headline batch workload at N ranks (frame-sharded, weak scaling) and at 1 rank on the same box;
#   3. BASELINE config 4 (one 7680x4320 frame, luma waveform, row / column bands) with the three cross-rank
#      steps: NCCL all-reduce + clamp, the fused peer-memory kernel two-shot and one-shot, and its NVLS form.
# Everything lands in gpurun_out/mg$N as it finishes.   usage: bash tools/run_multigpu.sh N
cd $GRAFT_REPO_ROOT
N=${1:-2}
O=gpurun_out/mg$N; mkdir -p $O
TR="timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29611
timeout -s KILL 400 python -m pytest -x -q -m gpu tests/test_gpu_multirank.py tests/test_gpu_z_peer_reduce.py > $O/pytest.full 2>&1
echo "exit $?" >> $O/pytest.full
timeout -s KILL 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/batch_n1.json 2>$O/batch_n1.err
$TR --master-port $((PORT++)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/batch_n$N.json 2>$O/batch_n$N.err
for bands in rows cols; do
  timeout -s KILL 120 python bench.py --workload roi-tiled-8k --bands $bands --steps 200 --warmup 20 > $O/tiled_${bands}_n1.json 2>/dev/null
  timeout -s KILL 120 python bench.py --workload roi-tiled-8k --bands $bands --graph --steps 200 --warmup 20 > $O/tiled_${bands}_graph_n1.json 2>$O/tiled_${bands}_graph_n1.err
  for red in nccl peers peers-one-shot nvls nvls-one-shot; do
    $TR --master-port $((PORT++)) bench.py --gpus $N --workload roi-tiled-8k --bands $bands --reduce $red --steps 200 --warmup 20 \
      > $O/tiled_${bands}_${red}_n$N.json 2>$O/tiled_${bands}_${red}_n$N.err
    # the same as one CUDA graph per frame: the device's number, without the Python harness's launch overhead
    [ $red != nccl ] && $TR --master-port $((PORT++)) bench.py --gpus $N --workload roi-tiled-8k --bands $bands --reduce $red --graph \
      --steps 200 --warmup 20 > $O/tiled_${bands}_${red}_graph_n$N.json 2>$O/tiled_${bands}_${red}_graph_n$N.err
  done
done
echo "== pytest"; tail -4 $O/pytest.full
for f in $O/*.json; do echo $f $(python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), d.get('ms_per_step'))" 2>&1 | tail -1); done
