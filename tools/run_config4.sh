# BASELINE config 4 only (one 8K frame, luma waveform, bands over N GPUs): N = 1 and N = $1, best forms.
#   usage (under gpurun --gpus N): bash tools/run_config4.sh N [tag]
cd $GRAFT_REPO_ROOT
N=${1:-2}; O=gpurun_out/${2:-cfg4_n$N}; mkdir -p $O
TR="timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29711
for bands in rows cols; do
  timeout -s KILL 120 python bench.py --workload roi-tiled-8k --bands $bands --reduce peers --graph --steps 400 --warmup 20 > $O/tiled_${bands}_n1.json 2>$O/tiled_${bands}_n1.err
done
if [ $N -gt 1 ]; then
  for cfg in "rows peers" "rows peers-one-shot" "rows nvls" "cols peers" "cols peers-one-shot"; do
    set -- $cfg
    $TR --master-port $((PORT++)) bench.py --gpus $N --workload roi-tiled-8k --bands $1 --reduce $2 --graph --steps 400 --warmup 20 \
      > $O/tiled_$1_$2_n$N.json 2>$O/tiled_$1_$2_n$N.err
  done
fi
for f in $O/*.json; do echo $f $(python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(1e3*d['ms_per_frame'],1), 'us/frame; band kernel', d.get('band_kernel_us'), 'us; graph', d.get('graph'), d.get('graph_error'), 'parity', (d.get('parity') or {}).get('mismatches'), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
