# BASELINE config 4 with F frames in flight (bench.py --in-flight): N ranks, column bands (the strip kernel's own peer stores)
# and row bands (fused peer reduce).   usage (under gpurun --gpus N): bash tools/run_config4_inflight.sh N [tag]
cd $GRAFT_REPO_ROOT
N=${1:-4}; O=gpurun_out/${2:-cfg4f_n$N}; mkdir -p $O
PORT=29811
run() { # ranks bands reduce in_flight
  timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((PORT++)) \
    bench.py --gpus $1 --workload roi-tiled-8k --bands $2 --reduce $3 --graph --steps 800 --warmup 24 --in-flight $4 \
    > $O/tiled_$2_$3_f$4_n$1.json 2>$O/tiled_$2_$3_f$4_n$1.err
}
timeout -s KILL 120 python bench.py --workload roi-tiled-8k --bands rows --reduce peers --graph --steps 400 --warmup 20 --in-flight 2 > $O/tiled_rows_peers_f2_n1.json 2>$O/tiled_rows_peers_f2_n1.err
for F in 2 4 8; do run $N cols peers $F; done
run $N rows peers 4
if [ $N -gt 2 ]; then for F in 4 8; do run 2 cols peers $F; done; fi
for f in $O/*.json; do echo $(basename $f) $(python -c "import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(1e3*d['ms_per_frame'],1), 'us/frame; band kernel', d.get('band_kernel_us'), 'us; graph', d.get('graph'), d.get('graph_error'), 'parity', (d.get('parity') or {}), d['clocks']['sm_mhz'])" 2>&1 | tail -1); done
