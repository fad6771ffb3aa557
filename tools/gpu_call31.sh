# wide end-of-strip write-out A/B + config 4 with F frames in flight (one GPU doing one rank's share of a 4-rank run)
cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-c31}; mkdir -p $O
timeout -s KILL 600 python -m pytest -x -q -m gpu tests/test_gpu_parity.py tests/test_gpu_headline_parity.py > $O/pytest.txt 2>&1; tail -2 $O/pytest.txt
B="timeout -s KILL 100 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-config4"
export SCOPE_BENCH_DIAGNOSTIC=1
for v in variants_tmp/*.so; do
  n=$(basename $v .so)
  SCOPE_LIB=$PWD/$v $B > $O/${n}_mixed.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content random > $O/${n}_random.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --content natural > $O/${n}_natural.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --width 1920 --height 1080 > $O/${n}_1080p.json 2>/dev/null
  SCOPE_LIB=$PWD/$v $B --width 7680 --height 4320 --frames-per-gpu 16 > $O/${n}_8k.json 2>/dev/null
done
for f in $O/*.json; do echo $(basename $f) $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['parity']['mismatches'])" 2>&1 | tail -1); done
C="timeout -s KILL 120 python bench.py --workload roi-tiled-8k --graph --reduce peers --steps 400 --warmup 20"
for F in 2 3 4; do $C --bands rows --in-flight $F > $O/cfg4_n1_rows_f$F.json 2>$O/cfg4_n1_rows_f$F.err; done
for F in 2 4 6 8; do $C --bands cols --in-flight $F --emulate-world 4 > $O/cfg4_e4_cols_f$F.json 2>$O/cfg4_e4_cols_f$F.err; done
for F in 2 4 6; do $C --bands rows --in-flight $F --emulate-world 4 > $O/cfg4_e4_rows_f$F.json 2>$O/cfg4_e4_rows_f$F.err; done
for F in 2 4; do $C --bands cols --in-flight $F --emulate-world 2 > $O/cfg4_e2_cols_f$F.json 2>$O/cfg4_e2_cols_f$F.err; done
for f in $O/cfg4_*.json; do echo $(basename $f) $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_frame']*1e3,1), 'us/frame', d['frames_in_flight'], d['band_kernel_us'], d['graph'], d.get('graph_error'))" 2>&1 | tail -1); done
