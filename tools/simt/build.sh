# Build the SIMT emulator libraries (host only, no GPU, no nvcc): one per set of kernel build flags.
#   usage: bash tools/simt/build.sh [name "flags"]...      default: the set tests/test_kernel_emulation.py uses
cd "$(dirname "$0")/../.."
mkdir -p tools/simt/_build
build() {  # name, flags
  g++ -std=c++17 -O1 -g -fPIC -shared -DSCOPE_EMULATE $2 -include tools/simt/cuda_emul.h \
      -Iobs-color-monitor_b200/csrc tools/simt/emul_main.cpp -o tools/simt/_build/libscope_emul_$1.so
}
if [ $# -ge 2 ]; then
  while [ $# -ge 2 ]; do build "$1" "$2" || exit 1; shift 2; done
  exit 0
fi
pids=""
build default "" & pids="$pids $!"
build w8 "-DSCOPE_TMA_WARPS=8" & pids="$pids $!"
build w12n6 "-DSCOPE_EXPERIMENT -DSCOPE_TMA_WARPS=12 -DSCOPE_TILE_ROWS=72" & pids="$pids $!"
build wide "-DSCOPE_WIDE_FUSED=1 -DSCOPE_IMMCOEF=1" & pids="$pids $!"
build straight "-DSCOPE_STRAIGHT=1" & pids="$pids $!"
build w8_straight "-DSCOPE_STRAIGHT=1 -DSCOPE_TMA_WARPS=8" & pids="$pids $!"
build nopipe "-DSCOPE_PIPELINE=0" & pids="$pids $!"
build base "-DSCOPE_LDSM=0 -DSCOPE_XORSWZ=0 -DSCOPE_DEFER=0 -DSCOPE_FADDR=0 -DSCOPE_FAST_EMIT=0" & pids="$pids $!"
rc=0
for p in $pids; do wait $p || rc=1; done
exit $rc
