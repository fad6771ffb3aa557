# Static budget of the fused kernel's steady-state loop for the in-tree library and every variants_tmp/*.so
# (no GPU): instructions per 32 pixels, by pipe, LSU instructions, ptxas stall sum, registers / stack.
#   usage: bash tools/variant_table.sh
cd "$(dirname "$0")/.."
K=scope_strip_kernel_tmaILi1ELb1ELb0
printf "%-20s %7s %6s %6s %6s %6s %9s %5s %6s\n" build instr alu imad ffma lsu stall/px regs stack
for lib in obs-color-monitor_b200/lib/libscope_b200.so variants_tmp/*.so; do
  n=$(basename $lib .so); [ "$n" = libscope_b200 ] && n="(in-tree)"
  KK=$K; case $n in *immcoef*|wide*) KK=${K}ELi2;; esac   # SCOPE_IMMCOEF builds: the BT.709 instance
  out=$(python tools/sass_budget.py $lib --kernel $KK 2>&1)
  tot=$(echo "$out" | sed -n 2p | sed -E 's/.*= ([0-9.]+) per 32 pixels/\1/')
  pw=$(echo "$out" | sed -n 2p | sed -E 's/.*per ([0-9]+) pixel-warps.*/\1/')
  alu=$(echo "$out" | grep " alu " | awk '{print $1}')
  imad=$(echo "$out" | grep "fma-heavy" | awk '{print $1}')
  ffma=$(echo "$out" | grep " fma (FFMA" | awk '{print $1}')
  lsu=$(echo "$out" | grep "^LSU" | awk '{print $3}')
  st=$(echo "$out" | grep "stall counts" | sed -E 's/.*fast path: ([0-9]+) cycles.*/\1/')
  ru=$(cuobjdump --dump-resource-usage $lib 2>/dev/null | grep -A1 $KK | grep -o "REG:[0-9]*\|STACK:[0-9]*" | cut -d: -f2 | paste - -)
  printf "%-20s %7s %6s %6s %6s %6s %9s %5s %6s\n" "$n" "$tot" "$alu" "$imad" "$ffma" "$lsu" "$(python -c "print(round($st/$pw,1))")" $ru
done
