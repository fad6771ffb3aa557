"""Print one visit of scope_fused_kernel_v3's lean loop from the built library's SASS (no GPU): from the tile wait in
front of the n-th ldmatrix to the branch behind the sixteen shared-memory atomics of the ordinary block.
   usage: python tools/sass_visit.py [LIB] [N]     (N = 1, 2, 3: the loop's three unrolled visits; default 1)
tests/test_sass_budget.py::test_headline_kernel_lean_visit_budget guards the same window."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "obs-color-monitor_b200", "lib", "libscope_b200.so")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
ins, on = [], False
for line in out.splitlines():
    if "Function :" in line:
        on = "scope_fused_kernel_v3ILi2E" in line
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
    if on and m:
        ins.append((m.group(1), m.group(2)))
ldsm = [i for i, (_, t) in enumerate(ins) if "LDSM" in t]
i = ldsm[n]
j = i
while "SYNCS.PHASECHK" not in ins[j][1]:
    j -= 1
k, atoms = i + 1, 0
while atoms < 16:
    atoms += "ATOMS" in ins[k][1]
    k += 1
while "BRA" not in ins[k][1]:
    k += 1
seg = ins[j - 1:k + 1]
ops = {}
for _, t in seg:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    op = op.split(".")[0] if not op.startswith(("ATOMS", "SYNCS")) else ".".join(op.split(".")[:2])
    ops[op] = ops.get(op, 0) + 1
print(f"# visit {n} of the lean loop: {len(seg)} instructions for 4 x 32 pixels = {len(seg) / 4:.1f} per 32 pixels "
      f"(kernel: {len(ins)} instructions)")
print("# " + ", ".join(f"{v} {k_}" for k_, v in sorted(ops.items(), key=lambda kv: -kv[1])))
for a, t in seg:
    print(a, t)
