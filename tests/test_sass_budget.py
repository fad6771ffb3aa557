"""Static guard on the fused kernel's steady-state loop (no GPU): the pass is bound by instruction issue and
by the LSU (DESIGN.md sections 5 and 8.1), so its instruction budget is a property worth pinning.  Read from
the built library with tools/sass_budget.py (nvdisasm); skipped where the CUDA binary tools are absent."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "obs-color-monitor_b200", "lib", "libscope_b200.so")
FUSED = "scope_strip_kernel_tmaILi1ELb1ELb0"     # <SRC_RGB, VSCOPE, fused>: the headline kernel


@pytest.mark.skipif(not (shutil.which("nvdisasm") and shutil.which("cuobjdump")), reason="CUDA binary tools not installed")
def test_fused_loop_budget():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_budget.py"), LIB, "--kernel", FUSED],
                         capture_output=True, text=True, check=True).stdout
    total = float(re.search(r"= ([0-9.]+) per 32 pixels", out).group(1))
    atoms = float(re.search(r"([0-9.]+)\s+lsu: shared atomics", out).group(1))
    lsu = float(re.search(r"LSU instructions: ([0-9.]+) per 32 pixels", out).group(1))
    # 4 scatter updates per pixel is what the formulation needs - not one more; one LDSM per 4 rows
    assert atoms == 4.0
    assert lsu <= 5.25
    # shipped: 39.8 instructions per 32 pixels on the fast path (27 per-pixel core + per-visit overhead / 4)
    assert total <= 40.5, out
    usage = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    m = re.search(FUSED + r".*?\n.*?REG:(\d+) STACK:(\d+)", usage)
    assert m, "fused kernel not found in the library"
    regs, stack = int(m.group(1)), int(m.group(2))
    assert stack == 0, "the fused kernel spills"
    assert regs * 544 <= 65536, "17 warps of the fused kernel no longer fit the register file"
