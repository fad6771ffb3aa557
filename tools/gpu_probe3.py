"""probe: the exact frame list of tests/test_gpu_parity.py::test_fused_all_scopes_host, one at a time"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import obs_color_monitor_b200 as pkg
from oracle.oracle import Oracle
from test_gpu_parity import _frames, _check
name = sys.argv[1]
o = Oracle(); eng = pkg.ScopeEngine(0)
f = _frames(pkg)[name]
print("start", name, f.shape, flush=True)
st = pkg.ScopeSettings(colorspace=1, wave_intensity=51, vscope_intensity=25)
res = eng.accumulate_host(f, settings=st)
print("  returned", flush=True)
_check(res, o, f, o.rgb_to_yuv(f, 1), st, name)
print("  ok", flush=True)
