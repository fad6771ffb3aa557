"""probe: python tools/gpu_probe2.py W H SCOPES [intensity]  (debug aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import obs_color_monitor_b200 as pkg
from oracle.oracle import Oracle
w, h, scopes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
inten = int(sys.argv[4]) if len(sys.argv) > 4 else 0
kind = sys.argv[5] if len(sys.argv) > 5 else "random"
o = Oracle(); eng = pkg.ScopeEngine(0)
f = pkg.frames.random(w, h, seed=1) if kind == "random" else pkg.frames.ramp(w, h)
yuv = o.rgb_to_yuv(f, 2)
print("start", w, h, scopes, inten, kind, "tma_off=", os.environ.get("SCOPE_DISABLE_TMA"), flush=True)
res = eng.accumulate_host(f, settings=pkg.ScopeSettings(scopes=scopes, wave_intensity=inten, vscope_intensity=inten))
ok = []
if "hist" in res: ok.append(("hist", np.array_equal(res["hist"], o.histogram_counts(7, f, yuv))))
if "wave" in res: ok.append(("wave", np.array_equal(res["wave"], o.waveform(7, f, yuv))))
if "vscope" in res: ok.append(("vscope", np.array_equal(res["vscope"], o.vectorscope(yuv))))
print("  result", ok, flush=True)
