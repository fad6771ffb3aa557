"""Step-by-step GPU probe with flushed progress lines (debug aid; writes gpurun_out/probe.log)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "a")


def say(*a):
    msg = " ".join(str(x) for x in a)
    print(msg, flush=True)
    LOG.write(msg + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())


step = sys.argv[1] if len(sys.argv) > 1 else "all"
say("== probe", step, "tma_disabled=", os.environ.get("SCOPE_DISABLE_TMA"))
t0 = time.time()
import numpy as np
import torch
say("torch imported", round(time.time() - t0, 1), "cuda", torch.cuda.is_available())
import obs_color_monitor_b200 as pkg
from oracle.oracle import Oracle
o = Oracle()
eng = pkg.ScopeEngine(0)
say("engine created, sms", eng.ctx.sm_count)

if step in ("table", "all"):
    got = eng.debug_yuv_table(2).cpu().numpy().view(np.uint32)
    exp, _ = o.rgb_to_yuv_table(2)
    say("yuv table mismatches:", int((got != exp).sum()))

f = pkg.frames.random(200, 150, seed=1)
yuv = o.rgb_to_yuv(f, 2)
for name, scopes in (("hist", 1), ("wave", 2), ("vscope", 4), ("all", 7)):
    if step not in (name, "all", "scopes"):
        continue
    say("running", name)
    res = eng.accumulate_host(f, settings=pkg.ScopeSettings(scopes=scopes))
    say("  returned", list(res))
    if "hist" in res:
        say("  hist ok:", np.array_equal(res["hist"], o.histogram_counts(7, f, yuv)))
    if "wave" in res:
        say("  wave ok:", np.array_equal(res["wave"], o.waveform(7, f, yuv)))
    if "vscope" in res:
        say("  vscope ok:", np.array_equal(res["vscope"], o.vectorscope(yuv)))
say("done")
