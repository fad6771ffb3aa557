import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import obs_color_monitor_b200 as pkg
from obs_color_monitor_b200 import frames_torch
from oracle.oracle import Oracle
orc = Oracle()
eng = pkg.ScopeEngine(0)
dev = torch.device("cuda", 0)
def check(w, h, n, content, reps=1):
    batch = frames_torch.mixed_batch(n, w, h, dev, content=content)
    for rep in range(reps):
        out = eng.accumulate_device(batch)
    torch.cuda.synchronize()
    for i in range(n):
        f = np.ascontiguousarray(batch[i].cpu().numpy())
        yuv = orc.rgb_to_yuv(f, 2)
        eh = orc.histogram_counts(7, f, yuv).ravel().astype(np.int64)
        gh = out["hist"][i].cpu().numpy().view(np.uint32).ravel().astype(np.int64)
        ew = orc.waveform(7, f, yuv); gw = out["wave"][i].cpu().numpy()
        ev = orc.vectorscope(yuv); gv = out["vscope"][i].cpu().numpy()
        d = np.nonzero(eh != gh)[0]
        print(f"{w}x{h} n={n} {content} frame {i}: hist differ {d.size} total diff {gh.sum()-eh.sum()} wave differ {(ew!=gw).sum()} vscope differ {(ev!=gv).sum()}", flush=True)
        if d.size:
            print("   hist idx", d[:12], "got-exp", (gh - eh)[d[:12]])
            wd = np.argwhere(ew != gw)
            print("   wave idx (level,x,ch)", wd[:8].tolist(), "got", [int(gw[tuple(t)]) for t in wd[:8]], "exp", [int(ew[tuple(t)]) for t in wd[:8]])
            cols = np.unique(wd[:, 1]); print("   wave columns affected:", cols.size, cols[:40])
            vd = np.argwhere(ev != gv); print("   vs idx", vd[:8].tolist(), [int(gv[tuple(t)]) for t in vd[:8]], [int(ev[tuple(t)]) for t in vd[:8]])
            break
for (w, h, n) in ((1920, 1080, 1), (3840, 2160, 1), (3840, 2160, 4), (3840, 2160, 16)):
    check(w, h, n, "ui")
check(3840, 2160, 64, "ui", reps=3)
check(7680, 4320, 16, "mixed", reps=5)
