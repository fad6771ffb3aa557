import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import obs_color_monitor_b200 as pkg
from obs_color_monitor_b200 import frames_torch
from oracle.oracle import Oracle
orc = Oracle()
eng = pkg.ScopeEngine(0)
dev = torch.device("cuda", 0)
def check(w, h, n, content, idx0=0):
    batch = frames_torch.mixed_batch(n, w, h, dev, first_index=idx0, content=content)
    out = eng.accumulate_device(batch)
    torch.cuda.synchronize()
    res = []
    for i in sorted(set([0, n - 1])):
        f = np.ascontiguousarray(batch[i].cpu().numpy())
        yuv = orc.rgb_to_yuv(f, 2)
        eh = orc.histogram_counts(7, f, yuv).ravel()
        gh = out["hist"][i].cpu().numpy().view(np.uint32).ravel()
        ew = orc.waveform(7, f, yuv); gw = out["wave"][i].cpu().numpy()
        ev = orc.vectorscope(yuv); gv = out["vscope"][i].cpu().numpy()
        res.append((i, int((eh != gh).sum()), int(gh.sum()) - int(eh.sum()), int((ew != gw).sum()), int((ev != gv).sum())))
    print(f"{w}x{h} n={n} {content}: (frame, hist bins differ, hist total diff, wave bytes differ, vscope bins differ) {res}", flush=True)
for (w, h, n) in ((7680, 4320, 1), (7680, 4320, 2), (7680, 4320, 16), (3840, 4320, 4), (7680, 2160, 4), (3840, 2160, 8)):
    for content in ("random", "natural"):
        check(w, h, n, content)
