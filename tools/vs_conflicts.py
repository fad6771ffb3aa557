"""Replay the vectorscope's shared-memory access pattern on the CPU: for every warp-wide atomic
(32 horizontally adjacent pixels of one row) count the serialisation passes = the largest number
of lanes that fall into one bank (lanes on the very same word serialise as well), for the bank
swizzles the kernel has had.  Usage: python tools/vs_conflicts.py [width height]"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obs_color_monitor_b200 as pkg
from oracle.oracle import Oracle

def passes(word):
    bank = (word & 31).reshape(-1, 32)
    counts = np.zeros((bank.shape[0], 32), np.int32)
    rows = np.arange(bank.shape[0])[:, None]
    np.add.at(counts, (rows, bank), 1)
    return counts.max(axis=1).mean()

def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 270)
    orc = Oracle()
    frames = {"random": pkg.frames.random(w, h, 0), "ramp": pkg.frames.ramp(w, h), "natural": pkg.frames.natural(w, h, 3)}
    print(f"{'content':10s} {'plain':>7s} {'add(U+4V)':>10s} {'xor':>7s}")
    for name, f in frames.items():
        yuv = orc.rgb_to_yuv(f, 2)
        idx = yuv[..., 0].astype(np.uint32) | (yuv[..., 2].astype(np.uint32) << 8)
        plain = idx & 0x7FFF
        add = (idx & 0x7FE0) | (((idx >> 8) * 4 + idx) & 31)
        xor = (idx ^ ((idx >> 6) & 0x1C)) & 0x7FFF
        print(f"{name:10s} {passes(plain):7.2f} {passes(add):10.2f} {passes(xor):7.2f}")

if __name__ == "__main__":
    main()
