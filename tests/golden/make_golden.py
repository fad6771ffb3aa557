#!/usr/bin/env python
"""Generate tests/golden/scope_golden.npz from the REFERENCE ITSELF.

Runs only in the build container (needs oracle/_ref/libref.so, i.e. the reference's own
src/{histogram,waveform,vectorscope}.c compiled from /root/reference by oracle/Makefile).
Inputs are small seeded frames; outputs are what the reference's loops produce for them.
The YUV planes fed to the reference are stored too (they are inputs of the integer path;
how they are derived from RGB is the separate, parity-unpinned transform).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import obs_color_monitor_b200 as pkg  # noqa: E402
from oracle.oracle import Oracle, Ref  # noqa: E402

fr = pkg.frames
CASES = {
    "ramp": fr.ramp(96, 64),
    "random": fr.random(67, 45, seed=1),
    "solid": fr.solid(40, 300, (12, 130, 250, 255)),   # H > 255: waveform saturation
    "alpha": fr.alpha_stripes(50, 37, seed=2),
    "natural": fr.natural(80, 60, seed=3),
    "pitched": fr.random(33, 20, seed=4),
}
COMPONENTS = [0x07, 0x20, 0x50, 0x70, 0x05, 0x42]


def main():
    ref, orc = Ref(), Oracle()
    out = {}
    for name, rgb in CASES.items():
        # an arbitrary-bytes "YUV" plane (surface mode does not care where it came from)
        yuv = orc.rgb_to_yuv(rgb, 2) if name != "alpha" else fr.alpha_stripes(50, 37, seed=9, period=4)
        width = rgb.shape[1]
        if name == "pitched":
            rgb_in, yuv_in = fr.with_pitch(rgb, width * 4 + 28), fr.with_pitch(yuv, width * 4 + 28)
        else:
            rgb_in, yuv_in = rgb, yuv
        out[f"{name}/rgb"] = rgb_in
        out[f"{name}/yuv"] = yuv_in
        out[f"{name}/width"] = np.int32(width)
        for comp in COMPONENTS:
            for log in (0, 1):
                flt, hi = ref.histogram(comp, rgb_in, yuv_in, width=width, logscale=bool(log))
                out[f"{name}/hist/{comp:02x}/log{log}"] = flt
                out[f"{name}/hist_max/{comp:02x}/log{log}"] = hi
            flt, hi = ref.histogram(comp, rgb_in, yuv_in, width=width, level_fixed=100)
            out[f"{name}/hist_max/{comp:02x}/fixed100"] = hi
            flt, hi = ref.histogram(comp, rgb_in, yuv_in, width=width, level_ratio=5)
            out[f"{name}/hist_max/{comp:02x}/ratio5"] = hi
            out[f"{name}/wave/{comp:02x}"] = ref.waveform(comp, rgb_in, yuv_in, width=width)
        out[f"{name}/vscope"] = ref.vectorscope(yuv_in, width=width)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scope_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
