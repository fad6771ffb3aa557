"""Small workload for compute-sanitizer (tools/run_sanitizer.sh): every scope combination of the shipped
kernels on small frames through the C-ABI, checked against the oracle.  numpy + ctypes only (no torch: the
sanitizer instruments every kernel of the process, and torch would add minutes of start-up)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import obs_color_monitor_b200 as pkg  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    eng, orc, fr = pkg.ScopeEngine(0), Oracle(), pkg.frames
    n = 0
    frames = [fr.random(200, 150, seed=1), fr.natural(161, 140, seed=2), fr.solid(96, 131, (9, 99, 199, 255)),
              fr.alpha_stripes(130, 70, seed=3), fr.ramp(64, 260)]
    for f in frames:
        yuv = orc.rgb_to_yuv(f, 2)
        for scopes in (1, 2, 4, 3, 5, 6, 7):
            for mode, comps in ((pkg.MODE_FUSED, (0x07, 0x07)), (pkg.MODE_FUSED, (0x70, 0x20)),
                                (pkg.MODE_SURFACE, (0x07, 0x70))):
                st = pkg.ScopeSettings(scopes=scopes, mode=mode, hist_components=comps[0], wave_components=comps[1],
                                       vscope_intensity=25 if scopes & 4 else 0)
                res = eng.accumulate_host(f, yuv if mode == pkg.MODE_SURFACE else None, settings=st)
                if "hist" in res:
                    assert np.array_equal(res["hist"], orc.histogram_counts(comps[0], f, yuv))
                if "wave" in res:
                    assert np.array_equal(res["wave"], orc.waveform(comps[1], f, yuv))
                if "vscope" in res:
                    assert np.array_equal(res["vscope"], orc.vectorscope(yuv))
                n += 1
    # tall frames: seven tiles per strip, so that scope_fused_kernel_v3 runs its lean visits (blocks inside the frame
    # with a successor), the 16-byte write-out and, on the solid frame, the flat-block and take-back paths inside them
    patched = fr.random(64, 700, seed=8)          # flat patches: lanes with four equal pixels (background-lane rule)
    patched[..., 3] = 255
    patched[8:12, 10:22] = (200, 10, 60, 255)
    patched[332:336, 0:31] = (5, 130, 250, 255)
    tall = [fr.random(96, 700, seed=6), fr.solid(64, 700, (200, 17, 90, 255)), fr.alpha_stripes(72, 650, seed=7),
            fr.ui(96, 700, 1), patched]
    for f in tall:
        yuv = orc.rgb_to_yuv(f, 2)
        res = eng.accumulate_host(f, settings=pkg.ScopeSettings(scopes=7, mode=pkg.MODE_FUSED))
        assert np.array_equal(res["hist"], orc.histogram_counts(0x07, f, yuv))
        assert np.array_equal(res["wave"], orc.waveform(0x07, f, yuv))
        assert np.array_equal(res["vscope"], orc.vectorscope(yuv))
        n += 1
    # pitched plane whose rows are not 16-byte multiples (plain-load kernel), ring slots
    odd = fr.random(75, 40, seed=5)
    res = eng.accumulate_host(odd)
    assert np.array_equal(res["vscope"], orc.vectorscope(orc.rgb_to_yuv(odd, 2)))
    for i in range(6):
        assert eng.submit_host(i % 3, frames[i % len(frames)]) is True
        got = eng.wait_host(i % 3)
        f = frames[i % len(frames)]
        assert np.array_equal(got["wave"], orc.waveform(7, f, orc.rgb_to_yuv(f, 2)))
    print(f"sanitize_driver ok: {n} scope combinations, {eng.launch_count} launches")
    eng.close()


if __name__ == "__main__":
    main()
