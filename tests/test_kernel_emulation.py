"""The strip kernels' LOGIC without a GPU: csrc/scope_kernels.cuh compiled for the host on a small SIMT emulator
(tools/simt: every CUDA thread a coroutine under a seeded random scheduler; warp collectives, named barriers,
mbarriers with transaction counts, TMA tile loads that land late and out of order, ldmatrix, the PTX arithmetic
helpers) and compared bit for bit with the CPU oracle.  It covers what the GPU tests cover - but also the
build-flag variants prepared for the next round, which have not run on a GPU yet, and
it explores far more interleavings than the hardware does.  Timing and bank conflicts are not modelled."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tools", "simt", "_build")
SOURCES = [os.path.join(ROOT, "tools", "simt", f) for f in ("cuda_emul.h", "emul_main.cpp", "build.sh")] + \
          [os.path.join(ROOT, "obs-color-monitor_b200", "csrc", f)
           for f in ("scope_kernels.cuh", "scope_kernels_experiments.cuh", "scope_fused_v3.cuh")]
# build-flag sets of the general kernel that are still exercised (round 2 measured the others and dropped them:
# profiles/r02/ab_round2.md); "default" also carries scope_fused_kernel_v3, the headline combination's own kernel
VARIANTS = ["default", "w8", "w12n6", "straight", "w8_straight", "wide", "nopipe", "base"]
SRC_NONE, SRC_RGB, SRC_YUV = 0, 1, 2
K_TMA, K_LDG, K_GROUP, K_V3 = 0, 1, 2, 3


class Request(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("yuv", C.c_void_p), ("linesize", C.c_uint32), ("width", C.c_uint32),
                ("height", C.c_uint32), ("n_frames", C.c_uint32), ("frame_stride", C.c_uint64),
                ("colorspace", C.c_int32), ("surface", C.c_int32), ("src", C.c_int32), ("vscope", C.c_int32),
                ("bins_mask", C.c_uint32), ("hist_mask", C.c_uint32), ("wave_mask", C.c_uint32),
                ("hist", C.c_void_p), ("wave", C.c_void_p), ("vs_acc", C.c_void_p), ("wave_pairs", C.c_void_p),
                ("out_width", C.c_uint32), ("x_offset", C.c_uint32), ("partial", C.c_uint32),
                ("kernel", C.c_int32), ("ctas", C.c_int32), ("seed", C.c_uint32), ("tma_land_percent", C.c_int32),
                ("steps", C.c_int64), ("error", C.c_char * 256)]


@pytest.fixture(scope="session")
def emul_libs():
    newest = max(os.path.getmtime(f) for f in SOURCES)
    paths = {v: os.path.join(BUILD, f"libscope_emul_{v}.so") for v in VARIANTS}
    if not all(os.path.exists(p) and os.path.getmtime(p) >= newest for p in paths.values()):
        subprocess.run(["bash", os.path.join(ROOT, "tools", "simt", "build.sh")], check=True)
    libs = {}
    for v, p in paths.items():
        lib = C.CDLL(p)
        lib.emul_run.argtypes = [C.POINTER(Request)]
        lib.emul_build_flags.restype = C.c_char_p
        libs[v] = lib
    return libs


def masks(comp):
    src = SRC_RGB if comp & 0x07 else (SRC_YUV if comp & 0x70 else SRC_NONE)
    m = (1 if comp & 0x11 else 0) | (2 if comp & 0x22 else 0) | (4 if comp & 0x44 else 0)
    return src, (m if src != SRC_NONE else 0)


def run(lib, frames, yuv=None, surface=False, hist_comp=0x07, wave_comp=0x07, vscope=True, kernel=K_TMA, ctas=2,
        seed=1, colorspace=2, land=30, width=None, partial_into=None, x_offset=0, out_width=None, rows=None):
    """frames: (n, H, W, 4) u8.  hist_comp and wave_comp must select the same plane (one launch).
    width < W reads only the first `width` pixels of every (then pitched) row; rows = (y0, y1) reads a row band;
    partial_into = (hist, wave_pairs, acc) accumulates a tile of a larger frame into shared partial results."""
    assert frames.flags["C_CONTIGUOUS"]
    n, h, w, _ = frames.shape
    pitch_w = w
    if width is not None:
        w = width
    base = frames.ctypes.data
    if rows is not None:
        base += rows[0] * pitch_w * 4
        h_band = rows[1] - rows[0]
    else:
        h_band = h
    hsrc, hmask = masks(hist_comp)
    wsrc, wmask = masks(wave_comp)
    src = hsrc if hsrc != SRC_NONE else wsrc
    assert wsrc in (SRC_NONE, src)
    ow = out_width or w
    wave = np.full((n, 256, ow, 4), 0xEE, np.uint8)     # the kernel must write every row it owns
    if partial_into is None:
        hist = np.zeros((n, 1024), np.uint32)
        acc = np.zeros((n, 65536), np.uint32)
        pairs = None
    else:
        hist, pairs, acc = partial_into
    rq = Request()
    rq.rgb = base
    rq.yuv = yuv.ctypes.data + (base - frames.ctypes.data) if yuv is not None else None
    rq.linesize, rq.width, rq.height, rq.n_frames, rq.frame_stride = pitch_w * 4, w, h_band, n, pitch_w * h * 4
    rq.colorspace, rq.surface, rq.src, rq.vscope = colorspace, int(surface), src, int(vscope)
    rq.bins_mask, rq.hist_mask, rq.wave_mask = hmask | wmask, hmask, wmask
    rq.hist, rq.wave, rq.vs_acc = hist.ctypes.data, wave.ctypes.data, acc.ctypes.data
    rq.wave_pairs = pairs.ctypes.data if pairs is not None else None
    rq.out_width, rq.x_offset, rq.partial = ow, x_offset, int(partial_into is not None)
    rq.kernel, rq.ctas, rq.seed, rq.tma_land_percent = kernel, ctas, seed, land
    rc = lib.emul_run(C.byref(rq))
    assert rc == 0, rq.error.decode()
    if src == SRC_NONE or not wmask:
        wave = None
    return hist, wave, np.minimum(acc, 255).astype(np.uint8).reshape(n, 256, 256), rq.steps


def check(oracle, frames, out, yuv_planes, hist_comp, wave_comp, vscope, what):
    hist, wave, vs, _ = out
    for i in range(frames.shape[0]):
        f, y = frames[i], yuv_planes[i]
        if masks(hist_comp)[1]:
            assert np.array_equal(hist[i], oracle.histogram_counts(hist_comp, f, y)), f"{what}: histogram, frame {i}"
        if wave is not None:
            exp = oracle.waveform(wave_comp, f, y)
            assert np.array_equal(wave[i][..., :3], exp[..., :3]) and not wave[i][..., 3].any(), f"{what}: waveform, frame {i}"
        if vscope:
            assert np.array_equal(vs[i], oracle.vectorscope(y)), f"{what}: vectorscope, frame {i}"


def small_batch(pkg):
    fr = pkg.frames
    w, h = 70, 150                      # 3 strips (the last one 6 pixels wide), 2 full tiles + a partial one
    return np.stack([fr.random(w, h, 3), fr.natural(w, h, 4), fr.alpha_stripes(w, h, 5), fr.ramp(w, h)])


@pytest.mark.parametrize("variant", VARIANTS)
def test_fused_all_scopes_every_variant(emul_libs, oracle, pkg, variant):
    """the headline combination (RGB column bins + vectorscope, transform in registers) on 4 frames x 3 strips,
    two CTAs sharing the work through the chunk counter, for every build variant"""
    lib = emul_libs[variant]
    frames = small_batch(pkg)
    yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
    for seed, land in ((1, 30), (2, 3)):
        out = run(lib, frames, seed=seed, land=land)
        check(oracle, frames, out, yuv, 0x07, 0x07, True, f"{variant} ({lib.emul_build_flags().decode()}) seed {seed}")


@pytest.mark.parametrize("variant", ["default", "w8_straight", "w12n6", "wide"])
def test_other_kernels_and_modes(emul_libs, oracle, pkg, variant):
    lib = emul_libs[variant]
    fr = pkg.frames
    frames = np.stack([fr.alpha_stripes(45, 131, 7), fr.natural(45, 131, 8)])
    yuv = [oracle.rgb_to_yuv(f, 1) for f in frames]
    yuv_arr = np.ascontiguousarray(np.stack(yuv))
    surface_ok = variant in ("default", "wide")   # the _x builds' two-plane ring does not fit (SCOPE_EXPERIMENT)
    cases = [dict(hist_comp=0x70, wave_comp=0x20, vscope=True),                    # fused, YUV bins + vectorscope
             dict(hist_comp=0x07, wave_comp=0x05, vscope=False),                   # no vectorscope: 2 CTAs per SM kernel
             dict(hist_comp=0x00, wave_comp=0x00, vscope=True),                    # vectorscope only
             dict(hist_comp=0x50, wave_comp=0x00, vscope=False)]                   # histogram only, two channels
    for kw in cases:
        for kernel in (K_TMA, K_LDG) + ((K_GROUP,) if variant == "default" else ()):
            out = run(lib, frames, colorspace=1, kernel=kernel, ctas=3, seed=5, **kw)
            check(oracle, frames, out, yuv, kw["hist_comp"], kw["wave_comp"], kw["vscope"], f"{variant} fused {kw} kernel {kernel}")
    if surface_ok:
        rng = np.random.default_rng(3)
        any_yuv = rng.integers(0, 256, yuv_arr.shape, dtype=np.uint8)      # surface mode takes the plane as it is
        for kw in (dict(hist_comp=0x07, wave_comp=0x07, vscope=True), dict(hist_comp=0x70, wave_comp=0x70, vscope=True),
                   dict(hist_comp=0x20, wave_comp=0x50, vscope=False)):
            for kernel in (K_TMA, K_LDG):
                out = run(lib, frames, yuv=any_yuv, surface=True, kernel=kernel, ctas=2, seed=9, **kw)
                check(oracle, frames, out, list(any_yuv), kw["hist_comp"], kw["wave_comp"], kw["vscope"],
                      f"{variant} surface {kw} kernel {kernel}")


@pytest.mark.parametrize("variant", ["default", "w8", "straight"])
def test_saturation_and_flat_blocks(emul_libs, oracle, pkg, variant):
    """a solid frame: one vectorscope bin takes every pixel (more than the 0x8000 a half-word bin may hold before
    adds are taken back), waveform bins saturate at 255, the flat-block path is the one that runs; plus a frame
    that is solid except for a few pixels, which defeats the flat path on some blocks"""
    lib = emul_libs[variant]
    w, h = 96, 400                       # 38 400 pixels > 32 768
    solid = pkg.frames.solid(w, h, (200, 17, 90, 255))
    speck = solid.copy()
    speck[::37, ::11] = (3, 250, 128, 255)
    speck[5::53, 7::13, 3] = 0           # and some transparent pixels
    frames = np.stack([solid, speck])
    yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
    out = run(lib, frames, ctas=1, seed=11)
    check(oracle, frames, out, yuv, 0x07, 0x07, True, variant)
    assert out[2][0].max() == 255 and out[1][0].max() == 255


@pytest.mark.parametrize("variant", ["default"])
def test_almost_flat_blocks(emul_libs, oracle, pkg, variant):
    """screen-like content: a flat background with text in it.  Most blocks are not flat but most lanes sit on
    the background's bin - the case SCOPE_BALLOT aggregates (and the shipped kernel must get right the slow way);
    96 x 400 keeps the background bin above 0x8000 so that the take-back of aggregated adds is exercised too"""
    lib = emul_libs[variant]
    frames = np.stack([pkg.frames.ui(96, 400, 1), pkg.frames.ui(96, 400, 2)])
    frames[1, 100:140, 10:50] = pkg.frames.random(40, 40, 3)        # a picture inside the window
    frames[1, ::29, ::7, 3] = 0                                      # and some transparent pixels
    yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
    out = run(lib, frames, ctas=2, seed=21)
    check(oracle, frames, out, yuv, 0x07, 0x07, True, variant)
    out = run(lib, frames, hist_comp=0, wave_comp=0, ctas=1, seed=22)      # vectorscope only
    check(oracle, frames, out, yuv, 0, 0, True, variant + " vectorscope only")


def test_scheduler_seeds_and_late_tma(emul_libs, oracle, pkg):
    """the shipped kernel under many interleavings, including TMA loads that almost never land promptly"""
    lib = emul_libs["default"]
    frames = small_batch(pkg)[:2]
    yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
    for seed in range(6):
        out = run(lib, frames, ctas=1 + seed % 3, seed=100 + seed, land=(1, 10, 60)[seed % 3])
        check(oracle, frames, out, yuv, 0x07, 0x07, True, f"seed {seed}")


@pytest.mark.parametrize("variant", ["default", "w8_straight"])
def test_pitched_rows_and_tile_sharded_frames(emul_libs, oracle, pkg, variant):
    """rows with padding behind them, the plain-load kernel on a plane that is only pixel-aligned, and one frame
    accumulated as two row bands + two column bands into shared partial results (the multi-GPU tile sharding:
    u32 histogram, u16-pair waveform planes, u32 vectorscope; saturation after the sum)"""
    lib = emul_libs[variant]
    rng = np.random.default_rng(8)
    wide = rng.integers(0, 256, (1, 140, 104, 4), dtype=np.uint8)       # 104-pixel pitch, garbage in the padding
    wide[..., 3] = np.where(rng.random((1, 140, 104)) < 0.1, 0, 255)
    w = 77
    f = np.ascontiguousarray(wide[:, :, :w])
    yuv = [oracle.rgb_to_yuv(f[0], 2)]
    out = run(lib, wide, width=w, ctas=2, seed=3)                        # pitch 416 bytes = 26 x 16: TMA
    check(oracle, f, out, yuv, 0x07, 0x07, True, f"{variant} pitched")
    # a crop that starts at an odd column is only 4-byte aligned: the plain-load kernel's job
    crop = wide.reshape(-1)[3 * 4:][: 139 * 104 * 4].reshape(1, 139, 104, 4)
    fc = np.ascontiguousarray(crop[:, :, :w])
    out = run(lib, crop, width=w, kernel=K_LDG, ctas=2, seed=4)
    check(oracle, fc, out, [oracle.rgb_to_yuv(fc[0], 2)], 0x07, 0x07, True, f"{variant} unaligned crop")
    # tile sharding: rows [0, 64) and [64, 140) of the full width, accumulated into the same partial buffers
    full = np.ascontiguousarray(wide[:, :, :96])
    hist = np.zeros((1, 1024), np.uint32)
    pairs = np.zeros((2, 256, 96), np.uint32)
    acc = np.zeros((1, 65536), np.uint32)
    for band in ((0, 64), (64, 140)):
        run(lib, wide, width=96, rows=band, partial_into=(hist, pairs, acc), ctas=2, seed=band[0] + 5)
    y = oracle.rgb_to_yuv(full[0], 2)
    assert np.array_equal(hist[0], oracle.histogram_counts(0x07, full[0], y))
    wave = np.stack([np.minimum(pairs[0] & 0xFFFF, 255), np.minimum(pairs[0] >> 16, 255), np.minimum(pairs[1] & 0xFFFF, 255)],
                    axis=-1).astype(np.uint8)
    assert np.array_equal(wave, oracle.waveform(0x07, full[0], y)[..., :3])
    assert np.array_equal(np.minimum(acc[0], 255).astype(np.uint8).reshape(256, 256), oracle.vectorscope(y))


@pytest.mark.parametrize("variant", ["default", "w12n6", "w8_straight", "wide"])
def test_extreme_geometries(emul_libs, oracle, pkg, variant):
    """one pixel, one row, one column, a narrow tall strip, exactly one tile, one row more than a tile"""
    lib = emul_libs[variant]
    for w, h in ((1, 1), (33, 1), (1, 97), (5, 300), (32, 64), (64, 65), (31, 63)):
        f = pkg.frames.random(w, h, seed=w * 1000 + h)[None]
        f = np.ascontiguousarray(f)
        yuv = [oracle.rgb_to_yuv(f[0], 2)]
        for kernel in (K_TMA, K_LDG):
            if kernel == K_TMA and (w * 4) % 16:
                continue                 # the host sends pitches that are not a multiple of 16 to the plain-load kernel
            out = run(lib, f, kernel=kernel, ctas=2, seed=w + h)
            check(oracle, f, out, yuv, 0x07, 0x07, True, f"{variant} {w}x{h} kernel {kernel}")


@pytest.mark.parametrize("case", ["ramp", "random", "solid", "alpha", "natural", "pitched"])
def test_surface_mode_matches_reference_golden_emulated(emul_libs, case):
    """The kernels (emulated) against the outputs of the reference's OWN loops - tests/golden/scope_golden.npz, made
    from the unmodified src/{histogram,waveform,vectorscope}.c - in strict drop-in mode: same planes in, every
    count and byte equal.  No oracle in between (the GPU twin of this test is in test_gpu_parity.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "scope_golden.npz"))
    rgb, yuv, width = np.ascontiguousarray(g[f"{case}/rgb"]), np.ascontiguousarray(g[f"{case}/yuv"]), int(g[f"{case}/width"])
    if rgb.ndim == 2:                                   # pitched: (H, linesize) bytes
        rgb, yuv = rgb.reshape(rgb.shape[0], -1, 4), yuv.reshape(yuv.shape[0], -1, 4)
    rgb, yuv = rgb[None], yuv[None]
    for variant in ("default", "wide"):
        lib = emul_libs[variant]
        for comp in (0x07, 0x20, 0x50, 0x70, 0x05, 0x42):
            for kernel in (K_TMA, K_LDG):
                if kernel == K_TMA and (rgb.shape[2] * 4) % 16:
                    continue
                hist, wave, vs, _ = run(lib, rgb, yuv=yuv, surface=True, hist_comp=comp, wave_comp=comp, vscope=True,
                                        kernel=kernel, ctas=2, seed=comp + kernel, width=width)
                what = (case, variant, hex(comp), kernel)
                assert np.array_equal(hist[0].astype(np.float32).view(np.uint32),
                                      g[f"{case}/hist/{comp:02x}/log0"].view(np.uint32)), what
                gw = g[f"{case}/wave/{comp:02x}"]
                assert np.array_equal(wave[0][..., :3], gw[..., :3]) and not wave[0][..., 3].any(), what
                assert np.array_equal(vs[0], g[f"{case}/vscope"]), what


def test_emulator_catches_injected_bugs(oracle, pkg, tmp_path):
    """The emulation tests must have teeth: a copy of the kernel header with one deliberate bug each - the
    vectorscope's un-swizzle off by one bit (arithmetic), one arrival too few on the ring's "empty" barriers
    (protocol), a mailbox of one entry (protocol) - has to FAIL against the oracle or be stopped by the emulator."""
    import shutil
    csrc = os.path.join(ROOT, "obs-color-monitor_b200", "csrc")
    bugs = {
        "unswizzle": ("return (word & 0xFFu) ^ ((v7 & 7u) << 2);", "return (word & 0xFFu) ^ ((v7 & 3u) << 2);"),
        "arrivals": ("mbar_init(bar_empty + 8 * s, EMPTY_ARRIVALS);", "mbar_init(bar_empty + 8 * s, EMPTY_ARRIVALS - 1);"),
        "mailbox": ("constexpr int kQueue = SCOPE_DEEP_RING ? 8 : 4;", "constexpr int kQueue = 1;"),
    }
    frames = small_batch(pkg)
    yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
    procs = {}
    for name, (good, bad) in bugs.items():
        d = tmp_path / name
        d.mkdir()
        for f in ("scope_kernels.cuh", "scope_kernels_experiments.cuh", "scope_fused_v3.cuh"):
            shutil.copy(os.path.join(csrc, f), d / f)
        text = (d / "scope_kernels.cuh").read_text()
        assert text.count(good) == 1, f"the line the '{name}' bug replaces has changed"
        (d / "scope_kernels.cuh").write_text(text.replace(good, bad))
        procs[name] = subprocess.Popen(
            ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-DSCOPE_EMULATE", "-include",
             os.path.join(ROOT, "tools", "simt", "cuda_emul.h"), f"-I{d}", os.path.join(ROOT, "tools", "simt", "emul_main.cpp"),
             "-o", str(d / "lib.so")])
    for name, p in procs.items():
        assert p.wait() == 0
        lib = C.CDLL(str(tmp_path / name / "lib.so"))
        lib.emul_run.argtypes = [C.POINTER(Request)]
        lib.emul_build_flags.restype = C.c_char_p
        caught = 0
        for seed, land in ((1, 30), (2, 2), (3, 10)):
            try:
                out = run(lib, frames, seed=seed, land=land)
                check(oracle, frames, out, yuv, 0x07, 0x07, True, name)
            except AssertionError:
                caught += 1
        assert caught > 0, f"injected bug '{name}' went unnoticed"


# ---------------------------------------------------------------------------
# scope_fused_kernel_v3 (csrc/scope_fused_v3.cuh): the headline combination's own kernel
# ---------------------------------------------------------------------------
def test_v3_headline_batch(emul_libs, oracle, pkg):
    """4 frames x 3 strips (the last one 6 pixels wide), partial last tile, transparent pixels, both colour spaces,
    several CTAs sharing the chunk counter, TMA loads that land promptly / late / almost never"""
    lib = emul_libs["default"]
    frames = small_batch(pkg)
    for cs in (2, 1):
        yuv = [oracle.rgb_to_yuv(f, cs) for f in frames]
        for seed, land, ctas in ((1, 30, 2), (2, 3, 3), (3, 60, 1)):
            out = run(lib, frames, kernel=K_V3, seed=seed, land=land, ctas=ctas, colorspace=cs)
            check(oracle, frames, out, yuv, 0x07, 0x07, True, f"v3 cs {cs} seed {seed}")
    out = run(lib, frames, kernel=K_V3, hist_comp=0, seed=9)              # waveform + vectorscope, no histogram
    check(oracle, frames, out, [oracle.rgb_to_yuv(f, 2) for f in frames], 0, 0x07, True, "v3 without histogram")
    # tall frames: 7 tiles of 108 rows, so that the strips' first visits run through the lean loop (blocks inside the
    # frame with a successor, no per-visit checks) - random, natural and transparent content, late and prompt loads
    fr = pkg.frames
    tall = np.stack([fr.random(72, 700, 13), fr.alpha_stripes(72, 700, 14), fr.natural(72, 700, 15)])   # 72: 16-byte rows, wide write-out
    yuv = [oracle.rgb_to_yuv(f, 2) for f in tall]
    for seed, land, ctas in ((31, 30, 2), (32, 3, 1)):
        out = run(lib, tall, kernel=K_V3, seed=seed, land=land, ctas=ctas)
        check(oracle, tall, out, yuv, 0x07, 0x07, True, f"v3 tall seed {seed}")


def test_v3_saturation_flat_and_almost_flat(emul_libs, oracle, pkg):
    """solid frame (one vectorscope half beyond 0x8000: adds taken back; the flat-block path), a solid frame with
    specks and transparent pixels, and screen-like content where most blocks are flat and some are not"""
    lib = emul_libs["default"]
    w, h = 96, 800                       # 76 800 pixels on one bin > 65 535
    solid = pkg.frames.solid(w, h, (200, 17, 90, 255))
    speck = solid.copy()
    speck[::37, ::11] = (3, 250, 128, 255)
    speck[5::53, 7::13, 3] = 0
    frames = np.stack([solid, speck])
    yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
    out = run(lib, frames, kernel=K_V3, ctas=1, seed=11)
    check(oracle, frames, out, yuv, 0x07, 0x07, True, "v3 solid")
    assert out[2][0].max() == 255 and out[1][0].max() == 255
    ui = np.stack([pkg.frames.ui(96, 400, 1), pkg.frames.ui(96, 400, 2)])
    ui[1, 100:140, 10:50] = pkg.frames.random(40, 40, 3)
    ui[1, ::29, ::7, 3] = 0
    yuv = [oracle.rgb_to_yuv(f, 2) for f in ui]
    for seed, land in ((21, 30), (22, 3)):
        out = run(lib, ui, kernel=K_V3, ctas=2, seed=seed, land=land)
        check(oracle, ui, out, yuv, 0x07, 0x07, True, f"v3 ui seed {seed}")


def test_v3_frame_affine_claiming(emul_libs, oracle, pkg):
    """one work counter per frame (StripParams::frame_affine): more CTAs than frames, more frames than CTAs, one frame,
    chunk sizes 1 and > 1 - every strip of every frame exactly once, whatever order the CTAs come in (seeds without
    bit 2 run the per-frame counters, the emulator's launcher turns them off for the others)"""
    lib = emul_libs["default"]
    fr = pkg.frames
    for n_frames, w, h, ctas, seeds in ((2, 100, 120, 5, (1, 2, 8)), (7, 40, 110, 2, (3, 16)), (1, 130, 109, 3, (9,)),
                                        (12, 64, 40, 3, (10, 17)),
                                        (2, 1024, 8, 3, (1, 8))):       # many strips per frame: guided chunks > 1
        frames = np.stack([fr.random(w, h, 50 + i) if i % 2 else fr.natural(w, h, 60 + i) for i in range(n_frames)])
        yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
        for seed in seeds:
            assert not seed & 4
            out = run(lib, frames, kernel=K_V3, seed=seed, land=30, ctas=ctas)
            check(oracle, frames, out, yuv, 0x07, 0x07, True, f"v3 affine {n_frames} frames, {ctas} CTAs, seed {seed}")


def test_v3_background_lanes(emul_libs, oracle, pkg):
    """v3_block_mixed: blocks in which 8 or more lanes hold four equal pixels (text on a flat background).  The patches
    are small enough that their vectorscope bins stay far below 255 - a wrong count of background lanes shows (on
    screen-like frames every such bin saturates and hides it): 12 lanes of one colour, 31 lanes, 7 lanes (below the
    threshold: ordinary block), flat lanes of two colours (only the first lane's colour is the background), a flat
    patch with transparent pixels (per-pixel path), inside the lean loop (tall frame) and in a strip's last tiles."""
    lib = emul_libs["default"]
    f = pkg.frames.random(64, 700, seed=77)
    f[..., 3] = 255

    def patch(y0, x0, n, colour):
        f[y0:y0 + 4, x0:x0 + n] = colour
    patch(8, 10, 12, (200, 10, 60, 255))
    patch(116, 33, 31, (12, 240, 99, 255))
    patch(224, 3, 7, (77, 77, 200, 255))
    patch(332, 0, 10, (5, 130, 250, 255))
    patch(332, 12, 9, (250, 130, 5, 255))
    patch(440, 40, 16, (90, 90, 90, 255))
    f[441, 44, 3] = 0
    patch(692, 2, 20, (33, 66, 99, 255))          # rows 692..695: the strip's last (partial) tile
    frames = np.ascontiguousarray(f[None])
    yuv = [oracle.rgb_to_yuv(frames[0], 2)]
    for seed, land in ((41, 30), (42, 3)):
        out = run(lib, frames, kernel=K_V3, seed=seed, land=land, ctas=2)
        check(oracle, frames, out, yuv, 0x07, 0x07, True, f"v3 background lanes seed {seed}")
    assert out[2][0].max() < 255


def test_v3_extreme_geometries(emul_libs, oracle, pkg):
    """one row, one tile exactly (23 warps x 4 rows = 92), one row more, fewer rows than one warp takes, narrow strips"""
    lib = emul_libs["default"]
    for w, h in ((4, 1), (36, 1), (4, 97), (8, 300), (32, 92), (64, 93), (32, 63), (40, 3)):
        f = np.ascontiguousarray(pkg.frames.random(w, h, seed=w * 1000 + h)[None])
        yuv = [oracle.rgb_to_yuv(f[0], 2)]
        out = run(lib, f, kernel=K_V3, ctas=2, seed=w + h)
        check(oracle, f, out, yuv, 0x07, 0x07, True, f"v3 {w}x{h}")


def test_v3_emulator_catches_injected_bugs(oracle, pkg, tmp_path):
    """one deliberate bug each in a copy of scope_fused_v3.cuh: the flush's division by 132 with a wrong magic number,
    the wrong bit of U selecting a word's half, one arrival too few on the "empty" barriers"""
    import shutil
    csrc = os.path.join(ROOT, "obs-color-monitor_b200", "csrc")
    bugs = {
        "div132": ("const uint32_t v = ((w >> 2) * 1986u) >> 16;", "const uint32_t v = ((w >> 2) * 1900u) >> 16;"),
        "half": ("const F2 t = f2_fma(f2_pack(idx & 0xFFFFu, ru & 0x80u), k.vs_mul, k.vs_add);",
                 "const F2 t = f2_fma(f2_pack(idx & 0xFFFFu, ru & 0x40u), k.vs_mul, k.vs_add);"),
        "arrivals": ("mbar_init(bar_empty + 8 * s, V3::kWarps);", "mbar_init(bar_empty + 8 * s, V3::kWarps - 1);"),
    }
    frames = small_batch(pkg)
    yuv = [oracle.rgb_to_yuv(f, 2) for f in frames]
    procs = {}
    for name, (good, bad) in bugs.items():
        d = tmp_path / name
        d.mkdir()
        for f in ("scope_kernels.cuh", "scope_kernels_experiments.cuh", "scope_fused_v3.cuh"):
            shutil.copy(os.path.join(csrc, f), d / f)
        text = (d / "scope_fused_v3.cuh").read_text()
        assert text.count(good) == 1, f"the line the '{name}' bug replaces has changed"
        (d / "scope_fused_v3.cuh").write_text(text.replace(good, bad))
        procs[name] = subprocess.Popen(
            ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-DSCOPE_EMULATE", "-include",
             os.path.join(ROOT, "tools", "simt", "cuda_emul.h"), f"-I{d}", os.path.join(ROOT, "tools", "simt", "emul_main.cpp"),
             "-o", str(d / "lib.so")])
    for name, p in procs.items():
        assert p.wait() == 0
        lib = C.CDLL(str(tmp_path / name / "lib.so"))
        lib.emul_run.argtypes = [C.POINTER(Request)]
        caught = 0
        for seed, land in ((1, 30), (2, 2)):
            try:
                out = run(lib, frames, kernel=K_V3, seed=seed, land=land)
                check(oracle, frames, out, yuv, 0x07, 0x07, True, name)
            except AssertionError:
                caught += 1
        assert caught > 0, f"injected bug '{name}' went unnoticed"
