"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle, bit-exact.

Sizes here are ones the oracle finishes in seconds; full-size (4K / 8K) checks that do not
need the oracle live in test_gpu_properties.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COMPS = [0x07, 0x20, 0x50, 0x70]


def _frames(pkg):
    fr = pkg.frames
    return {
        "ramp_1080p": fr.ramp(1920, 1080),
        "random_odd": fr.random(641, 363, seed=3),
        "solid": fr.solid(320, 300, (10, 200, 97, 255)),
        "alpha": fr.alpha_stripes(333, 257),
        "natural": fr.natural(500, 270, seed=4),
        "narrow": fr.random(5, 700, seed=9),
        "one_px": fr.random(1, 1, seed=2),
    }


def _check(res, o, rgb, yuv, st, name):
    if "hist" in res:
        exp = o.histogram_counts(st.hist_components, rgb, yuv, colorspace=st.colorspace)
        assert np.array_equal(res["hist"], exp), f"{name}: histogram counts differ"
        expf, exphi = o.histogram_post(st.hist_components, rgb.shape[1], rgb.shape[0], exp, st.level_fixed_value,
                                       st.level_ratio_value, st.logscale)
        assert np.array_equal(res["hist_max"], exphi), f"{name}: hi_max differs"
        assert np.array_equal(res["hist_float"].view(np.uint32), expf.view(np.uint32)), f"{name}: hist float differs"
    if "wave" in res:
        exp = o.waveform(st.wave_components, rgb, yuv, colorspace=st.colorspace)
        assert np.array_equal(res["wave"], exp), f"{name}: waveform differs"
        if "wave_display" in res:
            assert np.array_equal(res["wave_display"], o.apply_intensity(exp, st.wave_intensity))
    if "vscope" in res:
        exp = o.vectorscope(yuv, colorspace=st.colorspace)
        assert np.array_equal(res["vscope"], exp), f"{name}: vectorscope differs"
        if "vscope_display" in res:
            assert np.array_equal(res["vscope_display"], o.apply_intensity(exp, st.vscope_intensity))


@pytest.mark.parametrize("colorspace", [1, 2])
def test_fused_all_scopes_host(engine, oracle, pkg, colorspace):
    for name, f in _frames(pkg).items():
        st = pkg.ScopeSettings(colorspace=colorspace, wave_intensity=51, vscope_intensity=25)
        res = engine.accumulate_host(f, settings=st)
        yuv = oracle.rgb_to_yuv(f, colorspace)
        _check(res, oracle, f, yuv, st, f"{name}/cs{colorspace}")


@pytest.mark.parametrize("hc", COMPS)
@pytest.mark.parametrize("wc", COMPS)
def test_component_matrix_fused(engine, oracle, pkg, hc, wc):
    f = pkg.frames.alpha_stripes(257, 131, seed=5)
    yuv = oracle.rgb_to_yuv(f, 2)
    st = pkg.ScopeSettings(hist_components=hc, wave_components=wc)
    _check(engine.accumulate_host(f, settings=st), oracle, f, yuv, st, f"h{hc:x}w{wc:x}")


@pytest.mark.parametrize("hc,wc", [(0x07, 0x07), (0x70, 0x07), (0x20, 0x50), (0x05, 0x42), (0x77, 0x30)])
def test_surface_mode(engine, oracle, pkg, hc, wc):
    """Both planes supplied like cm_surface_data: integer-only path, arbitrary YUV bytes
    (including alpha == 0 in the YUV plane)."""
    f = pkg.frames.alpha_stripes(300, 200, seed=7)
    yuv = pkg.frames.alpha_stripes(300, 200, seed=8, period=5)
    st = pkg.ScopeSettings(mode=pkg.MODE_SURFACE, hist_components=hc, wave_components=wc)
    _check(engine.accumulate_host(f, yuv, settings=st), oracle, f, yuv, st, f"surface h{hc:x}w{wc:x}")


def test_single_scopes_and_pitch(engine, oracle, pkg):
    f = pkg.frames.random(203, 99, seed=11)
    yuv = oracle.rgb_to_yuv(f, 2)
    for ls in (203 * 4, 203 * 4 + 4, 203 * 4 + 52, 1024):
        buf = pkg.frames.with_pitch(f, ls)
        for scopes in (pkg.SCOPE_HIST, pkg.SCOPE_WAVE, pkg.SCOPE_VSCOPE, pkg.SCOPE_HIST | pkg.SCOPE_VSCOPE):
            st = pkg.ScopeSettings(scopes=scopes)
            res = engine.accumulate_host(buf, settings=st, width=203)
            _check(res, oracle, f, yuv, st, f"pitch{ls}/scopes{scopes}")


def test_hist_levels_and_log(engine, oracle, pkg):
    f = pkg.frames.natural(320, 180, seed=1)
    yuv = oracle.rgb_to_yuv(f, 2)
    for kw in (dict(level_fixed_value=500), dict(level_ratio_value=7), dict(logscale=True),
               dict(logscale=True, level_ratio_value=3)):
        st = pkg.ScopeSettings(scopes=pkg.SCOPE_HIST, **kw)
        _check(engine.accumulate_host(f, settings=st), oracle, f, yuv, st, str(kw))


def test_saturation_solid(engine, oracle, pkg):
    """solid colour: waveform bin = min(H,255) in every column, one vectorscope bin = 255,
    histogram bin = W*H (SURVEY.md §8(c) known-answer test)."""
    w, h = 640, 400
    f = pkg.frames.solid(w, h, (30, 60, 90, 255))
    res = engine.accumulate_host(f)
    assert res["hist"][90 * 4 + 0] == w * h and res["hist"][60 * 4 + 1] == w * h and res["hist"][30 * 4 + 2] == w * h
    assert res["hist"].sum() == 3 * w * h
    assert (res["wave"][255 - 30, :, 0] == 255).all() and res["wave"].astype(np.int64).sum() == 3 * 255 * w
    assert res["vscope"].max() == 255 and np.count_nonzero(res["vscope"]) == 1
    yuv = oracle.rgb_to_yuv(f, 2)
    _check(res, oracle, f, yuv, pkg.ScopeSettings(), "solid")


def test_transform_exhaustive(engine, oracle):
    """the kernel's in-register RGB->YUV for all 2^24 colours == the pinned oracle"""
    for cs in (1, 2):
        exp, clamp = oracle.rgb_to_yuv_table(cs)
        assert not clamp
        got = engine.debug_yuv_table(cs).cpu().numpy().view(np.uint32)
        bad = np.nonzero(got != exp)[0]
        assert bad.size == 0, f"colorspace {cs}: {bad.size} colours differ, first {bad[:5]}"


def test_transform_exhaustive_v3(engine, oracle):
    """the headline kernel's form of the transform (integer sums, division by 10^6 as funnel shift + f32x2 add +
    f32x2 fused multiply-add, scope_fused_v3.cuh) for all 2^24 colours == the pinned oracle"""
    for cs in (1, 2):
        exp, clamp = oracle.rgb_to_yuv_table(cs)
        assert not clamp
        exp_uv = (exp & 0xFF) | ((exp >> 16) & 0xFF) << 8
        got = engine.debug_uv_table_v3(cs).cpu().numpy().view(np.uint32)
        bad = np.nonzero(got != exp_uv)[0]
        assert bad.size == 0, f"colorspace {cs}: {bad.size} colours differ, first {bad[:5]}"


def test_device_batch(engine, oracle, pkg):
    import torch
    fr = pkg.frames
    w, h, n = 416, 240, 5
    frames = np.stack([fr.mixed(w, h, i) for i in range(n)])
    d = torch.from_numpy(frames).cuda()
    st = pkg.ScopeSettings(vscope_intensity=25)
    out = engine.accumulate_device(d, settings=st)
    torch.cuda.synchronize()
    for i in range(n):
        yuv = oracle.rgb_to_yuv(frames[i], 2)
        res = {"hist": out["hist"][i].cpu().numpy().view(np.uint32),
               "wave": out["wave"][i].cpu().numpy(), "vscope": out["vscope"][i].cpu().numpy()}
        assert np.array_equal(res["hist"], oracle.histogram_counts(0x07, frames[i], yuv))
        assert np.array_equal(res["wave"], oracle.waveform(0x07, frames[i], yuv))
        exp = oracle.vectorscope(yuv)
        assert np.array_equal(res["vscope"], exp)
        assert np.array_equal(out["vscope_display"][i].cpu().numpy(), oracle.apply_intensity(exp, 25))
        _, hi = oracle.histogram_post(0x07, w, h, res["hist"])
        assert np.array_equal(out["hist_max"][i, :3].cpu().numpy().view(np.uint32), hi)


def test_unaligned_device_pointer_paths(engine, oracle, pkg):
    """ROI-style crops: base pointer 4-byte but not 16-byte aligned (TMA x-offset path) and a
    pitch that is not a multiple of 16 (plain-load path)."""
    import torch
    fr = pkg.frames
    big = fr.random(300, 120, seed=21)
    d = torch.from_numpy(big).cuda()
    for x0, y0, w, h in [(1, 3, 100, 50), (2, 0, 257, 120), (3, 7, 33, 64), (0, 0, 299, 119)]:
        # pitch 1200 B (TMA ok), base pointer misaligned by x0 pixels
        host = np.ascontiguousarray(big[y0:y0 + h, x0:x0 + w])
        yuv = oracle.rgb_to_yuv(host, 2)
        out = engine.accumulate_device(_as_pitched(d, x0, y0, h), width=w)
        torch.cuda.synchronize()
        assert np.array_equal(out["hist"][0].cpu().numpy().view(np.uint32), oracle.histogram_counts(7, host, yuv))
        assert np.array_equal(out["wave"][0].cpu().numpy(), oracle.waveform(7, host, yuv))
        assert np.array_equal(out["vscope"][0].cpu().numpy(), oracle.vectorscope(yuv))
    odd = fr.random(75, 40, seed=5)                # pitch 300 B: not a multiple of 16
    out = engine.accumulate_device(torch.from_numpy(odd).cuda()[None])
    yuv = oracle.rgb_to_yuv(odd, 2)
    torch.cuda.synchronize()
    assert np.array_equal(out["wave"][0].cpu().numpy(), oracle.waveform(7, odd, yuv))
    assert np.array_equal(out["vscope"][0].cpu().numpy(), oracle.vectorscope(yuv))
    assert np.array_equal(out["hist"][0].cpu().numpy().view(np.uint32), oracle.histogram_counts(7, odd, yuv))


def _as_pitched(d, x0, y0, h):
    """(1, h, linesize) byte view of rows y0.. starting at column x0 of a (H, W, 4) tensor."""
    import torch
    H, W, _ = d.shape
    flat = d.reshape(-1)
    off = (y0 * W + x0) * 4
    n = (h - 1) * W * 4 + (W - x0) * 4
    return torch.as_strided(flat[off:off + n], (1, h, (W - x0) * 4), (0, W * 4, 1))


def test_partial_tiles_equal_whole_frame(engine, oracle, pkg):
    """row bands and column bands accumulated as tiles, then finalized == whole frame"""
    import torch
    f = pkg.frames.natural(260, 300, seed=2)
    f[:, :, 3] = 255
    f[10:40, 5:9, 3] = 0
    d = torch.from_numpy(f).cuda()
    yuv = oracle.rgb_to_yuv(f, 2)
    st = pkg.ScopeSettings()
    for bands in ("rows", "cols"):
        part = engine.alloc_partial(260)
        if bands == "rows":
            for y0, y1 in [(0, 77), (77, 200), (200, 300)]:
                engine.accumulate_partial(d[y0:y1], part, x_offset=0, full_width=260, settings=st)
        else:
            for x0, x1 in [(0, 64), (64, 65), (65, 260)]:
                tile = _as_pitched(d, x0, 0, 300)[0]
                engine.accumulate_partial(tile, part, x_offset=x0, full_width=260, settings=st, width=x1 - x0)
        out = engine.finalize_partial(part, full_width=260, full_height=300, settings=st)
        torch.cuda.synchronize()
        assert np.array_equal(out["hist"][0].cpu().numpy().view(np.uint32), oracle.histogram_counts(7, f, yuv))
        assert np.array_equal(out["wave"][0].cpu().numpy(), oracle.waveform(7, f, yuv)), bands
        assert np.array_equal(out["vscope"][0].cpu().numpy(), oracle.vectorscope(yuv))


def test_error_codes_and_limits(engine, pkg):
    """bad requests fail with the documented status instead of computing something else"""
    import torch
    E = pkg._ffi
    f = pkg.frames.random(64, 48, seed=1)
    with pytest.raises(pkg.ScopeError) as e:                      # vectorscope in surface mode needs the YUV plane
        engine.accumulate_host(f, None, settings=pkg.ScopeSettings(mode=pkg.MODE_SURFACE))
    assert e.value.code == E.SCOPE_ERR_INVALID
    with pytest.raises(pkg.ScopeError) as e:                      # linesize < width*4
        engine.accumulate_host(np.zeros((8, 16), np.uint8), settings=pkg.ScopeSettings(), width=8)
    assert e.value.code == E.SCOPE_ERR_INVALID
    tall = torch.zeros((1, 70000, 4, 4), dtype=torch.uint8, device="cuda")
    with pytest.raises(pkg.ScopeError) as e:                      # u16 column bins: height <= 65535
        engine.accumulate_device(tall)
    assert e.value.code == E.SCOPE_ERR_UNSUPPORTED
    # the context is still usable afterwards
    res = engine.accumulate_host(f)
    assert res["hist"].sum() == 3 * 64 * 48
    # components that select no plane: the reference leaves all-zero buffers (histogram.c:372-373)
    res = engine.accumulate_host(f, settings=pkg.ScopeSettings(hist_components=0x00, wave_components=0x08))
    assert res["hist"].sum() == 0 and res["wave"].sum() == 0 and res["vscope"].sum() > 0


def test_tall_and_wide_extremes(engine, oracle, pkg):
    """very tall (many tiles, u16 bins close to their limit is covered by 8K in the property tests),
    very wide single row, width not a multiple of the strip, height not a multiple of the tile"""
    for w, h in [(3, 5000), (4100, 1), (95, 129), (33, 65)]:
        f = pkg.frames.random(w, h, seed=w + h)
        yuv = oracle.rgb_to_yuv(f, 2)
        st = pkg.ScopeSettings()
        _check(engine.accumulate_host(f, settings=st), oracle, f, yuv, st, f"{w}x{h}")


GOLDEN_CASES = ["ramp", "random", "solid", "alpha", "natural", "pitched"]
GOLDEN_COMPONENTS = [0x07, 0x20, 0x50, 0x70, 0x05, 0x42]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_surface_mode_matches_reference_golden(engine, pkg, case):
    """CUDA path vs the outputs of the reference's OWN loops (tests/golden/scope_golden.npz, made
    by tests/golden/make_golden.py from the unmodified src/{histogram,waveform,vectorscope}.c):
    same planes in, strict drop-in mode, every byte / float bit equal.  No oracle in between."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scope_golden.npz"))
    rgb, yuv, width = np.ascontiguousarray(g[f"{case}/rgb"]), np.ascontiguousarray(g[f"{case}/yuv"]), int(g[f"{case}/width"])
    for comp in GOLDEN_COMPONENTS:
        for kw, key in ((dict(logscale=False), "log0"), (dict(logscale=True), "log1"),
                        (dict(level_fixed_value=100), "fixed100"), (dict(level_ratio_value=5), "ratio5")):
            st = pkg.ScopeSettings(mode=pkg.MODE_SURFACE, hist_components=comp, wave_components=comp, **kw)
            res = engine.accumulate_host(rgb, yuv, settings=st, width=width)
            assert np.array_equal(res["hist_max"], g[f"{case}/hist_max/{comp:02x}/{key}"]), (case, hex(comp), key)
            if key in ("log0", "log1"):
                assert np.array_equal(res["hist_float"].view(np.uint32),
                                      g[f"{case}/hist/{comp:02x}/{key}"].view(np.uint32)), (case, hex(comp), key)
        assert np.array_equal(res["wave"], g[f"{case}/wave/{comp:02x}"]), (case, hex(comp))
        assert np.array_equal(res["vscope"], g[f"{case}/vscope"]), case


def test_transform_exhaustive_fp32_strict(engine, oracle):
    """SCOPE_XFORM_FP32_STRICT (SURVEY.md 8(c)'s fp32 reading of data/common.effect:23-43: every product and sum
    rounded separately, no FMA) for all 2^24 colours == the oracle's "fp32_strict" table"""
    for cs in (1, 2):
        exp, _ = oracle.rgb_to_yuv_table(cs, "fp32_strict")
        got = engine.debug_yuv_table_strict(cs).cpu().numpy().view(np.uint32)
        bad = np.nonzero(got != exp)[0]
        assert bad.size == 0, f"colorspace {cs}: {bad.size} colours differ, first {bad[:5]}"


def test_fp32_strict_mode_scopes(engine, oracle, pkg):
    """xform = SCOPE_XFORM_FP32_STRICT through the host and the device entry points: every scope equals the oracle's
    loops run on the YUV plane the strict table gives (fused mode; YUV components so that the transform feeds all
    three scopes).  The frame holds colours on which strict and exact differ, so the mode is really selected."""
    import torch
    exact, _ = oracle.rgb_to_yuv_table(2, "exact")
    strict, _ = oracle.rgb_to_yuv_table(2, "fp32_strict")
    diff = np.nonzero(exact != strict)[0][:4000]
    assert diff.size > 100
    f = pkg.frames.natural(200, 120, seed=5)
    flat = f.reshape(-1, 4)
    flat[: diff.size, 0], flat[: diff.size, 1], flat[: diff.size, 2] = diff & 0xFF, (diff >> 8) & 0xFF, diff >> 16
    yuv = oracle.yuv_from_table(f, strict)
    assert not np.array_equal(yuv, oracle.rgb_to_yuv(f, 2))
    for comps in ((0x70, 0x70), (0x07, 0x07)):
        st = pkg.ScopeSettings(hist_components=comps[0], wave_components=comps[1], xform=1, vscope_intensity=25)
        res = engine.accumulate_host(f, settings=st)
        _check(res, oracle, f, yuv, st, f"strict host {comps}")
        out = engine.accumulate_device(torch.from_numpy(f[None]).cuda(), settings=st)
        assert np.array_equal(out["vscope"][0].cpu().numpy(), oracle.vectorscope(yuv))
        assert np.array_equal(out["wave"][0].cpu().numpy(), oracle.waveform(comps[1], f, yuv))


@pytest.mark.parametrize("scale", [2, 3, 4, 7])
def test_target_scale(engine, oracle, pkg, scale):
    """scope_params.target_scale (common.c:88-90,249-250): the scopes of the point-downsampled surface, host entry
    point (rows dropped by the copy), ring and device entry point (rows and columns picked by the kernel)"""
    import torch
    w, h = 333, 207                       # neither a multiple of the scale
    f = pkg.frames.natural(w, h, seed=scale)
    f[::5, ::3, 3] = 0                    # some transparent pixels
    small = oracle.downsample(f, scale)
    assert small.shape == (h // scale, w // scale, 4)
    yuv = oracle.rgb_to_yuv(small, 2)
    st = pkg.ScopeSettings(target_scale=scale, vscope_intensity=25, wave_intensity=51, level_ratio_value=500)
    res = engine.accumulate_host(f, settings=st)
    _check(res, oracle, small, yuv, st, f"scale {scale} host")
    assert engine.submit_host(1, f, settings=st) is True
    _check(engine.wait_host(1), oracle, small, yuv, st, f"scale {scale} ring")
    out = engine.accumulate_device(torch.from_numpy(np.stack([f, f])).cuda(), settings=st)
    for i in range(2):
        assert np.array_equal(out["hist"][i].cpu().numpy().view(np.uint32), oracle.histogram_counts(7, small, yuv).ravel())
        assert np.array_equal(out["wave"][i].cpu().numpy(), oracle.waveform(7, small, yuv))
        assert np.array_equal(out["vscope"][i].cpu().numpy(), oracle.vectorscope(yuv))
    # strict drop-in mode: both planes scaled the same way
    full_yuv = oracle.rgb_to_yuv(f, 2)
    st2 = pkg.ScopeSettings(mode=pkg.MODE_SURFACE, target_scale=scale, hist_components=0x70, wave_components=0x20)
    res = engine.accumulate_host(f, full_yuv, settings=st2)
    _check(res, oracle, small, oracle.downsample(full_yuv, scale), st2, f"scale {scale} surface mode")
