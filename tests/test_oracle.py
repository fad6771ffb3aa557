"""CPU tests of the checker itself (no GPU):
  * oracle restatement == the reference's own loops (oracle/_ref, when built here)
  * oracle restatement == the committed golden vectors (generated FROM the reference by
    tests/golden/make_golden.py) — this one also runs on the GPU box, where /root/reference
    and possibly oracle/_ref do not exist
  * hand-checkable known answers
  * the pinned transform's arithmetic facts the CUDA kernel relies on
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scope_golden.npz")
COMPONENTS = [0x07, 0x20, 0x50, 0x70, 0x05, 0x42]
CASES = ["ramp", "random", "solid", "alpha", "natural", "pitched"]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_golden(oracle, golden, case):
    rgb, yuv, width = golden[f"{case}/rgb"], golden[f"{case}/yuv"], int(golden[f"{case}/width"])
    height = rgb.shape[0]
    for comp in COMPONENTS:
        counts = oracle.histogram_counts(comp, rgb, yuv, width=width)
        for log in (0, 1):
            flt, hi = oracle.histogram_post(comp, width, height, counts, logscale=bool(log))
            assert np.array_equal(flt.view(np.uint32), golden[f"{case}/hist/{comp:02x}/log{log}"].view(np.uint32))
            assert np.array_equal(hi, golden[f"{case}/hist_max/{comp:02x}/log{log}"])
        assert np.array_equal(oracle.histogram_post(comp, width, height, counts, level_fixed=100)[1],
                              golden[f"{case}/hist_max/{comp:02x}/fixed100"])
        assert np.array_equal(oracle.histogram_post(comp, width, height, counts, level_ratio=5)[1],
                              golden[f"{case}/hist_max/{comp:02x}/ratio5"])
        assert np.array_equal(oracle.waveform(comp, rgb, yuv, width=width), golden[f"{case}/wave/{comp:02x}"])
    assert np.array_equal(oracle.vectorscope(yuv, width=width), golden[f"{case}/vscope"])


def test_oracle_matches_reference_loops(oracle, ref, pkg):
    """bigger, seeded, straight against the reference's compiled loops (build container only)"""
    fr = pkg.frames
    frames = [fr.ramp(1920, 1080), fr.random(641, 363, 3), fr.solid(320, 300), fr.alpha_stripes(333, 257),
              fr.natural(400, 300, 5)]
    for f in frames:
        yuv = oracle.rgb_to_yuv(f, 2)
        for comp in (0x07, 0x20, 0x50, 0x70, 0x13, 0x64):
            c = oracle.histogram_counts(comp, f, yuv)
            for kw in (dict(), dict(logscale=True), dict(level_fixed=77), dict(level_ratio=9)):
                post, hi = oracle.histogram_post(comp, f.shape[1], f.shape[0], c, **kw)
                rp, rhi = ref.histogram(comp, f, yuv, **kw)
                assert np.array_equal(post.view(np.uint32), rp.view(np.uint32))
                assert np.array_equal(hi, rhi)
            assert np.array_equal(oracle.waveform(comp, f, yuv), ref.waveform(comp, f, yuv))
        assert np.array_equal(oracle.vectorscope(yuv), ref.vectorscope(yuv))


def test_config1_ramp_histogram_closed_form(oracle, pkg):
    """BASELINE config 1: 1920x1080 ramp, histogram RGB, bit-exact against the closed form."""
    w, h = 1920, 1080
    f = pkg.frames.ramp(w, h)
    c = oracle.histogram_counts(0x07, f, None).reshape(256, 4)
    xs, ys = np.arange(w) & 255, np.arange(h) & 255
    exp_b = np.bincount(xs, minlength=256) * h
    exp_g = np.bincount(ys, minlength=256) * w
    exp_r = np.bincount(((np.arange(w)[None, :] + np.arange(h)[:, None]) & 255).ravel(), minlength=256)
    assert np.array_equal(c[:, 2], exp_b) and np.array_equal(c[:, 1], exp_g) and np.array_equal(c[:, 0], exp_r)
    assert (c[:, 3] == 0).all() and c.sum() == 3 * w * h


def test_known_answers_solid(oracle, pkg):
    w, h = 64, 300
    f = pkg.frames.solid(w, h, (7, 100, 200, 255))
    yuv = oracle.rgb_to_yuv(f, 2)
    c = oracle.histogram_counts(0x07, f, yuv)
    assert c[200 * 4 + 0] == w * h and c[100 * 4 + 1] == w * h and c[7 * 4 + 2] == w * h and c.sum() == 3 * w * h
    wv = oracle.waveform(0x07, f, yuv)
    assert (wv[255 - 7, :, 0] == 255).all() and (wv[255 - 100, :, 1] == 255).all() and (wv[:, :, 3] == 0).all()
    vs = oracle.vectorscope(yuv)
    u, v = int(yuv[0, 0, 0]), int(yuv[0, 0, 2])
    assert vs[255 - v, u] == 255 and np.count_nonzero(vs) == 1


def test_alpha_rule_and_null_planes(oracle, pkg):
    f = pkg.frames.random(31, 17, seed=1)
    f[:, ::2, 3] = 0
    yuv = oracle.rgb_to_yuv(f, 2)
    assert (yuv[..., 3] == 255).all()                       # shader writes a = 1
    assert oracle.histogram_counts(0x07, f, yuv).sum() == 3 * 17 * 15   # a == 0 skipped
    assert oracle.histogram_counts(0x70, f, yuv).sum() == 3 * 17 * 31   # YUV plane never skipped
    assert oracle.vectorscope(yuv).astype(int).sum() == 17 * 31         # no alpha test
    assert oracle.histogram_counts(0x70, f, None).sum() == 0            # missing plane -> zeros
    assert oracle.waveform(0x07, None, yuv, width=31, height=17).sum() == 0


def test_transform_facts(oracle):
    """facts the CUDA kernel's arithmetic relies on, and how far fp32 pipelines sit from the pin"""
    for cs in (1, 2):
        tab, out_of_range = oracle.rgb_to_yuv_table(cs)
        assert not out_of_range, "the value must stay inside the UNORM range (the kernel has no clamp)"
        u, v = tab & 0xFF, tab >> 16
        assert 15 <= u.min() and u.max() <= 239 and 16 <= v.min() and v.max() <= 240
        for variant in ("fp32_strict", "fp32_contracted"):
            ftab, clamp = oracle.rgb_to_yuv_table(cs, variant)
            assert not clamp
            for shift in (0, 8, 16):   # U, Y, V
                a = ((tab >> shift) & 0xFF).astype(np.int32)
                b = ((ftab >> shift) & 0xFF).astype(np.int32)
                assert np.abs(a - b).max() <= 1
                assert int((a != b).sum()) < 500      # < 3e-5 of all colours, each by one step
    assert oracle.calc_colorspace(0) == 2 and oracle.calc_colorspace(1) == 1 and oracle.calc_colorspace(7) == 2


def test_transform_is_the_exact_value(oracle):
    """the pinned table == exact rational evaluation of the effect file's expression (python
    integers, independent of the C code), and the kernel's multiply-high division is exact"""
    from fractions import Fraction as F
    coef = {1: (("-0.147643", "-0.289855", "0.437500"), ("0.299000", "0.587000", "0.114000"),
                ("0.437500", "-0.366351", "-0.071147")),
            2: (("-0.100643", "-0.338571", "0.439216"), ("0.212600", "0.715200", "0.072200"),
                ("0.439216", "-0.398941", "-0.040273"))}
    off = (F(1, 2) - F(1, 256), F(0), F(1, 2))
    r, g, b = np.meshgrid(*(np.arange(256, dtype=np.int64),) * 3, indexing="ij")
    magic = -(-2 ** 48 // 10 ** 6)
    assert magic == 281474977                     # kDivMagic in scope_kernels.cuh
    for cs in (1, 2):
        tab = oracle.rgb_to_yuv_table(cs)[0].reshape(256, 256, 256).astype(np.int64)
        for ch in range(3):
            # scale everything by 256e6 so the U offset (1/2 - 1/256) is an integer too
            scale = 256 * 10 ** 6
            c = [int(F(x) * scale) for x in coef[cs][ch]]
            k = (255 * off[ch] + F(1, 2)) * scale
            assert k.denominator == 1
            exact = (c[0] * r + c[1] * g + c[2] * b + int(k)) // scale
            assert np.array_equal((tab >> (8 * ch)) & 0xFF, exact)
            # the kernel's form: S in units of 1e-6, q = byte 2 of the high word of S * magic
            s = (c[0] * r + c[1] * g + c[2] * b) // 256 + int(k) // 256
            assert s.min() >= 0 and s.max() < 2 ** 28
            hi = (s * magic) >> 32
            assert np.array_equal(hi >> 16, exact) and (hi >> 24).max() == 0


def test_intensity_mapping(oracle):
    bins = np.arange(256, dtype=np.uint8)
    out = oracle.apply_intensity(bins, 25)
    assert out[0] == 0 and out[10] == 250 and out[11] == 255 and (out[11:] == 255).all()
    assert np.array_equal(oracle.apply_intensity(bins, 1), bins)


def test_oracle_matches_reference_loops_random_geometry(oracle, ref):
    """Property check against the reference's compiled loops (build container only): random small
    geometries (including one-pixel planes and pitched rows with garbage padding), random component
    masks, random alpha with many zeros, arbitrary bytes in the YUV plane."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None, derandomize=True)
    @given(w=st.integers(1, 70), h=st.integers(1, 40), pad=st.integers(0, 3), comp=st.integers(0, 0x77),
           seed=st.integers(0, 2 ** 31 - 1), log=st.booleans())
    def check(w, h, pad, comp, seed, log):
        rng = np.random.default_rng(seed)
        ls = w * 4 + 4 * pad
        rgb = rng.integers(0, 256, size=(h, ls), dtype=np.uint8)
        yuv = rng.integers(0, 256, size=(h, ls), dtype=np.uint8)
        a = rgb[:, 3:w * 4:4]
        a[rng.random(a.shape) < 0.3] = 0            # transparent pixels are skipped, a = 1..255 counted
        post, hi = oracle.draw_histogram(comp, rgb, yuv, width=w, logscale=log, hi_init=(7, 8, 9))
        rp, rhi = ref.histogram(comp, rgb, yuv, width=w, logscale=log, hi_init=(7, 8, 9))
        assert np.array_equal(post.view(np.uint32), rp.view(np.uint32)) and np.array_equal(hi, rhi)
        if comp & 0x77:   # the two-step form the GPU tests use agrees whenever a plane is selected
            c = oracle.histogram_counts(comp, rgb, yuv, width=w)
            post2, hi2 = oracle.histogram_post(comp, w, h, c, logscale=log)
            assert np.array_equal(post2.view(np.uint32), rp.view(np.uint32)) and np.array_equal(hi2, rhi)
        else:             # no plane selected: zeroed buffer, level pass never reached (histogram.c:366-373)
            assert not post.any() and tuple(hi) == (7, 8, 9)
        assert np.array_equal(oracle.waveform(comp, rgb, yuv, width=w), ref.waveform(comp, rgb, yuv, width=w))
        assert np.array_equal(oracle.vectorscope(yuv, width=w), ref.vectorscope(yuv, width=w))

    check()


def test_transform_coefficients_are_the_effect_files(oracle):
    """The 18 coefficients of data/common.effect:23-43, read from the reference's file itself (build container
    only), are the ones the product multiplies by (coef_for in csrc/scope_kernels.cuh, as integers x 10^6) and the ones
    the oracle's table is made from (checked through the table: one channel at a time, two colours each)."""
    import re
    effect = "/root/reference/data/common.effect"
    if not os.path.exists(effect):
        pytest.skip("reference tree not present")
    text = open(effect).read()
    got = {}
    for cs, name in ((1, "PSConvertRGB_YUV601"), (2, "PSConvertRGB_YUV709")):
        body = text[text.index(name):]
        body = body[:body.index("return")]
        rows = {}
        for comp, ch in (("z", "u"), ("y", "y"), ("x", "v")):     # uv00.z = U, .y = Y, .x = V
            m = re.search(r"uv00\." + comp + r"\s*=\s*([+-][0-9.]+)\s*\*\s*rgb\.x\s*([+-][0-9.]+)\s*\*\s*rgb\.y\s*"
                          r"([+-][0-9.]+)\s*\*\s*rgb\.z([^;]*);", body)
            rows[ch] = ([m.group(i) for i in (1, 2, 3)], m.group(4).replace(" ", ""))
        assert rows["u"][1] == "+0.5-1.0/256.0" and rows["y"][1] == "" and rows["v"][1] == "+0.5"
        got[cs] = [[round(float(c) * 10 ** 6) for c in rows[ch][0]] for ch in "uyv"]
        assert all(len(c.split(".")[1]) == 6 for ch in "uyv" for c in rows[ch][0]), "six decimals: S is an integer"
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "obs-color-monitor_b200", "csrc",
                            "scope_kernels.cuh")).read()
    for cs, tag in ((1, "k601"), (2, "k709")):
        m = re.search(tag + r"\[3\]\[3\]\s*=\s*\{(.*?)\};", src, re.S)
        product = [int(v) for v in re.findall(r"[+-]?\d+", m.group(1))]
        assert product == [c for row in got[cs] for c in row], tag
    for cs in (1, 2):
        tab = oracle.rgb_to_yuv_table(cs)[0]
        off = (127003906, 500000, 128000000)          # floor(10^6 (255 off + 1/2)), DESIGN.md section 3
        for rgb in ((255, 0, 0), (0, 255, 0), (0, 0, 255), (13, 200, 77)):
            word = int(tab[rgb[0] << 16 | rgb[1] << 8 | rgb[2]])
            for ch in range(3):
                s = sum(c * v for c, v in zip(got[cs][ch], rgb)) + off[ch]
                assert (word >> (8 * ch)) & 0xFF == s // 10 ** 6


def test_device_side_generators_match_frames_py(pkg):
    """the integer-only frame families are bit-identical on the torch side (bench.py's bulk data) and in
    frames.py (what the parity tests and the CPU baseline use)"""
    import torch
    from obs_color_monitor_b200 import frames_torch
    cpu = torch.device("cpu")
    for seed in (0, 7):
        assert np.array_equal(pkg.frames.ui(700, 130, seed),
                              frames_torch.mixed_batch(1, 700, 130, cpu, first_index=seed, content="ui")[0].numpy())
    assert np.array_equal(pkg.frames.ramp(300, 77), frames_torch.mixed_batch(1, 300, 77, cpu, content="ramp")[0].numpy())
    assert np.array_equal(pkg.frames.mixed(64, 48, 6), frames_torch.mixed_batch(1, 64, 48, cpu, first_index=6, content="solid")[0].numpy())


def test_fused_transform_ranges_the_headline_kernel_relies_on(oracle):
    """scope_fused_kernel_v3 sizes its vectorscope table for V in [16, 240] (csrc/scope_fused_v3.cuh: words
    132 * 16 .. 132 * 241): true for all 2^24 colours in both colour spaces, for the exact transform it evaluates"""
    for cs in (1, 2):
        t, clamp = oracle.rgb_to_yuv_table(cs)
        assert not clamp
        v = (t >> 16) & 0xFF
        u = t & 0xFF
        assert int(v.min()) >= 16 and int(v.max()) <= 240, (cs, int(v.min()), int(v.max()))
        assert int(u.min()) >= 15 and int(u.max()) <= 239
