"""The C shim (include/cm_shim.h) driven like the reference drives its callbacks, on the GPU:
same double buffers, same flip / no-flip behaviour, results == oracle; the fused ROI fan-out
gives the same bytes as the three separate callbacks."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from obs_color_monitor_b200 import shim as S  # noqa: E402  (ctypes mirror of include/cm_shim.h)

SurfaceData, His, Wvs, Vss = S.SurfaceData, S.HisSource, S.WvsSource, S.VssSource


def _arr(ptr, n, dtype):
    return np.frombuffer((C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr), dtype=dtype).copy()


def _sd(rgb, yuv, cs=2):
    s = SurfaceData()
    s.rgb_data = rgb.ctypes.data if rgb is not None else None
    s.yuv_data = yuv.ctypes.data if yuv is not None else None
    ref = rgb if rgb is not None else yuv
    s.linesize, s.width, s.height, s.colorspace = ref.shape[1] * 4, ref.shape[1], ref.shape[0], cs
    return s


@pytest.fixture()
def shim(pkg, engine):
    return S.load()


def test_callbacks_surface_mode_match_oracle_and_flip(shim, engine, oracle, pkg):
    f = pkg.frames.alpha_stripes(320, 200, seed=3)
    yuv = oracle.rgb_to_yuv(f, 2)
    his, wvs, vss = His(), Wvs(), Vss()
    ctx = engine.ctx.handle
    shim.b200_his_init(C.byref(his), ctx, 0x07)
    shim.b200_wvs_init(C.byref(wvs), ctx, 0x70)
    shim.b200_vss_init(C.byref(vss), ctx)
    sd = _sd(f, yuv)
    for cb, src in ((shim.b200_his_surface_cb, his), (shim.b200_wvs_surface_cb, wvs), (shim.b200_vss_surface_cb, vss)):
        assert src.w_tex_buf == 0
        cb(C.byref(src), C.byref(sd))
        assert src.w_tex_buf == 1                       # flipped: the reader takes buffer 0
    counts = oracle.histogram_counts(0x07, f, yuv)
    flt, hi = oracle.histogram_post(0x07, 320, 200, counts)
    assert np.array_equal(_arr(his.tex_buf[0], 1024, np.float32).view(np.uint32), flt.view(np.uint32))
    assert list(his.hi_max[0]) == list(hi)
    assert np.array_equal(_arr(wvs.tex_buf[0], 256 * 320 * 4, np.uint8).reshape(256, 320, 4),
                          oracle.waveform(0x70, f, yuv))
    assert wvs.tex_buf_width[0] == 320
    assert np.array_equal(_arr(vss.tex_buf[0], 65536, np.uint8).reshape(256, 256), oracle.vectorscope(yuv))
    assert vss.tex_cs[0] == 2
    # missing plane -> early return, NO flip, previous result stays (histogram.c:436-441 etc.)
    sd_bad = _sd(f, None)
    shim.b200_vss_surface_cb(C.byref(vss), C.byref(sd_bad))
    shim.b200_wvs_surface_cb(C.byref(wvs), C.byref(sd_bad))
    assert vss.w_tex_buf == 1 and wvs.w_tex_buf == 1
    shim.b200_his_surface_cb(C.byref(his), C.byref(sd_bad))   # RGB histogram only needs rgb_data
    assert his.w_tex_buf == 0


def test_roi_fanout_fused_equals_separate(shim, engine, oracle, pkg):
    f = pkg.frames.natural(256, 144, seed=9)
    ctx = engine.ctx.handle
    roi_obj = S.RoiSource()
    roi = C.byref(roi_obj)
    shim.b200_roi_init(roi, ctx, 0)                       # SCOPE_MODE_FUSED: only rgb_data is supplied
    his, wvs, vss = His(), Wvs(), Vss()
    shim.b200_his_init(C.byref(his), ctx, 0x07)
    shim.b200_wvs_init(C.byref(wvs), ctx, 0x07)
    shim.b200_vss_init(C.byref(vss), ctx)
    assert shim.b200_roi_register_his(roi, C.byref(his)) == 0
    assert shim.b200_roi_register_wvs(roi, C.byref(wvs)) == 0
    assert shim.b200_roi_register_vss(roi, C.byref(vss)) == 0
    sd = _sd(f, None)
    launches0 = engine.launch_count
    shim.b200_roi_surface_cb(roi, C.byref(sd))
    fused_launches = engine.launch_count - launches0
    assert (his.w_tex_buf, wvs.w_tex_buf, vss.w_tex_buf) == (1, 1, 1)
    yuv = oracle.rgb_to_yuv(f, 2)
    flt, _ = oracle.histogram_post(0x07, 256, 144, oracle.histogram_counts(0x07, f, yuv))
    assert np.array_equal(_arr(his.tex_buf[0], 1024, np.float32).view(np.uint32), flt.view(np.uint32))
    assert np.array_equal(_arr(wvs.tex_buf[0], 256 * 256 * 4, np.uint8).reshape(256, 256, 4), oracle.waveform(0x07, f, yuv))
    assert np.array_equal(_arr(vss.tex_buf[0], 65536, np.uint8).reshape(256, 256), oracle.vectorscope(yuv))
    assert fused_launches <= 3                            # one accumulation pass (+ finalize kernels)


def _results(his, wvs, vss, width):
    """the buffers a reader would take: tex_buf[w_tex_buf ^ 1]"""
    out = {}
    if his is not None:
        r = his.w_tex_buf ^ 1
        out["hist_float"] = _arr(his.tex_buf[r], 1024, np.float32)
        out["hist_max"] = list(his.hi_max[r])
    if wvs is not None:
        r = wvs.w_tex_buf ^ 1
        assert wvs.tex_buf_width[r] == width
        out["wave"] = _arr(wvs.tex_buf[r], 256 * width * 4, np.uint8).reshape(256, width, 4)
    if vss is not None:
        r = vss.w_tex_buf ^ 1
        out["vscope"] = _arr(vss.tex_buf[r], 65536, np.uint8).reshape(256, 256)
    return out


@pytest.mark.parametrize("zero_copy", [False, True])
@pytest.mark.parametrize("scale", [1, 2, 3])
def test_capture_core_on_the_gpu_ring(shim, engine, oracle, pkg, zero_copy, scale):
    """b200_cm_attach_gpu: queue slot i stages into ring slot i (one host copy, or none with zero_copy), the worker's
    ROI fan-out SUBMITS a surface and files the results of the one before it (the copy of frame n overlaps the
    kernels and the read-back of frame n - 1), target_scale travels in the hint and is applied by the copy (rows) and
    the kernel (columns); an ROI rectangle (in pixels of the scaled surface) crops what is staged.  Every filed
    result is compared with the oracle on the same scaled, cropped surface."""
    lib, ctx = shim, engine.ctx.handle
    W, H = 328, 204
    frames = [pkg.frames.natural(W, H, seed=s) for s in range(5)] + [pkg.frames.alpha_stripes(W, H, seed=7)]
    if zero_copy:                                   # page-locked, and alive until the worker is done with them
        import torch
        pinned = [torch.from_numpy(f.copy()).pin_memory() for f in frames]
        frames = [p.numpy() for p in pinned]
    roi_rect = (4, 2, 4 + 96, 2 + 50)               # x0, y0, x1, y1 on the scaled surface
    for crop in (False, True):
        cm, roi = S.CmSource(), S.RoiSource()
        his, wvs, vss = His(), Wvs(), Vss()
        lib.b200_cm_create(C.byref(cm))
        lib.b200_cm_attach_gpu(C.byref(cm), ctx, zero_copy)
        cm.target_scale = scale
        lib.b200_roi_init(C.byref(roi), ctx, 0)     # SCOPE_MODE_FUSED
        lib.b200_his_init(C.byref(his), ctx, 0x07)
        lib.b200_wvs_init(C.byref(wvs), ctx, 0x07)
        lib.b200_vss_init(C.byref(vss), ctx)
        for reg, src in ((lib.b200_roi_register_his, his), (lib.b200_roi_register_wvs, wvs), (lib.b200_roi_register_vss, vss)):
            assert reg(C.byref(roi), C.byref(src)) == 0
        cm.flags = lib.b200_roi_capture_flags(C.byref(roi)) & ~S.CM_FLAG_ROI
        cb = C.cast(lib.b200_roi_surface_cb, C.c_void_p)
        lib.b200_cm_request(C.byref(cm), cb, C.cast(C.byref(roi), C.c_void_p))
        sw, sh = W // scale, H // scale
        if crop:
            lib.b200_cm_set_roi(C.byref(cm), *roi_rect, sw, sh)
        expected = []
        for f in frames:
            small = oracle.downsample(f, scale)
            if crop:
                small = np.ascontiguousarray(small[roi_rect[1]:roi_rect[3], roi_rect[0]:roi_rect[2]])
            expected.append(small)
        filed = []
        for i, f in enumerate(frames + frames[:1]):          # one more render pushes the last frame to the worker
            lib.b200_cm_tick(C.byref(cm))
            assert lib.b200_cm_render_target(C.byref(cm), f.ctypes.data, None, W * 4, W, H) is True
            lib.b200_cm_drain(C.byref(cm))
            if roi.frames_filed > len(filed):                 # results of the surface before the one just submitted
                filed.append(_results(his, wvs, vss, expected[len(filed)].shape[1]))
        lib.b200_roi_finish(C.byref(roi))
        if roi.frames_filed > len(filed):
            filed.append(_results(his, wvs, vss, expected[len(filed)].shape[1]))
        assert len(filed) == len(frames) and cm.frames_dropped == 0
        for i, (got, small) in enumerate(zip(filed, expected)):
            yuv = oracle.rgb_to_yuv(small, 2)
            flt, hi = oracle.histogram_post(0x07, small.shape[1], small.shape[0], oracle.histogram_counts(0x07, small, yuv))
            what = (zero_copy, scale, crop, i)
            assert np.array_equal(got["hist_float"].view(np.uint32), flt.view(np.uint32)) and got["hist_max"] == list(hi), what
            assert np.array_equal(got["wave"], oracle.waveform(0x07, small, yuv)), what
            assert np.array_equal(got["vscope"], oracle.vectorscope(yuv)), what
        lib.b200_cm_destroy(C.byref(cm))
        lib.b200_roi_destroy(C.byref(roi))
        for d, src in ((lib.b200_his_destroy, his), (lib.b200_wvs_destroy, wvs), (lib.b200_vss_destroy, vss)):
            d(C.byref(src))


def test_obs_source_info_tick_render_order(shim, engine, oracle, pkg):
    """The outer plugin ABI's shape (histogram.c:580-595, waveform.c:402-417, vectorscope.c:484-519): the three
    exported tables, driven the way libobs drives a source - get_defaults, create, then per frame video_tick(data,
    seconds) followed by video_render(data, effect) - with the defaults of the reference (target_scale 2).  What a
    render would upload (tex_buf[w ^ 1]) equals the oracle on the scaled frame, one frame late like the reference."""
    lib, ctx = shim, engine.ctx.handle
    W, H = 200, 120
    frames = [pkg.frames.natural(W, H, seed=s) for s in (1, 2, 3)]
    cur = {"f": frames[0]}

    def get_frame(_opaque, rgb, yuv, linesize, width, height):
        f = cur["f"]
        rgb[0], yuv[0] = f.ctypes.data, None
        linesize[0], width[0], height[0] = W * 4, W, H
        return True

    target = S.Target(S.GET_FRAME(get_frame), None)
    for name, kind in (("histogram", "hist"), ("waveform", "wave"), ("vectorscope", "vscope")):
        info = S.source_info(lib, name)
        assert info.id == f"{name}_source".encode() and info.type == 0 and info.output_flags & 0x9 == 0x9
        assert info.get_name(None).decode().lower() == name
        st = S.Settings()
        info.get_defaults(C.byref(st))
        assert st.target_scale == 2                        # histogram.c:166, waveform.c:113, vectorscope.c:157
        st.ctx, st.mode, st.colorspace, st.gpu_ring = ctx, 0, 2, name != "waveform"
        data = info.create(C.byref(st), C.byref(target))
        assert data
        for i, f in enumerate(frames + frames[-1:]):
            cur["f"] = f
            info.video_tick(data, C.c_float(1 / 60))
            info.video_render(data, None)
            info.video_render(data, None)                  # a second render in the same tick is ignored
            lib.b200_source_drain(data)
        width, aux = C.c_uint32(0), C.POINTER(C.c_uint32)()
        buf = lib.b200_source_result(data, C.byref(width), C.byref(aux))
        assert buf
        small = oracle.downsample(frames[-1], 2)
        yuv = oracle.rgb_to_yuv(small, 2)
        if kind == "hist":
            flt, hi = oracle.histogram_post(0x07, W // 2, H // 2, oracle.histogram_counts(0x07, small, yuv))
            assert (st.level_mode, st.level_fixed_value, st.level_ratio_value) == (0, 1000, 10.0)   # histogram.c:164-172
            assert np.array_equal(_arr(buf, 1024, np.float32).view(np.uint32), flt.view(np.uint32))
            assert [aux[0], aux[1], aux[2]] == list(hi) and info.get_width(data) == 256 and info.get_height(data) == 200
        elif kind == "wave":
            assert width.value == W // 2 == info.get_width(data) and info.get_height(data) == 256
            assert np.array_equal(_arr(buf, 256 * (W // 2) * 4, np.uint8).reshape(256, W // 2, 4), oracle.waveform(0x07, small, yuv))
        else:
            assert np.array_equal(_arr(buf, 65536, np.uint8).reshape(256, 256), oracle.vectorscope(yuv)) and aux[0] == 2
        info.destroy(data)
