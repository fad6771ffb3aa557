"""The C shim (include/cm_shim.h) driven like the reference drives its callbacks, on the GPU:
same double buffers, same flip / no-flip behaviour, results == oracle; the fused ROI fan-out
gives the same bytes as the three separate callbacks."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class SurfaceData(C.Structure):
    _fields_ = [("rgb_data", C.c_void_p), ("yuv_data", C.c_void_p), ("linesize", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("colorspace", C.c_int), ("tex", C.c_void_p)]


class His(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("components", C.c_uint32), ("level_fixed_value", C.c_int),
                ("level_ratio_value", C.c_int), ("logscale", C.c_bool), ("tex_buf", C.c_void_p * 2),
                ("hi_max", (C.c_uint32 * 3) * 2), ("w_tex_buf", C.c_int)]


class Wvs(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("components", C.c_uint32), ("tex_buf", C.c_void_p * 2),
                ("tex_buf_width", C.c_uint32 * 2), ("w_tex_buf", C.c_int)]


class Vss(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("tex_buf", C.c_void_p * 2), ("tex_cs", C.c_int * 2),
                ("w_tex_buf", C.c_int)]


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
    lib = C.CDLL(pkg._ffi.SHIM_PATH)
    for n in ("b200_his_init", "b200_wvs_init"):
        getattr(lib, n).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.b200_vss_init.argtypes = [C.c_void_p, C.c_void_p]
    for n in ("b200_his_surface_cb", "b200_wvs_surface_cb", "b200_vss_surface_cb", "b200_roi_surface_cb"):
        getattr(lib, n).argtypes = [C.c_void_p, C.c_void_p]
    lib.b200_roi_init.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    return lib


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
    roi = C.create_string_buffer(512)
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
