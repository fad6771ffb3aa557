"""Host logic of the C shim's capture core (include/cm_shim.h), no GPU: the 3-slot queue, the
one-frame staging latency, drop-on-busy and once-per-tick rules of src/common.c."""
import ctypes as C
import threading
import time

import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from obs_color_monitor_b200 import shim as S  # noqa: E402  (ctypes mirror of include/cm_shim.h, sizes checked at load)

SurfaceData, CB = S.SurfaceData, S.SURFACE_CB
_His, _Wvs, _Vss, _Cm = S.HisSource, S.WvsSource, S.VssSource, S.CmSource


@pytest.fixture()
def shim(pkg):
    lib = S.load()
    lib.b200_cm_request.argtypes = [C.c_void_p, CB, C.c_void_p]
    return lib


def test_ring_orders_frames_and_drops_when_busy(shim, pkg):
    lib = shim
    Cm = S.CmSource
    cm = Cm()
    lib.b200_cm_create(C.byref(cm))
    assert (cm.i_write_queue, cm.i_staging_queue, cm.i_read_queue) == (0, 0, 2)
    cm.flags = 1  # CONVERT_RGB

    seen, gate = [], threading.Event()
    gate.set()

    def cb(_data, sd):
        gate.wait()
        s = sd.contents
        first = C.cast(s.rgb_data, C.POINTER(C.c_uint8))[0]
        seen.append((int(first), s.width, s.height, s.linesize, bool(s.yuv_data)))

    cfn = CB(cb)
    lib.b200_cm_request(C.byref(cm), cfn, None)
    w, h = 16, 4

    def frame(tag):
        return np.full((h, w * 4), tag, np.uint8)

    def render(tag):
        f = frame(tag)
        return lib.b200_cm_render_target(C.byref(cm), f.ctypes.data, None, w * 4, w, h)

    # once per tick
    lib.b200_cm_tick(C.byref(cm))
    assert render(1) is True
    assert render(99) is False
    # the staged frame is consumed only after the NEXT one is staged (one frame of latency)
    lib.b200_cm_drain(C.byref(cm))
    assert seen == []
    lib.b200_cm_tick(C.byref(cm))
    assert render(2) is True
    lib.b200_cm_drain(C.byref(cm))
    assert [s[0] for s in seen] == [1]
    assert seen[0][1:] == (w, h, w * 4, False)
    # block the worker inside the callback, keep rendering: frames get dropped, none reordered
    gate.clear()
    results = []
    for tag in range(3, 9):
        lib.b200_cm_tick(C.byref(cm))
        results.append(render(tag))
        time.sleep(0.01)
    assert results.count(False) >= 3 and cm.frames_dropped == results.count(False)
    gate.set()
    lib.b200_cm_drain(C.byref(cm))          # let the worker catch up before the next frames
    lib.b200_cm_tick(C.byref(cm))
    assert render(50) is True
    lib.b200_cm_tick(C.byref(cm))
    render(51)
    lib.b200_cm_drain(C.byref(cm))
    tags = [s[0] for s in seen]
    assert tags == sorted(tags) and tags[0] == 1 and 50 in tags
    assert cm.frames_processed == len(seen)
    lib.b200_cm_destroy(C.byref(cm))


def test_roi_interleave_pacing(shim, pkg):
    """n_interleave = 1 (the reference default): frames are staged on every other tick only."""
    lib = shim
    Roi = S.RoiSource
    roi = Roi()
    lib.b200_roi_init(C.byref(roi), None, 1)
    cm_obj = S.CmSource()
    cm = C.byref(cm_obj)
    lib.b200_cm_create(cm)
    seen = []
    cfn = CB(lambda _d, sd: seen.append(int(C.cast(sd.contents.rgb_data, C.POINTER(C.c_uint8))[0])))
    lib.b200_cm_request(cm, cfn, None)
    cm_obj.flags = 1            # CONVERT_RGB
    staged = []
    for interleave in (1, 0):
        roi.n_interleave, roi.i_interleave, roi.interleave_rendered = interleave, 0, False
        for tick in range(8):
            lib.b200_roi_tick(C.byref(roi), cm)
            f = np.full((4, 64), 10 * interleave + tick, np.uint8)
            before = cm_obj.i_write_queue
            lib.b200_roi_target_render(C.byref(roi), cm, f.ctypes.data, None, 64, 16, 4)
            staged.append((interleave, tick, cm_obj.i_write_queue != before))
            lib.b200_cm_drain(cm)
    on = [t for i, t, s in staged if i == 1 and s]
    assert on == [0, 2, 4, 6]                                        # every other tick
    assert all(s for i, t, s in staged if i == 0)                    # interleave off: every tick
    lib.b200_cm_destroy(cm)


def test_roi_crop_is_staged_like_the_reference(shim, pkg):
    """B200_CM_FLAG_ROI: the capture core stages only the ROI rectangle (common.c:272-291), clamped
    like roi_send_range (roi.c:478-500); the callback sees a cx x cy surface, RGB rows then YUV rows."""
    lib = shim
    W, H = 40, 24
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    yuv = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    cases = [((5, 3, 17, 20), (5, 3, 17, 20)),        # inside
             ((-4, -1, 100, 9), (0, 0, W, 9)),        # negative / too large ends snap to the border
             ((7, 2, -1, -1), (7, 2, W, H)),          # -1 = "to the end"
             ((9, 9, 9, 12), None)]                   # empty rectangle: whole frame (common.c:273 fails)
    for (x0, y0, x1, y1), want in cases:
        cm_obj = S.CmSource()
        cm = C.byref(cm_obj)
        lib.b200_cm_create(cm)
        got = []

        def cb(_d, sd):
            s = sd.contents
            n = s.linesize * s.height
            a = np.ctypeslib.as_array(C.cast(s.rgb_data, C.POINTER(C.c_uint8)), (n,)).copy()
            b = np.ctypeslib.as_array(C.cast(s.yuv_data, C.POINTER(C.c_uint8)), (n,)).copy()
            got.append((s.width, s.height, s.linesize, a, b))

        cfn = CB(cb)
        lib.b200_cm_request(cm, cfn, None)
        cm_obj.flags = 3   # CONVERT_RGB | CONVERT_YUV
        lib.b200_cm_set_roi(cm, x0, y0, x1, y1, W, H)
        assert cm_obj.flags == 3 | 8
        for _ in range(2):                                # one-frame staging latency: render twice
            lib.b200_cm_tick(cm)
            lib.b200_cm_render_target(cm, rgb.ctypes.data, yuv.ctypes.data, W * 4, W, H)
            lib.b200_cm_drain(cm)
        lib.b200_cm_destroy(cm)
        assert got, "callback never ran"
        w, h, ls, a, b = got[0]
        ex0, ey0, ex1, ey1 = want if want else (0, 0, W, H)
        assert (w, h) == (ex1 - ex0, ey1 - ey0)
        assert ls == (W * 4 if (w, h) == (W, H) else w * 4)
        assert np.array_equal(a.reshape(h, ls)[:, :w * 4], rgb[ey0:ey1, ex0:ex1].reshape(h, w * 4))
        assert np.array_equal(b.reshape(h, ls)[:, :w * 4], yuv[ey0:ey1, ex0:ex1].reshape(h, w * 4))


def test_callback_early_returns_match_the_reference(shim, ref):
    """The shim's b200_*_inputs_missing predicates (what makes a callback return without touching its
    buffers) against the reference's own his/wvs/vss_surface_cb compiled from its sources
    (oracle/_ref): the reference processed a surface iff it flipped w_tex_buf.  Whole truth table:
    components incl. none / both planes' bits, each plane present or NULL, empty surfaces."""
    lib, L = shim, ref.lib
    rgb = np.full((3, 16), 200, np.uint8)
    yuv = np.full((3, 16), 90, np.uint8)
    scratch = np.zeros(256 * 16 * 4 + 65536, np.uint8)
    aux = (C.c_uint32 * 4)()
    n_cases = 0
    for comp in (0x00, 0x07, 0x20, 0x50, 0x70, 0x77, 0x13, 0x08):
        for has_rgb in (False, True):
            for has_yuv in (False, True):
                for w, h in ((4, 3), (0, 3), (4, 0), (0, 0)):
                    sd = SurfaceData(rgb.ctypes.data if has_rgb else None, yuv.ctypes.data if has_yuv else None,
                                     16, w, h, 2, None)
                    args = (sd.rgb_data, sd.yuv_data, 16, w, h, 2, scratch.ctypes.data, C.cast(aux, C.c_void_p))
                    # reference: a fresh source starts with w_tex_buf = 0; 1 afterwards = processed
                    st = L.ref_his_new(comp, 0, 0, 0)
                    ref_his = L.ref_his_surface_cb(st, *args) == 1
                    L.ref_his_free(st)
                    st = L.ref_wvs_new(comp)
                    ref_wvs = L.ref_wvs_surface_cb(st, *args) == 1
                    L.ref_wvs_free(st)
                    st = L.ref_vss_new()
                    ref_vss = L.ref_vss_surface_cb(st, *args) == 1
                    L.ref_vss_free(st)
                    src = C.create_string_buffer(512)
                    lib.b200_his_init(src, None, comp)
                    assert lib.b200_his_inputs_missing(src, C.byref(sd)) == (not ref_his), (hex(comp), has_rgb, has_yuv, w, h)
                    lib.b200_wvs_init(src, None, comp)
                    assert lib.b200_wvs_inputs_missing(src, C.byref(sd)) == (not ref_wvs), (hex(comp), has_rgb, has_yuv, w, h)
                    lib.b200_vss_init(src, None)
                    assert lib.b200_vss_inputs_missing(src, C.byref(sd)) == (not ref_vss), (has_rgb, has_yuv, w, h)
                    n_cases += 1
    assert n_cases == 8 * 2 * 2 * 4


def test_empty_surfaces_match_the_reference(shim, ref):
    """Surfaces without rows (or without columns) never reach the GPU: the shim files the result the
    reference's loops leave when they do not iterate - zeroed buffer, level pass on zero counts, flip -
    or does nothing when the reference's callback returns early.  Compared field by field with the
    reference's own callbacks (oracle/_ref), standalone and through the ROI fan-out."""
    lib, L = shim, ref.lib
    rgb = np.full((3, 16), 200, np.uint8)
    yuv = np.full((3, 16), 90, np.uint8)
    n_flips = 0
    for via_roi in (False, True):
        for comp in (0x00, 0x07, 0x20, 0x70, 0x77):
            for has_rgb, has_yuv in ((True, True), (True, False), (False, True)):
                for w, h in ((4, 0), (0, 3), (0, 0)):
                    for fixed, ratio, log in ((0, 0, 0), (77, 0, 0), (0, 9, 1), (77, 9, 1)):
                        sd = SurfaceData(rgb.ctypes.data if has_rgb else None, yuv.ctypes.data if has_yuv else None,
                                         16, w, h, 1, None)
                        args = (sd.rgb_data, sd.yuv_data, 16, w, h, 1)
                        # --- the reference ---
                        hbuf, hmax = np.full(1024, 7, np.uint32), np.full(3, 7, np.uint32)
                        st = L.ref_his_new(comp, fixed, ratio, log)
                        r_his = L.ref_his_surface_cb(st, *args, hbuf.ctypes.data, hmax.ctypes.data)
                        L.ref_his_free(st)
                        wbuf, wwidth = np.full(256 * 16 * 4, 7, np.uint8), C.c_uint32(99)
                        st = L.ref_wvs_new(comp)
                        r_wvs = L.ref_wvs_surface_cb(st, *args, wbuf.ctypes.data, C.cast(C.byref(wwidth), C.c_void_p))
                        L.ref_wvs_free(st)
                        vbuf, vcs = np.full(65536, 7, np.uint8), C.c_int(99)
                        st = L.ref_vss_new()
                        r_vss = L.ref_vss_surface_cb(st, *args, vbuf.ctypes.data, C.cast(C.byref(vcs), C.c_void_p))
                        L.ref_vss_free(st)
                        # --- the shim (no GPU context: none of these cases may need one) ---
                        his, wvs, vss = _His(), _Wvs(), _Vss()
                        lib.b200_his_init(C.byref(his), None, comp)
                        his.level_fixed_value, his.level_ratio_value, his.logscale = fixed, ratio, bool(log)
                        lib.b200_wvs_init(C.byref(wvs), None, comp)
                        lib.b200_vss_init(C.byref(vss), None)
                        if via_roi:
                            roi_obj = S.RoiSource()
                            roi = C.byref(roi_obj)
                            lib.b200_roi_init(roi, None, 1)      # SCOPE_MODE_SURFACE
                            lib.b200_roi_register_his(roi, C.byref(his))
                            lib.b200_roi_register_wvs(roi, C.byref(wvs))
                            lib.b200_roi_register_vss(roi, C.byref(vss))
                            lib.b200_roi_surface_cb(roi, C.byref(sd))
                            lib.b200_roi_destroy(roi)
                        else:
                            lib.b200_his_surface_cb(C.byref(his), C.byref(sd))
                            lib.b200_wvs_surface_cb(C.byref(wvs), C.byref(sd))
                            lib.b200_vss_surface_cb(C.byref(vss), C.byref(sd))
                        case = (via_roi, hex(comp), has_rgb, has_yuv, w, h, fixed, ratio, log)
                        assert his.w_tex_buf == r_his and wvs.w_tex_buf == r_wvs and vss.w_tex_buf == r_vss, case
                        if r_his:
                            got = np.ctypeslib.as_array(C.cast(his.tex_buf[0], C.POINTER(C.c_uint32)), (1024,))
                            assert np.array_equal(got, hbuf) and list(his.hi_max[0]) == list(hmax), case
                        if r_wvs:
                            assert wvs.tex_buf_width[0] == wwidth.value == w, case
                            got = np.ctypeslib.as_array(C.cast(wvs.tex_buf[0], C.POINTER(C.c_uint8)), (256 * w * 4,))
                            assert np.array_equal(got, wbuf[:256 * w * 4]), case
                        if r_vss:
                            got = np.ctypeslib.as_array(C.cast(vss.tex_buf[0], C.POINTER(C.c_uint8)), (65536,))
                            assert np.array_equal(got, vbuf) and vss.tex_cs[0] == vcs.value == 1, case
                        n_flips += r_his + r_wvs + r_vss
                        lib.b200_his_destroy(C.byref(his))
                        lib.b200_wvs_destroy(C.byref(wvs))
                        lib.b200_vss_destroy(C.byref(vss))
    assert n_flips > 100


def test_roi_pacing_clamp_and_flags_match_the_reference(shim, ref):
    """The ROI source's host logic against the reference's own src/roi.c (compiled into oracle/_ref with
    cm_tick / cm_render_target replaced by counters): on which ticks the capture core is ticked and
    asked to stage (frame interleave, roi.c:266-277, 523-532), how a requested rectangle is clamped
    (roi_send_range, roi.c:478-500), which planes the registered scopes make it stage (roi.c:533-540)."""
    lib, L = shim, ref.lib
    L.ref_roi_new.restype = C.c_void_p
    L.ref_roi_new.argtypes = [C.c_int]
    L.ref_roi_free.argtypes = [C.c_void_p]
    L.ref_roi_tick.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_roi_target_render.argtypes = [C.c_void_p]
    L.ref_roi_add_consumer.restype = C.c_void_p
    L.ref_roi_add_consumer.argtypes = [C.c_void_p, C.c_uint32]
    L.ref_roi_remove_consumer.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_roi_send_range.argtypes = [C.c_int] * 4 + [C.c_uint32] * 2 + [C.c_void_p]
    Roi = S.RoiSource

    rng = np.random.default_rng(11)

    # --- pacing: random sequences of ticks with 0..2 renders each ---
    for n_interleave in (-1, 0, 1, 2, 3, 5):
        st = L.ref_roi_new(n_interleave)
        roi = Roi()
        lib.b200_roi_init(C.byref(roi), None, 1)
        roi.n_interleave = n_interleave
        cm_obj = S.CmSource()
        cm = C.byref(cm_obj)
        lib.b200_cm_create(cm)

        class _Flag:                                       # struct b200_cm_source.rendered
            value = property(lambda self: cm_obj.rendered, lambda self, v: setattr(cm_obj, "rendered", v))
        flag = _Flag()
        ticks = renders = 0
        for frame in range(60):
            flag.value = True                              # cm_tick resets it (common.c:216-221)
            lib.b200_roi_tick(C.byref(roi), cm)
            ticks += not flag.value
            assert L.ref_roi_tick(st, None) == ticks, (n_interleave, frame)
            for _ in range(int(rng.integers(0, 3))):
                flag.value = False                         # cm_render_target sets it (common.c:225-227)
                got = lib.b200_roi_target_render(C.byref(roi), cm, None, None, 0, 0, 0)
                renders += flag.value
                want = L.ref_roi_target_render(st)
                assert (bool(want >> 16), want & 0xFFFF) == (got, renders), (n_interleave, frame)
        assert ticks > 0 and renders > 0
        lib.b200_cm_destroy(cm)
        lib.b200_roi_destroy(C.byref(roi))
        L.ref_roi_free(st)

    # --- clamping of the requested rectangle ---
    out = (C.c_int * 4)()

    for _ in range(300):
        w, h = int(rng.integers(1, 5000)), int(rng.integers(1, 3000))
        r = [int(v) for v in rng.integers(-50, 5200, 4)]
        if rng.random() < 0.3:
            r[int(rng.integers(0, 4))] = -1
        t = S.CmSource()
        cm = C.byref(t)
        lib.b200_cm_create(cm)
        lib.b200_cm_set_roi(cm, r[0], r[1], r[2], r[3], w, h)
        L.ref_roi_send_range(r[0], r[1], r[2], r[3], w, h, out)
        assert (t.x0, t.y0, t.x1, t.y1) == tuple(out), (r, w, h)
        lib.b200_cm_destroy(cm)

    # --- capture flags from the registered scopes (surface mode = the reference's rule) ---
    def ref_flags(comps_his, comps_wvs, n_vss):
        st = L.ref_roi_new(1)
        cons = [L.ref_roi_add_consumer(st, (1 if c & 0x07 else 0) | (2 if c & 0x70 else 0)) for c in comps_his + comps_wvs]
        cons += [L.ref_roi_add_consumer(st, 2) for _ in range(n_vss)]
        f = C.c_uint32(0)
        L.ref_roi_tick(st, C.byref(f))
        for c in cons:
            L.ref_roi_remove_consumer(st, c)
        L.ref_roi_free(st)
        return f.value

    for comps_his, comps_wvs, n_vss in (([], [], 0), ([0x07], [], 0), ([0x20], [0x07], 0), ([], [0x70], 0), ([], [], 1),
                                        ([0x07], [0x07], 1), ([0x50, 0x07], [0x20], 2), ([0x00], [], 0)):
        for mode in (1, 0):    # SCOPE_MODE_SURFACE, SCOPE_MODE_FUSED
            roi = Roi()
            lib.b200_roi_init(C.byref(roi), None, mode)
            keep = []
            for c in comps_his:
                s = C.create_string_buffer(256)
                lib.b200_his_init(s, None, c)
                lib.b200_roi_register_his(C.byref(roi), s)
                keep.append(s)
            for c in comps_wvs:
                s = C.create_string_buffer(256)
                lib.b200_wvs_init(s, None, c)
                lib.b200_roi_register_wvs(C.byref(roi), s)
                keep.append(s)
            for _ in range(n_vss):
                s = C.create_string_buffer(256)
                lib.b200_vss_init(s, None)
                lib.b200_roi_register_vss(C.byref(roi), s)
                keep.append(s)
            got = lib.b200_roi_capture_flags(C.byref(roi))
            if mode == 1:
                assert got == ref_flags(comps_his, comps_wvs, n_vss), (comps_his, comps_wvs, n_vss)
            else:
                assert got == (8 | 4 | (1 if keep else 0))
            lib.b200_roi_destroy(C.byref(roi))


@pytest.mark.parametrize("flags,roi", [(1, None), (3, None), (2, None), (3, (5, 3, 29, 17)), (1, (0, 0, 40, 9))])
def test_capture_core_matches_the_reference(shim, pkg, flags, roi):
    """b200_cm_* against the reference's own capture core: src/common.c compiled unmodified on a software
    graphics layer (oracle/_ref/libref_common.so; the YUV shader is a stand-in that inverts B, G, R).
    Same schedule on both sides - ticks, renders, a worker that is sometimes held inside the callback -
    and the same things must come out: which renders stage a slot and which are dropped, the queue
    indices after every step, and every surface the callback sees (size, pitch, planes, bytes)."""
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_common.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_common.so not built (needs /root/reference at build time)")
    R, lib = C.CDLL(so), shim
    R.refc_new.restype = C.c_void_p
    R.refc_new.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, CB, C.c_void_p]
    for n in ("refc_free", "refc_tick", "refc_idle"):
        getattr(R, n).argtypes = [C.c_void_p]
    R.refc_render.argtypes = [C.c_void_p, C.c_void_p]
    R.refc_indices.argtypes = [C.c_void_p, C.c_void_p]
    R.refc_set_roi.argtypes = [C.c_void_p] + [C.c_int] * 4
    R.refc_callbacks.restype = C.c_long
    R.refc_callbacks.argtypes = [C.c_void_p]
    W, H = 40, 24
    gate = threading.Event()
    gate.set()
    seen = {"ref": [], "shim": []}
    inside = {"ref": 0, "shim": 0}

    def make_cb(who):
        def cb(_d, sd):
            inside[who] = 1
            gate.wait()
            s = sd.contents
            n = s.linesize * s.height
            planes = tuple(None if not p else bytes(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n,))
                                                    .reshape(s.height, s.linesize)[:, :s.width * 4])
                           for p in (s.rgb_data, s.yuv_data))
            seen[who].append((s.width, s.height, s.colorspace, planes))
            inside[who] = 0
        return CB(cb)

    cb_ref, cb_shim = make_cb("ref"), make_cb("shim")
    ref = R.refc_new(flags, 1, W, H, cb_ref, None)
    cm = _Cm()
    lib.b200_cm_create(C.byref(cm))
    cm.flags, cm.colorspace = flags, 1
    lib.b200_cm_request(C.byref(cm), cb_shim, None)
    if roi:
        R.refc_set_roi(ref, *roi)
        lib.b200_cm_set_roi(C.byref(cm), roi[0], roi[1], roi[2], roi[3], W, H)

    # settled = the worker has nothing left it could do: it waits for work, or - only while the gate is closed -
    # it sits inside the callback (with the gate open, "inside" is a state that is about to change)
    def ref_settled():
        return (inside["ref"] and not gate.is_set()) or (R.refc_idle(ref) and not inside["ref"])

    def shim_settled():
        nxt = (cm.i_read_queue + 1) % 3
        idle = (cm.i_write_queue == nxt or cm.i_staging_queue == nxt) and not cm.worker_busy and not inside["shim"]
        return (inside["shim"] and not gate.is_set()) or idle

    def settle():
        for who in (ref_settled, shim_settled):
            ok = 0
            for _ in range(40000):           # the worker is either waiting for work or inside the callback
                ok = ok + 1 if who() else 0
                if ok >= 3:
                    break
                time.sleep(0.0005)
            assert ok >= 3, "worker did not settle"

    rng = np.random.default_rng(flags * 7 + (1 if roi else 0))
    frames = []
    trace = []
    for step in range(40):
        if step in (12, 26):
            gate.clear()                      # hold the worker inside the next callback for a while
        if step in (19, 33):
            gate.set()
            settle()
        f = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        yuv = f.copy()
        yuv[..., :3] = 255 - f[..., :3]       # what the stand-in shader of the harness writes
        yuv[..., 3] = 255
        frames.append((f, yuv))               # keep alive while staged
        R.refc_tick(ref)
        lib.b200_cm_tick(C.byref(cm))
        n_render = 2 if step % 5 == 0 else 1  # a second render in the same tick is ignored (common.c:225-227)
        for _ in range(n_render):
            a = bool(R.refc_render(ref, f.ctypes.data))
            b = bool(lib.b200_cm_render_target(C.byref(cm), f.ctypes.data, yuv.ctypes.data, W * 4, W, H))
            settle()
            idx = (C.c_int * 3)()
            R.refc_indices(ref, idx)
            trace.append((step, a, b, tuple(idx), (cm.i_write_queue, cm.i_staging_queue, cm.i_read_queue)))
    gate.set()
    settle()
    for step, a, b, ia, ib in trace:
        assert a == b and ia == ib, (step, a, b, ia, ib)
    assert any(not a for _, a, *_ in trace) and sum(a for _, a, *_ in trace) > 15   # drops happened, and stages
    assert len(seen["ref"]) == len(seen["shim"]) > 10
    for i, (x, y) in enumerate(zip(seen["ref"], seen["shim"])):
        assert x[:3] == y[:3], (i, x[:3], y[:3])
        assert x[3] == y[3], f"surface {i}: plane bytes differ"
    want_planes = (bool(flags & 1), bool(flags & 2))
    assert all((p[3][0] is not None, p[3][1] is not None) == want_planes for p in seen["shim"])
    if roi:
        assert seen["shim"][0][:2] == (roi[2] - roi[0], roi[3] - roi[1])
    lib.b200_cm_destroy(C.byref(cm))
    R.refc_free(ref)


@pytest.mark.parametrize("scale", [2, 3, 5])
def test_target_scale_is_staged_point_sampled(shim, pkg, oracle, scale):
    """target_scale without a GPU attached: the staged surface is target size / scale (common.c:249-250), point-sampled
    at the texel centres - the rule oracle/scope_oracle.c pins (orc_point_downsample) - and the ROI rectangle is in
    pixels of that scaled surface (common.c:272-282)."""
    lib = shim
    W, H = 53, 38
    rng = np.random.default_rng(scale)
    rgb = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    yuv = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    small_rgb, small_yuv = oracle.downsample(rgb, scale), oracle.downsample(yuv, scale)
    sw, sh = W // scale, H // scale
    for rect in (None, (1, 2, sw - 1, sh - 1)):
        cm_obj = S.CmSource()
        cm = C.byref(cm_obj)
        lib.b200_cm_create(cm)
        cm_obj.flags, cm_obj.target_scale = 3, scale
        got = []

        def cb(_d, sd):
            s = sd.contents
            n = s.linesize * s.height
            a = np.ctypeslib.as_array(C.cast(s.rgb_data, C.POINTER(C.c_uint8)), (n,)).copy()
            b = np.ctypeslib.as_array(C.cast(s.yuv_data, C.POINTER(C.c_uint8)), (n,)).copy()
            got.append((s.width, s.height, s.linesize, a, b, s.tex))

        cfn = CB(cb)
        lib.b200_cm_request(cm, cfn, None)
        if rect:
            lib.b200_cm_set_roi(cm, *rect, sw, sh)
        for _ in range(2):
            lib.b200_cm_tick(cm)
            lib.b200_cm_render_target(cm, rgb.ctypes.data, yuv.ctypes.data, W * 4, W, H)
            lib.b200_cm_drain(cm)
        lib.b200_cm_destroy(cm)
        w, h, ls, a, b, tex = got[0]
        x0, y0, x1, y1 = rect if rect else (0, 0, sw, sh)
        assert (w, h, ls) == (x1 - x0, y1 - y0, (x1 - x0) * 4) and not tex       # no hint without a GPU
        assert np.array_equal(a.reshape(h, w, 4), small_rgb[y0:y1, x0:x1])
        assert np.array_equal(b.reshape(h, w, 4), small_yuv[y0:y1, x0:x1])
