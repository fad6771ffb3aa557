"""ctypes bindings for the CPU checker (TEST INFRASTRUCTURE ONLY).

Two libraries, both built by ``oracle/Makefile``:

* ``oracle/liboracle.so``  — ``Oracle``: the plain-C restatement of the reference
  loops (``oracle/scope_oracle.c``; each function there cites the reference
  file:line it follows).
* ``oracle/_ref/libref.so`` — ``Ref``: the reference's own
  ``src/{histogram,waveform,vectorscope}.c`` compiled unmodified against the mock
  libobs headers.  Exists when it was built in the container that has
  ``/root/reference``; it travels to the GPU box as a prebuilt file.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)


class _Surface(C.Structure):
    # mirror of struct orc_surface  (== cm_surface_data minus the texture, common.h:24-30)
    _fields_ = [
        ("rgb_data", C.c_void_p),
        ("yuv_data", C.c_void_p),
        ("linesize", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("colorspace", C.c_int),
    ]


def _ptr(a: Optional[np.ndarray]) -> Optional[int]:
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _plane_geometry(plane: np.ndarray, width: Optional[int], height: Optional[int]) -> Tuple[int, int, int]:
    """plane is (H, linesize) u8 or (H, W, 4) u8; returns (linesize, width, height)."""
    if plane.ndim == 3:
        h, w, c = plane.shape
        assert c == 4
        return w * 4, w if width is None else width, h if height is None else height
    h, ls = plane.shape
    assert width is not None
    return ls, width, h if height is None else height


def _surface(rgb, yuv, width, height, colorspace):
    ref = rgb if rgb is not None else yuv
    if ref is None:
        ls, w, h = 0, width or 0, height or 0
    else:
        ls, w, h = _plane_geometry(ref, width, height)
    s = _Surface(_ptr(rgb), _ptr(yuv), ls, w, h, colorspace)
    return s, ls, w, h


class Oracle:
    """The restated algorithm (oracle/scope_oracle.c)."""

    def __init__(self, path: str = ORACLE_SO):
        self.lib = L = C.CDLL(path)
        L.orc_rgb_to_yuv.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32]
        L.orc_rgb_to_yuv.restype = None
        L.orc_rgb_to_yuv_table.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_rgb_to_yuv_table.restype = C.c_int
        L.orc_calc_colorspace.argtypes = [C.c_int]
        L.orc_calc_colorspace.restype = C.c_int
        L.orc_histogram_counts.argtypes = [C.c_uint32, C.POINTER(_Surface), C.c_void_p]
        L.orc_histogram_counts.restype = None
        L.orc_histogram_post.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_histogram_post.restype = None
        L.orc_draw_histogram.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(_Surface), C.c_void_p,
                                         C.c_void_p]
        L.orc_draw_histogram.restype = None
        L.orc_waveform.argtypes = [C.c_uint32, C.POINTER(_Surface), C.c_void_p]
        L.orc_waveform.restype = None
        L.orc_vectorscope.argtypes = [C.POINTER(_Surface), C.c_void_p]
        L.orc_vectorscope.restype = None
        L.orc_apply_intensity.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_apply_intensity.restype = None

    # -- transform -------------------------------------------------------
    def rgb_to_yuv(self, bgra: np.ndarray, colorspace: int = 2, width: Optional[int] = None) -> np.ndarray:
        ls, w, h = _plane_geometry(bgra, width, None)
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.orc_rgb_to_yuv(_ptr(bgra), ls, w, h, colorspace, out.ctypes.data, w * 4)
        return out

    def downsample(self, plane: np.ndarray, scale: int, width: Optional[int] = None) -> np.ndarray:
        """target_scale: (H, W, 4) -> (H // scale, W // scale, 4), point-sampled at the texel centres
        (orc_point_downsample; src/common.c:249-250)"""
        ls, w, h = _plane_geometry(plane, width, None)
        s = max(1, int(scale))
        out = np.zeros((h // s, w // s, 4), np.uint8)
        self.lib.orc_point_downsample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                  C.c_void_p, C.c_uint32]
        self.lib.orc_point_downsample.restype = None
        self.lib.orc_point_downsample(_ptr(plane), ls, w, h, s, out.ctypes.data, (w // s) * 4)
        return out

    def yuv_from_table(self, bgra: np.ndarray, table: np.ndarray) -> np.ndarray:
        """[U, Y, V, 255] plane of a frame through one of rgb_to_yuv_table's 2^24-entry tables"""
        idx = (bgra[..., 2].astype(np.uint32) << 16) | (bgra[..., 1].astype(np.uint32) << 8) | bgra[..., 0]
        t = table[idx]
        out = np.empty(bgra.shape, np.uint8)
        out[..., 0], out[..., 1], out[..., 2], out[..., 3] = t & 0xFF, (t >> 8) & 0xFF, (t >> 16) & 0xFF, 255
        return out

    TRANSFORM_VARIANTS = {"exact": 0, "fp32_strict": 1, "fp32_contracted": 2}

    def rgb_to_yuv_table(self, colorspace: int, variant: str = "exact") -> Tuple[np.ndarray, bool]:
        """All 2^24 colours: out[r<<16|g<<8|b] = u | y<<8 | v<<16; second value = some value left
        the UNORM range.  variant: "exact" (the pinned definition) or one of the two fp32
        evaluations kept for comparison."""
        out = np.zeros(1 << 24, np.uint32)
        clamp = self.lib.orc_rgb_to_yuv_table(colorspace, self.TRANSFORM_VARIANTS[variant], out.ctypes.data)
        return out, bool(clamp)

    def calc_colorspace(self, cs: int) -> int:
        return self.lib.orc_calc_colorspace(cs)

    # -- scopes ----------------------------------------------------------
    def histogram_counts(self, components, rgb=None, yuv=None, width=None, height=None, colorspace=2):
        s, _, _, _ = _surface(rgb, yuv, width, height, colorspace)
        out = np.zeros(1024, np.uint32)
        self.lib.orc_histogram_counts(components, C.byref(s), out.ctypes.data)
        return out

    def histogram_post(self, components, width, height, counts, level_fixed=0, level_ratio=0, logscale=False):
        counts = np.ascontiguousarray(counts, np.uint32)
        out = np.zeros(1024, np.float32)
        hi = np.zeros(3, np.uint32)
        self.lib.orc_histogram_post(components, width, height, level_fixed, level_ratio, int(logscale),
                                    counts.ctypes.data, out.ctypes.data, hi.ctypes.data)
        return out, hi

    def draw_histogram(self, components, rgb=None, yuv=None, width=None, height=None, colorspace=2,
                       level_fixed=0, level_ratio=0, logscale=False, hi_init=(0, 0, 0)):
        """his_draw_histogram as a whole: (float32[1024], hi_max u32[3]); hi_max starts as ``hi_init`` and is
        left alone when no plane is selected (histogram.c:366-373)."""
        s, _, _, _ = _surface(rgb, yuv, width, height, colorspace)
        out = np.zeros(1024, np.float32)
        hi = np.array(hi_init, np.uint32)
        self.lib.orc_draw_histogram(components, level_fixed, level_ratio, int(logscale), C.byref(s),
                                    out.ctypes.data, hi.ctypes.data)
        return out, hi

    def waveform(self, components, rgb=None, yuv=None, width=None, height=None, colorspace=2):
        s, _, w, _ = _surface(rgb, yuv, width, height, colorspace)
        out = np.zeros((256, w, 4), np.uint8)
        self.lib.orc_waveform(components, C.byref(s), out.ctypes.data)
        return out

    def vectorscope(self, yuv, width=None, height=None, colorspace=2):
        s, _, _, _ = _surface(None, yuv, width, height, colorspace)
        out = np.zeros((256, 256), np.uint8)
        self.lib.orc_vectorscope(C.byref(s), out.ctypes.data)
        return out

    def apply_intensity(self, bins: np.ndarray, intensity: int) -> np.ndarray:
        b = np.ascontiguousarray(bins, np.uint8)
        out = np.zeros_like(b)
        self.lib.orc_apply_intensity(b.ctypes.data, b.size, intensity, out.ctypes.data)
        return out

    def fused(self, bgra, hist_components=0x07, wave_components=0x07, colorspace=2, width=None):
        """Fused-mode oracle: RGB plane in, YUV plane by the pinned transform,
        then the three reference loops.  Returns (hist u32[1024], wave u8[256,W,4],
        vscope u8[256,256])."""
        yuv = self.rgb_to_yuv(bgra, colorspace, width)
        if bgra.ndim == 2:
            h = bgra.shape[0]
            rgb = np.ascontiguousarray(bgra[:, : width * 4].reshape(h, width, 4))
        else:
            rgb = bgra
        return (self.histogram_counts(hist_components, rgb, yuv, colorspace=colorspace),
                self.waveform(wave_components, rgb, yuv, colorspace=colorspace),
                self.vectorscope(yuv, colorspace=colorspace))


class Ref:
    """The reference's own loops (oracle/_ref/libref.so)."""

    def __init__(self, path: str = REF_SO):
        self.lib = L = C.CDLL(path)
        L.ref_his_draw_histogram.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_his_draw_histogram.restype = C.c_int
        L.ref_wvs_draw_waveform.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_int, C.c_void_p]
        L.ref_wvs_draw_waveform.restype = C.c_int
        L.ref_vss_draw_vectorscope.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.ref_vss_draw_vectorscope.restype = C.c_int
        L.ref_his_new.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.ref_his_new.restype = C.c_void_p
        L.ref_his_free.argtypes = [C.c_void_p]
        L.ref_his_surface_cb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_int, C.c_void_p, C.c_void_p]
        L.ref_his_surface_cb.restype = C.c_int
        L.ref_wvs_new.argtypes = [C.c_uint32]
        L.ref_wvs_new.restype = C.c_void_p
        L.ref_wvs_free.argtypes = [C.c_void_p]
        L.ref_wvs_surface_cb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_int, C.c_void_p, C.c_void_p]
        L.ref_wvs_surface_cb.restype = C.c_int
        L.ref_vss_new.argtypes = []
        L.ref_vss_new.restype = C.c_void_p
        L.ref_vss_free.argtypes = [C.c_void_p]
        L.ref_vss_surface_cb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_int, C.c_void_p, C.c_void_p]
        L.ref_vss_surface_cb.restype = C.c_int

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def histogram(self, components, rgb=None, yuv=None, width=None, height=None, colorspace=2,
                  level_fixed=0, level_ratio=0, logscale=False, hi_init=(0, 0, 0)):
        """Returns (float32[1024] as the reference leaves tex_buf, hi_max u32[3] starting as hi_init)."""
        _, ls, w, h = _surface(rgb, yuv, width, height, colorspace)
        out = np.zeros(1024, np.float32)
        hi = np.array(hi_init, np.uint32)
        self.lib.ref_his_draw_histogram(components, level_fixed, level_ratio, int(logscale), _ptr(rgb), _ptr(yuv),
                                        ls, w, h, colorspace, out.ctypes.data, hi.ctypes.data)
        return out, hi

    def waveform(self, components, rgb=None, yuv=None, width=None, height=None, colorspace=2):
        _, ls, w, h = _surface(rgb, yuv, width, height, colorspace)
        out = np.zeros((256, w, 4), np.uint8)
        self.lib.ref_wvs_draw_waveform(components, _ptr(rgb), _ptr(yuv), ls, w, h, colorspace, out.ctypes.data)
        return out

    def vectorscope(self, yuv, width=None, height=None, colorspace=2):
        ls, w, h = _plane_geometry(yuv, width, height)
        out = np.zeros((256, 256), np.uint8)
        self.lib.ref_vss_draw_vectorscope(_ptr(yuv), ls, w, h, colorspace, out.ctypes.data)
        return out
