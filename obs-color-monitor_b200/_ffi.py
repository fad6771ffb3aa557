"""ctypes binding of ``libscope_b200.so`` (the C-ABI in ``include/scope_ffi.h``).

Everything here goes through the ``extern "C"`` entry points a C host would call;
there are no torch types in the signatures (pointers are passed as integers) and
there is no CPU fallback: if the library or a GPU is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCOPE_LIB") or os.path.join(_HERE, "lib", "libscope_b200.so")  # SCOPE_LIB: A/B builds
SHIM_PATH = os.path.join(_HERE, "lib", "libcm_shim.so")

SCOPE_OK = 0
SCOPE_ERR_INVALID = 1
SCOPE_ERR_NO_DEVICE = 2
SCOPE_ERR_CUDA = 3
SCOPE_ERR_UNSUPPORTED = 4
SCOPE_ERR_NOMEM = 5
SCOPE_ERR_BUSY = 6

SCOPE_HIST, SCOPE_WAVE, SCOPE_VSCOPE, SCOPE_ALL = 1, 2, 4, 7
COMP_RGB, COMP_Y, COMP_UV, COMP_YUV = 0x07, 0x20, 0x50, 0x70
MODE_FUSED, MODE_SURFACE = 0, 1
RING_SLOTS = 3
MAX_PEERS = 16
BAND_EXCLUSIVE = 1

# every symbol include/scope_ffi.h declares (tests check the library exports them all)
EXPORTED_SYMBOLS = [
    "scope_abi_version", "scope_ctx_create", "scope_ctx_destroy", "scope_last_error",
    "scope_launch_count", "scope_sm_count", "scope_accumulate_host", "scope_submit_host",
    "scope_wait_host", "scope_ring_input", "scope_accumulate_device", "scope_accumulate_partial", "scope_accumulate_band",
    "scope_finalize_partial", "scope_finalize_peers", "scope_finalize_multicast", "scope_host_alloc", "scope_host_free", "scope_debug_yuv_table", "scope_debug_uv_table_v3", "scope_debug_yuv_table_strict",
    "scope_wave_bytes", "scope_partial_wave_words", "scope_profile_enable", "scope_profile_read",
]


class Surface(C.Structure):
    """struct scope_surface (mirror of cm_surface_data, src/common.h:24-30)."""
    _fields_ = [
        ("rgb_data", C.c_void_p),
        ("yuv_data", C.c_void_p),
        ("linesize", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("colorspace", C.c_int32),
    ]


class Params(C.Structure):
    """struct scope_params."""
    _fields_ = [
        ("scopes", C.c_uint32),
        ("mode", C.c_uint32),
        ("hist_components", C.c_uint32),
        ("wave_components", C.c_uint32),
        ("level_fixed_value", C.c_int32),
        ("level_ratio_value", C.c_int32),
        ("logscale", C.c_int32),
        ("wave_intensity", C.c_int32),
        ("vscope_intensity", C.c_int32),
        ("target_scale", C.c_uint32),
        ("xform", C.c_uint32),
        ("reserved", C.c_uint32 * 1),
    ]


class OutHost(C.Structure):
    _fields_ = [
        ("hist_counts", C.c_void_p),
        ("hist_float", C.c_void_p),
        ("hist_max", C.c_void_p),
        ("wave", C.c_void_p),
        ("vscope", C.c_void_p),
        ("wave_display", C.c_void_p),
        ("vscope_display", C.c_void_p),
    ]


class OutDevice(C.Structure):
    _fields_ = [
        ("hist_counts", C.c_void_p),
        ("hist_max", C.c_void_p),
        ("wave", C.c_void_p),
        ("vscope", C.c_void_p),
        ("vscope_display", C.c_void_p),
        ("wave_display", C.c_void_p),
    ]


class PartialDevice(C.Structure):
    _fields_ = [
        ("hist_counts", C.c_void_p),
        ("wave_pairs", C.c_void_p),
        ("vscope_counts", C.c_void_p),
    ]


class ScopeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libscope_b200 error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libscope_b200.so (raises if it has not been built; never substitutes
    another implementation)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make product` (or __graft_entry__.build()). "
            "There is no CPU fallback for the scope kernels.")
    L = C.CDLL(LIB_PATH)
    L.scope_abi_version.restype = C.c_int
    L.scope_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.scope_ctx_create.restype = C.c_int
    L.scope_ctx_destroy.argtypes = [C.c_void_p]
    L.scope_ctx_destroy.restype = None
    L.scope_last_error.argtypes = [C.c_void_p]
    L.scope_last_error.restype = C.c_char_p
    L.scope_launch_count.argtypes = [C.c_void_p]
    L.scope_launch_count.restype = C.c_uint64
    L.scope_sm_count.argtypes = [C.c_void_p]
    L.scope_sm_count.restype = C.c_int
    L.scope_accumulate_host.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Surface), C.POINTER(OutHost)]
    L.scope_accumulate_host.restype = C.c_int
    L.scope_submit_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(Params), C.POINTER(Surface)]
    L.scope_submit_host.restype = C.c_int
    L.scope_wait_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(OutHost)]
    L.scope_wait_host.restype = C.c_int
    L.scope_ring_input.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    L.scope_ring_input.restype = C.c_int
    L.scope_accumulate_device.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Surface), C.c_uint32,
                                          C.c_size_t, C.POINTER(OutDevice), C.c_void_p]
    L.scope_accumulate_device.restype = C.c_int
    L.scope_accumulate_partial.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Surface), C.c_uint32,
                                           C.c_uint32, C.POINTER(PartialDevice), C.c_void_p]
    L.scope_accumulate_partial.restype = C.c_int
    L.scope_accumulate_band.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Surface), C.c_uint32, C.c_uint32,
                                        C.POINTER(PartialDevice), C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32,
                                        C.c_void_p]
    L.scope_accumulate_band.restype = C.c_int
    L.scope_finalize_partial.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.c_uint32,
                                         C.POINTER(PartialDevice), C.POINTER(OutDevice), C.c_void_p]
    L.scope_finalize_partial.restype = C.c_int
    L.scope_finalize_peers.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.c_uint32,
                                       C.POINTER(PartialDevice), C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.POINTER(OutDevice), C.c_uint32, C.c_void_p]
    L.scope_finalize_peers.restype = C.c_int
    L.scope_finalize_multicast.argtypes = [C.c_void_p, C.POINTER(Params), C.c_uint32, C.c_uint32,
                                           C.POINTER(PartialDevice), C.c_uint32, C.c_uint32,
                                           C.POINTER(OutDevice), C.POINTER(OutDevice), C.c_void_p]
    L.scope_finalize_multicast.restype = C.c_int
    L.scope_host_alloc.argtypes = [C.c_size_t]
    L.scope_host_alloc.restype = C.c_void_p
    L.scope_host_free.argtypes = [C.c_void_p]
    L.scope_host_free.restype = None
    L.scope_debug_yuv_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.scope_debug_yuv_table.restype = C.c_int
    L.scope_debug_uv_table_v3.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.scope_debug_uv_table_v3.restype = C.c_int
    L.scope_debug_yuv_table_strict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.scope_debug_yuv_table_strict.restype = C.c_int
    L.scope_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.scope_profile_enable.restype = C.c_int
    L.scope_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
    L.scope_profile_read.restype = C.c_int
    L.scope_wave_bytes.argtypes = [C.c_uint32]
    L.scope_wave_bytes.restype = C.c_size_t
    L.scope_partial_wave_words.argtypes = [C.c_uint32]
    L.scope_partial_wave_words.restype = C.c_size_t
    _lib = L
    return L


class Context:
    """RAII wrapper of ``scope_ctx`` (one per worker thread / GPU, like one worker per
    ``cm_source`` in the reference, src/common.c:375-403)."""

    def __init__(self, device: int = -1):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.scope_ctx_create(device, C.byref(h))
        if rc != SCOPE_OK:
            msg = self.lib.scope_last_error(None)
            raise ScopeError(rc, msg.decode() if msg else "scope_ctx_create failed")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.scope_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != SCOPE_OK:
            msg = self.lib.scope_last_error(self.handle)
            raise ScopeError(rc, msg.decode() if msg else "?")

    @property
    def launch_count(self) -> int:
        return int(self.lib.scope_launch_count(self.handle))

    def profile_enable(self, on: bool = True):
        self.check(self.lib.scope_profile_enable(self.handle, int(on)))

    def profile_read(self, max_entries: int = 4096):
        buf = (C.c_float * max_entries)()
        n = self.lib.scope_profile_read(self.handle, buf, max_entries)
        return [float(buf[i]) for i in range(max(n, 0))]

    @property
    def sm_count(self) -> int:
        return int(self.lib.scope_sm_count(self.handle))
