"""ctypes mirror of include/cm_shim.h (libcm_shim.so): the C host side behind the reference's scope seam -
``cm_surface_cb_t`` callbacks, the ROI fan-out, the capture core's 3-slot queue + worker, and the
``obs_source_info``-shaped tables.  Layouts are checked against the library's own ``sizeof``s at load time."""
from __future__ import annotations

import ctypes as C
import os

from . import _ffi

SHIM_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libcm_shim.so")
_MUTEX, _COND = C.c_byte * 40, C.c_byte * 48          # pthread_mutex_t / pthread_cond_t (x86-64 glibc)

CM_FLAG_CONVERT_RGB, CM_FLAG_CONVERT_YUV, CM_FLAG_RAW_TEXTURE, CM_FLAG_ROI = 1, 2, 4, 8
CM_HINT_MAGIC = 0xB200C0DE


class SurfaceData(C.Structure):        # struct cm_surface_data (src/common.h:24-30)
    _fields_ = [("rgb_data", C.c_void_p), ("yuv_data", C.c_void_p), ("linesize", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("colorspace", C.c_int), ("tex", C.c_void_p)]


class HisSource(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("components", C.c_uint32), ("level_fixed_value", C.c_int),
                ("level_ratio_value", C.c_int), ("logscale", C.c_bool), ("tex_buf", C.c_void_p * 2),
                ("hi_max", (C.c_uint32 * 3) * 2), ("w_tex_buf", C.c_int)]


class WvsSource(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("components", C.c_uint32), ("tex_buf", C.c_void_p * 2),
                ("tex_buf_width", C.c_uint32 * 2), ("w_tex_buf", C.c_int)]


class VssSource(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("tex_buf", C.c_void_p * 2), ("tex_cs", C.c_int * 2),
                ("w_tex_buf", C.c_int)]


class RoiPending(C.Structure):
    _fields_ = [("valid", C.c_bool), ("slot", C.c_int), ("his", C.c_void_p), ("wvs", C.c_void_p), ("vss", C.c_void_p),
                ("width", C.c_uint32), ("colorspace", C.c_int)]


class RoiSource(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("sources_mutex", _MUTEX),
                ("his", C.c_void_p * 8), ("wvs", C.c_void_p * 8), ("vss", C.c_void_p * 8),
                ("n_his", C.c_int), ("n_wvs", C.c_int), ("n_vss", C.c_int), ("wave_tmp", C.c_void_p),
                ("wave_tmp_width", C.c_uint32), ("n_interleave", C.c_int), ("i_interleave", C.c_int),
                ("interleave_rendered", C.c_bool), ("pending", RoiPending), ("frames_filed", C.c_ulong)]


class CmHint(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("slot", C.c_int), ("target_scale", C.c_uint32)]


class CmQueueItem(C.Structure):
    _fields_ = [("staged", C.c_void_p), ("staged_bytes", C.c_size_t), ("rgb", C.c_void_p), ("yuv", C.c_void_p),
                ("width", C.c_uint32), ("height", C.c_uint32), ("linesize", C.c_uint32), ("flags", C.c_uint32),
                ("colorspace", C.c_int), ("cb", C.c_void_p), ("cb_data", C.c_void_p)]


class CmSource(C.Structure):
    _fields_ = [("queue", CmQueueItem * 3), ("i_write_queue", C.c_int), ("i_staging_queue", C.c_int),
                ("i_read_queue", C.c_int), ("rendered", C.c_bool), ("pipeline_thread", C.c_ulong),
                ("pipeline_mutex", _MUTEX), ("pipeline_cond", _COND),
                ("pipeline_thread_running", C.c_bool), ("request_exit", C.c_bool), ("worker_busy", C.c_bool),
                ("callback", C.c_void_p), ("callback_data", C.c_void_p), ("flags", C.c_uint32), ("colorspace", C.c_int),
                ("frames_dropped", C.c_ulong), ("frames_processed", C.c_ulong),
                ("x0", C.c_int), ("x1", C.c_int), ("y0", C.c_int), ("y1", C.c_int),
                ("target_scale", C.c_int), ("gpu", C.c_void_p), ("zero_copy", C.c_bool), ("hints", CmHint * 3)]


class Settings(C.Structure):           # struct b200_settings: the obs_data keys the path reads
    _fields_ = [("ctx", C.c_void_p), ("mode", C.c_uint32), ("target_scale", C.c_int), ("colorspace", C.c_int),
                ("components", C.c_uint32), ("intensity", C.c_int), ("level_mode", C.c_int),
                ("level_fixed_value", C.c_int), ("level_ratio_value", C.c_double), ("logscale", C.c_bool),
                ("gpu_ring", C.c_bool), ("zero_copy", C.c_bool)]


GET_FRAME = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32),
                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))


class Target(C.Structure):             # struct b200_target: the source whose frame video_render captures
    _fields_ = [("get_frame", GET_FRAME), ("opaque", C.c_void_p)]


class SourceInfo(C.Structure):         # struct b200_source_info: the obs_source_info members the scopes fill in
    _fields_ = [("id", C.c_char_p), ("type", C.c_int), ("output_flags", C.c_uint32),
                ("get_name", C.CFUNCTYPE(C.c_char_p, C.c_void_p)),
                ("create", C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_void_p)),
                ("destroy", C.CFUNCTYPE(None, C.c_void_p)),
                ("update", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)),
                ("get_defaults", C.CFUNCTYPE(None, C.c_void_p)),
                ("get_properties", C.c_void_p),
                ("get_width", C.CFUNCTYPE(C.c_uint32, C.c_void_p)),
                ("get_height", C.CFUNCTYPE(C.c_uint32, C.c_void_p)),
                ("enum_active_sources", C.c_void_p),
                ("video_render", C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)),
                ("video_tick", C.CFUNCTYPE(None, C.c_void_p, C.c_float))]


SURFACE_CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(SurfaceData))   # cm_surface_cb_t (src/common.h:32)
_STRUCTS = [HisSource, WvsSource, VssSource, RoiSource, CmQueueItem, CmSource, Settings, SourceInfo, CmHint, SurfaceData,
            Target]
_lib = None


def load() -> C.CDLL:
    """libcm_shim.so with argument types set; raises if a mirrored layout disagrees with the library's sizeof."""
    global _lib
    if _lib is not None:
        return _lib
    _ffi.load()                                      # libscope_b200.so first (the shim links against it)
    L = C.CDLL(SHIM_PATH)
    L.b200_sizeof_struct.argtypes = [C.c_int]
    L.b200_sizeof_struct.restype = C.c_size_t
    for i, st in enumerate(_STRUCTS):
        assert L.b200_sizeof_struct(i) == C.sizeof(st), (st.__name__, L.b200_sizeof_struct(i), C.sizeof(st))
    vp, u32 = C.c_void_p, C.c_uint32
    for name, args, res in [
        ("b200_his_init", [vp, vp, u32], None), ("b200_his_destroy", [vp], None),
        ("b200_wvs_init", [vp, vp, u32], None), ("b200_wvs_destroy", [vp], None),
        ("b200_vss_init", [vp, vp], None), ("b200_vss_destroy", [vp], None),
        ("b200_his_surface_cb", [vp, vp], None), ("b200_wvs_surface_cb", [vp, vp], None),
        ("b200_vss_surface_cb", [vp, vp], None), ("b200_roi_surface_cb", [vp, vp], None),
        ("b200_his_inputs_missing", [vp, vp], C.c_bool), ("b200_wvs_inputs_missing", [vp, vp], C.c_bool),
        ("b200_vss_inputs_missing", [vp, vp], C.c_bool),
        ("b200_roi_init", [vp, vp, u32], None), ("b200_roi_destroy", [vp], None), ("b200_roi_finish", [vp], None),
        ("b200_roi_capture_flags", [vp], u32),
        ("b200_roi_register_his", [vp, vp], C.c_int), ("b200_roi_register_wvs", [vp, vp], C.c_int),
        ("b200_roi_register_vss", [vp, vp], C.c_int),
        ("b200_cm_create", [vp], None), ("b200_cm_destroy", [vp], None), ("b200_cm_request", [vp, vp, vp], None),
        ("b200_cm_attach_gpu", [vp, vp, C.c_bool], None), ("b200_cm_tick", [vp], None),
        ("b200_cm_tick_obs", [vp, C.c_float], None),
        ("b200_cm_render_target", [vp, vp, vp, u32, u32, u32], C.c_bool),
        ("b200_cm_set_roi", [vp, C.c_int, C.c_int, C.c_int, C.c_int, u32, u32], None),
        ("b200_roi_tick", [vp, vp], None), ("b200_roi_target_render", [vp, vp, vp, vp, u32, u32, u32], C.c_bool),
        ("b200_cm_drain", [vp], None), ("b200_source_drain", [vp], None),
        ("b200_source_result", [vp, C.POINTER(u32), C.POINTER(C.POINTER(u32))], vp),
    ]:
        fn = getattr(L, name)
        fn.argtypes, fn.restype = args, res
    _lib = L
    return L


def source_info(lib: C.CDLL, name: str) -> SourceInfo:
    """the exported table ``b200_colormonitor_<name>`` (histogram / waveform / vectorscope)"""
    return SourceInfo.in_dll(lib, f"b200_colormonitor_{name}")
