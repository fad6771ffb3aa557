"""Host-side mirror of the reference's scope interface, on top of the C-ABI.

The reference exposes, per scope, a callback ``*_surface_cb(data, cm_surface_data*)``
(src/histogram.c:432-450, src/waveform.c:272-289, src/vectorscope.c:248-265) whose
only inputs are the mapped surface planes and the source's ``components`` /
``colorspace`` settings.  ``ScopeEngine`` keeps those names and meanings:

* ``ScopeSettings.hist_components`` / ``wave_components`` = ``his_source.components`` /
  ``wvs_source.components`` (0x07 RGB, 0x20 luma, 0x50 chroma, 0x70 YUV)
* ``colorspace`` 1 = BT.601, 2 = BT.709, anything else -> 709 (src/util.c:25-41)
* outputs in the reference's exact buffer layouts (see include/scope_ffi.h)
* a plane the request needs but is ``None`` raises (the reference returns early and keeps
  the previous result; the C shim in csrc/cm_shim.c reproduces that no-flip behaviour)

PyTorch appears only as the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _ffi
from ._ffi import (COMP_RGB, MODE_FUSED, MODE_SURFACE, SCOPE_ALL, SCOPE_HIST, SCOPE_VSCOPE, SCOPE_WAVE, Context,
                   OutDevice, OutHost, Params, PartialDevice, Surface)


@dataclass
class ScopeSettings:
    scopes: int = SCOPE_ALL
    mode: int = MODE_FUSED
    hist_components: int = COMP_RGB   # default of histogram.c:167
    wave_components: int = COMP_RGB   # default of waveform.c:115
    colorspace: int = 2
    level_fixed_value: int = 0
    level_ratio_value: int = 0
    logscale: bool = False
    wave_intensity: int = 0           # 0 = no display image; reference default 51 (waveform.c:114)
    vscope_intensity: int = 0         # reference default 25 (vectorscope.c:158)
    target_scale: int = 1             # point-downsample first (common.c:88-90,249-250); the reference's default is 2
    xform: int = 0                    # 0 = exact transform, 1 = SCOPE_XFORM_FP32_STRICT

    def to_c(self) -> Params:
        p = Params()
        p.scopes = self.scopes
        p.mode = self.mode
        p.hist_components = self.hist_components
        p.wave_components = self.wave_components
        p.level_fixed_value = self.level_fixed_value
        p.level_ratio_value = self.level_ratio_value
        p.logscale = int(self.logscale)
        p.wave_intensity = self.wave_intensity
        p.vscope_intensity = self.vscope_intensity
        p.target_scale = self.target_scale
        p.xform = self.xform
        return p


def _plane_info(plane: np.ndarray, width: Optional[int]):
    assert plane.dtype == np.uint8 and plane.flags["C_CONTIGUOUS"]
    if plane.ndim == 3:
        h, w, c = plane.shape
        assert c == 4
        return w * 4, (w if width is None else width), h
    h, ls = plane.shape
    assert width is not None, "pitched (H, linesize) planes need an explicit width"
    return ls, width, h


class ScopeEngine:
    """One ``scope_ctx`` = one worker (the reference runs one "color-monitor" worker
    per ``cm_source``, src/common.c:375-403)."""

    def __init__(self, device: int = -1):
        self.ctx = Context(device)
        self.lib = self.ctx.lib

    def close(self):
        self.ctx.close()

    @property
    def launch_count(self) -> int:
        return self.ctx.launch_count

    # ------------------------------------------------------------------
    # host buffers (numpy) -> host results: the surface-callback drop-in
    # ------------------------------------------------------------------
    def _host_out(self, st: ScopeSettings, width: int):
        res: Dict[str, np.ndarray] = {}
        out = OutHost()
        if st.scopes & SCOPE_HIST:
            res["hist"] = np.zeros(1024, np.uint32)
            res["hist_float"] = np.zeros(1024, np.float32)
            res["hist_max"] = np.zeros(3, np.uint32)
            out.hist_counts = res["hist"].ctypes.data
            out.hist_float = res["hist_float"].ctypes.data
            out.hist_max = res["hist_max"].ctypes.data
        if st.scopes & SCOPE_WAVE:
            res["wave"] = np.zeros((256, width, 4), np.uint8)
            out.wave = res["wave"].ctypes.data
            if st.wave_intensity > 0:
                res["wave_display"] = np.zeros((256, width, 4), np.uint8)
                out.wave_display = res["wave_display"].ctypes.data
        if st.scopes & SCOPE_VSCOPE:
            res["vscope"] = np.zeros((256, 256), np.uint8)
            out.vscope = res["vscope"].ctypes.data
            if st.vscope_intensity > 0:
                res["vscope_display"] = np.zeros((256, 256), np.uint8)
                out.vscope_display = res["vscope_display"].ctypes.data
        return res, out

    def _host_surface(self, rgb, yuv, width, st: ScopeSettings) -> Surface:
        ref = rgb if rgb is not None else yuv
        if ref is None:
            raise ValueError("no plane given")
        ls, w, h = _plane_info(ref, width)
        s = Surface()
        s.rgb_data = rgb.ctypes.data if rgb is not None else None
        s.yuv_data = yuv.ctypes.data if yuv is not None else None
        s.linesize, s.width, s.height, s.colorspace = ls, w, h, st.colorspace
        return s

    def accumulate_host(self, rgb: Optional[np.ndarray], yuv: Optional[np.ndarray] = None, *,
                        settings: Optional[ScopeSettings] = None, width: Optional[int] = None):
        """Synchronous: one mapped surface in, the scopes' buffers out (numpy)."""
        st = settings or ScopeSettings()
        s = self._host_surface(rgb, yuv, width, st)
        res, out = self._host_out(st, s.width // max(1, st.target_scale))
        p = st.to_c()
        self.ctx.check(self.lib.scope_accumulate_host(self.ctx.handle, C.byref(p), C.byref(s), C.byref(out)))
        return res

    def submit_host(self, slot: int, rgb, yuv=None, *, settings=None, width=None) -> bool:
        """Enqueue into ring slot ``slot`` (0..2).  Returns False when the slot is still in
        flight (the reference drops the frame in that case, src/common.c:260-268)."""
        st = settings or ScopeSettings()
        s = self._host_surface(rgb, yuv, width, st)
        p = st.to_c()
        rc = self.lib.scope_submit_host(self.ctx.handle, slot, C.byref(p), C.byref(s))
        if rc == _ffi.SCOPE_ERR_BUSY:
            return False
        self.ctx.check(rc)
        self._pending = getattr(self, "_pending", {})
        self._pending[slot] = (st, s.width // max(1, st.target_scale))
        return True

    def wait_host(self, slot: int):
        st, width = self._pending.pop(slot)
        res, out = self._host_out(st, width)
        self.ctx.check(self.lib.scope_wait_host(self.ctx.handle, slot, C.byref(out)))
        return res

    # ------------------------------------------------------------------
    # device buffers (torch tensors own the memory), batched
    # ------------------------------------------------------------------
    def alloc_device_out(self, n_frames: int, width: int, st: ScopeSettings, device=None):
        import torch

        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        res = {}
        if st.scopes & SCOPE_HIST:
            res["hist"] = torch.empty((n_frames, 1024), dtype=torch.int32, device=dev)
            res["hist_max"] = torch.empty((n_frames, 4), dtype=torch.int32, device=dev)
        if st.scopes & SCOPE_WAVE:
            res["wave"] = torch.empty((n_frames, 256, width, 4), dtype=torch.uint8, device=dev)
            if st.wave_intensity > 0:
                res["wave_display"] = torch.empty((n_frames, 256, width, 4), dtype=torch.uint8, device=dev)
        if st.scopes & SCOPE_VSCOPE:
            res["vscope"] = torch.empty((n_frames, 256, 256), dtype=torch.uint8, device=dev)
            if st.vscope_intensity > 0:
                res["vscope_display"] = torch.empty((n_frames, 256, 256), dtype=torch.uint8, device=dev)
        return res

    def accumulate_device(self, rgb, yuv=None, *, settings: Optional[ScopeSettings] = None, out=None,
                          width: Optional[int] = None, stream: Optional[int] = None):
        """``rgb`` / ``yuv``: CUDA uint8 tensors shaped (N, H, W, 4) or (N, H, linesize);
        frames are ``tensor.stride(0)`` bytes apart.  Asynchronous on ``stream`` (a raw
        cudaStream_t handle; default: torch's current stream).  Returns device tensors."""
        import torch

        st = settings or ScopeSettings()
        ref = rgb if rgb is not None else yuv
        assert ref.is_cuda and ref.dtype == torch.uint8
        if ref.dim() == 4:
            n, h, w, c = ref.shape
            assert c == 4 and ref.stride(3) == 1 and ref.stride(2) == 4
            width = w if width is None else width
        else:
            n, h, _ = ref.shape
            assert ref.stride(2) == 1 and width is not None
        linesize = ref.stride(1)
        frame_stride = ref.stride(0) if n > 1 else linesize * h
        if rgb is not None and yuv is not None:
            assert rgb.stride() == yuv.stride() and rgb.shape == yuv.shape
        s = Surface()
        s.rgb_data = rgb.data_ptr() if rgb is not None else None
        s.yuv_data = yuv.data_ptr() if yuv is not None else None
        s.linesize, s.width, s.height, s.colorspace = linesize, width, h, st.colorspace
        if out is None:
            out = self.alloc_device_out(n, width // max(1, st.target_scale), st, ref.device)
        od = OutDevice()
        od.hist_counts = out["hist"].data_ptr() if "hist" in out else None
        od.hist_max = out["hist_max"].data_ptr() if "hist_max" in out else None
        od.wave = out["wave"].data_ptr() if "wave" in out else None
        od.vscope = out["vscope"].data_ptr() if "vscope" in out else None
        od.wave_display = out["wave_display"].data_ptr() if "wave_display" in out else None
        od.vscope_display = out["vscope_display"].data_ptr() if "vscope_display" in out else None
        if stream is None:
            stream = torch.cuda.current_stream(ref.device).cuda_stream
        p = st.to_c()
        self.ctx.check(self.lib.scope_accumulate_device(self.ctx.handle, C.byref(p), C.byref(s), n, frame_stride,
                                                        C.byref(od), C.c_void_p(stream)))
        return out

    # ------------------------------------------------------------------
    # tile-sharded frames: partial accumulators -> (all-reduce) -> finalize
    # ------------------------------------------------------------------
    def alloc_partial(self, full_width: int, device=None):
        import torch

        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        return {
            "hist": torch.zeros(1024, dtype=torch.int32, device=dev),
            "wave_pairs": torch.zeros((2, 256, full_width), dtype=torch.int32, device=dev),
            "vscope": torch.zeros(65536, dtype=torch.int32, device=dev),
        }

    def accumulate_partial(self, rgb_tile, partial, *, x_offset: int, full_width: int, yuv_tile=None,
                           settings: Optional[ScopeSettings] = None, width: Optional[int] = None,
                           stream: Optional[int] = None):
        """Accumulate one tile (H_tile, W_tile, 4) / (H_tile, linesize) CUDA tensor into the
        caller-zeroed partial accumulators."""
        import torch

        st = settings or ScopeSettings()
        ref = rgb_tile if rgb_tile is not None else yuv_tile
        if ref.dim() == 3 and ref.shape[-1] == 4 and width is None:
            h, w, _ = ref.shape
            assert ref.stride(2) == 1 and ref.stride(1) == 4
            width = w
        else:
            h = ref.shape[0]
            assert width is not None
        s = Surface()
        s.rgb_data = rgb_tile.data_ptr() if rgb_tile is not None else None
        s.yuv_data = yuv_tile.data_ptr() if yuv_tile is not None else None
        s.linesize, s.width, s.height, s.colorspace = ref.stride(0), width, h, st.colorspace
        pd = PartialDevice()
        pd.hist_counts = partial["hist"].data_ptr()
        pd.wave_pairs = partial["wave_pairs"].data_ptr()
        pd.vscope_counts = partial["vscope"].data_ptr()
        if stream is None:
            stream = torch.cuda.current_stream(ref.device).cuda_stream
        p = st.to_c()
        self.ctx.check(self.lib.scope_accumulate_partial(self.ctx.handle, C.byref(p), C.byref(s), x_offset,
                                                         full_width, C.byref(pd), C.c_void_p(stream)))

    def accumulate_band(self, tile, partial=None, *, x_offset: int, full_width: int, wave_outs=None,
                        exclusive: bool = False, yuv_tile=None, settings: Optional[ScopeSettings] = None,
                        width: Optional[int] = None, stream: Optional[int] = None):
        """``scope_accumulate_band``: one band of a tile-sharded frame.

        ``exclusive=True``: this call is the only writer of its columns of ``partial["wave_pairs"]`` (a row band
        accumulated in one call): pairs are stored, not added - no zero-fill, no atomics.
        ``wave_outs``: list of waveform images (CUDA tensors shaped like ``alloc_device_out(1, ...)["wave"]`` or plain
        device addresses, e.g. peer mappings) that receive this band's FINAL columns - only for bands that span the
        full height (column bands); the waveform then needs no partial and no reduce."""
        import torch

        st = settings or ScopeSettings()
        ref = tile if tile is not None else yuv_tile
        if ref.dim() == 3 and ref.shape[-1] == 4 and width is None:
            h, w, _ = ref.shape
            assert ref.stride(2) == 1 and ref.stride(1) == 4
            width = w
        else:
            h = ref.shape[0]
            assert width is not None
        s = Surface()
        s.rgb_data = tile.data_ptr() if tile is not None else None
        s.yuv_data = yuv_tile.data_ptr() if yuv_tile is not None else None
        s.linesize, s.width, s.height, s.colorspace = ref.stride(0), width, h, st.colorspace
        pd = None
        if partial is not None:
            pd = PartialDevice()
            pd.hist_counts = partial["hist"].data_ptr() if "hist" in partial else None
            pd.wave_pairs = partial["wave_pairs"].data_ptr() if "wave_pairs" in partial else None
            pd.vscope_counts = partial["vscope"].data_ptr() if "vscope" in partial else None
        outs, n_outs = None, 0
        if wave_outs:
            n_outs = len(wave_outs)
            outs = (C.c_void_p * n_outs)(*[(o.data_ptr() if hasattr(o, "data_ptr") else int(o)) for o in wave_outs])
        if stream is None:
            stream = torch.cuda.current_stream(ref.device).cuda_stream
        p = st.to_c()
        self.ctx.check(self.lib.scope_accumulate_band(self.ctx.handle, C.byref(p), C.byref(s), x_offset, full_width,
                                                      C.byref(pd) if pd is not None else None, outs, n_outs,
                                                      _ffi.BAND_EXCLUSIVE if exclusive else 0, C.c_void_p(stream)))

    def finalize_partial(self, partial, *, full_width: int, full_height: int,
                         settings: Optional[ScopeSettings] = None, stream: Optional[int] = None):
        import torch

        st = settings or ScopeSettings()
        dev = partial["hist"].device
        out = self.alloc_device_out(1, full_width, st, dev)
        pd = PartialDevice()
        pd.hist_counts = partial["hist"].data_ptr()
        pd.wave_pairs = partial["wave_pairs"].data_ptr()
        pd.vscope_counts = partial["vscope"].data_ptr()
        od = OutDevice()
        od.hist_counts = out["hist"].data_ptr() if "hist" in out else None
        od.hist_max = out["hist_max"].data_ptr() if "hist_max" in out else None
        od.wave = out["wave"].data_ptr() if "wave" in out else None
        od.vscope = out["vscope"].data_ptr() if "vscope" in out else None
        od.wave_display = out["wave_display"].data_ptr() if "wave_display" in out else None
        od.vscope_display = out["vscope_display"].data_ptr() if "vscope_display" in out else None
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        p = st.to_c()
        self.ctx.check(self.lib.scope_finalize_partial(self.ctx.handle, C.byref(p), full_width, full_height,
                                                       C.byref(pd), C.byref(od), C.c_void_p(stream)))
        return out

    def finalize_peers(self, partials, outs, *, full_width: int, full_height: int,
                       settings: Optional[ScopeSettings] = None, slice_index: int = 0, slice_count: int = 1,
                       stream: Optional[int] = None):
        """Reduce + saturate over peer memory in one kernel (``scope_finalize_peers``).

        ``partials``: one entry per rank, each a dict like ``alloc_partial`` returns whose values are
        CUDA tensors **or plain device addresses** (ints: peer mappings such as a symmetric-memory
        handle's ``buffer_ptrs``).  ``outs``: one dict per receiving rank in the layout of
        ``alloc_device_out(1, ...)`` (tensors or addresses); ``outs[0]`` is this rank's own output and
        the only one that receives the histogram.  Cross-rank synchronisation is the caller's."""
        import torch

        st = settings or ScopeSettings()

        def addr(v):
            if v is None:
                return None
            return v.data_ptr() if hasattr(v, "data_ptr") else int(v)

        pds = (PartialDevice * len(partials))()
        for i, part in enumerate(partials):
            pds[i].hist_counts = addr(part.get("hist"))
            pds[i].wave_pairs = addr(part.get("wave_pairs"))
            pds[i].vscope_counts = addr(part.get("vscope"))
        ods = (OutDevice * len(outs))()
        for i, out in enumerate(outs):
            ods[i].hist_counts = addr(out.get("hist"))
            ods[i].hist_max = addr(out.get("hist_max"))
            ods[i].wave = addr(out.get("wave"))
            ods[i].vscope = addr(out.get("vscope"))
            ods[i].wave_display = addr(out.get("wave_display"))
            ods[i].vscope_display = addr(out.get("vscope_display"))
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        p = st.to_c()
        self.ctx.check(self.lib.scope_finalize_peers(self.ctx.handle, C.byref(p), full_width, full_height, pds,
                                                     len(partials), slice_index, slice_count, ods, len(outs),
                                                     C.c_void_p(stream)))
        return outs[0]

    def finalize_multicast(self, mc_partial, local_out, mc_images=None, *, full_width: int, full_height: int,
                           settings: Optional[ScopeSettings] = None, slice_index: int = 0, slice_count: int = 1,
                           stream: Optional[int] = None):
        """The NVLS form (``scope_finalize_multicast``): ``mc_partial`` holds the multicast addresses of the partial
        arrays, ``mc_images`` (optional) those of the result images; ``local_out`` is this rank's own output."""
        import torch

        st = settings or ScopeSettings()

        def addr(v):
            if v is None:
                return None
            return v.data_ptr() if hasattr(v, "data_ptr") else int(v)

        pd = PartialDevice()
        pd.hist_counts, pd.wave_pairs, pd.vscope_counts = (addr(mc_partial.get(k)) for k in ("hist", "wave_pairs", "vscope"))

        def out_struct(d):
            od = OutDevice()
            od.hist_counts, od.hist_max = addr(d.get("hist")), addr(d.get("hist_max"))
            od.wave, od.vscope = addr(d.get("wave")), addr(d.get("vscope"))
            od.wave_display, od.vscope_display = addr(d.get("wave_display")), addr(d.get("vscope_display"))
            return od

        lo = out_struct(local_out)
        mi = out_struct(mc_images) if mc_images is not None else None
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        p = st.to_c()
        self.ctx.check(self.lib.scope_finalize_multicast(self.ctx.handle, C.byref(p), full_width, full_height,
                                                         C.byref(pd), slice_index, slice_count, C.byref(lo),
                                                         C.byref(mi) if mi is not None else None, C.c_void_p(stream)))
        return local_out

    def debug_yuv_table_strict(self, colorspace: int):
        """test hook: SCOPE_XFORM_FP32_STRICT for all 2^24 colours, u | y<<8 | v<<16"""
        import torch

        out = torch.empty(1 << 24, dtype=torch.int32, device="cuda")
        self.ctx.check(self.lib.scope_debug_yuv_table_strict(self.ctx.handle, colorspace, out.data_ptr(),
                                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    def debug_uv_table_v3(self, colorspace: int):
        """test hook: U | V << 8 of the headline kernel's own transform for all 2^24 colours"""
        import torch

        out = torch.empty(1 << 24, dtype=torch.int32, device="cuda")
        self.ctx.check(self.lib.scope_debug_uv_table_v3(self.ctx.handle, colorspace, out.data_ptr(),
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out

    # test hook
    def debug_yuv_table(self, colorspace: int):
        import torch

        out = torch.empty(1 << 24, dtype=torch.int32, device="cuda")
        self.ctx.check(self.lib.scope_debug_yuv_table(self.ctx.handle, colorspace, out.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out
