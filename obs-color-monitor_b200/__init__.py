"""obs-color-monitor_b200 — B200-native scope accumulation (histogram / waveform /
vectorscope + BT.601/709 transform) behind obs-color-monitor's surface-callback seam.

The directory name carries a hyphen (it is the name the task fixes); import it as
``import obs_color_monitor_b200`` (alias module at the repo root) or with
``importlib.import_module("obs-color-monitor_b200")``.

Nothing in this package imports ``oracle/``: the CUDA library is the only implementation,
and loading fails loudly when it has not been built.
"""
from . import _ffi, frames, sharding  # noqa: F401
from ._ffi import (COMP_RGB, COMP_UV, COMP_Y, COMP_YUV, MODE_FUSED, MODE_SURFACE, SCOPE_ALL, SCOPE_HIST,  # noqa: F401
                   SCOPE_VSCOPE, SCOPE_WAVE, ScopeError)
from .scopes import ScopeEngine, ScopeSettings  # noqa: F401

__all__ = ["ScopeEngine", "ScopeSettings", "ScopeError", "frames", "COMP_RGB", "COMP_Y", "COMP_UV", "COMP_YUV",
           "MODE_FUSED", "MODE_SURFACE", "SCOPE_ALL", "SCOPE_HIST", "SCOPE_WAVE", "SCOPE_VSCOPE"]
