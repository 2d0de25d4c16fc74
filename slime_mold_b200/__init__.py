"""slime_mold_b200 -- B200-native engine for the slime-mold step loop of Velfi/slime-mold.

Only the hot path of the reference is here (src/compute.wgsl: agent sense/rotate/move/
deposit + trail decay/diffusion) behind the reference's own configuration surface
(`Settings`, presets, `SimSizeUniform`).  Device code: hand-written sm_100a CUDA in
csrc/, reached through the C ABI of include/slime_b200.h.  No CPU fallback.
"""
from .settings import Settings, SimSizeUniform  # noqa: F401
from .presets import Preset, PresetManager, init_preset_manager  # noqa: F401
from ._lib import SlimeError, SM_FLAG_GAUSSIAN_BLUR, SM_FLAG_NO_SORT, SM_FLAG_SEM_INPLACE  # noqa: F401
from .backend import CudaBackend, device_count  # noqa: F401
from .lut_manager import LutData, LutManager, write_png  # noqa: F401

__all__ = ["Settings", "SimSizeUniform", "Preset", "PresetManager", "init_preset_manager", "CudaBackend",
           "SlimeError", "device_count", "LutData", "LutManager", "write_png", "SM_FLAG_GAUSSIAN_BLUR", "SM_FLAG_NO_SORT", "SM_FLAG_SEM_INPLACE"]
