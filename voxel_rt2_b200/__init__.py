"""voxel_rt2_b200 — B200-native rendering hot path for voxel-rt2 scenes.

Python host (this package) -> ctypes -> libvoxelrt.so (hand-written sm_100a CUDA).
There is no CPU fallback: importing `Renderer` without the built CUDA library raises.
"""
from .camera import look_at, perspective, default_camera_matrices  # noqa: F401
from .materials import material_table  # noqa: F401
from .renderer import Renderer  # noqa: F401

__all__ = ["Renderer", "look_at", "perspective", "default_camera_matrices", "material_table"]
