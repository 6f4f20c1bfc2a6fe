"""Drop-in replacement for voxel-rt2's scene.py: `from scene import Scene` in the reference's
example*.py / main.py resolves here, and `import taichi as ti` in those scripts then resolves to
the pure-Python shim (voxel_rt2_b200/compat/taichi) — no Taichi at run time.
Run an unmodified example with:  PYTHONPATH=/path/to/this/repo python example6.py"""
from voxel_rt2_b200.scene import Camera, Scene, save_image  # noqa: F401
