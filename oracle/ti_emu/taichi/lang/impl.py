"""taichi.lang.impl of the emulator: scene.py:22 imports _ti_core (used by the GGUI camera only)."""
_ti_core = None
