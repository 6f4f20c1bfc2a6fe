"""taichi.lang of the emulator: only what scene.py imports at module level."""
