"""Compatibility layer so the reference's example*.py scene scripts run unchanged without Taichi.

`install()` registers the pure-Python `taichi` shim of this directory in sys.modules (only if the
real Taichi has not been imported). The root-level `scene.py` calls it before the example script
reaches its own `import taichi as ti` line — every example imports `scene` first."""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    if "taichi" in sys.modules:
        return sys.modules["taichi"]
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    return importlib.import_module("taichi")
