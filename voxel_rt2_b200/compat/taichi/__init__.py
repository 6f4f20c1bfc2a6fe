"""Minimal pure-Python stand-in for the part of Taichi that voxel-rt2's example scene scripts
use. Kernels are ordinary Python functions; `@ti.kernel` / `@ti.func`
re-write them so that an outermost `ti.ndrange` / `ti.grouped` loop runs all its iterations at
once on NumPy arrays (`_simd.py`; functions the pass does not understand stay as they are).
`ti.random()` is a counter-based generator keyed by (seed, loop launch, iteration, draw), the same
in the vectorised and the plain execution (seed: env VRT_SEED, default 0; VRT_SHIM_RNG=legacy brings
back the sequential Python RNG the committed fixture scenes were generated with). Vectors are
`taichi.math` Vec objects. Nothing here touches the GPU; the scene script only fills the host voxel
arrays through scene.set_voxel."""
import itertools
import math as _m
import os as _os
import random as _random

import numpy as _np

from . import math  # noqa: F401  (ti.math)
from .math import Vec, _lift1, _np1

f32 = "f32"
i32 = "i32"
u8 = "u8"
i8 = "i8"
cpu = "cpu"
gpu = "gpu"
vulkan = "vulkan"
cuda = "cuda"

_LEGACY_RNG = _os.environ.get("VRT_SHIM_RNG", "") == "legacy"
_rng = _random.Random(int(_os.environ.get("VRT_SEED", "0")))

from . import _simd  # noqa: E402


def init(*args, **kwargs):
    return None


def seed(s):
    """Not part of Taichi's surface: reseed ti.random() (used by fixtures/tests)."""
    _rng.seed(int(s))
    _simd.RNG.seed = int(s)
    _simd.RNG.reset()


def kernel(fn):
    return _simd.decorate(fn, True)


def func(fn):
    return _simd.decorate(fn, False)


def static(x, *rest):
    return x if not rest else (x,) + rest


def random(dtype=float):
    if _LEGACY_RNG:
        if dtype in (int, i32):
            return _rng.getrandbits(31)
        return _rng.random()
    r = _simd.random_scalar()
    if dtype in (int, i32):
        return int(r * 2147483648.0)
    return r


def _random_simd(m, dtype=float):
    if m is True:
        return random(dtype)
    r = _simd.random_lanes(m)
    if dtype in (int, i32):
        return (r * 2147483648.0).astype(_np.int64)
    return r


random.__simd__ = _random_simd


def ndrange(*dims):
    bounds = []
    for d in dims:
        if isinstance(d, (tuple, list, Vec)):
            lo, hi = d
        else:
            lo, hi = 0, d
        bounds.append((math.int(lo), math.int(hi)))
    if any(type(b) is _np.ndarray for lh in bounds for b in lh):
        return _NDRange(None, bounds)  # per-lane bounds: only inside a vectorised loop
    return _NDRange([range(lo, hi) for lo, hi in bounds], bounds)


class _NDRange:
    """A 1-D ndrange yields scalars in a plain `for`, index vectors under ti.grouped; higher dimensions yield tuples."""

    def __init__(self, ranges, bounds):
        self.ranges, self.bounds = ranges, bounds
        self.varying = ranges is None
        self.one_d = len(bounds) == 1

    def product(self):
        return itertools.product(*self.ranges)

    def total(self):
        n = 1
        for r in self.ranges:
            n *= len(r)
        return n

    def __iter__(self):
        if self.varying:
            raise TypeError("ndrange with per-lane bounds iterated outside a vectorised loop")
        if _LEGACY_RNG:
            return iter(self.ranges[0]) if self.one_d else self.product()
        return _simd.plain_lanes(iter(self.ranges[0]) if self.one_d else self.product())


class _Grouped:
    def __init__(self, r):
        self.r = r

    def __iter__(self):
        r = self.r
        if isinstance(r, _NDRange):
            it = (Vec(t) for t in r.product())
            return it if _LEGACY_RNG else _simd.plain_lanes(it)
        return (Vec(t) if isinstance(t, tuple) else t for t in r)


def grouped(r):
    return _Grouped(r)


def Vector(vals, dt=None):
    v = Vec(vals)
    if dt in (int, i32):
        return Vec([int(x) for x in v])
    if dt in (float, f32):
        return Vec([float(x) for x in v])
    return v


def cast(x, dt):
    if dt in (int, i32, i8, u8):
        return math.int(x)
    return math.float(x)


def _round_half_away(x):
    return float(_m.floor(x + 0.5)) if x >= 0 else float(_m.ceil(x - 0.5))


def _round_lanes(x):
    return _np.where(x >= 0, _np.floor(x + 0.5), _np.ceil(x - 0.5))


sin = math.sin
cos = math.cos
tan = math.tan
sqrt = math.sqrt
exp = math.exp
log = math.log
floor = math.floor
ceil = math.ceil
round = _lift1(_np1(_round_half_away, _round_lanes))
abs = _lift1(lambda x: _np.abs(x) if type(x) is _np.ndarray else (-x if x < 0 else x))
atan2 = math.atan2
pow = math.pow


def min(*a):
    return math.min(*a)


def max(*a):
    return math.max(*a)


class _Tools:
    pass


tools = _Tools()
