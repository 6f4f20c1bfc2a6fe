"""Minimal pure-Python stand-in for the part of Taichi that voxel-rt2's example scene scripts
use. Kernels run as ordinary Python: `@ti.kernel` / `@ti.func` are
pass-through decorators, `ti.random()` is a seeded Python RNG (seed: env VRT_SEED, default 0),
vectors are `taichi.math` Vec objects. Nothing here touches the GPU; the scene script only
fills the host voxel arrays through scene.set_voxel."""
import itertools
import math as _m
import os as _os
import random as _random

from . import math  # noqa: F401  (ti.math)
from .math import Vec, _lift1, _lift2

f32 = "f32"
i32 = "i32"
u8 = "u8"
i8 = "i8"
cpu = "cpu"
gpu = "gpu"
vulkan = "vulkan"
cuda = "cuda"

_rng = _random.Random(int(_os.environ.get("VRT_SEED", "0")))


def init(*args, **kwargs):
    return None


def seed(s):
    """Not part of Taichi's surface: reseed ti.random() (used by fixtures/tests)."""
    _rng.seed(int(s))


def kernel(fn):
    return fn


def func(fn):
    return fn


def static(x, *rest):
    return x if not rest else (x,) + rest


def random(dtype=float):
    if dtype in (int, i32):
        return _rng.getrandbits(31)
    return _rng.random()


def ndrange(*dims):
    ranges = []
    for d in dims:
        if isinstance(d, (tuple, list, Vec)):
            lo, hi = d
            ranges.append(range(int(lo), int(hi)))
        else:
            ranges.append(range(int(d)))
    if len(ranges) == 1:
        return _NDRange1(ranges[0])
    return _NDRange(ranges)


class _NDRange:
    def __init__(self, ranges):
        self.ranges = ranges

    def __iter__(self):
        return itertools.product(*self.ranges)


class _NDRange1(_NDRange):
    """A 1-D ndrange yields scalars in a plain `for`, index vectors under ti.grouped."""

    def __init__(self, r):
        self.ranges = [r]

    def __iter__(self):
        return iter(self.ranges[0])


def grouped(r):
    if isinstance(r, _NDRange):
        for t in itertools.product(*r.ranges):
            yield Vec(t)
    else:
        for t in r:
            yield Vec(t) if isinstance(t, tuple) else t


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


sin = _lift1(_m.sin)
cos = _lift1(_m.cos)
tan = _lift1(_m.tan)
sqrt = _lift1(_m.sqrt)
exp = _lift1(_m.exp)
log = _lift1(_m.log)
floor = _lift1(lambda x: float(_m.floor(x)))
ceil = _lift1(lambda x: float(_m.ceil(x)))
round = _lift1(_round_half_away)
abs = _lift1(lambda x: -x if x < 0 else x)
atan2 = _lift2(_m.atan2)
pow = _lift2(lambda a, b: a ** b)


def min(*a):
    return math.min(*a)


def max(*a):
    return math.max(*a)


class _Tools:
    pass


tools = _Tools()
