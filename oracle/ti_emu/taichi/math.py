"""taichi.math of the emulator (TEST INFRASTRUCTURE, see taichi/__init__.py). The GLSL-style helpers
are written as Taichi's python/taichi/math/mathimpl.py composes them from primitive operations, so
the float32 rounding sequence is the same."""
import math as _m

import numpy as _np

from . import (Matrix, VectorType, _float_in, _raw, _wrap, abs_, acos, asin, atan2, binop, ceil, cmpop, cos, exp, f32, floor, i32, log,
               max, min, power, round, sin, sqrt, tan, tanh, u32)
from . import cast as _cast
from . import select as _select

pi = _m.pi
e = _m.e
inf = float("inf")
nan = float("nan")

vec2 = VectorType(2, f32)
vec3 = VectorType(3, f32)
vec4 = VectorType(4, f32)
ivec2 = VectorType(2, i32)
ivec3 = VectorType(3, i32)
ivec4 = VectorType(4, i32)
uvec2 = VectorType(2, u32)
uvec3 = VectorType(3, u32)
uvec4 = VectorType(4, u32)


class _MatType(VectorType):
    def __call__(self, *args):
        rows = list(args[0]) if len(args) == 1 else list(args)
        if len(rows) == self.n and isinstance(rows[0], (Matrix, list, tuple)):
            return Matrix([r if isinstance(r, Matrix) else Matrix(list(r)) for r in rows], self.dtype)  # given vectors are rows
        flat = []
        for r in rows:
            flat.extend(list(r) if isinstance(r, (Matrix, list, tuple)) else [r])
        return Matrix([flat[i * self.m:(i + 1) * self.m] for i in range(self.n)], self.dtype)


mat2 = _MatType(2, f32, 2)
mat3 = _MatType(3, f32, 3)
mat4 = _MatType(4, f32, 4)

pow = power
abs = abs_


def mix(x, y, a):
    return binop("+", binop("*", x, binop("-", 1.0, a)), binop("*", y, a))


def clamp(x, xmin, xmax):
    return max(xmin, min(xmax, x))


def fract(x):
    return binop("-", x, floor(x))


def sign(x):
    a, _ = _float_in(x)
    return _wrap(_np.sign(a).astype(a.dtype))


def step(edge, x):
    return _cast(cmpop(">=", x, edge), f32) if isinstance(x, Matrix) or isinstance(edge, Matrix) else (f32(1.0) if cmpop(">=", x, edge) else f32(0.0))


def smoothstep(edge0, edge1, x):
    t = clamp(binop("/", binop("-", x, edge0), binop("-", edge1, edge0)), 0.0, 1.0)
    return binop("*", binop("*", t, t), binop("-", 3.0, binop("*", 2.0, t)))


def mod(x, y):
    return binop("-", x, binop("*", y, floor(binop("/", x, y))))


def dot(x, y):
    return x.dot(y)


def cross(x, y):
    return x.cross(y)


def normalize(x):
    return x.normalized()


def length(x):
    return x.norm()


def distance(x, y):
    return binop("-", x, y).norm()


def reflect(x, n):
    return binop("-", x, binop("*", binop("*", 2.0, x.dot(n)), n))


def refract(x, n, eta):
    dxn = x.dot(n)
    k = binop("-", 1.0, binop("*", binop("*", eta, eta), binop("-", 1.0, binop("*", dxn, dxn))))
    if k < 0:
        return binop("*", x, 0.0)
    return binop("-", binop("*", eta, x), binop("*", binop("+", binop("*", eta, dxn), sqrt(k)), n))


def isnan(x):
    a, _ = _float_in(x)
    r = _np.isnan(a)
    return bool(r) if r.ndim == 0 else Matrix(r.astype(_np.int32), _noconv=True)


def isinf(x):
    a, _ = _float_in(x)
    r = _np.isinf(a)
    return bool(r) if r.ndim == 0 else Matrix(r.astype(_np.int32), _noconv=True)


def inverse(m):
    return m.inverse()


def transpose(m):
    return m.transpose()


def radians(x):
    return binop("*", x, pi / 180.0)


def degrees(x):
    return binop("*", x, 180.0 / pi)


def log2(x):
    a, _ = _float_in(x)
    return _wrap(_np.log2(a).astype(a.dtype))


def select(c, a, b):
    return _select(c, a, b)
