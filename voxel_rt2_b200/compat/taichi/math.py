"""taichi.math subset used by the example scenes: vec2/3/4, ivec2/3/4, mix, fract, dot, pi, any,
abs, ... plus `int` / `float` that also accept vectors. The examples do
`from taichi.math import *`, so these two names shadow the builtins inside the scene script —
that is how `int(vec2(...))` (legal inside a Taichi kernel) keeps working in plain Python."""
import builtins as _b
import math as _m

import operator as _operator

import numpy as _np

pi = _m.pi
e = _m.e
inf = float("inf")


def _is_vec(x):
    return isinstance(x, Vec)


_set = object.__setattr__
_SCALARS = (_b.int, _b.float, _b.bool)
_ND = _np.ndarray

# --- lane arrays. Inside a vectorised loop (compat/taichi/_simd.py) the loop indices, and everything computed
# from them, are NumPy arrays with one element per iteration ("lane"); a Vec then holds arrays as components.
# Every function below therefore accepts arrays next to Python numbers. Functions whose NumPy implementation is
# not bit-identical to libm's (exp, log, pow, tan, atan, acos, ... differ in the last ulp) go through the math
# module element by element, so that a vectorised loop computes exactly what the plain Python loop computes.


def _exact1(f):
    def g(x):
        if type(x) is _ND:
            return _np.fromiter(map(f, x.ravel().tolist()), _np.float64, x.size).reshape(x.shape)
        return f(x)

    return g


def _exact2(f):
    def g(a, b):
        if type(a) is _ND or type(b) is _ND:
            a, b = _np.broadcast_arrays(a, b)
            return _np.fromiter(map(f, a.ravel().tolist(), b.ravel().tolist()), _np.float64, a.size).reshape(a.shape)
        return f(a, b)

    return g


def _pow_any(a, b):
    """a ** b with Python's semantics (libm pow for floats, exact integers for int ** non-negative int)."""
    if type(a) is _ND or type(b) is _ND:
        ka = a.dtype.kind if type(a) is _ND else ("i" if type(a) is _b.int else "f")
        kb = b.dtype.kind if type(b) is _ND else ("i" if type(b) is _b.int else "f")
        if ka in "iub" and kb in "iub" and _np.all(_np.asarray(b) >= 0):
            return _np.power(_np.asarray(a, _np.int64), _np.asarray(b, _np.int64))
        return _pow_f(a, b)
    return a ** b


def _np1(f, nf):
    def g(x):
        if type(x) is _ND:
            return nf(x)
        return f(x)

    return g


def _bi(x):
    """int(bool) of a comparison result (Taichi comparisons yield integers)."""
    return x.astype(_np.int64) if type(x) is _ND else _b.int(x)


def _ii(x):
    return x.astype(_np.int64) if type(x) is _ND else _b.int(x)


_sqrt = _np1(_m.sqrt, _np.sqrt)
_pow_f = _exact2(_operator.pow)
_pow = _pow_any


class Vec:
    """Small fixed-size numeric vector with GLSL-style swizzles and element-wise operators."""

    __slots__ = ("v",)
    __array_ufunc__ = None  # lane_array * Vec must reach Vec.__rmul__ instead of NumPy treating the Vec as a sequence
    _SW = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}

    def __init__(self, vals):
        _set(self, "v", vals if type(vals) is list else list(vals))

    # --- container protocol
    def __len__(self):
        return len(self.v)

    def __iter__(self):
        return iter(self.v)

    def __getitem__(self, i):
        return self.v[i]

    def __setitem__(self, i, val):
        self.v[i] = val

    def __repr__(self):
        return "Vec(%r)" % (self.v,)

    def __getattr__(self, name):
        sw = Vec._SW
        try:
            if len(name) == 1:
                return self.v[sw[name]]
            return Vec([self.v[sw[c]] for c in name])
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, val):
        if name == "v":
            object.__setattr__(self, name, val)
            return
        sw = Vec._SW
        if len(name) == 1 and name in sw:
            self.v[sw[name]] = val
        elif all(c in sw for c in name):
            for c, x in zip(name, val):
                self.v[sw[c]] = x
        else:
            raise AttributeError(name)

    # --- arithmetic
    def _bin(self, o, f):
        if isinstance(o, Vec):
            return Vec([f(a, b) for a, b in zip(self.v, o.v)])
        if isinstance(o, (tuple, list)):
            return Vec([f(a, b) for a, b in zip(self.v, o)])
        return Vec([f(a, o) for a in self.v])

    def _rbin(self, o, f):
        if isinstance(o, (tuple, list)):
            return Vec([f(b, a) for a, b in zip(self.v, o)])
        return Vec([f(o, a) for a in self.v])

    def __add__(self, o):
        a = self.v
        if type(o) is Vec:
            return Vec([x + y for x, y in zip(a, o.v)])
        if type(o) in _SCALARS:
            return Vec([x + o for x in a])
        return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._rbin(o, lambda a, b: a + b)
    def __sub__(self, o):
        a = self.v
        if type(o) is Vec:
            return Vec([x - y for x, y in zip(a, o.v)])
        if type(o) in _SCALARS:
            return Vec([x - o for x in a])
        return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._rbin(o, lambda a, b: a - b)
    def __mul__(self, o):
        a = self.v
        if type(o) is Vec:
            return Vec([x * y for x, y in zip(a, o.v)])
        if type(o) in _SCALARS:
            return Vec([x * o for x in a])
        return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o):
        if type(o) in _SCALARS:
            return Vec([o * x for x in self.v])
        return self._rbin(o, lambda a, b: a * b)
    def __truediv__(self, o):
        if type(o) in _SCALARS:
            return Vec([x / o for x in self.v])
        return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._rbin(o, lambda a, b: a / b)
    def __floordiv__(self, o): return self._bin(o, lambda a, b: a // b)
    def __rfloordiv__(self, o): return self._rbin(o, lambda a, b: a // b)
    def __mod__(self, o): return self._bin(o, lambda a, b: a % b)
    def __rmod__(self, o): return self._rbin(o, lambda a, b: a % b)
    def __pow__(self, o): return self._bin(o, _pow)
    def __rpow__(self, o): return self._rbin(o, _pow)
    def __and__(self, o): return self._bin(o, lambda a, b: _ii(a) & _ii(b))
    def __rand__(self, o): return self._rbin(o, lambda a, b: _ii(a) & _ii(b))
    def __or__(self, o): return self._bin(o, lambda a, b: _ii(a) | _ii(b))
    def __ror__(self, o): return self._rbin(o, lambda a, b: _ii(a) | _ii(b))
    def __xor__(self, o): return self._bin(o, lambda a, b: _ii(a) ^ _ii(b))
    def __rxor__(self, o): return self._rbin(o, lambda a, b: _ii(a) ^ _ii(b))
    def __neg__(self): return Vec([-a for a in self.v])
    def __pos__(self): return Vec(self.v)
    def __abs__(self): return Vec([_b.abs(a) for a in self.v])  # abs() of an array is element-wise
    # comparisons are element-wise and return integer vectors, as in Taichi
    def __eq__(self, o): return self._bin(o, lambda a, b: _bi(a == b))
    def __ne__(self, o): return self._bin(o, lambda a, b: _bi(a != b))
    def __lt__(self, o): return self._bin(o, lambda a, b: _bi(a < b))
    def __le__(self, o): return self._bin(o, lambda a, b: _bi(a <= b))
    def __gt__(self, o): return self._bin(o, lambda a, b: _bi(a > b))
    def __ge__(self, o): return self._bin(o, lambda a, b: _bi(a >= b))
    __hash__ = None

    # --- methods used by the examples
    def dot(self, o):
        a, b = self.v, (o.v if type(o) is Vec else o)
        if len(a) == 3:
            return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
        return _b.sum(x * y for x, y in zip(a, b))

    def norm(self, eps=0.0):
        a = self.v
        if len(a) == 3:
            return _sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + eps)
        return _sqrt(_b.sum(x * x for x in a) + eps)

    def norm_sqr(self):
        return _b.sum(a * a for a in self.v)

    def normalized(self, eps=0.0):
        inv = 1.0 / (self.norm() + eps)
        return Vec([a * inv for a in self.v])

    def sum(self):
        return _b.sum(self.v)

    def max(self):
        return _minmax(_b.max, self.v)

    def min(self):
        return _minmax(_b.min, self.v)

    def cross(self, o):
        a, b = self.v, list(o)
        return Vec([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    def cast(self, dt):
        return int(self) if dt in (_b.int, "i32") else float(self)

    def to_list(self):
        return list(self.v)


def _flatten(args):
    out = []
    for a in args:
        if isinstance(a, (Vec, tuple, list)):
            out.extend(a)
        else:
            out.append(a)
    return out


def _make_ctor(n, conv):
    def ctor(*args):
        # fast path: n plain scalars (by far the most common call in scene scripts)
        if len(args) == n:
            for a in args:
                if type(a) not in _SCALARS:
                    break
            else:
                return Vec([conv(a) for a in args])
        vals = _flatten(args)
        if len(vals) == 1:
            vals = vals * n
        if len(vals) != n:
            raise TypeError("expected %d components, got %d" % (n, len(vals)))
        return Vec([conv(x) for x in vals])

    return ctor


def _to_f(x):
    return x.astype(_np.float64) if type(x) is _ND else _b.float(x)


def _to_i(x):
    return x.astype(_np.int64) if type(x) is _ND else _b.int(x)  # truncation toward zero, like a Taichi i32 cast


_vec2g, _vec3g, _vec4g = _make_ctor(2, _to_f), _make_ctor(3, _to_f), _make_ctor(4, _to_f)
_ivec3g = _make_ctor(3, _to_i)
_new = object.__new__


def vec3(*args):
    if len(args) == 3:
        a, b, c = args
        if type(a) in _SCALARS and type(b) in _SCALARS and type(c) in _SCALARS:
            r = _new(Vec)
            _set(r, "v", [_b.float(a), _b.float(b), _b.float(c)])
            return r
    return _vec3g(*args)


def vec2(*args):
    if len(args) == 2:
        a, b = args
        if type(a) in _SCALARS and type(b) in _SCALARS:
            r = _new(Vec)
            _set(r, "v", [_b.float(a), _b.float(b)])
            return r
    return _vec2g(*args)


def ivec3(*args):
    if len(args) == 3:
        a, b, c = args
        if type(a) in _SCALARS and type(b) in _SCALARS and type(c) in _SCALARS:
            r = _new(Vec)
            _set(r, "v", [_b.int(a), _b.int(b), _b.int(c)])
            return r
    return _ivec3g(*args)


vec4 = _vec4g
ivec2, ivec4 = _make_ctor(2, _to_i), _make_ctor(4, _to_i)
uvec2, uvec3, uvec4 = ivec2, ivec3, ivec4


def int(x=0, *a):  # noqa: A001 - deliberate shadow, see module docstring
    if isinstance(x, Vec):
        return Vec([_to_i(c) for c in x.v])
    if type(x) is _ND:
        return x.astype(_np.int64)
    return _b.int(x, *a)


def float(x=0.0):  # noqa: A001
    if isinstance(x, Vec):
        return Vec([_to_f(c) for c in x.v])
    if type(x) is _ND:
        return x.astype(_np.float64)
    return _b.float(x)


def _lift1(f):
    def g(x):
        if isinstance(x, Vec):
            return Vec([f(c) for c in x.v])
        return f(x)

    return g


def _lift2(f):
    def g(a, b):
        if isinstance(a, Vec):
            return a._bin(b, f)
        if isinstance(b, Vec):
            return b._rbin(a, f)
        return f(a, b)

    return g


def mix(x, y, a):
    return x * (1 - a) + y * a


_floor = _np1(lambda x: _b.float(_m.floor(x)), _np.floor)
_ceil = _np1(lambda x: _b.float(_m.ceil(x)), _np.ceil)
fract = _lift1(lambda x: x - _floor(x))  # int - float for an integer argument, as before
floor = _lift1(_floor)
ceil = _lift1(_ceil)
sign = _lift1(lambda x: _bi(x > 0) - _bi(x < 0))
sqrt = _lift1(_sqrt)
sin = _lift1(_exact1(_m.sin))  # (NumPy's sin / cos agree with libm on this machine, but its SIMD dispatch depends on the CPU)
cos = _lift1(_exact1(_m.cos))
tan = _lift1(_exact1(_m.tan))
exp = _lift1(_exact1(_m.exp))
log = _lift1(_exact1(_m.log))
acos = _lift1(_exact1(_m.acos))
asin = _lift1(_exact1(_m.asin))
atan2 = _lift2(_exact2(_m.atan2))
pow = _lift2(_pow)
mod = _lift2(lambda a, b: a - b * _floor(a / b))
step = _lift2(lambda edge, x: _np.where(x < edge, 0.0, 1.0) if (type(x) is _ND or type(edge) is _ND) else (0.0 if x < edge else 1.0))


def abs(x):  # noqa: A001
    return x.__abs__() if isinstance(x, Vec) else _b.abs(x)


def _mm2(f):
    nf = _np.minimum if f is _b.min else _np.maximum

    def g(a, b):
        if type(a) is _ND or type(b) is _ND:
            return nf(a, b)
        return f(a, b)

    return g


_MM = {_b.min: _mm2(_b.min), _b.max: _mm2(_b.max)}


def _minmax(f, args):
    if len(args) == 1 and isinstance(args[0], Vec):
        args = args[0].v
    elif len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    g = _MM[f]
    r = args[0]
    for o in args[1:]:
        if isinstance(r, Vec) or isinstance(o, Vec):
            r = _lift2(g)(r, o)
        else:
            r = g(r, o)
    return r


def min(*args):  # noqa: A001
    return _minmax(_b.min, args)


def max(*args):  # noqa: A001
    return _minmax(_b.max, args)


def clamp(x, lo, hi):
    return max(lo, min(hi, x))


def smoothstep(e0, e1, x):
    t = clamp((x - e0) / (e1 - e0), 0.0, 1.0)
    return t * t * (3.0 - 2.0 * t)


def dot(a, b):
    return Vec(a).dot(b)


def cross(a, b):
    return Vec(a).cross(b)


def length(a):
    return Vec(a).norm()


def distance(a, b):
    return (Vec(a) - b).norm()


def normalize(a):
    return Vec(a).normalized()


def _truthy(c):
    return (c != 0) if type(c) is _ND else bool(c)


def any(x):  # noqa: A001
    if isinstance(x, Vec):
        r = _truthy(x.v[0])
        for c in x.v[1:]:
            r = r | _truthy(c)
        return _bi(r)
    return _bi(_truthy(x))


def all(x):  # noqa: A001
    if isinstance(x, Vec):
        r = _truthy(x.v[0])
        for c in x.v[1:]:
            r = r & _truthy(c)
        return _bi(r)
    return _bi(_truthy(x))


def reflect(i, n):
    return i - 2.0 * dot(i, n) * n


__all__ = [
    "pi", "e", "inf", "Vec", "vec2", "vec3", "vec4", "ivec2", "ivec3", "ivec4", "uvec2", "uvec3", "uvec4", "int", "float", "mix", "fract",
    "floor", "ceil", "sign", "sqrt", "sin", "cos", "tan", "exp", "log", "acos", "asin", "atan2", "pow", "mod", "step", "abs", "min", "max",
    "clamp", "smoothstep", "dot", "cross", "length", "distance", "normalize", "any", "all", "reflect",
]
