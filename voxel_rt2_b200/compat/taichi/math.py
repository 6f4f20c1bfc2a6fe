"""taichi.math subset used by the example scenes: vec2/3/4, ivec2/3/4, mix, fract, dot, pi, any,
abs, ... plus `int` / `float` that also accept vectors. The examples do
`from taichi.math import *`, so these two names shadow the builtins inside the scene script —
that is how `int(vec2(...))` (legal inside a Taichi kernel) keeps working in plain Python."""
import builtins as _b
import math as _m

pi = _m.pi
e = _m.e
inf = float("inf")


def _is_vec(x):
    return isinstance(x, Vec)


_set = object.__setattr__
_SCALARS = (_b.int, _b.float, _b.bool)


class Vec:
    """Small fixed-size numeric vector with GLSL-style swizzles and element-wise operators."""

    __slots__ = ("v",)
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
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __rpow__(self, o): return self._rbin(o, lambda a, b: a ** b)
    def __and__(self, o): return self._bin(o, lambda a, b: _b.int(a) & _b.int(b))
    def __rand__(self, o): return self._rbin(o, lambda a, b: _b.int(a) & _b.int(b))
    def __or__(self, o): return self._bin(o, lambda a, b: _b.int(a) | _b.int(b))
    def __ror__(self, o): return self._rbin(o, lambda a, b: _b.int(a) | _b.int(b))
    def __xor__(self, o): return self._bin(o, lambda a, b: _b.int(a) ^ _b.int(b))
    def __rxor__(self, o): return self._rbin(o, lambda a, b: _b.int(a) ^ _b.int(b))
    def __neg__(self): return Vec([-a for a in self.v])
    def __pos__(self): return Vec(self.v)
    def __abs__(self): return Vec([_b.abs(a) for a in self.v])
    # comparisons are element-wise and return integer vectors, as in Taichi
    def __eq__(self, o): return self._bin(o, lambda a, b: _b.int(a == b))
    def __ne__(self, o): return self._bin(o, lambda a, b: _b.int(a != b))
    def __lt__(self, o): return self._bin(o, lambda a, b: _b.int(a < b))
    def __le__(self, o): return self._bin(o, lambda a, b: _b.int(a <= b))
    def __gt__(self, o): return self._bin(o, lambda a, b: _b.int(a > b))
    def __ge__(self, o): return self._bin(o, lambda a, b: _b.int(a >= b))
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
            return _m.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + eps)
        return _m.sqrt(_b.sum(x * x for x in a) + eps)

    def norm_sqr(self):
        return _b.sum(a * a for a in self.v)

    def normalized(self, eps=0.0):
        inv = 1.0 / (self.norm() + eps)
        return Vec([a * inv for a in self.v])

    def sum(self):
        return _b.sum(self.v)

    def max(self):
        return _b.max(self.v)

    def min(self):
        return _b.min(self.v)

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
    return _b.float(x)


def _to_i(x):
    return _b.int(x)  # truncation toward zero, like a Taichi i32 cast


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
        return Vec([_b.int(c) for c in x.v])
    return _b.int(x, *a)


def float(x=0.0):  # noqa: A001
    if isinstance(x, Vec):
        return Vec([_b.float(c) for c in x.v])
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


fract = _lift1(lambda x: x - _m.floor(x))
floor = _lift1(lambda x: _b.float(_m.floor(x)))
ceil = _lift1(lambda x: _b.float(_m.ceil(x)))
sign = _lift1(lambda x: (x > 0) - (x < 0))
sqrt = _lift1(_m.sqrt)
sin = _lift1(_m.sin)
cos = _lift1(_m.cos)
tan = _lift1(_m.tan)
exp = _lift1(_m.exp)
log = _lift1(_m.log)
acos = _lift1(_m.acos)
asin = _lift1(_m.asin)
atan2 = _lift2(_m.atan2)
pow = _lift2(lambda a, b: a ** b)
mod = _lift2(lambda a, b: a - b * _m.floor(a / b))
step = _lift2(lambda edge, x: 0.0 if x < edge else 1.0)


def abs(x):  # noqa: A001
    return x.__abs__() if isinstance(x, Vec) else _b.abs(x)


def _minmax(f, args):
    if len(args) == 1 and isinstance(args[0], Vec):
        return f(args[0].v)
    r = args[0]
    for o in args[1:]:
        if isinstance(r, Vec) or isinstance(o, Vec):
            r = _lift2(lambda a, b: f(a, b))(r, o)
        else:
            r = f(r, o)
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


def any(x):  # noqa: A001
    if isinstance(x, Vec):
        return _b.int(_b.any(x.v))
    return _b.int(bool(x))


def all(x):  # noqa: A001
    if isinstance(x, Vec):
        return _b.int(_b.all(x.v))
    return _b.int(bool(x))


def reflect(i, n):
    return i - 2.0 * dot(i, n) * n


__all__ = [
    "pi", "e", "inf", "Vec", "vec2", "vec3", "vec4", "ivec2", "ivec3", "ivec4", "uvec2", "uvec3", "uvec4", "int", "float", "mix", "fract",
    "floor", "ceil", "sign", "sqrt", "sin", "cos", "tan", "exp", "log", "acos", "asin", "atan2", "pow", "mod", "step", "abs", "min", "max",
    "clamp", "smoothstep", "dot", "cross", "length", "distance", "normalize", "any", "all", "reflect",
]
