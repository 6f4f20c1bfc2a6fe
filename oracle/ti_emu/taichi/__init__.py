"""TEST INFRASTRUCTURE — a float32-faithful, pure-Python emulator of the Taichi surface used by
voxel-rt2's `renderer/*.py`, so that the reference's OWN source text can be executed in this
container (Taichi itself is not installable here) to generate golden vectors that pin the CPU
oracle. Only `tests/golden/make_ref_vectors.py` and tests import it; the product never does.

What is emulated (and how):
  * scalars are numpy scalars (f32 / i32 / u32 / u8 / i8 / f16); Python literals are "weak"
    constants that adopt the type of the other operand (default f32 / i32), as in Taichi;
  * every arithmetic node of a @ti.func / @ti.kernel body is rewritten (AST) into a call that
    applies Taichi's C-like type promotion and rounds to the result type after each operation
    (no FMA contraction, IEEE division) — Taichi's `fast_math` is therefore NOT emulated;
  * local variables are type-stable (re-assignment casts to the declared type), assignment of
    vectors / structs copies, ti.func arguments are passed by value except `ti.template()`
    parameters, which are passed by reference (written back at statement-level call sites);
  * Matrix helpers follow Taichi's python implementations: normalized() = (1 / norm) * v,
    dot / norm_sqr / matmul accumulate left to right, mix(x, y, a) = x * (1 - a) + y * a,
    clamp(x, lo, hi) = max(lo, min(x, hi)), reflect(x, n) = x - 2 * dot(x, n) * n; ti.max / ti.min
    ignore a NaN operand (fmax / fmin, as the LLVM backends lower them);
  * fields are numpy arrays (SNode dense layouts only), struct-for loops run sequentially,
    textures are float arrays with UNORM8 quantisation for rgba8;
  * ti.random() pops from a host-supplied source (`set_random_source`) so a harness can feed the
    same numbers to the reference code and to the oracle.
Anything else raises, loudly."""
import ast
import builtins
import inspect
import itertools
import sys
import textwrap
import types as _pytypes

import numpy as np

np.seterr(all="ignore")

# ------------------------------------------------------------------------------------ dtypes
f32 = np.float32
f64 = np.float64
f16 = np.float16
i8 = np.int8
i16 = np.int16
i32 = np.int32
i64 = np.int64
u8 = np.uint8
u16 = np.uint16
u32 = np.uint32
u64 = np.uint64

cpu = "cpu"
gpu = "gpu"
vulkan = "vulkan"
cuda = "cuda"

_WF = "weak_float"
_WI = "weak_int"
_F64 = np.dtype(np.float64)
_I64 = np.dtype(np.int64)
_F32 = np.dtype(np.float32)
_I32 = np.dtype(np.int32)

_scope_depth = 0  # > 0 while a @ti.func / @ti.kernel body is executing ("Taichi scope")


def in_taichi_scope():
    return _scope_depth > 0


def _as_dtype(dt):
    if dt is float:
        return _F32
    if dt is int:
        return _I32
    if isinstance(dt, VectorType):
        return dt.dtype
    return np.dtype(dt)


def _kind(x):
    """Type of an operand: a numpy dtype for typed values, _WF / _WI for weak Python constants."""
    if isinstance(x, Matrix):
        d = x.data.dtype
        return _WF if d == _F64 else (_WI if d == _I64 else d)
    if isinstance(x, (bool, np.bool_, int)):
        return _WI
    if isinstance(x, float):
        return _WF
    if isinstance(x, np.generic):
        d = x.dtype
        return _WF if d == _F64 else (_WI if d == _I64 else d)
    if isinstance(x, (list, tuple)):
        return _kind(Matrix(x))
    if isinstance(x, np.ndarray):
        d = x.dtype
        return _WF if d == _F64 else (_WI if d == _I64 else d)
    raise TypeError("taichi emulator: unsupported operand %r" % (type(x),))


def _is_float(k):
    return k == _WF or (k != _WI and np.issubdtype(k, np.floating))


def _promote(ka, kb):
    """Taichi's promoted_type: C-like, with weak Python constants adopting the typed side."""
    wa, wb = ka in (_WF, _WI), kb in (_WF, _WI)
    if wa and wb:
        return _WF if _WF in (ka, kb) else _WI
    if wa or wb:
        t, w = (kb, ka) if wa else (ka, kb)
        if w == _WF and not np.issubdtype(t, np.floating):
            return _F32
        return t
    fa, fb = np.issubdtype(ka, np.floating), np.issubdtype(kb, np.floating)
    if fa and fb:
        return ka if ka.itemsize >= kb.itemsize else kb
    if fa or fb:
        return ka if fa else kb
    if ka.itemsize != kb.itemsize:
        return ka if ka.itemsize > kb.itemsize else kb
    if ka == kb:
        return ka
    return ka if np.issubdtype(ka, np.unsignedinteger) else kb  # same width: unsigned wins


def _np_dtype(k):
    return _F64 if k == _WF else (_I64 if k == _WI else k)


def _raw(x, dt):
    if isinstance(x, Matrix):
        return x.data if x.data.dtype == dt else x.data.astype(dt)
    if isinstance(x, (list, tuple)):
        return _raw(Matrix(x), dt)
    if type(x) is int and np.issubdtype(dt, np.integer):  # wrap-around conversion of Python integers
        return np.asarray(x & 0xFFFFFFFFFFFFFFFF, dtype=np.uint64).astype(dt)
    return np.asarray(x).astype(dt)


def _wrap(r):
    """numpy result -> emulator value (0-d: scalar, weak kinds become Python numbers)."""
    if isinstance(r, np.ndarray) and r.ndim > 0:
        return Matrix(r, _noconv=True)
    r = np.asarray(r)
    if r.dtype == _F64:
        return float(r)
    if r.dtype == _I64:
        return int(r)
    if r.dtype == np.bool_:
        return bool(r)
    return r[()]


_BIN = {
    "+": np.add, "-": np.subtract, "*": np.multiply, "%": np.mod, "&": np.bitwise_and, "|": np.bitwise_or,
    "^": np.bitwise_xor, "<<": np.left_shift, ">>": np.right_shift, "//": np.floor_divide, "**": np.power,
}
_CMP = {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal, "==": np.equal, "!=": np.not_equal}
_PY_BIN = {
    "+": lambda a, b: a + b, "-": lambda a, b: a - b, "*": lambda a, b: a * b, "/": lambda a, b: a / b, "%": lambda a, b: a % b,
    "&": lambda a, b: a & b, "|": lambda a, b: a | b, "^": lambda a, b: a ^ b, "<<": lambda a, b: a << b, ">>": lambda a, b: a >> b,
    "//": lambda a, b: a // b, "**": lambda a, b: a ** b, "@": lambda a, b: a @ b,
}
_PYNUM = (bool, int, float)


def binop(op, a, b):
    if type(a) in _PYNUM and type(b) in _PYNUM:  # compile-time constants: Python semantics
        try:
            return _PY_BIN[op](a, b)
        except ZeroDivisionError:
            return float("inf") if a > 0 else (float("-inf") if a < 0 else float("nan"))
    if op == "@":
        return matmul(a, b)
    if isinstance(a, (Struct, np.ndarray)) or isinstance(b, (Struct, np.ndarray)) or a is None or b is None:
        if isinstance(a, Matrix) and isinstance(b, np.ndarray):
            return binop(op, a, Matrix(b))
        if isinstance(b, Matrix) and isinstance(a, np.ndarray):
            return binop(op, Matrix(a), b)
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            return _PY_BIN[op](a, b)  # Python-scope numpy code of the reference (np_rotate_matrix, ...)
        raise TypeError("taichi emulator: bad operands for %s: %r, %r" % (op, type(a), type(b)))
    ka, kb = _kind(a), _kind(b)
    if op in ("<<", ">>"):
        k = ka if ka not in (_WF, _WI) else (kb if kb not in (_WF, _WI) else _WI)
    else:
        k = _promote(ka, kb)
    if op == "/":
        if not _is_float(k):
            k = _WF if k == _WI else _F32
        dt = _np_dtype(k)
        return _wrap(np.divide(_raw(a, dt), _raw(b, dt)))
    dt = _np_dtype(k)
    A, B = _raw(a, dt), _raw(b, dt)
    if op == "//" and _is_float(k):
        return _wrap(np.floor(np.divide(A, B)))
    if op == "**" and not _is_float(k):
        return _wrap(np.power(A.astype(np.int64), B.astype(np.int64)).astype(dt))
    return _wrap(_BIN[op](A, B))


def cmpop(op, a, b):
    if type(a) in _PYNUM and type(b) in _PYNUM:
        return bool(_CMP[op](a, b))
    if op in ("is", "is not", "in", "not in") or a is None or b is None or isinstance(a, str) or isinstance(b, str):
        return {"is": lambda: a is b, "is not": lambda: a is not b, "in": lambda: a in b, "not in": lambda: a not in b,
                "==": lambda: a == b, "!=": lambda: a != b}[op]()
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return _CMP[op](a, b)
    dt = _np_dtype(_promote(_kind(a), _kind(b)))
    r = _CMP[op](_raw(a, dt), _raw(b, dt))
    if r.ndim > 0:
        return Matrix(r.astype(np.int32), _noconv=True)
    return bool(r)


def unop(op, a):
    if op == "not":
        return not truth(a)
    if type(a) in _PYNUM:
        return {"-": lambda: -a, "+": lambda: a, "~": lambda: ~a}[op]()
    if isinstance(a, np.ndarray):
        return -a if op == "-" else (~a if op == "~" else a)
    dt = _np_dtype(_kind(a))
    A = _raw(a, dt)
    return _wrap({"-": np.negative, "+": np.positive, "~": np.invert}[op](A))


def truth(x):
    if isinstance(x, Matrix):
        raise TypeError("taichi emulator: truth value of a vector")
    return bool(x)


def logical_and(*xs):
    r = True
    for x in xs:
        r = truth(x) and r
    return r


def logical_or(*xs):
    r = False
    for x in xs:
        r = truth(x) or r
    return r


# ------------------------------------------------------------------------------------ Matrix
_SWZ = {c: i for s in ("xyzw", "rgba", "stpq") for i, c in enumerate(s)}


class Matrix:
    """Vector (1-D) or matrix (2-D). dtype float64 / int64 == Python-scope ("weak") values."""

    __slots__ = ("data",)
    __array_ufunc__ = None  # numpy scalars defer to our reflected operators instead of iterating us

    def __init__(self, vals, dt=None, _noconv=False):
        if _noconv:
            object.__setattr__(self, "data", vals)
            return
        if isinstance(vals, Matrix):
            arr = vals.data.copy()
        elif isinstance(vals, np.ndarray):
            arr = vals.copy()
        else:
            rows = [v for v in vals]
            if rows and isinstance(rows[0], (Matrix, list, tuple, np.ndarray)):
                rows = [Matrix(r) if not isinstance(r, Matrix) else r for r in rows]
                k = None
                for r in rows:
                    k = _kind(r) if k is None else _promote(k, _kind(r))
                arr = np.stack([_raw(r, _np_dtype(k)) for r in rows])
            else:
                k = None
                for v in rows:
                    k = _kind(v) if k is None else _promote(k, _kind(v))
                if k is None:
                    k = _WF
                arr = np.array([_raw(v, _np_dtype(k))[()] for v in rows], dtype=_np_dtype(k))
        if dt is not None:
            arr = arr.astype(_as_dtype(dt))
        elif in_taichi_scope():
            arr = _typed_arr(arr)
        object.__setattr__(self, "data", arr)

    # --- structure
    @property
    def n(self):
        return self.data.shape[0]

    @property
    def m(self):
        return self.data.shape[1] if self.data.ndim > 1 else 1

    @property
    def shape(self):
        return self.data.shape

    def __len__(self):
        return self.data.shape[0]

    def __iter__(self):
        for i in range(self.data.shape[0]):
            yield self[i]

    def copy(self):
        return Matrix(self.data.copy(), _noconv=True)

    def get_shape(self):
        return self.data.shape

    def to_numpy(self):
        return self.data.copy()

    def to_list(self):
        return self.data.tolist()

    # --- element access
    def _elem(self, v):
        if isinstance(v, np.ndarray) and v.ndim > 0:
            return Matrix(v, _noconv=True)  # a view: writes go through (field[None][i, j] = ...)
        return _wrap(v)

    def __getitem__(self, k):
        if isinstance(k, tuple):
            k = tuple(int(i) for i in k)
        elif isinstance(k, Matrix):
            k = tuple(int(i) for i in k.data)
        else:
            k = int(k)
        return self._elem(self.data[k])

    def __setitem__(self, k, v):
        if isinstance(k, tuple):
            k = tuple(int(i) for i in k)
        else:
            k = int(k)
        self.data[k] = _raw(v, self.data.dtype)

    def __getattr__(self, name):
        try:
            idx = [_SWZ[c] for c in name]
        except KeyError:
            raise AttributeError(name)
        if len(idx) == 1:
            return _wrap(self.data[idx[0]])
        return Matrix(self.data[idx], _noconv=True)  # fancy indexing copies

    def __setattr__(self, name, v):
        try:
            idx = [_SWZ[c] for c in name]
        except KeyError:
            raise AttributeError(name)
        self.data[idx] = _raw(v, self.data.dtype)

    # --- arithmetic (Python-scope code and un-rewritten expressions)
    def __add__(self, o): return binop("+", self, o)
    def __radd__(self, o): return binop("+", o, self)
    def __sub__(self, o): return binop("-", self, o)
    def __rsub__(self, o): return binop("-", o, self)
    def __mul__(self, o): return binop("*", self, o)
    def __rmul__(self, o): return binop("*", o, self)
    def __truediv__(self, o): return binop("/", self, o)
    def __rtruediv__(self, o): return binop("/", o, self)
    def __floordiv__(self, o): return binop("//", self, o)
    def __mod__(self, o): return binop("%", self, o)
    def __pow__(self, o): return power(self, o)
    def __rpow__(self, o): return power(o, self)
    def __lshift__(self, o): return binop("<<", self, o)
    def __rshift__(self, o): return binop(">>", self, o)
    def __and__(self, o): return binop("&", self, o)
    def __or__(self, o): return binop("|", self, o)
    def __xor__(self, o): return binop("^", self, o)
    def __matmul__(self, o): return matmul(self, o)
    def __neg__(self): return unop("-", self)
    def __abs__(self): return abs_(self)
    def __lt__(self, o): return cmpop("<", self, o)
    def __le__(self, o): return cmpop("<=", self, o)
    def __gt__(self, o): return cmpop(">", self, o)
    def __ge__(self, o): return cmpop(">=", self, o)
    def __eq__(self, o): return cmpop("==", self, o)
    def __ne__(self, o): return cmpop("!=", self, o)
    __hash__ = None

    # --- Taichi's Matrix methods (python/taichi/lang/matrix.py semantics)
    def cast(self, dt):
        return cast(self, dt)

    def sum(self):
        r = self.data.reshape(-1)
        acc = r[0]
        for v in r[1:]:
            acc = acc + v
        return _wrap(acc)

    def dot(self, o):
        return (self * o).sum()

    def norm_sqr(self):
        return (self * self).sum()

    def norm(self, eps=0):
        return sqrt(binop("+", self.norm_sqr(), eps))

    def normalized(self, eps=0):
        invlen = binop("/", 1, binop("+", self.norm(), eps))
        return binop("*", invlen, self)

    def cross(self, o):
        a, b = self, o
        return Matrix([a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x])

    def transpose(self):
        return Matrix(self.data.T.copy(), _noconv=True)

    def min(self):
        return _wrap(self.data.min())

    def max(self):
        return _wrap(self.data.max())

    def inverse(self):
        return Matrix(np.linalg.inv(self.data.astype(np.float64)).astype(self.data.dtype if self.data.dtype != _I64 else _F64), _noconv=True)

    def fill(self, v):
        self.data[...] = _raw(v, self.data.dtype)

    def __repr__(self):
        return "Matrix(%r)" % (self.data,)

    # ti.Vector.field / ti.Matrix.field
    @staticmethod
    def field(n, m=None, dtype=f32, shape=None, **kw):
        if m is not None and not isinstance(m, int):  # ti.Vector.field(n, dtype, shape) positional
            dtype, m = m, None
        return Field(dtype, shape, (n,) if m is None else (n, m))


def _typed_arr(arr):
    if arr.dtype == _F64:
        return arr.astype(np.float32)
    if arr.dtype == _I64:
        return arr.astype(np.int32)
    return arr


def _typed(x):
    """A Python-scope (weak) vector entering Taichi scope becomes an f32 / i32 vector."""
    if isinstance(x, Matrix) and x.data.dtype in (_F64, _I64):
        return Matrix(_typed_arr(x.data), _noconv=True)
    return x


def Vector(vals, dt=None):
    return Matrix(vals, dt)


Vector.field = lambda n, dtype=f32, shape=None, **kw: Field(dtype, shape, (n,))


def matmul(a, b):
    a, b = (Matrix(a) if not isinstance(a, Matrix) else a), (Matrix(b) if not isinstance(b, Matrix) else b)
    if a.data.ndim != 2:
        raise TypeError("matmul: left operand must be a matrix")
    rows = []
    if b.data.ndim == 1:
        for i in range(a.data.shape[0]):
            acc = binop("*", a[i, 0], b[0])
            for k in range(1, a.data.shape[1]):
                acc = binop("+", acc, binop("*", a[i, k], b[k]))
            rows.append(acc)
        return Matrix(rows)
    for i in range(a.data.shape[0]):
        row = []
        for j in range(b.data.shape[1]):
            acc = binop("*", a[i, 0], b[0, j])
            for k in range(1, a.data.shape[1]):
                acc = binop("+", acc, binop("*", a[i, k], b[k, j]))
            row.append(acc)
        rows.append(row)
    return Matrix(rows)


# ------------------------------------------------------------------------------ math functions
def _float_in(x):
    """Operand of a float function as (ndarray, weak?)."""
    k = _kind(x)
    if k in (_WF, _WI):
        if in_taichi_scope():
            return _raw(x, _F32), False
        return _raw(x, _F64), True
    if not np.issubdtype(k, np.floating):
        k = _F32
    return _raw(x, k), False


def _f1(fn):
    def g(x):
        a, _ = _float_in(x)
        return _wrap(fn(a).astype(a.dtype))
    return g


sqrt = _f1(np.sqrt)
sin = _f1(np.sin)
cos = _f1(np.cos)
tan = _f1(np.tan)
asin = _f1(np.arcsin)
acos = _f1(np.arccos)
exp = _f1(np.exp)
log = _f1(np.log)
tanh = _f1(np.tanh)
floor = _f1(np.floor)
ceil = _f1(np.ceil)
rsqrt = _f1(lambda a: 1 / np.sqrt(a))


def round(x):  # ti.round: half away from zero (C roundf), exact: floor(|a| + 0.5) would round 0.49999997f up
    a, _ = _float_in(x)
    m = np.abs(a)
    f = np.floor(m)
    r = f + ((m - f) >= a.dtype.type(0.5)).astype(a.dtype)  # m - f is exact in floating point
    return _wrap(np.copysign(r, a).astype(a.dtype))


def abs_(x):
    if type(x) in _PYNUM:
        return builtins.abs(x)
    dt = _np_dtype(_kind(x))
    return _wrap(np.abs(_raw(x, dt)))


abs = abs_


def atan2(y, x):
    k = _promote(_kind(y), _kind(x))
    if not _is_float(k) or k in (_WF, _WI):
        k = _F32 if in_taichi_scope() or k not in (_WF, _WI) else _WF
    dt = _np_dtype(k)
    return _wrap(np.arctan2(_raw(y, dt), _raw(x, dt)).astype(dt))


def power(a, b):
    if type(a) in _PYNUM and type(b) in _PYNUM:
        return a ** b
    k = _promote(_kind(a), _kind(b))
    dt = _np_dtype(k)
    if _is_float(k):
        return _wrap(np.power(_raw(a, dt), _raw(b, dt)).astype(dt))
    return _wrap(np.power(_raw(a, np.dtype(np.int64)), _raw(b, np.dtype(np.int64))).astype(dt))


pow = power


def _minmax(fn):
    def g(*xs):
        if len(xs) == 1 and isinstance(xs[0], Matrix):
            return _wrap(fn.reduce(xs[0].data))
        acc = xs[0]
        for x in xs[1:]:
            if type(acc) in _PYNUM and type(x) in _PYNUM:
                acc = builtins.max(acc, x) if fn is np.fmax else builtins.min(acc, x)
                continue
            dt = _np_dtype(_promote(_kind(acc), _kind(x)))
            acc = _wrap(fn(_raw(acc, dt), _raw(x, dt)))
        return acc
    return g


# fmax / fmin semantics (a NaN operand is ignored), as Taichi's LLVM backends lower ti.max / ti.min
max = _minmax(np.fmax)
min = _minmax(np.fmin)


def select(c, a, b):
    if isinstance(c, Matrix):
        dt = _np_dtype(_promote(_kind(a), _kind(b)))
        return _wrap(np.where(c.data != 0, _raw(a, dt), _raw(b, dt)))
    r = a if truth(c) else b
    o = b if truth(c) else a
    if isinstance(r, Matrix):
        return r.copy()
    if type(r) in _PYNUM and type(o) not in _PYNUM:  # the discarded side still fixes the type
        return _raw(r, _np_dtype(_promote(_kind(a), _kind(b))))[()]
    return r


def cast(x, dt):
    if isinstance(dt, VectorType):
        return Matrix(x, dt.dtype)
    d = _as_dtype(dt)
    if isinstance(x, Matrix):
        src = x.data
    else:
        src = np.asarray(x)
    if np.issubdtype(d, np.integer) and np.issubdtype(src.dtype, np.floating):
        src = np.trunc(src)  # C cast: toward zero
        wide = src.astype(np.int64)
        return _wrap(wide.astype(d))
    if np.issubdtype(d, np.integer) and src.dtype == np.bool_:
        return _wrap(src.astype(d))
    return _wrap(src.astype(d))


def static(x, *rest):
    if rest:
        return (x,) + rest
    if isinstance(x, (range, _NDRange)):
        return _StaticIter(x)
    return x


class _StaticIter:
    def __init__(self, it):
        self.it = it

    def __iter__(self):
        if isinstance(self.it, _NDRange):
            return self.it.python_iter()
        return iter(self.it)


class _NDRange:
    def __init__(self, ranges):
        self.ranges = ranges

    def python_iter(self):
        if len(self.ranges) == 1:
            return iter(self.ranges[0])
        return itertools.product(*self.ranges)

    def __iter__(self):
        if len(self.ranges) == 1:
            return (np.int32(i) for i in self.ranges[0])
        return (tuple(np.int32(i) for i in t) for t in itertools.product(*self.ranges))


def ndrange(*dims):
    rs = []
    for d in dims:
        if isinstance(d, (tuple, list, Matrix)):
            lo, hi = d
            rs.append(range(int(lo), int(hi)))
        else:
            rs.append(range(int(d)))
    return _NDRange(rs)


def grouped(x):
    if isinstance(x, _NDRange):
        return (Matrix(np.array(t, dtype=np.int32), _noconv=True) for t in itertools.product(*x.ranges))
    if isinstance(x, (Field, StructField)):
        return (Matrix(np.array(t, dtype=np.int32), _noconv=True) for t in x._indices())
    raise TypeError("ti.grouped(%r)" % (type(x),))


def ti_iter(x):
    """Iterable of a (non-static) Taichi for loop: loop variables are i32."""
    if isinstance(x, range):
        return (np.int32(i) for i in x)
    if isinstance(x, (Field, StructField, _NdArg)):
        return x._struct_for()
    if isinstance(x, np.ndarray):
        return (np.int32(i) for i in range(x.shape[0]))
    return x


def loop_config(**kw):
    return None


def init(*a, **kw):
    return None


# ------------------------------------------------------------------------------------- random
_random_source = None


def set_random_source(fn, with_frame=False):
    """fn(caller_function_name[, caller_frame]) -> float in [0, 1). None restores the default
    generator. With the frame a harness can walk f_back to the kernel and read its loop variables
    (pixel, depth), i.e. map each draw to a dimension of a counter-based sampler."""
    global _random_source, _random_with_frame
    _random_source, _random_with_frame = fn, with_frame


_random_with_frame = False


_default_rng = np.random.default_rng(0)


def random(dtype=float):
    if _random_source is not None:
        fr = sys._getframe(1)
        return np.float32(_random_source(fr.f_code.co_name, fr) if _random_with_frame else _random_source(fr.f_code.co_name))
    return np.float32(_default_rng.random(dtype=np.float32))


# ------------------------------------------------------------------------------------- fields
def _norm_shape(shape):
    if shape is None:
        return None
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


class Field:
    def __init__(self, dtype, shape=None, elem_shape=()):
        self.dtype = _as_dtype(dtype)
        self.elem_shape = tuple(elem_shape)
        self.offset = None
        self.arr = None
        self.oob_zero = False
        shape = _norm_shape(shape)
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape, offset=None):
        self.shape = tuple(shape)
        self.arr = np.zeros(self.shape + self.elem_shape, dtype=self.dtype)
        self.offset = tuple(int(o) for o in offset) if offset is not None else None

    def _key(self, k):
        if k is None:
            return ()
        if isinstance(k, Matrix):
            k = tuple(int(i) for i in k.data)
        elif isinstance(k, (tuple, list)):
            k = tuple(int(i) for i in k)
        else:
            k = (int(k),)
        if self.offset is not None:
            k = tuple(i - o for i, o in zip(k, self.offset))
        for i, n in zip(k, self.shape):
            if i < 0 or i >= n:
                if self.oob_zero:
                    return None
                raise IndexError("field index %r out of range %r" % (k, self.shape))
        return k

    def __getitem__(self, k):
        k = self._key(k)
        if k is None:  # harness opt-in (Field.oob_zero): out-of-range reads return zeros
            return Matrix(np.zeros(self.elem_shape, self.dtype), _noconv=True) if self.elem_shape else self.dtype.type(0)
        v = self.arr[k]
        if self.elem_shape:
            return Matrix(v, _noconv=True)  # view into the field storage
        return _wrap(v)

    def __setitem__(self, k, v):
        self.arr[self._key(k)] = _raw(v, self.dtype) if not isinstance(v, np.ndarray) else v.astype(self.dtype)

    def _indices(self):
        off = self.offset or (0,) * len(self.shape)
        for t in itertools.product(*[range(n) for n in self.shape]):
            yield tuple(i + o for i, o in zip(t, off))

    def _struct_for(self):
        if len(self.shape) == 1:
            return (np.int32(t[0]) for t in self._indices())
        return (tuple(np.int32(i) for i in t) for t in self._indices())

    def fill(self, v):
        self.arr[...] = _raw(v, self.dtype)

    def from_numpy(self, a):
        self.arr[...] = np.asarray(a).astype(self.dtype).reshape(self.arr.shape)

    def to_numpy(self):
        return self.arr.copy()

    def get_field_members(self):
        return [self]


def field(dtype, shape=None, **kw):
    return Field(dtype, shape)


class _Axes:
    def __init__(self, ids):
        self.ids = ids


i = _Axes([0])
j = _Axes([1])
k = _Axes([2])
ij = _Axes([0, 1])
ijk = _Axes([0, 1, 2])
ik = _Axes([0, 2])
jk = _Axes([1, 2])


class _SNode:
    def __init__(self, dims=None):
        self.dims = dict(dims or {})

    def dense(self, axes, shape):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),) * len(axes.ids)
        d = dict(self.dims)
        for ax, s in zip(axes.ids, shape):
            d[ax] = d.get(ax, 1) * int(s)
        return _SNode(d)

    def place(self, *fields, offset=None):
        shape = tuple(self.dims[a] for a in sorted(self.dims))
        for f in fields:
            f._alloc(shape, offset)
        return self


class _Root:
    def dense(self, axes, shape):
        return _SNode().dense(axes, shape)


root = _Root()


# ------------------------------------------------------------------------------------ structs
class Struct:
    _members = {}

    def __init__(self, *args, **kw):
        names = list(self._members)
        vals = dict(zip(names, args))
        vals.update(kw)
        for n, t in self._members.items():
            object.__setattr__(self, n, _zero_of(t))
            if n in vals:
                setattr(self, n, vals[n])

    def __setattr__(self, n, v):
        t = self._members.get(n)
        if t is None:
            raise AttributeError("struct %s has no member %s" % (type(self).__name__, n))
        object.__setattr__(self, n, _convert_to(t, v))

    def copy(self):
        c = object.__new__(type(self))
        for n in self._members:
            v = getattr(self, n)
            object.__setattr__(c, n, v.copy() if isinstance(v, (Matrix, Struct)) else v)
        return c

    @classmethod
    def field(cls, shape=None, **kw):
        return StructField(cls, shape)


class VectorType:
    def __init__(self, n, dtype, m=None):
        self.n, self.m, self.dtype = n, m, _as_dtype(dtype)

    def __call__(self, *args):
        if len(args) == 1 and isinstance(args[0], (list, tuple, Matrix, np.ndarray)):
            vals = list(args[0]) if not isinstance(args[0], Matrix) else args[0]
        elif len(args) == 1 and self.n > 1:
            vals = [args[0]] * self.n
        else:
            vals = []
            for a in args:
                if isinstance(a, Matrix):
                    vals.extend(list(a))
                else:
                    vals.append(a)
        if self.m is not None and not isinstance(vals, Matrix) and len(vals) == self.n and isinstance(vals[0], Matrix):
            return Matrix(vals)  # rows
        dt = self.dtype
        if not in_taichi_scope():
            dt = _F64 if np.issubdtype(dt, np.floating) else _I64  # Python-scope vectors hold Python numbers
        m = Matrix(vals, dt)
        if m.data.shape[0] != self.n:
            raise TypeError("vec%d built from %d components" % (self.n, m.data.shape[0]))
        return m


def _zero_of(t):
    if isinstance(t, VectorType):
        shape = (t.n,) if t.m is None else (t.n, t.m)
        return Matrix(np.zeros(shape, dtype=t.dtype), _noconv=True)
    if isinstance(t, type) and issubclass(t, Struct):
        return t()
    return _as_dtype(t).type(0)


def _convert_to(t, v):
    if isinstance(t, VectorType):
        return Matrix(_raw(v, t.dtype).copy() if isinstance(v, Matrix) else np.broadcast_to(_raw(v, t.dtype), (t.n,)).copy(), _noconv=True)
    if isinstance(t, type) and issubclass(t, Struct):
        return v.copy()
    return cast(v, t) if not isinstance(v, Matrix) else v


def _member_type(t):
    if t is float:
        return f32
    if t is int:
        return i32
    return t


def _make_struct(name, members, methods=None):
    ns = {"_members": {n: _member_type(t) for n, t in members.items()}}
    ns.update(methods or {})
    return type(name, (Struct,), ns)


def dataclass(cls):
    methods = {n: v for n, v in cls.__dict__.items() if callable(v)}
    return _make_struct(cls.__name__, dict(cls.__dict__.get("__annotations__", {})), methods)


class StructField:
    def __init__(self, cls, shape=None):
        self.cls = cls
        self.fields = {}
        for n, t in cls._members.items():
            if isinstance(t, VectorType):
                self.fields[n] = Field(t.dtype, None, (t.n,) if t.m is None else (t.n, t.m))
            elif isinstance(t, type) and issubclass(t, Struct):
                self.fields[n] = StructField(t, None)
            else:
                self.fields[n] = Field(t, None)
        shape = _norm_shape(shape)
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape, offset=None):
        self.shape = tuple(shape)
        for f in self.fields.values():
            f._alloc(shape, offset)

    def __getitem__(self, k):
        s = object.__new__(self.cls)
        for n, f in self.fields.items():
            v = f[k]
            object.__setattr__(s, n, v.copy() if isinstance(v, (Matrix, Struct)) else v)
        return s

    def __setitem__(self, k, s):
        for n, f in self.fields.items():
            f[k] = getattr(s, n)

    def _indices(self):
        return next(iter(self.fields.values()))._indices()

    def _struct_for(self):
        return next(iter(self.fields.values()))._struct_for()


# ----------------------------------------------------------------------------------- textures
class Format:
    rgba8 = "rgba8"
    rgba32f = "rgba32f"
    rgba16f = "rgba16f"
    r32f = "r32f"


class Texture:
    def __init__(self, fmt, shape):
        self.fmt = fmt
        self.shape = tuple(int(s) for s in shape)
        self.arr = np.zeros(self.shape + (4,), dtype=np.float32)

    def _key(self, c):
        return tuple(int(v) for v in (c.data if isinstance(c, Matrix) else c))

    def store(self, coord, value):
        v = _raw(value, _F32)
        if self.fmt == Format.rgba8:  # UNORM8: round to nearest code
            v = (np.rint(np.clip(v, 0.0, 1.0) * np.float32(255.0)) / np.float32(255.0)).astype(np.float32)
        self.arr[self._key(coord)] = v

    def fetch(self, coord, lod=0):
        return Matrix(self.arr[self._key(coord)].copy(), _noconv=True)

    load = fetch


# ------------------------------------------------------------------------------------ atomics
def atomic_or(x, v):  # rewritten by the AST pass into a read-modify-write of the target expression
    return binop("|", x, v)


def atomic_min(x, v):
    return min(x, v)


def atomic_max(x, v):
    return max(x, v)


def atomic_add(x, v):
    return binop("+", x, v)


# --------------------------------------------------------------------------------- ti.types / misc
class _Template:
    pass


def template():
    return _Template()


class _NdarrayType:
    def __init__(self, element_dim=0, **kw):
        self.element_dim = element_dim


class _TextureType:
    def __init__(self, *a, **kw):
        pass


class _Types:
    @staticmethod
    def vector(n, dtype):
        return VectorType(n, dtype)

    @staticmethod
    def matrix(n, m, dtype):
        return VectorType(n, dtype, m)

    @staticmethod
    def struct(**members):
        return _make_struct("struct", members)

    @staticmethod
    def ndarray(*a, **kw):
        return _NdarrayType(*a, **kw)

    @staticmethod
    def texture(*a, **kw):
        return _TextureType()

    @staticmethod
    def rw_texture(*a, **kw):
        return _TextureType()


types = _Types()


class _NdArg:
    """Kernel ndarray argument with element_dim=1: data[i] is a vector."""

    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, k):
        return Matrix(self.arr[int(k)].copy(), _noconv=True)

    def _struct_for(self):
        return (np.int32(i) for i in range(self.arr.shape[0]))


class _Tools:
    @staticmethod
    def imread(path, channels=0):
        from PIL import Image

        im = np.asarray(Image.open(path).convert("RGB"))
        return np.ascontiguousarray(np.transpose(im[::-1], (1, 0, 2)))  # (w, h, c), y flipped like ti.tools.imread

    class image:
        @staticmethod
        def imwrite(img, path):
            raise NotImplementedError


tools = _Tools()


def data_oriented(cls):
    return cls


# ------------------------------------------------------------------------- AST rewriting of bodies
_UNDEF = type("Undefined", (), {"__repr__": lambda s: "<undefined>"})()
_NOCH = type("NoChange", (), {})()

_BINOPS = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/", ast.Mod: "%", ast.BitAnd: "&", ast.BitOr: "|", ast.BitXor: "^",
           ast.LShift: "<<", ast.RShift: ">>", ast.FloorDiv: "//", ast.Pow: "**", ast.MatMult: "@"}
_CMPOPS = {ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">=", ast.Eq: "==", ast.NotEq: "!=", ast.Is: "is", ast.IsNot: "is not",
           ast.In: "in", ast.NotIn: "not in"}
_UNOPS = {ast.USub: "-", ast.UAdd: "+", ast.Invert: "~", ast.Not: "not"}
_ATOMICS = ("atomic_or", "atomic_min", "atomic_max", "atomic_add")


def store(old, new):
    """Type-stable assignment to a local variable."""
    if isinstance(new, (Matrix, Struct)):
        if isinstance(old, Matrix) and isinstance(new, Matrix) and old.data.shape == new.data.shape and old.data.dtype not in (_F64, _I64):
            return Matrix(_raw(new, old.data.dtype).copy(), _noconv=True)
        return _typed(new).copy() if isinstance(new, Matrix) else new.copy()
    if isinstance(new, (tuple, list, Field, StructField, Texture, _NdArg, np.ndarray, str, _pytypes.FunctionType)) or new is None:
        return new
    if old is _UNDEF or isinstance(old, (Matrix, Struct)) or old is None:
        t = type(new)
        if t is float:
            return np.float32(new)
        if t is int or t is bool or t is np.bool_:
            return np.int32(new)
        return new
    if isinstance(old, np.generic) and isinstance(new, (np.generic,) + _PYNUM):
        return cast(new, old.dtype)
    return new


def getattr_(obj, name):
    v = getattr(obj, name)
    if isinstance(v, Matrix) and v.data.dtype in (_F64, _I64) and in_taichi_scope():
        return _typed(v)
    return v


def writeback(old, finals, idx):
    if finals is None or idx >= len(finals) or finals[idx] is _NOCH:
        return old
    return finals[idx]


def call_wb(f, args, kwargs):
    tf = getattr(f, "__func__", f)
    inner = getattr(tf, "__ti_inner__", None)
    if inner is None:
        return f(*args, **kwargs), None
    if hasattr(f, "__self__"):
        ret, finals = inner((f.__self__,) + tuple(args), kwargs, True)
        return ret, finals[1:]
    return inner(tuple(args), kwargs, True)


class _Rewriter(ast.NodeTransformer):
    def __init__(self, params, local_names=()):
        self.params = params
        self.locals = set(params) | set(local_names)
        self.tmp = 0

    def _call(self, fn, *args):
        return ast.Call(func=ast.Name(id=fn, ctx=ast.Load()), args=list(args), keywords=[])

    def visit_BinOp(self, n):
        self.generic_visit(n)
        return self._call("__ti_bin", ast.Constant(_BINOPS[type(n.op)]), n.left, n.right)

    def visit_UnaryOp(self, n):
        self.generic_visit(n)
        return self._call("__ti_un", ast.Constant(_UNOPS[type(n.op)]), n.operand)

    def visit_BoolOp(self, n):
        self.generic_visit(n)
        return self._call("__ti_and" if isinstance(n.op, ast.And) else "__ti_or", *n.values)

    def visit_Compare(self, n):
        self.generic_visit(n)
        left, parts = n.left, []
        for op, right in zip(n.ops, n.comparators):
            parts.append(self._call("__ti_cmp", ast.Constant(_CMPOPS[type(op)]), left, right))
            left = right
        return parts[0] if len(parts) == 1 else self._call("__ti_and", *parts)

    def visit_Attribute(self, n):
        self.generic_visit(n)
        if isinstance(n.ctx, ast.Load):
            return self._call("__ti_getattr", n.value, ast.Constant(n.attr))
        return n

    def _load(self, target):
        t = ast.parse(ast.unparse(target), mode="eval").body  # a Load-context copy of the target
        return self.visit(t)

    def _store_name(self, name, value):
        return ast.Assign(targets=[ast.Name(id=name, ctx=ast.Store())],
                          value=self._call("__ti_store", ast.Name(id=name, ctx=ast.Load()), value))

    def _assign_to(self, target, value_expr):
        """Statements assigning an (already rewritten) value expression to one target."""
        if isinstance(target, ast.Name):
            return [self._store_name(target.id, value_expr)]
        if isinstance(target, (ast.Tuple, ast.List)):
            self.tmp += 1
            tmp = "__ti_t%d" % self.tmp
            out = [ast.Assign(targets=[ast.Name(id=tmp, ctx=ast.Store())], value=value_expr)]
            for i, el in enumerate(target.elts):
                out += self._assign_to(el, ast.Subscript(value=ast.Name(id=tmp, ctx=ast.Load()), slice=ast.Constant(i), ctx=ast.Load()))
            return out
        tgt = target
        if isinstance(tgt, ast.Attribute):
            tgt = ast.Attribute(value=self.visit(tgt.value), attr=tgt.attr, ctx=ast.Store())
        elif isinstance(tgt, ast.Subscript):
            tgt = ast.Subscript(value=self.visit(tgt.value), slice=self.visit(tgt.slice), ctx=ast.Store())
        return [ast.Assign(targets=[tgt], value=value_expr)]

    def _wb_call(self, call):
        """call -> (statements, result expression) with by-reference write-back of Name arguments."""
        call = ast.Call(func=self.visit(call.func), args=[self.visit(a) for a in call.args],
                        keywords=[ast.keyword(arg=k.arg, value=self.visit(k.value)) for k in call.keywords])
        self.tmp += 1
        tmp = "__ti_r%d" % self.tmp
        kw = ast.Dict(keys=[ast.Constant(k.arg) for k in call.keywords], values=[k.value for k in call.keywords])
        stmts = [ast.Assign(targets=[ast.Name(id=tmp, ctx=ast.Store())],
                            value=self._call("__ti_callwb", call.func, ast.Tuple(elts=call.args, ctx=ast.Load()), kw))]
        fin = ast.Subscript(value=ast.Name(id=tmp, ctx=ast.Load()), slice=ast.Constant(1), ctx=ast.Load())
        for i, a in enumerate(call.args):
            if isinstance(a, ast.Name) and a.id != "self" and a.id in self.locals:
                stmts.append(ast.Assign(targets=[ast.Name(id=a.id, ctx=ast.Store())],
                                        value=self._call("__ti_wb", ast.Name(id=a.id, ctx=ast.Load()), fin, ast.Constant(i))))
        return stmts, ast.Subscript(value=ast.Name(id=tmp, ctx=ast.Load()), slice=ast.Constant(0), ctx=ast.Load())

    def _is_plain_call(self, v):
        if not isinstance(v, ast.Call) or any(isinstance(a, ast.Starred) for a in v.args):
            return False
        f = v.func
        if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "ti":
            return False
        return any(isinstance(a, ast.Name) and a.id in self.locals for a in v.args)

    def visit_Assign(self, n):
        if self._is_plain_call(n.value):
            stmts, res = self._wb_call(n.value)
        else:
            stmts, res = [], self.visit(n.value)
        if len(n.targets) > 1:
            self.tmp += 1
            tmp = "__ti_t%d" % self.tmp
            stmts.append(ast.Assign(targets=[ast.Name(id=tmp, ctx=ast.Store())], value=res))
            res = ast.Name(id=tmp, ctx=ast.Load())
        for t in n.targets:
            stmts += self._assign_to(t, res)
        return stmts

    def visit_AnnAssign(self, n):
        if n.value is None:
            return []
        return self._assign_to(n.target, self._call("__ti_cast", self.visit(n.value), self.visit(n.annotation)))

    def visit_AugAssign(self, n):
        val = self._call("__ti_bin", ast.Constant(_BINOPS[type(n.op)]), self._load(n.target), self.visit(n.value))
        return self._assign_to(n.target, val)

    def visit_Expr(self, n):
        v = n.value
        if isinstance(v, ast.Call) and isinstance(v.func, ast.Attribute) and v.func.attr in _ATOMICS and isinstance(v.func.value, ast.Name) \
                and v.func.value.id == "ti":
            target = v.args[0]
            val = self._call("__ti_" + v.func.attr, self._load(target), self.visit(v.args[1]))
            return self._assign_to(target, val)
        if self._is_plain_call(v):
            stmts, res = self._wb_call(v)
            return stmts + [ast.Expr(value=res)]
        return ast.Expr(value=self.visit(v))

    def visit_For(self, n):
        it = n.iter
        n.iter = self._call("__ti_iter", self.visit(it))
        n.body = self._body(n.body)
        n.orelse = self._body(n.orelse)
        return n

    def _body(self, stmts):
        out = []
        for s in stmts:
            r = self.visit(s)
            if isinstance(r, list):
                out += r
            elif r is not None:
                out.append(r)
        return out

    def visit_If(self, n):
        n.test = self.visit(n.test)
        n.body = self._body(n.body)
        n.orelse = self._body(n.orelse)
        return n

    def visit_While(self, n):
        n.test = self.visit(n.test)
        n.body = self._body(n.body)
        n.orelse = self._body(n.orelse)
        return n

    def visit_Return(self, n):
        val = self.visit(n.value) if n.value is not None else ast.Constant(None)
        fin = ast.Tuple(elts=[ast.Name(id=p, ctx=ast.Load()) for p in self.params], ctx=ast.Load())
        return ast.Return(value=ast.Tuple(elts=[val, fin], ctx=ast.Load()))


def _assigned_names(fn_node):
    names = []
    for node in ast.walk(fn_node):
        tgts = []
        if isinstance(node, ast.Assign):
            tgts = node.targets
        elif isinstance(node, (ast.AugAssign, ast.AnnAssign)):
            tgts = [node.target]
        for t in tgts:
            for nn in ast.walk(t):
                if isinstance(nn, ast.Name) and isinstance(nn.ctx, ast.Store) and nn.id not in names:
                    names.append(nn.id)
    return names


_HELPERS = {
    "__ti_bin": binop, "__ti_un": unop, "__ti_cmp": cmpop, "__ti_and": logical_and, "__ti_or": logical_or, "__ti_store": store,
    "__ti_iter": ti_iter, "__ti_callwb": call_wb, "__ti_wb": writeback, "__ti_getattr": getattr_, "__ti_cast": cast, "__ti_UNDEF": _UNDEF,
    "__ti_atomic_or": atomic_or, "__ti_atomic_min": atomic_min, "__ti_atomic_max": atomic_max, "__ti_atomic_add": atomic_add,
}


def _compile(fn):
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = tree.body[0]
    fdef.decorator_list = []
    params = [a.arg for a in fdef.args.args]
    for a in fdef.args.args:
        a.annotation = None
    fdef.returns = None
    local_names = [n for n in _assigned_names(fdef) if n not in params]
    rw = _Rewriter(params, local_names)
    fdef.args.defaults = [rw.visit(d) for d in fdef.args.defaults]
    body = rw._body(fdef.body)
    init = [ast.Assign(targets=[ast.Name(id=n, ctx=ast.Store())], value=ast.Name(id="__ti_UNDEF", ctx=ast.Load())) for n in local_names]
    fin = ast.Tuple(elts=[ast.Name(id=p, ctx=ast.Load()) for p in params], ctx=ast.Load())
    fdef.body = init + body + [ast.Return(value=ast.Tuple(elts=[ast.Constant(None), fin], ctx=ast.Load()))]
    ast.fix_missing_locations(tree)
    g = fn.__globals__
    g.update(_HELPERS)
    ns = {}
    code = compile(tree, "<ti_emu:%s:%s>" % (inspect.getsourcefile(fn), fn.__qualname__), "exec")
    exec(code, g, ns)
    return ns[fdef.name], params


def _decorate(fn, is_kernel):
    ann = dict(getattr(fn, "__annotations__", {}))
    sig = inspect.signature(fn)
    state = {}

    def inner(args, kwargs, want_finals):
        global _scope_depth
        if "f" not in state:
            state["f"], state["params"] = _compile(fn)
        params = state["params"]
        bound = sig.bind(*args, **kwargs)
        n_given = len(args)
        vals, by_ref = [], []
        for idx, p in enumerate(params):
            given = p in bound.arguments
            if given:
                v = bound.arguments[p]
            else:
                d = sig.parameters[p].default
                v = d
            a = ann.get(p)
            ref = isinstance(a, _Template) or p == "self"
            if not ref:
                if isinstance(v, (list, tuple)) and not isinstance(a, (_NdarrayType, _TextureType)):
                    v = Matrix(list(v))
                if isinstance(v, (Matrix, Struct)):
                    v = _typed(v).copy() if isinstance(v, Matrix) else v.copy()
                if isinstance(a, _NdarrayType):
                    v = _NdArg(np.asarray(v)) if a.element_dim == 1 else np.asarray(v)
                elif a is not None and not isinstance(a, (_TextureType, _Template)) and given:
                    v = cast(v, a) if not isinstance(a, type) or not issubclass(a, Struct) else v
                elif is_kernel and type(v) in _PYNUM:
                    v = store(_UNDEF, v)
            vals.append(v)
            by_ref.append(ref)
        _scope_depth += 1
        try:
            ret, finals = state["f"](*vals)
        finally:
            _scope_depth -= 1
        if want_finals:
            out = tuple(f if (r and i < n_given) else _NOCH for i, (f, r) in enumerate(zip(finals, by_ref)))
            return ret, out[:n_given]
        return ret

    def wrapper(*args, **kwargs):
        return inner(args, kwargs, False)

    wrapper.__ti_inner__ = inner
    wrapper.__name__ = fn.__name__
    wrapper.__qualname__ = fn.__qualname__
    wrapper.__wrapped__ = fn
    return wrapper


def func(fn):
    return _decorate(fn, False)


def kernel(fn):
    return _decorate(fn, True)


# ------------------------------------------------------------------------------------- ui stubs
class _Unavailable:
    def __getattr__(self, n):
        raise RuntimeError("taichi emulator: ti.ui / GGUI is not emulated")


ui = _Unavailable()
simt = _Unavailable()
profiler = _Unavailable()

from . import math  # noqa: E402,F401

Matrix.field = staticmethod(lambda n, m, dtype=f32, shape=None, **kw: Field(dtype, shape, (n, m)))
