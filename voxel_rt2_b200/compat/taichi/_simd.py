"""Vectorised execution of the scene scripts' kernels (the f1 row of SURVEY.md §8: authoring speed).

A scene script is plain Python for the shim: `for i, j, k in ti.ndrange(...)` used to run one iteration at a
time (example5: 12.5 million iterations, 24 s). Here every `@ti.kernel` / `@ti.func` is re-written at decoration
time (an AST pass, `vectorise`) into a MASKED form in which an outermost `ti.ndrange` / `ti.grouped` loop runs all
its iterations at once: the loop indices become NumPy arrays with one element per iteration (a "lane"), `if` turns
into lane masks, assignments into selects, inner loops with per-lane bounds into masked loops over the union of the
bounds, `scene.set_voxel` into a logged scatter that is applied in iteration order when the loop ends.

The contract is that the vectorised loop leaves EXACTLY the scene the plain loop leaves:
 * `ti.random()` is a counter-based generator keyed by (seed, launch, lane, draw number of the lane) in both modes,
   and a lane only consumes a draw where the plain loop would have evaluated the call (masks follow `if`, `and` /
   `or` short-circuits and conditional expressions);
 * arithmetic is the same IEEE double arithmetic; functions whose NumPy version differs from libm in the last ulp
   are evaluated through the math module (taichi/math.py);
 * voxel writes are replayed sorted by (lane, time), i.e. in the order the sequential loop issues them.
`while`, `break` and `continue` are masked like `if`; `scene.get_voxel` reads the grid as it is when the loop starts and
the loop is re-run sequentially if it also WROTE one of the voxels it read. A loop whose iterations depend on each other (a reduction into a name or object defined outside the loop, or a value the
code after the loop reads) is recognised statically and runs one index at a time; so does a loop with fewer than MIN_LANES
iterations. Anything the pass does not understand
(closures, global statements, early returns, ...) leaves the function as it was — it then runs as plain Python, also
when called from a vectorised loop (one call per active lane, its writes joining the loop's log). A failure inside a
vectorised loop re-runs that loop sequentially; if part of it has already been applied, the scene is rolled back and
the whole kernel re-runs in plain Python.
VRT_SHIM_VECTORIZE=0 switches the pass off.
"""
import ast
import builtins as _b
import inspect
import os
import textwrap
import types

import numpy as np

from . import math as tm
from .math import Vec

ENABLED = os.environ.get("VRT_SHIM_VECTORIZE", "1") != "0" and os.environ.get("VRT_SHIM_RNG", "") != "legacy"
MAX_LANES = 1 << 16  # lanes per chunk of a vectorised loop: temporaries that stay in cache (example5: 2.7 s at 2^22, 1.4 s at 2^16)
MIN_LANES = 16       # a loop with fewer iterations runs one index at a time (its inner loops would otherwise walk arrays of a few lanes)
_ND = np.ndarray
_M64 = (1 << 64) - 1
_C0, _C1, _C2, _C3 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB, 0xD6E8FEB86659FD93


class Unsupported(Exception):
    pass


# ------------------------------------------------------------------------------------------ random numbers
class _RngState:
    """Counter-based ti.random(): value = mix(seed, launch, lane, draw). `launch` counts the outermost ndrange /
    grouped loops executed so far (plus one pseudo-launch for code outside any loop), `lane` is the iteration
    index inside that loop, `draw` the number of values the lane has consumed."""

    def __init__(self):
        self.seed = int(os.environ.get("VRT_SEED", "0"))
        self.reset()

    def reset(self):
        self.launch = 0      # launches started so far
        self.depth = 0       # > 0 while inside a lane-defining loop (plain mode)
        self.base = self._base(0, 0)
        self.ctr = 0         # draws of the current plain-mode lane (or of the code outside loops)
        self.outside_ctr = 0
        self.cur_launch = self.cur_lane = 0

    def _base(self, launch, lane):
        return (self.seed * _C0 + launch * _C1 + lane * _C2) & _M64


RNG = _RngState()
LANES = None  # the running vectorised loop (a _Lanes) or None


def _mix_int(x):
    x ^= x >> 30
    x = (x * _C1) & _M64
    x ^= x >> 27
    x = (x * _C2) & _M64
    x ^= x >> 31
    return x


def _mix_arr(x):
    x = x ^ (x >> np.uint64(30))
    x = x * np.uint64(_C1)
    x = x ^ (x >> np.uint64(27))
    x = x * np.uint64(_C2)
    x = x ^ (x >> np.uint64(31))
    return x


def random_scalar():
    r = RNG
    if r.base is None:  # first draw of this plain-mode lane
        r.base = r._base(r.cur_launch, r.cur_lane)
    x = _mix_int((r.base + r.ctr * _C3) & _M64)
    r.ctr += 1
    return (x >> 11) * (1.0 / 9007199254740992.0)


def random_lanes(mask):
    L = LANES
    with np.errstate(over="ignore"):
        x = _mix_arr(L.base + L.ctr * np.uint64(_C3))
    L.ctr += mask.astype(np.uint64)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def plain_lanes(it):
    """Plain-mode iteration of an outermost ndrange / grouped loop: every iteration is a lane of a new launch."""
    r = RNG
    if r.depth > 0 or LANES is not None:
        yield from it
        return
    r.launch += 1
    r.cur_launch = r.launch
    r.outside_ctr = r.ctr
    r.depth = 1
    try:
        for lane, t in enumerate(it):
            r.cur_lane = lane
            r.base = None  # computed by the lane's first draw, if any
            r.ctr = 0
            yield t
    finally:
        r.depth = 0
        r.base = r._base(0, 0)
        r.ctr = r.outside_ctr


# ------------------------------------------------------------------------------------------ lane context
class _Lanes:
    def __init__(self, launch, first_lane, n):
        self.n = n
        self.first = first_lane
        lane = np.arange(first_lane, first_lane + n, dtype=np.uint64)
        with np.errstate(over="ignore"):
            self.base = np.uint64((RNG.seed * _C0 + launch * _C1) & _M64) + lane * np.uint64(_C2)
        self.ctr = np.zeros(n, np.uint64)
        self.all = np.ones(n, bool)
        self.writes = []  # (scene, lanes, flat voxel index, material i8, rgb u8[., 3]) in time order
        self.reads = []   # (scene, flat voxel indices) read through scene.get_voxel while the writes are still in the log


UNDO = None  # while a top-level kernel call runs vectorised: (scene, voxel indices, old materials, old colours) of every applied launch


def _flush(L):
    """Replay the voxel writes of a finished vectorised loop in the order the sequential loop issues them:
    by lane, then by time (a stable sort of the time-ordered log by lane)."""
    by_scene = {}
    for w in L.writes:
        by_scene.setdefault(id(w[0]), []).append(w)
    for sc, idx in L.reads:
        # a voxel both read and written inside one launch: the sequential loop orders the two, the log does not
        ws = by_scene.get(id(sc))
        if ws is not None and np.isin(idx, np.concatenate([w[2] for w in ws])).any():
            raise Unsupported("a voxel is read and written inside one vectorised loop")
    for ws in by_scene.values():
        scene = ws[0][0]
        lanes = np.concatenate([w[1] for w in ws])
        order = np.argsort(lanes, kind="stable")
        idx = np.concatenate([w[2] for w in ws])[order]
        mat = np.concatenate([w[3] for w in ws])[order]
        rgb = np.concatenate([w[4] for w in ws])[order]
        # several writes to one voxel: the last one in this order wins (reverse + first occurrence)
        rev = idx[::-1]
        _, first = np.unique(rev, return_index=True)
        keep = len(idx) - 1 - first
        vm, vc = scene.voxel_material.reshape(-1), scene.voxel_color.reshape(-1, 3)
        if UNDO is not None:  # inside a kernel call: what the launch overwrites, for the rollback of a failed vectorised run
            UNDO.append((scene, idx[keep], vm[idx[keep]], vc[idx[keep]]))
        vm[idx[keep]] = mat[keep]
        vc[idx[keep]] = rgb[keep]


# ------------------------------------------------------------------------------------------ masks and selects
def live(m):
    return m is True or (m is not False and bool(m.any()))


def truth(c):
    if type(c) is _ND:
        return c if c.dtype == bool else c != 0
    if isinstance(c, Vec):
        raise Unsupported("vector used as a condition")
    return bool(c)


def m_and(m, c):
    """mask & condition; masks are True (plain code), False (nothing runs) or a bool array."""
    if c is True:
        return m
    if c is False or m is False:
        return False
    if m is True:
        return c
    return m & c


def m_or(a, b):
    if a is True or b is True:
        return True
    if a is False:
        return b
    if b is False:
        return a
    return a | b


def m_test(m, th):
    """Loop test of a while: evaluated for the lanes still in the loop (not at all when none is left)."""
    if not live(m):
        return False
    return m_and(m, truth(th(m)))


def brk(st, b, m):
    """`break` under mask m. Leaving an outermost loop that runs as a vectorised launch would cancel every later
    iteration, draws and writes included: that loop is re-run one index at a time instead."""
    if st is not None and st[0] == 1:
        raise Unsupported("break in a loop that runs as a vectorised launch")
    return m_or(b, m)


def m_andnot(m, c):
    if c is True or m is False:
        return False
    if c is False:
        return m
    if m is True:
        return ~c
    return m & ~c


def _sel(m, new, old):
    if isinstance(new, Vec) or isinstance(old, Vec):
        n = new.v if isinstance(new, Vec) else [new] * len(old.v)
        o = old.v if isinstance(old, Vec) else [old] * len(n)
        if len(n) != len(o):
            raise Unsupported("select between vectors of different sizes")
        return Vec([_sel(m, a, b) for a, b in zip(n, o)])
    if isinstance(new, (tuple, list)) and isinstance(old, (tuple, list)) and len(new) == len(old):
        return type(new)(_sel(m, a, b) for a, b in zip(new, old))
    if new is old:
        return new
    if not (isinstance(new, (int, float, _ND, np.generic)) and isinstance(old, (int, float, _ND, np.generic))):
        raise Unsupported("select between %s and %s" % (type(new).__name__, type(old).__name__))
    return np.where(m, new, old)


def assign(m, new, old_thunk):
    """x = new under mask m (lanes outside m keep the old value)."""
    if m is True:
        return new
    try:
        old = old_thunk()
    except NameError:  # first definition of the name: lanes outside the mask can never read it
        return new
    if _stale(old, m.shape[0]):  # left over from the previous chunk of a loop that runs in several chunks
        return new
    return _sel(m, new, old)


def _stale(x, n):
    if type(x) is _ND:
        return x.ndim > 0 and x.shape[0] != n
    if isinstance(x, Vec):
        return _b.any(_stale(c, n) for c in x.v)
    if isinstance(x, (tuple, list)):
        return _b.any(_stale(c, n) for c in x)
    return False


def setitem(m, obj, i, val):
    if m is True:
        obj[i] = val
    else:
        obj[i] = _sel(m, val, obj[i])


def setattr_(m, obj, name, val):
    if m is True:
        setattr(obj, name, val)
    else:
        setattr(obj, name, _sel(m, val, getattr(obj, name)))


def not_(x):
    t = truth(x)
    return ~t if type(t) is _ND else (not t)


def boolop(m, is_and, first, *rest):
    """`a and b and ...` / `a or b or ...`: operands after the first are thunks taking the mask under which they
    are evaluated (the lanes for which the sequential code would have reached them)."""
    acc = first
    for th in rest:
        if type(acc) is not _ND:
            if bool(acc) != is_and:  # short-circuit, Python semantics for plain values
                return acc
            acc = th(m)
            continue
        t = truth(acc)
        sub = m_and(m, t if is_and else ~t)
        if sub is False or (sub is not True and not sub.any()):
            return t
        v = truth(th(sub))
        acc = (t & v) if is_and else (t | v)
    return acc


def compare_chain(vals, ops):
    acc = None
    for k, op in enumerate(ops):
        r = op(vals[k], vals[k + 1])
        if isinstance(r, Vec):
            raise Unsupported("chained comparison of vectors")
        acc = r if acc is None else (truth(acc) & truth(r) if (type(acc) is _ND or type(r) is _ND) else (acc and r))
    return acc


def where(m, test, a, b):
    t = truth(test)
    if type(t) is not _ND:
        return a(m) if t else b(m)
    ma, mb = m_and(m, t), m_andnot(m, t)
    if not live(ma):
        return b(mb)
    if not live(mb):
        return a(ma)
    return _sel(t, a(ma), b(mb))


# ------------------------------------------------------------------------------------------ calls
def _exact_round(x, *a):
    if type(x) is _ND:
        raise Unsupported("builtin round() of a lane array")
    return _b.round(x, *a)


_BUILTIN_MAP = {_b.int: tm.int, _b.float: tm.float, _b.abs: tm.abs, _b.min: tm.min, _b.max: tm.max, _b.pow: tm.pow, _b.round: _exact_round,
                _b.any: tm.any, _b.all: tm.all}


def _has_lanes(x):
    if type(x) is _ND:
        return True
    if isinstance(x, Vec):
        return _b.any(type(c) is _ND for c in x.v)
    if isinstance(x, (tuple, list)):
        return _b.any(_has_lanes(c) for c in x)
    return False


def _lane_of(x, k):
    if type(x) is _ND:
        return x[k].item()
    if isinstance(x, Vec):
        return Vec([_lane_of(c, k) for c in x.v])
    if isinstance(x, (tuple, list)):
        return type(x)(_lane_of(c, k) for c in x)
    return x


def _stack(vals, n, active):
    """Per-lane results of a scalarised call -> lane arrays (lanes outside `active` get zeros)."""
    v0 = vals[0]
    if isinstance(v0, Vec):
        return Vec([_stack([v.v[c] for v in vals], n, active) for c in range(len(v0.v))])
    if isinstance(v0, (tuple, list)):
        return type(v0)(_stack([v[c] for v in vals], n, active) for c in range(len(v0)))
    if v0 is None:
        return None
    out = np.zeros(n, np.float64 if _b.any(isinstance(v, float) for v in vals) else np.int64)
    out[active] = vals
    return out


class _PlainLog:
    def __init__(self):
        self.lane = 0
        self.rows = {}
        self.reads = {}

    def read(self, scene, idx):
        self.reads.setdefault(id(scene), (scene, []))[1].append(idx)

    def add(self, scene, idx, mat, rgb):
        row = self.rows.get(id(scene))
        if row is None:
            row = self.rows[id(scene)] = (scene, [], [], [], [])
        row[1].append(self.lane), row[2].append(idx), row[3].append(mat), row[4].append(rgb)


PLAIN_LOG = None  # set while a plain-Python function runs lane by lane inside a vectorised loop


def _scalarise(m, f, a, k):
    """A plain-Python function called from a vectorised loop: one call per active lane, in lane order."""
    global LANES
    global PLAIN_LOG
    L = LANES
    active = np.flatnonzero(m)
    if len(active) == 0:
        return None
    vals = []
    r = RNG
    LANES = None
    saved = (r.depth, r.base, r.ctr)
    r.depth = 1
    log = PLAIN_LOG = _PlainLog()  # voxel writes of the plain calls join the loop's log (set_voxel), reads are refused (get_voxel)
    try:
        for lane in active.tolist():
            r.base = int(L.base[lane])
            r.ctr = int(L.ctr[lane])
            log.lane = lane
            vals.append(f(*[_lane_of(x, lane) for x in a], **{n: _lane_of(x, lane) for n, x in k.items()}))
            L.ctr[lane] = r.ctr
    finally:
        r.depth, r.base, r.ctr = saved
        LANES = L
        PLAIN_LOG = None
    for sc, idx in log.reads.values():
        L.reads.append((sc, np.array(idx, np.int64)))
    for scene, rows in log.rows.items():
        sc, lanes, idx, mat, rgb = rows
        L.writes.append((sc, np.array(lanes, np.int64), np.array(idx, np.int64), np.array(mat, np.int8), np.array(rgb, np.uint8).reshape(-1, 3)))
    return _stack(vals, L.n, active)


def call(m, f, *a, **k):
    if m is False:
        raise Unsupported("call under an empty mask")
    if type(f) is types.MethodType:  # (attribute access on a bound method falls through to its function: test this first)
        s = getattr(f.__func__, "__simd__", None)
        if s is not None:
            return s(f.__self__, m, *a, **k)
    else:
        s = getattr(f, "__simd__", None)
        if s is not None:
            return s(m, *a, **k)
    if m is True:
        return f(*a, **k)
    g = _BUILTIN_MAP.get(f) if isinstance(f, types.BuiltinFunctionType) or isinstance(f, type) else None
    if g is not None:
        return g(*a, **k)
    if getattr(f, "__plain__", False):  # a @ti.func the pass could not vectorise
        return _scalarise(m, f, a, k)
    return f(*a, **k)  # shim maths, Vec methods, undecorated helper functions of the script: array-aware or pure expressions


# ------------------------------------------------------------------------------------------ loops
class VRange:
    """range() / 1-D ndrange with per-lane bounds inside a vectorised loop."""

    def __init__(self, lo, hi, step=1):
        if type(step) is _ND or step != 1:
            raise Unsupported("per-lane range with a step")
        self.lo, self.hi = lo, hi


def range_(*a):
    if _b.any(type(x) is _ND for x in a):
        if len(a) == 1:
            return VRange(0, tm.int(a[0]))
        return VRange(tm.int(a[0]), tm.int(a[1]), *a[2:])
    return range(*[_b.int(x) for x in a])


def _iter_vrange(m, vr):
    lo, hi = vr.lo, vr.hi
    act = m if m is not True else LANES.all
    if not act.any():
        return
    lo_a, hi_a = np.broadcast_to(lo, act.shape), np.broadcast_to(hi, act.shape)
    k0, k1 = int(lo_a[act].min()), int(hi_a[act].max())
    for k in range(k0, k1):
        mk = act & (lo_a <= k) & (k < hi_a)
        if mk.any():
            yield mk, k


def loop(m, it, st):
    """Iterate `it` under mask m: yields (mask, value) pairs. In plain code (m is True) an ndrange / grouped object
    becomes ONE iteration whose value holds every index as a lane array. `st` is the retry token of the loop
    statement: st[0] = 1 while this statement runs as a vectorised launch, 2 = run it one index at a time."""
    global LANES
    from . import _NDRange, _Grouped  # late: the package imports this module

    if m is False:
        return
    grouped = isinstance(it, _Grouped)
    nd = it.r if grouped else it
    if isinstance(nd, _NDRange):
        if nd.varying:
            if m is True:
                raise Unsupported("per-lane ndrange outside a vectorised loop")
            yield from _iter_ndrange_varying(m, nd, grouped)
            return
        if m is True and LANES is None and ENABLED and RNG.depth == 0 and st[0] != 2 and nd.total() >= MIN_LANES:
            st[0] = 1
            yield from _launch(nd, grouped, st)
            st[0] = 0
            return
        if m is True:
            for t in it:  # plain iteration (handles the lane bookkeeping of ti.random itself)
                yield True, t
            return
        # uniform bounds inside a vectorised loop: every active lane walks the same index space
        for t in nd.product():
            yield m, (Vec(list(t)) if grouped else (t[0] if nd.one_d else t))
        return
    if isinstance(it, VRange):
        if m is True:
            raise Unsupported("per-lane range outside a vectorised loop")
        yield from _iter_vrange(m, it)
        return
    for t in it:
        yield m, t


def _iter_ndrange_varying(m, nd, grouped):
    def rec(mask, d, prefix):
        if d == len(nd.bounds):
            yield mask, (Vec(list(prefix)) if grouped else (prefix[0] if nd.one_d else tuple(prefix)))
            return
        lo, hi = nd.bounds[d]
        for mk, k in _iter_vrange(mask, VRange(lo, hi)):
            yield from rec(mk, d + 1, prefix + [k])

    yield from rec(m, 0, [])


def retry_plain(st, e):
    """An exception left the body of a loop statement. If the statement was running as a vectorised launch that has
    not applied any writes yet, forget the launch and let the statement run again one index at a time."""
    if st[0] != 1 or not isinstance(e, (Unsupported, TypeError, ValueError, IndexError, AttributeError)):
        return False
    if len(st) > 1 and st[1] > 0:  # an earlier chunk of the loop has already been applied: only the kernel-level rollback is exact
        return False
    RNG.launch -= 1  # the plain run is the same launch
    st[0] = 2
    if os.environ.get("VRT_SHIM_DEBUG"):
        print("[taichi shim] loop falls back to plain Python: %s: %s" % (type(e).__name__, e))
    return True


def _launch(nd, grouped, st):
    global LANES
    shape = [len(r) for r in nd.ranges]
    total = 1
    for s in shape:
        total *= s
    RNG.launch += 1
    launch = RNG.launch
    if total == 0:
        return
    first = 0
    while first < total:
        n = _b.min(MAX_LANES, total - first)
        flat = np.arange(first, first + n, dtype=np.int64)
        idx = []
        rem = flat
        for d in range(len(shape) - 1, -1, -1):
            idx.append(rem % shape[d] * nd.ranges[d].step + nd.ranges[d].start)
            rem = rem // shape[d]
        idx.reverse()
        L = _Lanes(launch, first, n)
        LANES = L
        try:
            yield L.all, (Vec(idx) if grouped else (idx[0] if nd.one_d else tuple(idx)))
        finally:
            LANES = None
        _flush(L)
        if len(st) == 1:
            st.append(0)
        st[1] += 1
        first += n


# ------------------------------------------------------------------------------------------ the AST pass
_CMP = {ast.Lt: "lt", ast.LtE: "le", ast.Gt: "gt", ast.GtE: "ge", ast.Eq: "eq", ast.NotEq: "ne"}


class _Expr(ast.NodeTransformer):
    def __init__(self, mask):
        self.mask = mask

    def _vz(self, name):
        return ast.Attribute(ast.Name("_vz", ast.Load()), name, ast.Load())

    def _thunk(self, node):
        inner = _Expr("_mq").visit(node)
        return ast.Lambda(ast.arguments(posonlyargs=[], args=[ast.arg("_mq")], kwonlyargs=[], kw_defaults=[], defaults=[]), inner)

    def visit_Call(self, node):
        if any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords):
            raise Unsupported("star arguments")
        f = self.visit(node.func)
        args = [self.visit(a) for a in node.args]
        kws = [ast.keyword(k.arg, self.visit(k.value)) for k in node.keywords]
        if isinstance(node.func, ast.Name) and node.func.id == "range":
            return ast.Call(self._vz("range_"), args, kws)
        return ast.Call(self._vz("call"), [ast.Name(self.mask, ast.Load()), f] + args, kws)

    def visit_BoolOp(self, node):
        first = self.visit(node.values[0])
        rest = [self._thunk(v) for v in node.values[1:]]
        return ast.Call(self._vz("boolop"), [ast.Name(self.mask, ast.Load()), ast.Constant(isinstance(node.op, ast.And)), first] + rest, [])

    def visit_UnaryOp(self, node):
        if isinstance(node.op, ast.Not):
            return ast.Call(self._vz("not_"), [self.visit(node.operand)], [])
        return self.generic_visit(node)

    def visit_BinOp(self, node):
        if isinstance(node.op, ast.Pow):  # a ** b of lane arrays goes through libm pow, as the plain loop does
            return ast.Call(ast.Attribute(ast.Name("_vz", ast.Load()), "pow_", ast.Load()), [self.visit(node.left), self.visit(node.right)], [])
        return self.generic_visit(node)

    def visit_Compare(self, node):
        if len(node.ops) == 1:
            return self.generic_visit(node)
        if not all(type(o) in _CMP for o in node.ops):
            raise Unsupported("chained comparison with is / in")
        vals = [self.visit(node.left)] + [self.visit(c) for c in node.comparators]
        ops = [ast.Attribute(ast.Name("_op", ast.Load()), _CMP[type(o)], ast.Load()) for o in node.ops]
        return ast.Call(self._vz("compare_chain"), [ast.List(vals, ast.Load()), ast.List(ops, ast.Load())], [])

    def visit_IfExp(self, node):
        return ast.Call(self._vz("where"), [ast.Name(self.mask, ast.Load()), self.visit(node.test), self._thunk(node.body), self._thunk(node.orelse)], [])

    def visit_ListComp(self, node):
        # the element expression and the iterables are transformed, the comprehension itself stays plain Python
        if any(g.is_async for g in node.generators):
            raise Unsupported("async comprehension")
        return self.generic_visit(node)

    def _no(self, node):
        raise Unsupported(type(node).__name__)

    visit_Lambda = visit_GeneratorExp = visit_DictComp = visit_SetComp = visit_Await = visit_Yield = visit_YieldFrom = visit_NamedExpr = _no


def pow_(a, b):
    if isinstance(a, Vec) or isinstance(b, Vec):
        return a ** b
    return tm._pow(a, b)


def _has_jump(node):
    """Does the statement contain a break / continue of the loop it sits in (not of a loop nested inside it)?"""
    if isinstance(node, (ast.Break, ast.Continue)):
        return True
    if isinstance(node, (ast.For, ast.While, ast.FunctionDef, ast.Lambda)):
        return False
    return _b.any(_has_jump(c) for c in ast.iter_child_nodes(node))


def _target_names(t):
    if isinstance(t, ast.Name):
        return {t.id}
    if isinstance(t, (ast.Tuple, ast.List)):
        out = set()
        for e in t.elts:
            out |= _target_names(e)
        return out
    return set()


def _loads(node):
    return {n.id for n in ast.walk(node) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}


def _stored(stmts):
    return {n.id for n in ast.walk(ast.Module(list(stmts), [])) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store)}


def _exposed(stmts, defined):
    """Names a statement list reads before it has definitely assigned them (conservative: an assignment counts only if it is
    unconditional or sits on both arms of an if) -> (exposed names, names definitely assigned afterwards, stores into objects
    that were not created by these statements?)."""
    defined = set(defined)
    exposed = set()
    outer_store = False
    for st in stmts:
        if isinstance(st, ast.Assign):
            exposed |= _loads(st.value) - defined
            for t in st.targets:
                for sub in ([t] if not isinstance(t, (ast.Tuple, ast.List)) else t.elts):
                    if isinstance(sub, (ast.Subscript, ast.Attribute)):
                        exposed |= _loads(sub) - defined
                        base = sub
                        while isinstance(base, (ast.Subscript, ast.Attribute)):
                            base = base.value
                        if not (isinstance(base, ast.Name) and base.id in defined):
                            outer_store = True
                defined |= _target_names(t)
        elif isinstance(st, ast.AugAssign):
            exposed |= (_loads(st.value) | _loads(st.target) | _target_names(st.target)) - defined
            base = st.target
            while isinstance(base, (ast.Subscript, ast.Attribute)):
                base = base.value
            if isinstance(st.target, (ast.Subscript, ast.Attribute)) and not (isinstance(base, ast.Name) and base.id in defined):
                outer_store = True
        elif isinstance(st, ast.If):
            exposed |= _loads(st.test) - defined
            e1, d1, o1 = _exposed(st.body, defined)
            e2, d2, o2 = _exposed(st.orelse, defined)
            exposed |= e1 | e2
            outer_store |= o1 or o2
            defined |= d1 & d2
        elif isinstance(st, (ast.For, ast.While)):
            exposed |= _loads(st.iter if isinstance(st, ast.For) else st.test) - defined
            e, _, o = _exposed(st.body, defined | (_target_names(st.target) if isinstance(st, ast.For) else set()))
            exposed |= e
            outer_store |= o
        else:
            exposed |= _loads(st) - defined
    return exposed, defined, outer_store


def _carried(loop):
    """Does an iteration of this for-loop read what an earlier iteration wrote (`count += 1`, `acc[0] = ...` on something
    defined outside the loop)? Such a loop is sequential by nature and never runs as a vectorised launch."""
    exposed, _, outer_store = _exposed(loop.body, _target_names(loop.target))
    return outer_store or bool(exposed & _stored(loop.body))


class _Stmts:
    def __init__(self):
        self.n = 0
        self.loops = []  # (break-mask name, continue-mask name, retry-token name or None) of the enclosing loops

    def fresh(self, stem):
        self.n += 1
        return "_%s%d" % (stem, self.n)

    def expr(self, node, mask):
        return _Expr(mask).visit(node)

    def vz(self, name, *args):
        return ast.Call(ast.Attribute(ast.Name("_vz", ast.Load()), name, ast.Load()), list(args), [])

    def old(self, name):
        return ast.Lambda(ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]), ast.Name(name, ast.Load()))

    def store(self, target, value, mask):
        """Statements for `target = value` under `mask` (value is an already transformed expression)."""
        m = ast.Name(mask, ast.Load())
        if isinstance(target, ast.Name):
            return [ast.Assign([ast.Name(target.id, ast.Store())], self.vz("assign", m, value, self.old(target.id)))]
        if isinstance(target, ast.Subscript):
            if isinstance(target.slice, ast.Slice):
                raise Unsupported("slice assignment")
            return [ast.Expr(self.vz("setitem", m, self.expr(target.value, mask), self.expr(target.slice, mask), value))]
        if isinstance(target, ast.Attribute):
            return [ast.Expr(self.vz("setattr_", m, self.expr(target.value, mask), ast.Constant(target.attr), value))]
        if isinstance(target, (ast.Tuple, ast.List)):
            if any(isinstance(e, ast.Starred) for e in target.elts):
                raise Unsupported("starred assignment")
            tmp = self.fresh("t")
            out = [ast.Assign([ast.Name(tmp, ast.Store())], ast.Call(ast.Name("tuple", ast.Load()), [value], []))]
            for k, e in enumerate(target.elts):
                out += self.store(e, ast.Subscript(ast.Name(tmp, ast.Load()), ast.Constant(k), ast.Load()), mask)
            return out
        raise Unsupported("assignment target %s" % type(target).__name__)

    def _store_target(self, t):
        if isinstance(t, ast.Name):
            return ast.Name(t.id, ast.Store())
        if isinstance(t, (ast.Tuple, ast.List)) and all(isinstance(e, ast.Name) for e in t.elts):
            return ast.Tuple([ast.Name(e.id, ast.Store()) for e in t.elts], ast.Store())
        raise Unsupported("loop target")

    def block(self, stmts, mask, top=False, after=frozenset()):
        """`after`: the names that whatever runs after this block (in the enclosing blocks) reads before assigning them."""
        out = []
        for k, s in enumerate(stmts):
            follow = _exposed(stmts[k + 1:], set())[0] | after
            out += self.stmt(s, mask, last=top and k == len(stmts) - 1, after=follow)
            if self.loops and k + 1 < len(stmts) and not isinstance(s, (ast.For, ast.While)) and _has_jump(s):
                # lanes that hit a break / continue inside s skip the rest of the block
                b, c, _ = self.loops[-1]
                m2 = self.fresh("m")
                out.append(ast.Assign([ast.Name(m2, ast.Store())],
                                      self.vz("m_andnot", ast.Name(mask, ast.Load()), self.vz("m_or", ast.Name(c, ast.Load()), ast.Name(b, ast.Load())))))
                out.append(ast.If(self.vz("live", ast.Name(m2, ast.Load())), self.block(stmts[k + 1:], m2, after=after), []))
                break
        return out or [ast.Pass()]

    def stmt(self, s, mask, last=False, after=frozenset()):
        m = ast.Name(mask, ast.Load())
        if isinstance(s, ast.Expr):
            return [ast.Expr(self.expr(s.value, mask))]
        if isinstance(s, ast.Pass):
            return [s]
        if isinstance(s, ast.Break):
            b, _, st = self.loops[-1]
            return [ast.Assign([ast.Name(b, ast.Store())], self.vz("brk", ast.Name(st, ast.Load()) if st else ast.Constant(None), ast.Name(b, ast.Load()), m))]
        if isinstance(s, ast.Continue):
            _, c, _ = self.loops[-1]
            return [ast.Assign([ast.Name(c, ast.Store())], self.vz("m_or", ast.Name(c, ast.Load()), m))]
        if isinstance(s, ast.Assign):
            val = self.expr(s.value, mask)
            if len(s.targets) == 1:
                return self.store(s.targets[0], val, mask)
            tmp = self.fresh("t")
            out = [ast.Assign([ast.Name(tmp, ast.Store())], val)]
            for t in s.targets:
                out += self.store(t, ast.Name(tmp, ast.Load()), mask)
            return out
        if isinstance(s, ast.AugAssign):
            if isinstance(s.target, ast.Name):
                load = ast.Name(s.target.id, ast.Load())
            elif isinstance(s.target, ast.Subscript) and not isinstance(s.target.slice, ast.Slice):
                load = ast.Subscript(s.target.value, s.target.slice, ast.Load())
            elif isinstance(s.target, ast.Attribute):
                load = ast.Attribute(s.target.value, s.target.attr, ast.Load())
            else:
                raise Unsupported("augmented assignment target")
            return self.store(s.target, self.expr(ast.BinOp(load, s.op, s.value), mask), mask)
        if isinstance(s, ast.If):
            c, m1, m2 = self.fresh("c"), self.fresh("m"), self.fresh("m")
            out = [ast.Assign([ast.Name(c, ast.Store())], self.vz("truth", self.expr(s.test, mask))),
                   ast.Assign([ast.Name(m1, ast.Store())], self.vz("m_and", m, ast.Name(c, ast.Load()))),
                   ast.If(self.vz("live", ast.Name(m1, ast.Load())), self.block(s.body, m1, after=after), [])]
            if s.orelse:
                out += [ast.Assign([ast.Name(m2, ast.Store())], self.vz("m_andnot", m, ast.Name(c, ast.Load()))),
                        ast.If(self.vz("live", ast.Name(m2, ast.Load())), self.block(s.orelse, m2, after=after), [])]
            return out
        if isinstance(s, ast.For):
            if s.orelse:
                raise Unsupported("for ... else")
            m1, v = self.fresh("m"), self.fresh("v")
            st, ex = self.fresh("s"), self.fresh("e")
            jumps = _b.any(_has_jump(x) for x in s.body)
            bk, cn = self.fresh("b"), self.fresh("k")
            self.loops.append((bk, cn, st))
            inner = self.block(s.body, m1, after=after | _exposed(s.body, set())[0])  # (what follows the body includes its next iteration)
            self.loops.pop()
            # the loop variable is stored unmasked: a lane that does not take part in an iteration never reads it, and
            # the index of a masked inner loop stays a plain number
            body = [ast.Assign([self._store_target(s.target)], ast.Name(v, ast.Load()))] + inner
            if jumps:
                # per iteration: nobody has continued yet; lanes that broke out earlier stay out (plain code really leaves)
                body = [ast.Assign([ast.Name(cn, ast.Store())], ast.Constant(False)),
                        ast.If(ast.Compare(ast.Name(bk, ast.Load()), [ast.Is()], [ast.Constant(True)]), [ast.Break()], []),
                        ast.Assign([ast.Name(m1, ast.Store())], self.vz("m_andnot", ast.Name(m1, ast.Load()), ast.Name(bk, ast.Load()))),
                        ast.If(self.vz("live", ast.Name(m1, ast.Load())), body, [])]
            tgt = ast.Tuple([ast.Name(m1, ast.Store()), ast.Name(v, ast.Store())], ast.Store())
            # st = [0]
            # while True:
            #     try:
            #         for m1, v in _vz.loop(m, iter, st): body
            #         break
            #     except Exception as e:
            #         if not _vz.retry_plain(st, e): raise
            # a loop whose iterations depend on each other starts with its retry token at 2: never a vectorised launch
            # ... and so does a loop that leaves a value behind for the code after it (the last iteration's, sequentially)
            first = 2 if (_carried(s) or (after & _stored(s.body))) else 0
            loop = ast.For(tgt, self.vz("loop", m, self.expr(s.iter, mask), ast.Name(st, ast.Load())), body, [])
            handler = ast.ExceptHandler(ast.Name("Exception", ast.Load()), ex,
                                        [ast.If(ast.UnaryOp(ast.Not(), self.vz("retry_plain", ast.Name(st, ast.Load()), ast.Name(ex, ast.Load()))), [ast.Raise(None, None)], [])])
            # (a retry starts from a clean break mask: the assignment sits inside the retry loop)
            return [ast.Assign([ast.Name(st, ast.Store())], ast.List([ast.Constant(first)], ast.Load())),
                    ast.While(ast.Constant(True), [ast.Assign([ast.Name(bk, ast.Store())], ast.Constant(False)),
                                                   ast.Try([loop, ast.Break()], [handler], [], [])], [])]
        if isinstance(s, ast.While):
            if s.orelse:
                raise Unsupported("while ... else")
            # mw = m
            # while True:
            #     mw = m_and(mw, truth(test under mw))      lanes leave the loop one by one
            #     if not live(mw): break
            #     body under mw
            mw = self.fresh("m")
            bk, cn = self.fresh("b"), self.fresh("k")
            self.loops.append((bk, cn, None))
            inner = self.block(s.body, mw, after=after | _exposed(s.body, set())[0] | _loads(s.test))
            self.loops.pop()
            head = [ast.Assign([ast.Name(cn, ast.Store())], ast.Constant(False)),
                    ast.Assign([ast.Name(mw, ast.Store())], self.vz("m_andnot", ast.Name(mw, ast.Load()), ast.Name(bk, ast.Load()))),
                    ast.Assign([ast.Name(mw, ast.Store())], self.vz("m_test", ast.Name(mw, ast.Load()), _Expr(mw)._thunk(s.test))),
                    ast.If(ast.UnaryOp(ast.Not(), self.vz("live", ast.Name(mw, ast.Load()))), [ast.Break()], [])]
            return [ast.Assign([ast.Name(mw, ast.Store())], m), ast.Assign([ast.Name(bk, ast.Store())], ast.Constant(False)),
                    ast.While(ast.Constant(True), head + inner, [])]
        if isinstance(s, ast.Return):
            if not last:
                raise Unsupported("return before the end of the function")
            return [ast.Return(self.expr(s.value, mask) if s.value is not None else None)]
        raise Unsupported(type(s).__name__)


def _names_stored(tree):
    return {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store)}


def vectorise(fn):
    """The masked form of fn: `fn__simd(mask, *args)`, or Unsupported."""
    if fn.__closure__:
        raise Unsupported("closure")
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fd = tree.body[0]
    if not isinstance(fd, ast.FunctionDef):
        raise Unsupported("not a plain function")
    a = fd.args
    if a.vararg or a.kwarg or a.kwonlyargs or a.posonlyargs:
        raise Unsupported("argument list")
    if {"_vz", "_op", "_mq"} & (_names_stored(fd) | {x.arg for x in a.args}):
        raise Unsupported("reserved name")
    body = _Stmts().block(fd.body, "_m0", top=True)
    new_args = ast.arguments(posonlyargs=[], args=[ast.arg("_m0")] + a.args, kwonlyargs=[], kw_defaults=[], defaults=a.defaults)
    nf = ast.FunctionDef(fd.name + "__simd", new_args, body, [], fd.returns)
    if hasattr(fd, "type_params"):
        nf.type_params = []
    mod = ast.Module([nf], [])
    ast.copy_location(nf, fd)
    ast.fix_missing_locations(mod)
    ast.increment_lineno(mod, fn.__code__.co_firstlineno - 1)  # tracebacks point at the script's own lines
    code = compile(mod, inspect.getsourcefile(fn) or "<shim>", "exec")
    import operator

    g = fn.__globals__
    g.setdefault("_vz", __import__(__name__, fromlist=["x"]))
    g.setdefault("_op", operator)
    loc = {}
    exec(code, g, loc)
    return loc[fd.name + "__simd"]


def _snapshot():
    """Start of a top-level kernel call: the applied launches are recorded from here on (an undo log, not a copy of the grid)."""
    global UNDO
    UNDO = []
    return (RNG.launch, RNG.ctr, RNG.base, RNG.depth)


def _restore(snap):
    """Undo every launch the failed vectorised run has applied (newest first) and put the random stream back. Writes the
    kernel issued directly from plain code are not undone: the sequential re-run issues them again, in the same order."""
    global LANES, UNDO
    LANES = None
    for scene, idx, mat, rgb in reversed(UNDO or []):
        scene.voxel_material.reshape(-1)[idx] = mat
        scene.voxel_color.reshape(-1, 3)[idx] = rgb
    UNDO = None
    RNG.launch, RNG.ctr, RNG.base, RNG.depth = snap


_DEPTH = [0]
_WARNED = set()


def decorate(fn, is_kernel):
    """@ti.kernel / @ti.func."""
    if not ENABLED:
        return fn
    try:
        simd = vectorise(fn)
    except (Unsupported, OSError, TypeError, SyntaxError) as e:
        fn.__plain__ = True
        fn.__why_plain__ = str(e)
        return fn

    def entry(*a, **k):
        if LANES is not None or _DEPTH[0] > 0:
            return simd(True if LANES is None else LANES.all, *a, **k)  # (a direct call from plain code inside a lane context cannot happen: calls are rewritten)
        global UNDO
        snap = _snapshot()
        _DEPTH[0] += 1
        try:
            r = simd(True, *a, **k)
            UNDO = None
            return r
        except Exception as e:  # noqa: BLE001 - anything the masked form cannot do: the plain function can
            _restore(snap)
            if fn.__name__ not in _WARNED:
                _WARNED.add(fn.__name__)
                if os.environ.get("VRT_SHIM_DEBUG"):
                    import traceback

                    traceback.print_exc()
                print("[taichi shim] %s: vectorised run failed (%s: %s); running it as plain Python" % (fn.__name__, type(e).__name__, e))
            return _plain_everything(fn, a, k)
        finally:
            _DEPTH[0] -= 1

    entry.__simd__ = simd
    entry.__wrapped__ = fn
    entry.__name__ = fn.__name__
    entry.__doc__ = fn.__doc__
    return entry


def _plain_everything(fn, a, k):
    """Re-run a kernel with the pass switched off for everything it calls."""
    global ENABLED
    saved = ENABLED
    ENABLED = False  # loop() then iterates every ndrange one index at a time, also inside the masked forms fn calls
    try:
        return fn(*a, **k)
    finally:
        ENABLED = saved
