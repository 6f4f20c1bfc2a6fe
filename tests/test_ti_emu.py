"""Self-tests of oracle/ti_emu (the float32 Taichi emulator that executes the reference's source to
produce tests/golden/ref_*.npz): the Taichi semantics the golden vectors depend on, checked on small
kernels, plus — when /root/reference is mounted — a regeneration of three vector sections that must
reproduce the committed files bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "oracle", "ti_emu")


@pytest.fixture(scope="module")
def ti():
    # the product shim is also called `taichi`: drop it from sys.modules while the emulator is loaded
    for k in [k for k in sys.modules if k == "taichi" or k.startswith("taichi.")]:
        saved = sys.modules.pop(k)
        assert saved is not None
    sys.path.insert(0, EMU)
    try:
        import taichi as t
    finally:
        sys.path.remove(EMU)
    assert t.__file__.startswith(EMU)
    yield t
    for k in [k for k in sys.modules if k == "taichi" or k.startswith("taichi.")]:
        del sys.modules[k]


def test_float32_rounding_after_every_operation(ti, tmp_path):
    src = tmp_path / "k_round.py"
    src.write_text(
        "import taichi as ti\n"
        "@ti.func\n"
        "def f(a, b, c):\n"
        "    return a * b + c\n"
        "@ti.func\n"
        "def g(x):\n"
        "    y = 0\n"
        "    y = x * 2.5\n"        # type-stable local: stays i32, value truncated
        "    z = 1 / 3\n"           # compile-time constants fold in Python (double), then become f32
        "    return y, x / 2, z * x\n")
    sys.path.insert(0, str(tmp_path))
    import k_round

    a, b, c = np.float32(1.0000001), np.float32(3.0000002), np.float32(-3.0)
    assert k_round.f(a, b, c) == np.float32(np.float32(a * b) + c)          # no FMA contraction
    assert k_round.f(a, b, c) != np.float32(np.float64(a) * np.float64(b) + np.float64(c))
    y, h, zx = k_round.g(np.int32(3))
    assert y == 7 and y.dtype == np.int32
    assert h == np.float32(1.5) and h.dtype == np.float32                    # int / int is a float division
    assert zx == np.float32(np.float32(1 / 3) * np.float32(3))


def test_value_semantics_and_template_references(ti, tmp_path):
    src = tmp_path / "k_vals.py"
    src.write_text(
        "import taichi as ti\n"
        "@ti.func\n"
        "def bump(v, out: ti.template()):\n"
        "    v.x += 1.0\n"            # by value: the caller's vector is untouched
        "    out = v\n"               # ti.template(): written back to the caller
        "@ti.func\n"
        "def run(a):\n"
        "    b = a\n"                 # assignment copies
        "    b[1] = 5.0\n"
        "    r = ti.Vector([0.0, 0.0, 0.0])\n"
        "    bump(a, r)\n"
        "    return a, b, r\n")
    sys.path.insert(0, str(tmp_path))
    import k_vals

    a, b, r = k_vals.run(ti.Vector([1.0, 2.0, 3.0], ti.f32))
    assert list(a.data) == [1.0, 2.0, 3.0] and list(b.data) == [1.0, 5.0, 3.0] and list(r.data) == [2.0, 2.0, 3.0]


def test_integer_wraparound_promotion_and_matrix_helpers(ti):
    x = ti.u32(0xFFFFFFF0)
    assert ti.binop("+", x, 0x20) == np.uint32(0x10)                         # u32 wraps; the literal adopts u32
    assert ti.binop("<<", 1, np.int32(31)) == np.int32(-2147483648)
    assert ti.binop("*", np.int32(3), 0.5).dtype == np.float32               # int * float literal -> f32
    assert ti.binop("+", np.uint8(200), np.int32(100)) == np.int32(300)
    assert ti.cast(np.float32(-2.7), ti.i32) == -2                           # C truncation
    assert ti.round(np.float32(0.49999997)) == 0.0 and ti.round(np.float32(2.5)) == 3.0 and ti.round(np.float32(-2.5)) == -3.0  # roundf
    assert np.isnan(ti.max(np.float32("nan"), np.float32("nan"))) and ti.max(np.float32("nan"), np.float32(2.0)) == 2.0
    v = ti.Vector([3.0, 4.0, 12.0], ti.f32)
    n = v.normalized()
    inv = np.float32(1.0) / np.sqrt(np.float32(np.float32(np.float32(9.0) + np.float32(16.0)) + np.float32(144.0)))
    assert np.array_equal(n.data, (inv * v.data).astype(np.float32))          # (1 / norm) * v, sums left to right
    assert v.zyx.data.tolist() == [12.0, 4.0, 3.0]
    m = ti.math.mat3(ti.Vector([1.0, 0.0, 0.0], ti.f32), ti.Vector([0.0, 2.0, 0.0], ti.f32), ti.Vector([0.0, 0.0, 3.0], ti.f32))
    assert (m @ v).data.tolist() == [3.0, 8.0, 36.0]                         # vectors given to mat3 are rows


def test_dense_fields_with_offsets_and_struct_for(ti):
    f = ti.Vector.field(3, dtype=ti.u8)
    g = ti.field(dtype=ti.i8)
    ti.root.dense(ti.ijk, 4).place(f, g, offset=(-2, -2, -2))
    g[-2, 1, 0] = 7
    f[ti.Vector([1, 1, -2], ti.i32)] = (1, 2, 300)                            # stored as u8 (300 wraps to 44)
    assert g.arr[0, 3, 2] == 7 and f.arr[3, 3, 0].tolist() == [1, 2, 44]
    assert len(list(ti.grouped(g))) == 64 and list(next(iter(ti.grouped(g))).data) == [-2, -2, -2]
    with pytest.raises(IndexError):
        g[2, 0, 0]


@pytest.mark.skipif(not os.path.isdir("/root/reference/renderer"), reason="needs the reference tree (authoring container only)")
def test_regenerating_vectors_from_the_reference_reproduces_the_committed_files(tmp_path):
    """make_ref_vectors.py math + bsdf + reservoir re-run against /root/reference: identical arrays."""
    gold = os.path.join(ROOT, "tests", "golden")
    before = {n: dict(np.load(os.path.join(gold, n + ".npz"))) for n in ("ref_math", "ref_bsdf", "ref_reservoir")}
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(gold, "make_ref_vectors.py"), "math", "bsdf", "reservoir"], capture_output=True, text=True, env=env,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    for n, old in before.items():
        new = np.load(os.path.join(gold, n + ".npz"))
        assert set(new.files) == set(old)
        for k in old:
            assert np.array_equal(old[k], new[k], equal_nan=old[k].dtype.kind == "f"), (n, k)
