"""Host-side logic that needs no GPU: camera matrices, the Scene API surface, the taichi shim."""
import math
import os
import runpy

import numpy as np
import pytest


def test_camera_matrices_glm_conventions(vrt):
    """glm lookAt (RH) and perspective (GL -1..1 depth)."""
    pos, view, proj = vrt.default_camera_matrices(1920, 1080)
    assert pos.tolist() == [np.float32(0.4), np.float32(0.5), np.float32(2.0)]
    v = view.astype(np.float64)
    # the eye maps to the origin, the target onto the -z axis at its distance
    assert np.allclose(v @ np.array([0.4, 0.5, 2.0, 1.0]), [0, 0, 0, 1], atol=1e-6)
    d = math.sqrt(0.4 ** 2 + 0.5 ** 2 + 2.0 ** 2)
    assert np.allclose(v @ np.array([0, 0, 0, 1.0]), [0, 0, -d, 1], atol=1e-6)
    assert np.allclose(v[:3, :3] @ v[:3, :3].T, np.eye(3), atol=1e-6)
    p = proj.astype(np.float64)
    g = 1.0 / math.tan(math.radians(50.0) / 2)
    assert np.isclose(p[1, 1], g) and np.isclose(p[0, 0], g / (1920 / 1080))
    near = p @ np.array([0, 0, -0.01, 1.0])
    far = p @ np.array([0, 0, -10.0, 1.0])
    assert np.isclose(near[2] / near[3], -1.0) and np.isclose(far[2] / far[3], 1.0)


def test_scene_api_surface_matches_reference():
    """scene.py:112-169: names, positional order and defaults of the public methods."""
    import inspect

    from voxel_rt2_b200.scene import Scene

    sig = inspect.signature(Scene.__init__)
    assert list(sig.parameters)[:3] == ["self", "voxel_edges", "exposure"]
    assert sig.parameters["voxel_edges"].default == 0.06 and sig.parameters["exposure"].default == 3
    assert list(inspect.signature(Scene.set_floor).parameters) == ["self", "height", "color", "material"]
    assert inspect.signature(Scene.set_floor).parameters["material"].default == 1
    assert list(inspect.signature(Scene.set_directional_light).parameters) == ["self", "direction", "direction_noise", "color"]
    for name in ("set_voxel", "get_voxel", "set_background_color", "set_use_physical_sky", "set_use_clouds", "finish", "round_idx"):
        assert hasattr(Scene, name)


def test_set_get_voxel_semantics():
    """round_idx (scene.py:131-137), colour clamp + u8 truncation (math_utils.py:86-100),
    i8 material wrap (pathtracer.py:1327), index offset R/2 (voxel_world.py:14-18)."""
    from voxel_rt2_b200.scene import Scene

    s = Scene()
    from taichi.math import vec3

    s.set_voxel(vec3(0, 0, 0), 2, vec3(0.9, 0.1, 0.1))
    assert s.voxel_material[64, 64, 64] == 2
    assert s.voxel_color[64, 64, 64].tolist() == [229, 25, 25]
    s.set_voxel(vec3(1.5, -2.5, 0.49), 1, vec3(1.7, -0.3, 0.5))  # rounds half away from zero; colour clamps
    assert s.voxel_material[66, 61, 64] == 1
    assert s.voxel_color[66, 61, 64].tolist() == [255, 0, 127]
    mat, col = s.get_voxel(vec3(2, -3, 0))
    assert mat == 1 and np.allclose(list(col), [1.0, 0.0, 127 / 255.0])
    s.set_voxel(vec3(-64, -64, -64), 200, vec3(1, 1, 1))          # i8 wrap: 200 -> -56 (= empty)
    assert s.voxel_material[0, 0, 0] == -56
    s.set_voxel(vec3(64, 0, 0), 1, vec3(1, 1, 1))                  # outside [-64, 64): ignored
    assert int((s.voxel_material != 0).sum()) == 3


def test_finish_drives_the_renderer_in_reference_order(tmp_path):
    from voxel_rt2_b200.scene import Scene

    calls = []

    class Rec:
        def __init__(self, **kw):
            calls.append(("init", kw["image_res"], kw["grid_res"], kw["sky_res"]))

        def __getattr__(self, n):
            def f(*a, **k):
                calls.append((n,) + tuple(x for x in a if not isinstance(x, np.ndarray)))
                if n == "fetch_image":
                    return np.zeros((8, 8, 4), np.float32)

            return f

    s = Scene(voxel_edges=0, exposure=2, renderer_factory=Rec)
    s.set_floor(-0.85, (1.0, 1.0, 1.0))
    s.set_directional_light((1, 1, -1), 0.025, (1.3, 1.2, 1.2))
    s.set_use_physical_sky(True)
    s.set_use_clouds(True)
    out = tmp_path / "o.png"
    s.finish(spp=10, out=str(out))
    names = [c[0] for c in calls]
    assert names[0] == "init" and calls[0][3] > 0
    assert names.index("set_voxels") < names.index("prepare_data") < names.index("accumulate") < names.index("fetch_image")
    assert ("set_use_physical_sky", True, True) in calls
    assert sum(c[1] for c in calls if c[0] == "accumulate") == 10
    assert out.exists()


def test_taichi_shim_vector_semantics():
    from voxel_rt2_b200 import compat

    compat.install()
    import taichi as ti
    from taichi.math import int as tint, ivec3, mix, vec2, vec3, vec4

    v = vec3(0.7)
    assert list(v) == [0.7, 0.7, 0.7]
    assert list(ivec3(3, 65 - 35 * 0.5, 3)) == [3, 47, 3]          # example6.py:45: float -> int truncation
    d = tint(vec4(1, 0, 1, 0))
    assert list(d.yzwx) == [0, 1, 0, 1] and d.sum() == 2            # example7.py:34
    assert ((d.x | d.z) ^ (d.y | d.w)) & 1 == 1
    assert list((ivec3(5, -3, 7) + 60) // 15) == [4, 3, 4] and list(ivec3(5, -3, 7) % 4) == [1, 1, 3]
    assert list(mix(vec3(1, 2, 3), vec3(3, 2, 1), 0.5)) == [2.0, 2.0, 2.0]
    assert mix(10, 0, True) == 0 and mix(10, 0, False) == 10        # bool t (example7.py:52)
    assert list(tint(vec2(0.9, 3.9) * 2)) == [1, 7]
    assert any((abs(vec2(3, 5) - vec2(4, 9)) == 1) | (abs(vec2(3, 5) - vec2(7, 7)) == 1))
    assert ti.max(3, 7) == 7 and ti.min(vec2(1, 8), vec2(4, 2)).v == [1.0, 2.0]
    assert [tuple(I) for I in ti.grouped(ti.ndrange((1, 3), 2))] == [(1, 0), (1, 1), (2, 0), (2, 1)]
    assert list(ti.ndrange((2, 5))) == [2, 3, 4]                    # 1-D ndrange yields scalars (example5.py:38)
    assert ti.round(2.5) == 3.0 and ti.round(-2.5) == -3.0
    ti.seed(5)
    a = [ti.random() for _ in range(3)]
    ti.seed(5)
    assert a == [ti.random() for _ in range(3)] and all(0 <= x < 1 for x in a)
    w = vec3(1, 2, 3)
    w[0] *= -1
    w.y = 9
    assert list(w) == [-1.0, 9, 3.0] and list(w.zx) == [3.0, -1.0]


@pytest.mark.parametrize("example,occupied", [("main", 1), ("example1", 3539), ("example3", 13168)])
def test_reference_examples_run_unchanged_through_the_shim(example, occupied):
    """The scene scripts of the reference execute unmodified (authoring only: the renderer is a
    stub because this container has no GPU). Needs the mounted reference tree."""
    path = "/root/reference/%s.py" % example
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    import voxel_rt2_b200.scene as S

    class Stub:
        def __init__(self, **kw):
            pass

        def __getattr__(self, n):
            return lambda *a, **k: np.zeros((4, 4, 4), np.float32) if n == "fetch_image" else None

    orig, save = S.Scene.__init__, S.save_image

    def patched(self, *a, **k):
        k["renderer_factory"] = Stub
        orig(self, *a, **k)

    S.Scene.__init__ = patched
    S.save_image = lambda img, p: None
    try:
        import taichi

        taichi.seed(0)
        g = runpy.run_path(path, run_name="__main__")
    finally:
        S.Scene.__init__, S.save_image = orig, save
    assert int((g["scene"].voxel_material > 0).sum()) == occupied


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference (CPU oracle arm) on a tiny configuration: stdout is exactly one JSON
    line carrying the keys the driver reads."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: the CPU arm must still use all host cores
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--res", "64x32",
                        "--grid", "32", "--sky-res", "16"], capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["value"] > 0 and j["cpu_baseline"]["kind"] == "port"
    assert j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0


def test_finish_modes_dispatch_through_the_renderer_surface(tmp_path, monkeypatch):
    """VRT_MODE / VRT_CHECKPOINT of Scene.finish with a recording stub renderer (no GPU): `hits` calls trace_primary and
    writes an .npz, `restir` switches temporal reuse on and calls accumulate_restir, a checkpoint is written and a second
    run resumes from it instead of starting over."""
    from voxel_rt2_b200.renderer import HIT_DTYPE
    from voxel_rt2_b200.scene import Scene

    calls = []

    class Rec:
        def __init__(self, **kw):
            self.spp = 0

        def get_accumulation(self):
            return np.full((8, 8, 4), float(self.spp), np.float32), self.spp

        def set_accumulation(self, sums, spp):
            calls.append(("set_accumulation", int(spp)))
            self.spp = int(spp)

        def accumulate(self, n):
            calls.append(("accumulate", n))
            self.spp += n

        def __getattr__(self, n):
            def f(*a, **k):
                calls.append((n,) + tuple(x for x in a if not isinstance(x, np.ndarray)))
                if n == "fetch_image":
                    return np.zeros((8, 8, 4), np.float32)
                if n == "trace_primary":
                    return np.zeros((8, 8), HIT_DTYPE)

            return f

    monkeypatch.setenv("VRT_RES", "8x8")
    monkeypatch.setenv("VRT_MODE", "hits")
    s = Scene(renderer_factory=Rec)
    h = s.finish(out=str(tmp_path / "h.npz"))
    assert h.dtype == HIT_DTYPE and "trace_primary" in [c[0] for c in calls] and "accumulate" not in [c[0] for c in calls]
    assert set(np.load(tmp_path / "h.npz").files) == {"t", "cell", "normal", "flags"}
    calls.clear()
    monkeypatch.setenv("VRT_MODE", "restir")
    Scene(renderer_factory=Rec).finish(spp=5, out=str(tmp_path / "r.png"))
    assert ("set_restir_temporal", True) in calls and sum(c[1] for c in calls if c[0] == "accumulate_restir") == 5
    calls.clear()
    monkeypatch.setenv("VRT_MODE", "pt")
    monkeypatch.setenv("VRT_CHECKPOINT", str(tmp_path / "ck.npz"))
    monkeypatch.setenv("VRT_CHECKPOINT_EVERY", "8")
    Scene(renderer_factory=Rec).finish(spp=16, out=str(tmp_path / "a.png"))
    z = np.load(tmp_path / "ck.npz")
    assert int(z["spp"]) == 16 and float(z["sums"][0, 0, 0]) == 16.0
    calls.clear()
    Scene(renderer_factory=Rec).finish(spp=24, out=str(tmp_path / "b.png"))   # resumes at 16, renders 8 more
    assert ("set_accumulation", 16) in calls and sum(c[1] for c in calls if c[0] == "accumulate") == 8
    with pytest.raises(ValueError):
        monkeypatch.setenv("VRT_MODE", "nope")
        Scene(renderer_factory=Rec).finish(spp=1)
