"""Helpers shared by the parity tests."""
import numpy as np


def make_pair(vrt, oracle, *, image_res, grid_res, dx=None, sky_res=0, cloud_passes=2, seed=7, jitter=True, max_depth=4,
              voxel_edges=0.06, exposure=3.0, sky_format="f32"):
    """A CUDA Renderer and an OracleRenderer with identical configuration."""
    from voxel_rt2_b200.materials import material_table
    import os

    dx = dx if dx is not None else 2.0 / grid_res
    kw = dict(dx=dx, image_res=image_res, voxel_edges=voxel_edges, exposure=exposure, grid_res=grid_res, max_depth=max_depth,
              sky_res=sky_res, cloud_passes=cloud_passes, seed=seed, jitter=jitter)
    g = vrt.Renderer(sky_format=sky_format, **kw)  # float sky tables: the oracle reads float tables (f16 has its own budget test)
    tex = np.load(os.path.join(os.path.dirname(vrt.__file__), "assets", "cloud_texture.npz"))["tex"]
    o = oracle.OracleRenderer(materials=material_table(), cloud_tex=tex, **kw)
    return g, o


def apply_both(objs, name, *args, **kw):
    for x in objs:
        getattr(x, name)(*args, **kw)


def rel_rmse(a, b):
    """sqrt(mean((a-b)^2)) / mean(b) on the rgb channels."""
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-12))


def renderer_from_reference_fixture(factory, z, **kw):
    """A Renderer / OracleRenderer configured like the reference Renderer that produced
    tests/golden/ref_render*.npz (tests/golden/make_ref_vectors.py section_render)."""
    W, H = int(z["W"]), int(z["H"])
    R = z["material"].shape[0]
    sky_res = int(z["sky_res"]) if "sky_res" in z else 0
    if factory.__name__ == "Renderer":
        kw.setdefault("sky_format", "f32")  # reference vectors are held to the float-table path
    r = factory(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=sky_res, seed=int(z["seed"]), jitter=False,
                voxel_edges=float(z["cfg_voxel_edges"]), **kw)
    r.set_voxels(z["material"], z["color"])
    r.set_floor(float(z["cfg_floor_height"]), z["cfg_floor_color"], int(z["cfg_floor_material"]))
    r.set_directional_light(z["cfg_light_dir"], float(z["cfg_light_cone"]), z["cfg_light_color"])
    r.set_background_color(z["cfg_background"])
    r.set_view_proj(z["cam_pos"], z["view"], z["proj"])
    return r


def reference_hit_fields(h):
    """trace_primary() records -> the arrays the reference-derived fixture stores."""
    kind = h["flags"] & 255
    t = np.where(kind > 0, h["t"], np.inf).astype(np.float32)
    return dict(t=t, normal=h["normal"] + 0.0, mat=(h["flags"] >> 16) & 255, light=(h["flags"] >> 24) & 255, shadow=(h["flags"] >> 8) & 255)


def reference_radiance(z, s):
    """Sample s of the reference render(): diffuse + specular with the NaN / negative scrub the
    reference applies in its temporal filter (pathtracer.py:1068-1075)."""
    def scrub(c):
        bad = ~np.isfinite(c).all(-1) | (c < 0).any(-1)
        c = c.copy()
        c[bad] = 0
        return c
    return scrub(z["render_diffuse"][s]) + scrub(z["render_specular"][s])


def example1_renderer(factory, **kw):
    """The example1.py scene (shim seed 0) in a Renderer configured as the reference ships it
    (128^3, dx = 1/64, default camera, fov 50 deg) at 64 x 64: the setup of
    tests/golden/ref_example1_hits_64.npz."""
    import os

    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z, h = np.load(os.path.join(g, "example1_seed0.npz")), np.load(os.path.join(g, "ref_example1_hits_64.npz"))
    r = factory(dx=1.0 / 64.0, image_res=(int(h["W"]), int(h["H"])), grid_res=128, sky_res=0, jitter=False, voxel_edges=float(z["voxel_edges"]), **kw)
    r.set_voxels(z["material"], z["color"])
    r.set_floor(float(z["floor_height"]), z["floor_color"], int(z["floor_material"]))
    r.set_directional_light(z["light_dir"], float(z["light_noise"]), z["light_color"])
    r.set_view_proj(h["cam_pos"], h["view"], h["proj"])
    r.prepare_data()
    return r, h


def assert_hits_equal_reference(hits, h):
    got = reference_hit_fields(hits)
    assert np.array_equal(got["t"].view(np.uint32), h["hit_t"].view(np.uint32))
    hit = np.isfinite(h["hit_t"])
    assert np.array_equal(got["normal"][hit], h["hit_normal"][hit] + 0.0)
    assert np.array_equal(got["mat"][hit], h["hit_mat"][hit])
    assert np.array_equal(got["light"][hit], h["hit_light"][hit])
    assert np.array_equal(got["shadow"], h["hit_shadow"])
    return hit
