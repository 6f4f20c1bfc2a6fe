"""Generates the committed fixtures under tests/golden/. Run in the authoring container (needs
/root/reference for the example scripts; they are executed unchanged through the taichi shim):

    python tests/golden/make_golden.py

The reference itself holds no golden vectors and cannot run without Taichi, so
these pin OUR restatement: scenes come from the shim's seeded RNG, hit buffers and radiance means
from the CPU oracle. GPU tests compare the CUDA path against them; CPU tests re-check the oracle
against them so the oracle cannot drift silently."""
import hashlib
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def example_scene(name, seed=0):
    """Run example<name>.py unchanged with a stub renderer and return its Scene settings."""
    import voxel_rt2_b200.scene as S

    class Stub:
        def __init__(self, **kw):
            pass

        def __getattr__(self, n):
            return lambda *a, **k: np.zeros((4, 4, 4), np.float32) if n == "fetch_image" else None

    orig = S.Scene.__init__

    def patched(self, *a, **k):
        k["renderer_factory"] = Stub
        orig(self, *a, **k)

    S.Scene.__init__ = patched
    save = S.save_image
    S.save_image = lambda img, path: None
    try:
        import taichi
        from taichi import _simd

        # the committed <example>_seed0.npz scenes come from the shim's first, sequential RNG: keep them reproducible
        # (tests/test_host.py::test_legacy_rng_reproduces_the_committed_fixture_scenes holds this mode to the files)
        taichi._LEGACY_RNG, _simd.ENABLED = True, False
        taichi.seed(seed)
        cwd = os.getcwd()
        os.chdir("/tmp")
        g = runpy.run_path(os.path.join(REF, name + ".py"), run_name="__main__")
        os.chdir(cwd)
    finally:
        S.Scene.__init__ = orig
        S.save_image = save
    sc = g["scene"]
    return dict(material=sc.voxel_material.copy(), color=sc.voxel_color.copy(), floor_height=np.float32(sc._floor[0]),
                floor_color=np.asarray(sc._floor[1], np.float32), floor_material=np.int32(sc._floor[2]),
                light_dir=np.asarray(sc._light[0], np.float32), light_noise=np.float32(sc._light[1]),
                light_color=np.asarray(sc._light[2], np.float32), background=np.asarray(sc._background, np.float32),
                physical_sky=np.int32(sc._physical_sky), clouds=np.int32(sc._clouds), voxel_edges=np.float32(sc.voxel_edges),
                exposure=np.float32(sc.exposure))


def oracle_for(sc, res, **kw):
    from oracle.binding import OracleRenderer
    from voxel_rt2_b200.materials import material_table

    R = sc["material"].shape[0]
    o = OracleRenderer(dx=2.0 / R, image_res=res, grid_res=R, voxel_edges=float(sc["voxel_edges"]), exposure=float(sc["exposure"]),
                       materials=material_table(), **kw)
    o.set_voxels(sc["material"], sc["color"])
    o.set_floor(float(sc["floor_height"]), sc["floor_color"], int(sc["floor_material"]))
    o.set_directional_light(sc["light_dir"], float(sc["light_noise"]), sc["light_color"])
    o.set_background_color(sc["background"])
    return o


def pack_hits(h):
    return dict(t_bits=h["t"].view(np.uint32), cell=h["cell"].astype(np.int16), normal=h["normal"].astype(np.int8), flags=h["flags"])


def main():
    out = {}
    for name in ("example1", "example6"):
        sc = example_scene(name, 0)
        np.savez_compressed(os.path.join(HERE, name + "_seed0.npz"), **sc)
        out[name] = sc
        print(name, "occupied", int((sc["material"] > 0).sum()))
    # config 1: example1, 640x640, primary rays + sun shadow ray
    o = oracle_for(out["example1"], (640, 640), sky_res=0, jitter=False)
    o.prepare_data()
    h = o.trace_primary()
    digest = hashlib.sha256(h.tobytes()).hexdigest()
    np.savez_compressed(os.path.join(HERE, "hits_example1_640.npz"), sha256=np.array(digest), **pack_hits(h))
    print("hits_example1_640 sha256", digest)
    # main.py scene at 256x256: tiny KAT buffer
    import scenes

    from oracle.binding import OracleRenderer
    from voxel_rt2_b200.materials import material_table

    o = OracleRenderer(dx=1 / 64, image_res=(256, 256), grid_res=128, sky_res=0, jitter=False, exposure=10, materials=material_table())
    o.set_voxels(*scenes.main_scene())
    o.set_floor(-0.05, (1, 1, 1))
    o.set_background_color((1.0, 0, 0))
    o.prepare_data()
    h = o.trace_primary()
    np.savez_compressed(os.path.join(HERE, "hits_main_256.npz"), sha256=np.array(hashlib.sha256(h.tobytes()).hexdigest()), **pack_hits(h))
    # converged radiance means: example1 (emissive voxels only, black sun) 96x64 crop-size image, 1024 spp
    o = oracle_for(out["example1"], (96, 64), sky_res=0, jitter=True, seed=11)
    o.prepare_data()
    o.accumulate(1024)
    np.savez_compressed(os.path.join(HERE, "radiance_example1_96x64_1024spp.npz"), hdr=o.fetch_hdr().astype(np.float32), seed=np.int32(11))
    # example6 with a background-colour sky and its sun (physical sky tables are GPU-computed; they are
    # covered by the table parity test instead)
    sc = dict(out["example6"])
    o = oracle_for(sc, (96, 64), sky_res=0, jitter=True, seed=12)
    o.set_background_color((0.5, 0.6, 0.8))
    o.prepare_data()
    o.accumulate(1024)
    np.savez_compressed(os.path.join(HERE, "radiance_example6_bg_96x64_1024spp.npz"), hdr=o.fetch_hdr().astype(np.float32), seed=np.int32(12))
    # example6 as authored (physical sky + clouds), sky tables precomputed BY THE ORACLE at 64^2 with
    # 4 cloud passes: end-to-end pin of precompute + lookups + path tracing
    import os as _os

    tex = np.load(_os.path.join(ROOT, "voxel_rt2_b200", "assets", "cloud_texture.npz"))["tex"]
    o = oracle_for(sc, (96, 64), sky_res=64, cloud_passes=4, jitter=True, seed=13, cloud_tex=tex)
    o.set_use_physical_sky(True, True)
    o.prepare_data()
    o.accumulate(1024)
    np.savez_compressed(os.path.join(HERE, "radiance_example6_sky64_96x64_1024spp.npz"), hdr=o.fetch_hdr().astype(np.float32), seed=np.int32(13))
    print("done")


if __name__ == "__main__":
    main()
