"""Quick device-time probe of the path kernel on the config-3 workload (dense random 256^3)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import scenes  # noqa: E402
import voxel_rt2_b200 as vrt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--R", type=int, default=256)
ap.add_argument("--res", default="1920x1080")
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--sky", type=int, default=0)
ap.add_argument("--sky-res", type=int, default=3840)
ap.add_argument("--scene", default="dense")
ap.add_argument("--restir", type=int, default=0, help="also time N ReSTIR frames")
ap.add_argument("--temporal", type=int, default=0, help="ReSTIR: temporal reservoir reuse before the spatial pass")
a = ap.parse_args()
W, H = [int(x) for x in a.res.split("x")]
t0 = time.time()
r = vrt.Renderer(dx=2.0 / a.R, image_res=(W, H), grid_res=a.R, sky_res=a.sky_res if a.sky else 0, exposure=2.0, seed=1)
if a.scene == "dense":
    mat, col = scenes.random_grid(a.R, 0.5, 1234)
    r.set_floor(-1e5, (1, 1, 1))
elif a.scene == "example3":
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "example3_seed0.npz"))
    mat, col = z["material"], z["color"]
    r.set_floor(0.0, (1, 1, 1))
elif a.scene == "example6":
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "example6_seed0.npz"))
    mat, col = z["material"], z["color"]
    r.set_floor(float(z["floor_height"]), z["floor_color"], int(z["floor_material"]))
else:
    mat, col = scenes.city(a.R, 0, 50)
    r.set_floor(-0.05, (1, 1, 1))
r.set_voxels(mat, col)
if a.scene == "example3":
    r.set_directional_light((1, 1, 1), 0.1, (0.0, 0.0, 0.0))
else:
    r.set_directional_light((1, 1, -1) if a.scene == "example6" else (1, 1, 1), 0.025, (1.3, 0.949 * 1.3, 0.937 * 1.3))
r.set_background_color((0.3, 0.4, 0.6))
if a.sky:
    r.set_use_physical_sky(True, True)
t1 = time.time()
r.prepare_data()
t2 = time.time()
print("setup %.2fs prepare %.2fs sky_ms %.1f" % (t1 - t0, t2 - t1, r.stats()["sky_precompute_ms"]))
r.accumulate(1, stats=True)
s = r.stats()
print("stats/path:", {k: round(s[k] / max(s["paths"], 1), 3) for k in ("rays", "steps", "queries", "hits", "sky_escapes", "nee_visible", "vertices")})
for spp in (1, a.spp):
    best = 1e9
    for i in range(a.iters):
        r.accumulate(spp)
        best = min(best, r.stats()["last_render_ms"])
    print("spp/launch=%d: %.3f ms/launch, %.3f ms/frame, %.3f Gpaths/s" % (spp, best, best / spp, W * H * spp / best / 1e6))
if a.restir:
    r.reset_framebuffer()
    r.set_restir_temporal(bool(a.temporal))
    r.accumulate_restir(2)
    best = (1e9, 0, 0, 0)
    for i in range(a.iters):
        r.accumulate_restir(a.restir)
        st = r.stats()
        tot = (st["last_render_ms"] + st["last_gris_ms"] + st["last_temporal_ms"]) / a.restir
        best = min(best, (tot, st["last_render_ms"] / a.restir, st["last_temporal_ms"] / a.restir, st["last_gris_ms"] / a.restir))
    print("restir%s: %.3f ms/frame (path+reservoir %.3f, temporal %.3f, gris %.3f), %.3f Gpaths/s" % (" temporal" if a.temporal else "", best[0], best[1], best[2], best[3], W * H / best[0] / 1e6))
img = r.fetch_image()
print("resolve ms", r.stats()["last_resolve_ms"], "mean ldr", img[..., :3].mean())
os.makedirs("gpurun_out", exist_ok=True)
try:
    from PIL import Image
    Image.fromarray((np.clip(img[::-1, :, :3], 0, 1) * 255).astype(np.uint8)).save("gpurun_out/probe_%s_%d%s.png" % (a.scene, a.sky, "_restir" if a.restir else ""))
except Exception as e:
    print("no png:", e)
