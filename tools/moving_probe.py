"""Device time of the moving-camera frame (accumulate() with camera_is_moving = 1, scene.py:214-228):
half-resolution path kernel with albedo demodulation + reprojecting temporal filters + upsample, example6 scene with the
physical sky at 1920x1080, the camera translating a little every frame."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voxel_rt2_b200 as vrt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=40)
ap.add_argument("--scale", type=float, default=0.5)
ap.add_argument("--sky-res", type=int, default=3840)
a = ap.parse_args()
W, H = 1920, 1080
z = np.load(os.path.join(ROOT, "tests", "golden", "example6_seed0.npz"))
R = z["material"].shape[0]
r = vrt.Renderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=a.sky_res, exposure=2.0, seed=1)
r.set_voxels(z["material"], z["color"])
r.set_floor(float(z["floor_height"]), z["floor_color"], int(z["floor_material"]))
r.set_directional_light((1, 1, -1), 0.025, (1.3, 0.949 * 1.3, 0.937 * 1.3))
r.set_background_color((0.3, 0.4, 0.6))
r.set_use_physical_sky(True, True)
r.prepare_data()
ms, t0 = [], None
for k in range(a.frames + 5):
    if k == 5:
        r.synchronize()
        t0 = time.perf_counter()
    r.set_camera_pos(0.4 + 0.004 * k, 0.5, 2.0 - 0.002 * k)
    r.accumulate_moving(a.scale, 50.0)
    if k >= 5:
        ms.append(r.stats()["last_render_ms"])
r.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / a.frames
img = r.fetch_image()
print("moving camera, render_scale %.2f: %.3f ms/frame device (path kernel + filters; min %.3f), %.3f ms/frame wall incl. history copies; mean ldr %.4f"
      % (a.scale, float(np.mean(ms)), float(np.min(ms)), wall, float(img[..., :3].mean())))
