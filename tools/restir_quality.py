"""Image quality of the ReSTIR mode on the GPU (BASELINE config 4 scenes, 960x540): per-frame error against a
4096-spp path-traced mean for (a) plain path tracing at 1 spp, (b) the reference's spatial resampling, (c) temporal +
spatial resampling, and the bias of the accumulated image after 32 frames. Prints a markdown table."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import voxel_rt2_b200 as vrt  # noqa: E402

W, H, FRAMES = 960, 540, 32


def make(scene, seed):
    z = np.load(os.path.join(ROOT, "tests", "golden", "%s_seed0.npz" % scene))
    sky = scene == "example6"
    r = vrt.Renderer(dx=1.0 / 64, image_res=(W, H), grid_res=128, sky_res=1024 if sky else 0, cloud_passes=8, seed=seed,
                     voxel_edges=float(z["voxel_edges"]), exposure=float(z["exposure"]), sky_format="f32")
    r.set_voxels(z["material"], z["color"])
    r.set_floor(float(z["floor_height"]), z["floor_color"], int(z["floor_material"]))
    r.set_directional_light(z["light_dir"], float(z["light_noise"]), z["light_color"])
    r.set_background_color(z["background"])
    if sky:
        r.set_use_physical_sky(True, True)
    r.prepare_data()
    return r


print("| scene | estimator | per-frame mean abs error / mean | per-frame rel-RMSE | image mean after %d frames / path-traced mean |" % FRAMES)
print("|---|---|---|---|---|")
for scene in ("example3", "example6"):
    ref = make(scene, 1)
    for _ in range(16):
        ref.accumulate(256)
    m = ref.fetch_hdr()[..., :3]
    geo = (ref.trace_primary()["flags"] & 255) > 0
    mm = m[geo].mean()
    for name in ("path tracing, 1 spp", "spatial resampling (reference)", "temporal + spatial resampling"):
        g = make(scene, 5)
        if name.startswith("temporal"):
            g.set_restir_temporal(True)
        prev, ea, er = np.zeros_like(m), [], []
        for k in range(FRAMES):
            if name.startswith("path"):
                g.accumulate(1)
            else:
                g.accumulate_restir(1)
            cur = g.fetch_hdr()[..., :3] * (k + 1)
            d = ((cur - prev) - m)[geo]
            ea.append(np.abs(d).mean() / mm)
            er.append(np.sqrt((d ** 2).mean()) / mm)
            prev = cur
        print("| %s | %s | %.3f | %.2f | %.4f |" % (scene, name, np.mean(ea[4:]), np.mean(er[4:]), (cur / FRAMES)[geo].mean() / mm))
