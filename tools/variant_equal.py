"""Render three small scenes with the library selected by VRT_LIB and write the HDR images to an .npz; with --compare A B
check two such files for bit-identity (kernel variants that only change scheduling must not change a single bit)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    ok = True
    for k in a.files:
        same = np.array_equal(a[k], b[k])
        ok &= same
        print(k, "identical" if same else "DIFFERENT: max abs %.3e, %d values" % (np.abs(a[k] - b[k]).max(), int((a[k] != b[k]).sum())))
    sys.exit(0 if ok else 1)
import scenes  # noqa: E402
import voxel_rt2_b200 as vrt  # noqa: E402

out = {}
for name, R, scene, sky in (("dense", 128, scenes.random_grid(128, 0.5, 1234), True), ("zoo", 64, scenes.material_zoo(64), False),
                            ("city", 128, scenes.city(128, 0, 50), False)):
    r = vrt.Renderer(dx=2.0 / R, image_res=(640, 360), grid_res=R, sky_res=256 if sky else 0, cloud_passes=2, seed=3)
    r.set_voxels(*scene)
    r.set_floor(-1e5 if name == "dense" else -0.05, (1, 1, 1))
    r.set_directional_light((1, 1, 1), 0.025, (1.3, 1.2, 1.2))
    r.set_background_color((0.3, 0.4, 0.6))
    if sky:
        r.set_use_physical_sky(True, True)
    r.prepare_data()
    r.accumulate(8)
    r.accumulate(3)
    out[name] = r.fetch_hdr()
    r.accumulate_restir(2)
    out[name + "_restir"] = r.fetch_hdr()
np.savez(sys.argv[1], **out)
print("wrote", sys.argv[1])
