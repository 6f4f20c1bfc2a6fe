import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, scenes
import voxel_rt2_b200 as vrt
R = 32
g = vrt.Renderer(dx=2.0 / R, image_res=(64, 32), grid_res=R, sky_res=32, cloud_passes=1, seed=4)
g.set_voxels(*scenes.random_grid(R, 0.3, 9, materials=(1, 2, 11, 50, 54)))
g.set_floor(-0.6, (0.8, 0.8, 0.8))
g.set_directional_light((1, 1, 0.5), 0.05, (1.2, 1.1, 1.0))
g.set_use_physical_sky(True, True)
g.prepare_data()
g.trace_primary()
g.accumulate(3)
g.reset_framebuffer(); g.accumulate(2)
g.set_restir_temporal(True); g.accumulate_restir(3)
g.set_accum_slot(1); g.reset_framebuffer(); g.accumulate(1); g.merge_slice([], 0, 64 * 32)
g.accumulate_moving(0.5, 50.0)
img = g.fetch_image()
print("sanitizer probe ok", float(img[..., :3].mean()))
