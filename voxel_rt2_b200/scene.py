"""Scene API of voxel-rt2 (scene.py:112-169), kept intact, on top of the B200 renderer.

    Scene(voxel_edges=0.06, exposure=3); set_voxel / get_voxel / set_floor /
    set_directional_light / set_background_color / set_use_physical_sky / set_use_clouds / finish

Differences that are out of scope by design: there is no window, no
interactive camera and no GUI. `finish()` is a headless driver: it uploads the voxels, runs the
start-up precompute, accumulates VRT_SPP samples per pixel, writes the tonemapped image and
returns. Everything the reference hard-codes is an environment override so example scripts stay
unchanged:
    VRT_RES=1920x1080  VRT_GRID=128  VRT_SPP=64  VRT_SKY_RES=3840  VRT_OUT=path.png
    VRT_DEVICE=0  VRT_SEED=0  VRT_BATCH=8 (samples per launch)
Voxels live in host NumPy arrays (material int8[R,R,R], colour uint8[R,R,R,3], index + R/2)
until finish() uploads them once (voxel_world.py:6-25 semantics: colour clamp + u8 truncation,
material cast to int8)."""
import ctypes
import math
import os
import time
from datetime import datetime

import numpy as np

from . import compat

compat.install()

from .renderer import Renderer  # noqa: E402

HELP_MSG = """
====================================================
voxel_rt2_b200 headless renderer (no window):
* VRT_SPP samples per pixel, image written to VRT_OUT or screenshot/
====================================================
"""


def _f32(x):
    """Round a Python number to float32 (Taichi's default_fp)."""
    return ctypes.c_float(x).value


def _u8(c):
    x = _f32(c)
    x = 0.0 if x < 0.0 else (1.0 if x > 1.0 else x)
    return int(_f32(x * 255.0))  # x has 24 significant bits: the double product is exact, then one f32 rounding


def _env_res():
    w, h = os.environ.get("VRT_RES", "1920x1080").lower().split("x")
    return int(w), int(h)


class Camera:
    """Camera state of scene.py:25-109 without the window: position, look-at, up."""

    def __init__(self, up=(0, 1, 0)):
        self._camera_pos = np.array((0.4, 0.5, 2.0))
        self._lookat_pos = np.array((0.0, 0.0, 0.0))
        self._up = np.asarray(up, np.float64) / np.linalg.norm(np.asarray(up, np.float64))

    @property
    def position(self):
        return self._camera_pos

    @property
    def look_at(self):
        return self._lookat_pos


class Scene:
    def __init__(self, voxel_edges=0.06, exposure=3, *, renderer_factory=None):
        self.grid_res = int(os.environ.get("VRT_GRID", "128"))
        self.voxel_dx = 2.0 / self.grid_res  # VOXEL_DX = 1/64 at 128^3 (scene.py:11): world box [-1,1)^3
        self.image_res = _env_res()
        self.voxel_edges = voxel_edges
        self.exposure = exposure
        R = self.grid_res
        self.voxel_material = np.zeros((R, R, R), np.int8)
        self.voxel_color = np.zeros((R, R, R, 3), np.uint8)
        self.camera = Camera()
        # deferred renderer construction: authoring a scene needs no GPU, finish() does
        self._renderer_factory = renderer_factory
        self._renderer = None
        self._floor = (0.0, (1.0, 1.0, 1.0), 1)                 # pathtracer.py:91-93
        self._light = ((1, 1, 1), 0.1, (0.0, 0.0, 0.0))         # scene.py:127
        self._background = (0.0, 0.0, 0.0)
        self._physical_sky = False
        self._clouds = False
        self.last_image = None
        self.last_stats = None

    # ------------------------------------------------------------------ voxel authoring
    @staticmethod
    def round_idx(idx_):  # scene.py:131-137: f32 cast, ti.round, i32
        out = []
        for c in idx_:
            if type(c) is int:
                out.append(c)  # exact in float32 for every index that can address the grid
                continue
            f = _f32(c)
            out.append(int(math.floor(f + 0.5)) if f >= 0 else int(math.ceil(f - 0.5)))
        return out

    def set_voxel(self, idx, mat, color):  # scene.py:139-141 -> pathtracer.py:1325-1328
        i, j, k = self.round_idx(idx)
        h = self.grid_res // 2
        i += h
        j += h
        k += h
        R = self.grid_res
        if not (0 <= i < R and 0 <= j < R and 0 <= k < R):
            return  # the reference writes out of bounds silently; ignored here
        m = int(mat)
        self.voxel_material[i, j, k] = ((m + 128) % 256) - 128  # ti.cast(mat, ti.i8) wraps
        # rgb32f_to_rgb8: clamp, u8(c * 255) truncation in float32 (math_utils.py:86-92)
        r, g, b = color
        self.voxel_color[i, j, k] = (_u8(r), _u8(g), _u8(b))

    def get_voxel(self, idx):  # scene.py:143-146 -> pathtracer.py:1330-1334
        from taichi.math import vec3

        i, j, k = self.round_idx(idx)
        h = self.grid_res // 2
        i += h
        j += h
        k += h
        R = self.grid_res
        if not (0 <= i < R and 0 <= j < R and 0 <= k < R):
            return 0, vec3(0.0)
        c = self.voxel_color[i, j, k]
        return int(self.voxel_material[i, j, k]), vec3(c[0] / 255.0, c[1] / 255.0, c[2] / 255.0)

    # ------------------------------------------------------------------ scene settings
    def set_floor(self, height, color, material=1):
        self._floor = (float(height), tuple(float(x) for x in color), int(material))

    def set_directional_light(self, direction, direction_noise, color):
        self._light = (tuple(float(x) for x in direction), float(direction_noise), tuple(float(x) for x in color))

    def set_background_color(self, color):
        self._background = tuple(float(x) for x in color)

    def set_use_physical_sky(self, use):
        self._physical_sky = bool(use)

    def set_use_clouds(self, use):
        self._clouds = bool(use)

    # ------------------------------------------------------------------ rendering
    @property
    def renderer(self):
        if self._renderer is None:
            factory = self._renderer_factory or Renderer
            self._renderer = factory(
                dx=self.voxel_dx, image_res=self.image_res, up=(0, 1, 0), voxel_edges=self.voxel_edges, exposure=self.exposure,
                grid_res=self.grid_res, sky_res=int(os.environ.get("VRT_SKY_RES", "3840")) if self._physical_sky else 0,
                device=int(os.environ.get("VRT_DEVICE", "0")), seed=int(os.environ.get("VRT_SEED", "0")))
        return self._renderer

    def _configure(self, r):
        r.set_voxels(self.voxel_material, self.voxel_color)
        r.set_floor(*self._floor)
        r.set_directional_light(*self._light)
        r.set_background_color(self._background)
        r.set_use_physical_sky(self._physical_sky, self._clouds)
        r.set_camera_pos(*self.camera.position)
        r.set_look_at(*self.camera.look_at)

    def finish(self, spp=None, out=None):
        """Headless replacement of the frame loop (scene.py:171-297)."""
        print(HELP_MSG)
        spp = int(spp if spp is not None else os.environ.get("VRT_SPP", "64"))
        batch = max(1, int(os.environ.get("VRT_BATCH", "8")))
        r = self.renderer
        self._configure(r)
        t0 = time.time()
        r.prepare_data()
        t_prep = time.time() - t0
        t0 = time.time()
        done = 0
        while done < spp:
            n = min(batch, spp - done)
            r.accumulate(n)
            done += n
        img = r.fetch_image()
        dt = time.time() - t0
        W, H = self.image_res
        print("%d samples took %.3f s (%.3f ms/frame, %.1f Mpaths/s); prepare %.2f s" % (spp, dt, 1e3 * dt / spp, W * H * spp / dt / 1e6, t_prep))
        self.last_image = img
        self.last_stats = {"spp": spp, "seconds": dt, "prepare_seconds": t_prep}
        out = out or os.environ.get("VRT_OUT")
        if out is None:
            import __main__

            os.makedirs("screenshot", exist_ok=True)
            main_filename = os.path.split(getattr(__main__, "__file__", "scene"))[1]
            out = os.path.join("screenshot", "%s-%s.png" % (main_filename, datetime.today().strftime("%Y-%m-%d-%H%M%S")))
        if out:
            save_image(img, out)
            print("Screenshot has been saved to %s" % out)
        return img


def save_image(img, path):
    """float32 [H,W,4] (row 0 = bottom, as the renderer's v axis) -> 8-bit image file."""
    from PIL import Image

    a = (np.clip(img[::-1, :, :3], 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    Image.fromarray(a).save(path)
