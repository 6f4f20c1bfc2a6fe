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
    VRT_MODE=pt|restir|hits   pt: path tracing (default); restir: the USE_RESTIR_PT mode (pathtracer.py:15), one
                              reservoir frame per sample, with temporal reuse unless VRT_RESTIR_TEMPORAL=0;
                              hits: BASELINE config 1, the primary-hit buffer (+ sun shadow bit) written as .npz
    VRT_CHECKPOINT=file.npz   (pt mode, one GPU) save the accumulation sums + sample counter every VRT_CHECKPOINT_EVERY
                              (256) samples and resume from the file if it exists
    VRT_GPUS=N                run the script on N GPUs of this box: it is re-launched under torch.distributed.run
                              (one process per GPU); a script already started by torchrun is recognised by RANK /
                              WORLD_SIZE. VRT_SHARD=tiles (default: interleaved 8x4 tiles of one frame) | rows
                              (contiguous strips; the default in ReSTIR mode: ONE reservoir chain over all GPUs, each
                              rank renders a 24-pixel halo for the spatial pass) | samples (rank r renders sample
                              indices r, r+N, ...; in ReSTIR mode one independent chain per GPU). The partial buffers
                              are merged by the fused peer-memory reduce-scatter + tonemap kernel
                              (parallel.FusedMerge); rank 0 writes the image.
Voxels live in host NumPy arrays (material int8[R,R,R], colour uint8[R,R,R,3], index + R/2)
until finish() uploads them once (voxel_world.py:6-25 semantics: colour clamp + u8 truncation,
material cast to int8)."""
import array
import math
import os
import time
from datetime import datetime

import numpy as np

from . import compat

compat.install()
try:  # the shim's vectorised kernels (compat/taichi/_simd.py); absent when a real Taichi is installed
    from taichi import _simd as _vz
except ImportError:
    _vz = None

from .renderer import Renderer  # noqa: E402

HELP_MSG = """
====================================================
voxel_rt2_b200 headless renderer (no window):
* VRT_SPP samples per pixel, image written to VRT_OUT or screenshot/
====================================================
"""


_F32_CELL = array.array("f", [0.0])


def _f32(x):
    """Round a Python number to float32 (Taichi's default_fp). A one-element float array does the rounding
    (1.5x faster than ctypes.c_float; example9 calls this 14 million times)."""
    _F32_CELL[0] = x
    return _F32_CELL[0]


_U8_CACHE = {}


def _u8(c):
    """rgb32f_to_rgb8 of one channel (math_utils.py:86-92): f32 cast, clamp to [0, 1], u8(c * 255) truncation in
    float32. Scene scripts paint large regions with a handful of constants, so results are memoised (bounded)."""
    r = _U8_CACHE.get(c)
    if r is not None:
        return r
    cell = _F32_CELL
    cell[0] = c
    x = cell[0]
    x = 0.0 if x < 0.0 else (1.0 if x > 1.0 else x)
    cell[0] = x * 255.0  # x has 24 significant bits: the double product is exact, then one f32 rounding
    r = int(cell[0])
    if len(_U8_CACHE) < 65536:
        _U8_CACHE[c] = r
    return r


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


def _relaunch_on_gpus():
    """VRT_GPUS=N outside torchrun: replace this process by `python -m torch.distributed.run ... <script>` so that
    the unchanged example script runs once per GPU (scene authoring is deterministic, every rank builds the same
    scene). Done when the Scene is constructed, before any authoring work."""
    n = int(os.environ.get("VRT_GPUS", "1") or "1")
    if n <= 1 or "RANK" in os.environ or "WORLD_SIZE" in os.environ:
        return
    import socket
    import sys

    import __main__

    script = getattr(__main__, "__file__", None)
    if not script:
        raise RuntimeError("VRT_GPUS needs a script to re-launch (interactive sessions: start it under torchrun yourself)")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    argv = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
            "--master-port", str(port), script] + sys.argv[1:]
    os.execv(sys.executable, argv)


class Scene:
    def __init__(self, voxel_edges=0.06, exposure=3, *, renderer_factory=None):
        if renderer_factory is None:
            _relaunch_on_gpus()
        self.grid_res = int(os.environ.get("VRT_GRID", "128"))
        self.voxel_dx = 2.0 / self.grid_res  # VOXEL_DX = 1/64 at 128^3 (scene.py:11): world box [-1,1)^3
        self.image_res = _env_res()
        self.voxel_edges = voxel_edges
        self.exposure = exposure
        R = self.grid_res
        self.voxel_material = np.zeros((R, R, R), np.int8)
        self.voxel_color = np.zeros((R, R, R, 3), np.uint8)
        # flat byte views of the two arrays for set_voxel / get_voxel: element access through a memoryview is several
        # times cheaper than NumPy scalar indexing, and the scene scripts issue millions of single-voxel writes
        self._mat_mv = memoryview(self.voxel_material).cast("B").cast("b")
        self._col_mv = memoryview(self.voxel_color).cast("B")
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
        i, j, k = idx
        if not (type(i) is int and type(j) is int and type(k) is int):
            i, j, k = self.round_idx((i, j, k))
        R = self.grid_res
        h = R >> 1
        i += h
        j += h
        k += h
        if not (0 <= i < R and 0 <= j < R and 0 <= k < R):
            return  # the reference writes out of bounds silently; ignored here
        o = (i * R + j) * R + k
        r, g, b = color
        if _vz is not None and _vz.PLAIN_LOG is not None:
            # a plain-Python function running lane by lane inside a vectorised loop: the write joins the loop's ordered log
            _vz.PLAIN_LOG.add(self, o, ((int(mat) + 128) & 255) - 128, (_u8(r), _u8(g), _u8(b)))
            return
        self._mat_mv[o] = ((int(mat) + 128) & 255) - 128  # ti.cast(mat, ti.i8) wraps
        # rgb32f_to_rgb8: clamp, u8(c * 255) truncation in float32 (math_utils.py:86-92)
        o *= 3
        cm = self._col_mv
        cm[o] = _u8(r)
        cm[o + 1] = _u8(g)
        cm[o + 2] = _u8(b)

    def _set_voxel_lanes(self, m, idx, mat, color):
        """set_voxel for every lane of mask m of a vectorised loop (compat/taichi/_simd.py): the same rounding, clamping
        and truncation as above on arrays; the writes are logged and applied in iteration order when the loop ends."""
        if m is True:
            return self.set_voxel(idx, mat, color)
        L = _vz.LANES
        sel, o = self._lane_index(m, idx)
        if len(sel) == 0:
            return
        take = lambda x: x[sel] if type(x) is np.ndarray else x  # noqa: E731
        mi = take(mat)
        mi = mi.astype(np.int64) if type(mi) is np.ndarray else int(mi)
        m8 = np.broadcast_to(((mi + 128) & 255) - 128, sel.shape).astype(np.int8)
        rgb = np.empty((len(sel), 3), np.uint8)
        comps = list(color)
        if len(comps) != 3:
            raise ValueError("set_voxel: colour must have 3 components")
        for ch, c in enumerate(comps):
            if type(c) is np.ndarray:
                x = np.clip(c[sel].astype(np.float32), np.float32(0.0), np.float32(1.0))
                rgb[:, ch] = (x * np.float32(255.0)).astype(np.uint8)
            else:
                rgb[:, ch] = _u8(c)
        L.writes.append((self, sel, o, m8, rgb))

    set_voxel.__simd__ = _set_voxel_lanes

    def get_voxel(self, idx):  # scene.py:143-146 -> pathtracer.py:1330-1334
        from taichi.math import vec3

        i, j, k = idx
        if not (type(i) is int and type(j) is int and type(k) is int):
            i, j, k = self.round_idx((i, j, k))
        R = self.grid_res
        h = R >> 1
        i += h
        j += h
        k += h
        if not (0 <= i < R and 0 <= j < R and 0 <= k < R):
            return 0, vec3(0.0)
        o = (i * R + j) * R + k
        if _vz is not None and _vz.PLAIN_LOG is not None:
            _vz.PLAIN_LOG.read(self, o)  # checked against the loop's logged writes when it ends
        cm = self._col_mv
        return self._mat_mv[o], vec3(cm[3 * o] / 255.0, cm[3 * o + 1] / 255.0, cm[3 * o + 2] / 255.0)

    def _lane_index(self, m, idx):
        """round_idx for the lanes of mask m: (lane numbers, flat voxel indices) of those that address the grid."""
        R = self.grid_res
        h = R >> 1
        sel = np.flatnonzero(m)
        inb = np.ones(len(sel), bool)
        ijk = []
        for c in idx:
            if type(c) is np.ndarray:
                c = c[sel]
                if c.dtype.kind == "f":
                    f = c.astype(np.float32).astype(np.float64)
                    c = np.where(f >= 0, np.floor(f + 0.5), np.ceil(f - 0.5)).astype(np.int64)
                else:
                    c = c.astype(np.int64)
            else:
                c = c if type(c) is int else self.round_idx((c,))[0]
            c = c + h
            inb &= (c >= 0) & (c < R)
            ijk.append(c)
        if len(ijk) != 3:
            raise ValueError("voxel index must have 3 components")
        o = np.broadcast_to((ijk[0] * R + ijk[1]) * R + ijk[2], sel.shape)
        return sel[inb], o[inb].astype(np.int64)

    def _get_voxel_lanes(self, m, idx):
        """get_voxel for the lanes of mask m. The voxels are read now, while the loop's own writes are still in its
        log: the read indices are recorded and the loop is re-run sequentially if it also wrote one of them."""
        if m is True:
            return self.get_voxel(idx)
        from taichi.math import Vec

        L = _vz.LANES
        sel, o = self._lane_index(m, idx)
        L.reads.append((self, o))
        mat = np.zeros(L.n, np.int64)
        mat[sel] = self.voxel_material.reshape(-1)[o]
        col = np.zeros((L.n, 3), np.float64)
        col[sel] = self.voxel_color.reshape(-1, 3)[o] / 255.0
        return mat, Vec([col[:, 0].copy(), col[:, 1].copy(), col[:, 2].copy()])

    get_voxel.__simd__ = _get_voxel_lanes

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
                device=int(os.environ.get("LOCAL_RANK", os.environ.get("VRT_DEVICE", "0"))), seed=int(os.environ.get("VRT_SEED", "0")))
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
        """Headless replacement of the frame loop (scene.py:171-297). Returns the tonemapped image (float32 [H, W, 4]),
        or the primary-hit records in VRT_MODE=hits; ranks other than 0 of a multi-GPU run return None."""
        mode = os.environ.get("VRT_MODE", "pt").lower()
        if mode not in ("pt", "restir", "hits"):
            raise ValueError("VRT_MODE must be pt, restir or hits")
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        if rank == 0:
            print(HELP_MSG)
        spp = int(spp if spp is not None else os.environ.get("VRT_SPP", "64"))
        batch = max(1, int(os.environ.get("VRT_BATCH", "8")))
        if world > 1:
            import torch
            import torch.distributed as dist

            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        r = self.renderer
        self._configure(r)
        out = out or os.environ.get("VRT_OUT")
        if out is None:
            import __main__

            os.makedirs("screenshot", exist_ok=True)
            main_filename = os.path.split(getattr(__main__, "__file__", "scene"))[1]
            out = os.path.join("screenshot", "%s-%s.%s" % (main_filename, datetime.today().strftime("%Y-%m-%d-%H%M%S"), "npz" if mode == "hits" else "png"))
        if mode == "hits":  # BASELINE config 1: primary hit + sun shadow ray on the cone axis, no TAA jitter (vrt_trace_primary)
            t0 = time.time()
            r.prepare_data()
            hits = r.trace_primary() if rank == 0 else None
            if rank == 0:
                print("hit buffer %dx%d took %.3f s including the scene build" % (self.image_res[0], self.image_res[1], time.time() - t0))
                if out:
                    np.savez_compressed(out, t=hits["t"], cell=hits["cell"], normal=hits["normal"], flags=hits["flags"])
                    print("Hit buffer has been saved to %s" % out)
            self.last_image = hits
            return hits
        fm = stream_ctx = None
        if world > 1:
            import contextlib

            from . import parallel

            shard = os.environ.get("VRT_SHARD", "rows" if mode == "restir" else "tiles").lower()
            if mode == "restir" and shard == "tiles":
                shard = "rows"  # interleaved tiles cannot carry the 24-pixel neighbourhood of the resampling passes
            if shard == "tiles":
                parallel.shard_tiles(r, rank, world)
                n_local = spp
            elif shard == "rows":
                n_local = spp  # the strips are cut after prepare_data (they are balanced by a primary-hit pass)
            else:
                parallel.shard_samples(r, rank, world)
                n_local = (spp - rank + world - 1) // world  # sample indices rank, rank + world, ... below spp
            r.set_sky_shard(rank, world)  # 1/N of the sky-table rows per GPU + one all-gather per table
            stream = torch.cuda.Stream()
            r.set_stream(stream.cuda_stream)
            stream_ctx = torch.cuda.stream(stream)
        else:
            n_local = spp
        t0 = time.time()
        r.prepare_data()
        if world > 1 and shard == "rows":
            parallel.shard_rows(r, rank, world)
        t_prep = time.time() - t0
        if mode == "restir":
            r.set_restir_temporal(os.environ.get("VRT_RESTIR_TEMPORAL", "1") != "0")
        t0 = time.time()
        with (stream_ctx if stream_ctx is not None else _null_context()):
            if world > 1:
                fm = parallel.FusedMerge(r)
                fm.begin(0)
            done = 0
            ckpt = os.environ.get("VRT_CHECKPOINT") if (world == 1 and mode == "pt") else None
            every = max(1, int(os.environ.get("VRT_CHECKPOINT_EVERY", "256")))
            if ckpt and os.path.exists(ckpt):  # resume: the accumulation sums and the sample counter of an earlier run
                z = np.load(ckpt)
                if tuple(z["sums"].shape) == (self.image_res[1], self.image_res[0], 4) and int(z["spp"]) <= n_local:
                    r.set_accumulation(z["sums"], int(z["spp"]))
                    done = int(z["spp"])
                    print("resumed from %s at %d samples" % (ckpt, done))
            last_saved = done
            while done < n_local:
                n = min(batch, n_local - done)
                if mode == "restir":
                    r.accumulate_restir(n)
                else:
                    r.accumulate(n)
                done += n
                if ckpt and (done - last_saved >= every or done == n_local):
                    sums, _ = r.get_accumulation()
                    np.savez(ckpt + ".tmp.npz", sums=sums, spp=np.int64(done))
                    os.replace(ckpt + ".tmp.npz", ckpt)
                    last_saved = done
            if fm is not None:
                import torch

                fm.merge()
                host = torch.empty((self.image_res[1], self.image_res[0], 4), dtype=torch.float32, pin_memory=True).numpy() if rank == 0 else None
                fm.finish(host)
                img = host
            else:
                img = r.fetch_image()
        dt = time.time() - t0
        if fm is not None:
            fm.close()
        W, H = self.image_res
        self.last_stats = {"spp": spp, "seconds": dt, "prepare_seconds": t_prep, "mode": mode, "gpus": world}
        if rank != 0:
            return None
        print("%d samples took %.3f s (%.3f ms/frame, %.1f Mpaths/s)%s; prepare %.2f s" % (
            spp, dt, 1e3 * dt / spp, W * H * spp / dt / 1e6, " on %d GPUs" % world if world > 1 else "", t_prep))
        self.last_image = img
        if out:
            save_image(img, out)
            print("Screenshot has been saved to %s" % out)
        return img


class _null_context:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def save_image(img, path):
    """float32 [H,W,4] (row 0 = bottom, as the renderer's v axis) -> 8-bit image file."""
    from PIL import Image

    a = (np.clip(img[::-1, :, :3], 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    Image.fromarray(a).save(path)
