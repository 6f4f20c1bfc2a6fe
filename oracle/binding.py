"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. `OracleRenderer` mirrors voxel_rt2_b200.Renderer's surface so parity tests
call both sides symmetrically."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
HIT_DTYPE = np.dtype([("t", "<f4"), ("cell", "<i4", (3,)), ("normal", "<f4", (3,)), ("flags", "<u4")])

_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("oracle.cpp", "ovec.h", "otrace.h", "obsdf.h", "osky.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        r = subprocess.run(["make", "-C", HERE, "-B" if force else "-s", "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.orc_create.restype = P
    lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_uint32, C.c_int, C.c_int]
    lib.orc_destroy.argtypes = [P]
    lib.orc_upload_voxels.argtypes = [P, P, P]
    lib.orc_set_camera.argtypes = [P, fp, fp, fp]
    lib.orc_set_camera.restype = C.c_int
    lib.orc_set_light.argtypes = [P, fp, C.c_float, fp]
    lib.orc_set_floor.argtypes = [P, C.c_float, fp, C.c_int]
    lib.orc_set_background.argtypes = [P, fp]
    lib.orc_set_sky.argtypes = [P, C.c_int, C.c_int]
    lib.orc_set_materials.argtypes = [P, fp]
    lib.orc_set_cloud_texture.argtypes = [P, P]
    lib.orc_precompute_sky.argtypes = [P, C.c_int]
    lib.orc_set_sky_tables.argtypes = [P, C.c_int, fp, fp]
    lib.orc_get_sky_tables.argtypes = [P, fp, fp]
    lib.orc_get_trans_lut.argtypes = [P, P]
    lib.orc_get_cloud_ambient.argtypes = [P, fp]
    lib.orc_sample_skybox.argtypes = [P, C.c_int, fp, fp, fp, fp]
    lib.orc_shift_probe.argtypes = [P, C.c_int, fp, fp]
    lib.orc_reservoir_probe.argtypes = [C.c_int, fp, fp]
    lib.orc_restir_render_probe.argtypes = [P, C.c_int, fp, fp, fp, fp]
    lib.orc_gris_probe.argtypes = [P, C.c_uint32, fp, fp, fp, fp, C.c_int, ip, fp]
    lib.orc_set_tile_shard.argtypes = [P, C.c_int, C.c_int]
    lib.orc_trace_primary.argtypes = [P, P]
    lib.orc_accumulate.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_accumulate_restir.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_get_reservoirs.argtypes = [P, P]
    lib.orc_oct_round_trip.argtypes = [C.c_int, fp, fp, fp]
    lib.orc_hash3.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    lib.orc_hash3.restype = C.c_uint32
    lib.orc_math_probe.argtypes = [C.c_int, C.c_int, fp, fp, fp]
    lib.orc_bsdf_lobewise_probe.argtypes = [P, C.c_int, ip, fp, fp, fp, fp, fp, fp]
    lib.orc_accumulate_moving.argtypes = [P, C.c_int, C.c_float, C.c_float]
    lib.orc_fetch_hdr_moving.argtypes = [P, fp]
    lib.orc_reset_moving.argtypes = [P]
    lib.orc_last_ms.argtypes = [P]
    lib.orc_last_ms.restype = C.c_double
    lib.orc_reset.argtypes = [P]
    lib.orc_set_restir_temporal.argtypes = [P, C.c_int]
    lib.orc_drop_restir_history.argtypes = [P]
    lib.orc_get_counters.argtypes = [P, C.POINTER(C.c_uint64)]
    lib.orc_fetch_hdr.argtypes = [P, fp]
    lib.orc_fetch_ldr.argtypes = [P, fp]
    lib.orc_tonemap.argtypes = [P, fp, fp]
    lib.orc_raytrace.argtypes = [P, C.c_int, fp, fp, fp, ip, fp, ip]
    lib.orc_occupancy.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_occupancy.restype = C.c_int
    lib.orc_bsdf_probe.argtypes = [P, C.c_int, ip, fp, fp, fp, fp, fp, fp]
    lib.orc_project_sky.argtypes = [C.c_int, C.c_int, fp, fp]
    lib.orc_unproject_sky.argtypes = [C.c_int, C.c_int, fp, fp]
    lib.orc_sample_sky_trans.argtypes = [P, C.c_int, fp, fp]
    lib.orc_rnd.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.orc_rnd.restype = C.c_float
    lib.orc_f32_to_f16.argtypes = [C.c_float]
    lib.orc_f32_to_f16.restype = C.c_uint16
    lib.orc_f16_to_f32.argtypes = [C.c_uint16]
    lib.orc_f16_to_f32.restype = C.c_float
    lib.orc_num_threads.restype = C.c_int
    lib.orc_set_num_threads.argtypes = [C.c_int]
    lib.orc_set_num_threads.restype = None
    _lib = lib
    return lib


def _f32(a, n=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    if n is not None and a.size != n:
        raise ValueError("expected %d floats" % n)
    return a


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleRenderer:
    """Same surface as voxel_rt2_b200.Renderer, evaluated by the CPU restatement."""

    def __init__(self, dx=1 / 64, image_res=(1920, 1080), up=(0, 1, 0), voxel_edges=0.06, exposure=3, *, grid_res=128,
                 max_depth=4, sky_res=3840, cloud_passes=32, device=0, seed=0, jitter=True, materials=None, cloud_tex=None):
        self._lib = load()
        self.image_res = (int(image_res[0]), int(image_res[1]))
        self.voxel_grid_res = int(grid_res)
        self.sky_res = int(sky_res)
        self.up = tuple(float(x) for x in up)
        self.current_spp = 0
        self.sample_stride, self.sample_offset = 1, 0
        self._h = C.c_void_p(self._lib.orc_create(self.image_res[0], self.image_res[1], self.voxel_grid_res, float(dx),
                                                  float(voxel_edges), float(exposure), int(max_depth), int(seed) & 0xFFFFFFFF,
                                                  1 if jitter else 0, int(cloud_passes)))
        self.fov = math.radians(50.0)
        self._camera_pos = np.array((0.4, 0.5, 2.0), np.float64)
        self._look_at = np.array((0.0, 0.0, 0.0), np.float64)
        self._dirty_camera = True
        self.use_physical_atmosphere, self.use_clouds = 0, 0
        self.n_threads = 0
        if materials is not None:
            self.set_materials(materials)
        if cloud_tex is not None:
            self.set_cloud_texture(cloud_tex)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_voxels(self, material, color):
        R = self.voxel_grid_res
        material = np.ascontiguousarray(material, dtype=np.int8)
        color = np.ascontiguousarray(color, dtype=np.uint8)
        assert material.shape == (R, R, R) and color.shape == (R, R, R, 3)
        self._lib.orc_upload_voxels(self._h, material.ctypes.data_as(C.c_void_p), color.ctypes.data_as(C.c_void_p))

    def set_directional_light(self, direction, light_cone_angle, light_color):
        self._lib.orc_set_light(self._h, _fp(_f32(direction, 3)), C.c_float(light_cone_angle), _fp(_f32(light_color, 3)))

    def set_floor(self, height, color, material=1):
        self._lib.orc_set_floor(self._h, C.c_float(height), _fp(_f32(color, 3)), int(material))

    def set_background_color(self, color):
        self._lib.orc_set_background(self._h, _fp(_f32(color, 3)))

    def set_use_physical_sky(self, use, clouds=None):
        self.use_physical_atmosphere = 1 if use else 0
        if clouds is not None:
            self.use_clouds = 1 if clouds else 0
        self._lib.orc_set_sky(self._h, self.use_physical_atmosphere, self.use_clouds)

    def set_use_clouds(self, use):
        self.use_clouds = 1 if use else 0
        self._lib.orc_set_sky(self._h, self.use_physical_atmosphere, self.use_clouds)

    def set_materials(self, t):
        self._lib.orc_set_materials(self._h, _fp(_f32(t, 128 * 14)))

    def set_cloud_texture(self, tex):
        tex = np.ascontiguousarray(tex, dtype=np.uint8)
        assert tex.shape == (256, 256, 3)
        self._lib.orc_set_cloud_texture(self._h, tex.ctypes.data_as(C.c_void_p))

    def set_camera_pos(self, x, y, z):
        self._camera_pos = np.array((x, y, z), np.float64)
        self._dirty_camera = True

    def set_look_at(self, x, y, z):
        self._look_at = np.array((x, y, z), np.float64)
        self._dirty_camera = True

    def set_fov(self, fov):
        self.fov = float(fov)
        self._dirty_camera = True

    def set_view_proj(self, pos, view, proj):
        rc = self._lib.orc_set_camera(self._h, _fp(_f32(pos, 3)), _fp(_f32(view, 16)), _fp(_f32(proj, 16)))
        assert rc == 0
        self._dirty_camera = False

    def _sync_camera(self):
        if self._dirty_camera:
            # same host-side matrix construction as the product's Python host (camera.py restated)
            eye, c, up = self._camera_pos, self._look_at, np.asarray(self.up, np.float64)
            f = (c - eye) / np.linalg.norm(c - eye)
            s = np.cross(f, up)
            s /= np.linalg.norm(s)
            u = np.cross(s, f)
            view = np.eye(4)
            view[0, :3], view[1, :3], view[2, :3] = s, u, -f
            view[0, 3], view[1, 3], view[2, 3] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
            g = 1.0 / math.tan(self.fov / 2.0)
            n, fa = 0.01, 10.0
            proj = np.zeros((4, 4))
            proj[0, 0] = g / (self.image_res[0] / self.image_res[1])
            proj[1, 1] = g
            proj[2, 2] = -(fa + n) / (fa - n)
            proj[2, 3] = -(2 * fa * n) / (fa - n)
            proj[3, 2] = -1.0
            self.set_view_proj(eye, view, proj)

    def prepare_data(self):
        self._sync_camera()
        if self.use_physical_atmosphere and not getattr(self, "_external_sky", False):
            self._lib.orc_precompute_sky(self._h, self.sky_res)

    def set_sky_tables(self, scattering, transmittance):
        S = self.sky_res
        self._lib.orc_set_sky_tables(self._h, S, _fp(_f32(scattering, S * S * 3)), _fp(_f32(transmittance, S * S * 3)))
        self._external_sky = True

    def get_sky_tables(self):
        S = self.sky_res
        a, b = np.empty((S, S, 3), np.float32), np.empty((S, S, 3), np.float32)
        self._lib.orc_get_sky_tables(self._h, _fp(a), _fp(b))
        return a, b

    def get_trans_lut(self):
        a = np.empty((256, 128, 3), np.float16)
        self._lib.orc_get_trans_lut(self._h, a.ctypes.data_as(C.c_void_p))
        return a

    def shift_probe(self, rows):
        """orc_shift_probe: rows[n][28] float32 (see oracle.cpp) -> [n][7]."""
        self._sync_camera()
        rows = np.ascontiguousarray(rows, np.float32)
        out = np.empty((rows.shape[0], 7), np.float32)
        self._lib.orc_shift_probe(self._h, rows.shape[0], _fp(rows), _fp(out))
        return out

    def gris_probe(self, frame, samples, gbuf, col_d, col_s, pixels):
        """orc_gris_probe: spatial_GRIS on caller-built reservoirs / G-buffer (see oracle.cpp) -> [n][6]."""
        self._sync_camera()
        npx = self.image_res[0] * self.image_res[1]
        samples, gbuf = _f32(samples, 23 * npx), _f32(gbuf, 7 * npx)
        col_d, col_s = _f32(col_d, 3 * npx), _f32(col_s, 3 * npx)
        pixels = np.ascontiguousarray(pixels, np.int32)
        out = np.empty((pixels.size, 6), np.float32)
        self._lib.orc_gris_probe(self._h, int(frame), _fp(samples), _fp(gbuf), _fp(col_d), _fp(col_s), pixels.size, _ip(pixels), _fp(out))
        return out

    def restir_render_probe(self, sample):
        """orc_restir_render_probe: (reservoirs before encode [npx][23], gbuf [npx][7], col_d, col_s) of one sample."""
        self._sync_camera()
        npx = self.image_res[0] * self.image_res[1]
        samples, gbuf = np.empty((npx, 23), np.float32), np.empty((npx, 7), np.float32)
        col_d, col_s = np.empty((npx, 3), np.float32), np.empty((npx, 3), np.float32)
        self._lib.orc_restir_render_probe(self._h, int(sample), _fp(samples), _fp(gbuf), _fp(col_d), _fp(col_s))
        return samples, gbuf, col_d, col_s

    def get_cloud_ambient(self):
        a = np.empty(3, np.float32)
        self._lib.orc_get_cloud_ambient(self._h, _fp(a))
        return a

    def sample_sky_trans(self, dirs):
        d = _f32(dirs)
        out = np.empty((d.size // 3, 3), np.float32)
        self._lib.orc_sample_sky_trans(self._h, d.size // 3, _fp(d), _fp(out))
        return out

    def sample_skybox(self, dirs, jitter):
        d, j = _f32(dirs), _f32(jitter)
        sc, tr = np.empty((d.size // 3, 3), np.float32), np.empty((d.size // 3, 3), np.float32)
        self._lib.orc_sample_skybox(self._h, d.size // 3, _fp(d), _fp(j), _fp(sc), _fp(tr))
        return sc, tr

    def set_tile_shard(self, rank, n):
        self._lib.orc_set_tile_shard(self._h, int(rank), int(n))

    def set_sample_shard(self, rank, n):
        self.sample_offset, self.sample_stride = int(rank), int(n)

    def accumulate(self, spp=1, stats=False):
        self._sync_camera()
        first = self.sample_offset + self.current_spp * self.sample_stride
        self._lib.orc_accumulate(self._h, first, int(spp), self.sample_stride, 1 if stats else 0, int(self.n_threads))
        self.current_spp += int(spp)

    def set_restir_temporal(self, enable):
        """Temporal reservoir reuse before the spatial pass (see temporal_reuse_pixel in oracle.cpp)."""
        self._lib.orc_set_restir_temporal(self._h, 1 if enable else 0)

    def accumulate_restir(self, frames=1):
        """accumulate() with USE_RESTIR_PT = True (pathtracer.py:1310-1319): render + spatial_GRIS per frame."""
        self._sync_camera()
        first = self.sample_offset + self.current_spp * self.sample_stride
        self._lib.orc_accumulate_restir(self._h, first, int(frames), self.sample_stride, int(self.n_threads))
        self.current_spp += int(frames)

    def accumulate_moving(self, render_scale=0.5, max_accum=50.0):
        """One frame of accumulate() with camera_is_moving = 1 (scene.py:214-228): half-resolution
        render, reprojected temporal filters, copy_prev_matrices."""
        self._sync_camera()
        self._lib.orc_accumulate_moving(self._h, self.sample_offset + self.current_spp * self.sample_stride, float(render_scale), float(max_accum))
        self.current_spp += 1

    def fetch_hdr_moving(self):
        out = np.empty((self.image_res[1], self.image_res[0], 4), np.float32)
        self._lib.orc_fetch_hdr_moving(self._h, _fp(out))
        return out

    def get_reservoirs(self):
        out = np.empty((self.image_res[1], self.image_res[0], 56), np.uint8)
        self._lib.orc_get_reservoirs(self._h, out.ctypes.data_as(C.c_void_p))
        return out

    def last_ms(self):
        return float(self._lib.orc_last_ms(self._h))

    def reset_framebuffer(self):
        self.current_spp = 0
        self._lib.orc_reset(self._h)
        self._lib.orc_reset_moving(self._h)

    def fetch_hdr(self):
        out = np.empty((self.image_res[1], self.image_res[0], 4), np.float32)
        self._lib.orc_fetch_hdr(self._h, _fp(out))
        return out

    def fetch_image(self):
        out = np.empty((self.image_res[1], self.image_res[0], 4), np.float32)
        self._lib.orc_fetch_ldr(self._h, _fp(out))
        return out

    def tonemap(self, hdr):
        hdr = np.ascontiguousarray(hdr, np.float32)
        out = np.empty_like(hdr)
        self._lib.orc_tonemap(self._h, _fp(hdr), _fp(out))
        return out

    def trace_primary(self):
        self._sync_camera()
        out = np.empty((self.image_res[1], self.image_res[0]), HIT_DTYPE)
        self._lib.orc_trace_primary(self._h, out.ctypes.data_as(C.c_void_p))
        return out

    def counters(self):
        a = (C.c_uint64 * 8)()
        self._lib.orc_get_counters(self._h, a)
        return dict(zip(("paths", "rays", "steps", "queries", "hits", "sky_escapes", "nee_visible", "vertices"), [int(x) for x in a]))

    # --- unit probes
    def raytrace(self, origins, dirs):
        o, d = _f32(origins), _f32(dirs)
        n = o.size // 3
        t = np.empty(n, np.float32)
        cell = np.empty((n, 3), np.int32)
        nrm = np.empty((n, 3), np.float32)
        it = np.empty(n, np.int32)
        self._lib.orc_raytrace(self._h, n, _fp(o), _fp(d), _fp(t), _ip(cell), _fp(nrm), _ip(it))
        return t, cell, nrm, it

    def occupancy(self, x, y, z, lod):
        return self._lib.orc_occupancy(self._h, int(x), int(y), int(z), int(lod))

    def bsdf_lobewise_probe(self, mat_id, albedo, v, n, l):
        mat_id = np.ascontiguousarray(mat_id, np.int32)
        k = mat_id.size
        out, lw = np.empty((k, 3, 7), np.float32), np.empty((k, 3), np.float32)
        self._lib.orc_bsdf_lobewise_probe(self._h, k, _ip(mat_id), _fp(_f32(albedo)), _fp(_f32(v)), _fp(_f32(n)), _fp(_f32(l)), _fp(out), _fp(lw))
        return out, lw

    def bsdf_probe(self, mat_id, albedo, v, n, l, u3):
        mat_id = np.ascontiguousarray(mat_id, np.int32)
        k = mat_id.size
        out = np.empty((k, 15), np.float32)
        self._lib.orc_bsdf_probe(self._h, k, _ip(mat_id), _fp(_f32(albedo)), _fp(_f32(v)), _fp(_f32(n)), _fp(_f32(l)), _fp(_f32(u3)),
                                 _fp(out))
        return out


def math_probe(kind, a, b=None, out_per=3):
    """orc_math_probe: see oracle.cpp for the kinds."""
    lib = load()
    a = np.ascontiguousarray(a, np.float32)
    n = a.shape[0]
    b = np.ascontiguousarray(b if b is not None else np.zeros((n, 3)), np.float32)
    out = np.zeros((n, out_per) if out_per > 1 else (n,), np.float32)
    lib.orc_math_probe(int(kind), n, _fp(a), _fp(b), _fp(out))
    return out


def reservoir_probe(rows):
    """orc_reservoir_probe: rows[n][53] float32 (see oracle.cpp) -> [n][28]."""
    lib = load()
    rows = np.ascontiguousarray(rows, np.float32)
    out = np.empty((rows.shape[0], 28), np.float32)
    lib.orc_reservoir_probe(rows.shape[0], _fp(rows), _fp(out))
    return out


def project_sky(dirs, S):
    lib = load()
    d = _f32(dirs)
    out = np.empty((d.size // 3, 2), np.float32)
    lib.orc_project_sky(d.size // 3, int(S), _fp(d), _fp(out))
    return out


def unproject_sky(uv, S):
    lib = load()
    u = _f32(uv)
    out = np.empty((u.size // 2, 3), np.float32)
    lib.orc_unproject_sky(u.size // 2, int(S), _fp(u), _fp(out))
    return out
