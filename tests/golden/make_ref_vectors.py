"""Golden vectors computed by the REFERENCE'S OWN SOURCE (/root/reference/renderer/*.py), executed
here through the float32 Taichi emulator in oracle/ti_emu (Taichi itself cannot be installed in
this container). Run in the authoring container:

    python tests/golden/make_ref_vectors.py [section ...]

Output: tests/golden/ref_*.npz (committed; /root/reference is not needed to run the tests).
tests/test_reference_vectors.py checks the CPU oracle against them.

Deviations from "the reference exactly as shipped", each a documented quirk:
  A1  the occupancy field is allocated 2 R^3 bits so the reference's LOD base offsets stay inside
      the field (as shipped they overrun the allocation; the intended bits are the same);
  A3  a query for a cell outside the grid (the reference reads an aliased / out-of-range word there)
      is answered "empty" and the ray is recorded with flag 1: only its hit distance is compared
      (the oracle's pin is "leaving the grid is a miss");
  A4  directions with an exactly-zero component are not generated (0 * inf in the reference).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "ti_emu"))
sys.path.insert(0, REF)
os.chdir(REF)  # the reference opens default_material_set.csv / textures by relative path

import taichi as ti  # noqa: E402  (the emulator)

F = np.float32


def vec(a):
    return ti.Matrix(np.asarray(a, dtype=np.float32), _noconv=True)


def unit(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


# ------------------------------------------------------------------------------- traversal
def grids():
    rng = np.random.default_rng(20260101)
    g = {}
    g["rand16_15"] = (rng.random((16, 16, 16)) < 0.15).astype(np.int8)
    g["rand32_40"] = (rng.random((32, 32, 32)) < 0.40).astype(np.int8)
    s = np.zeros((32, 32, 32), np.int8)  # sparse structured: a slab, a column and isolated voxels
    s[4:28, 3, 4:28] = 1
    s[16, 4:30, 16] = 1
    s[29, 29, 29] = s[0, 0, 0] = s[31, 31, 0] = s[7, 20, 9] = 1
    g["struct32"] = s
    return g


def section_raytrace():
    from renderer.math_utils import eps, inf
    from renderer.raytracer import VoxelOctreeRaytracer

    out = {}
    for name, occ in grids().items():
        R = occ.shape[0]
        rt = VoxelOctreeRaytracer(R)
        rt.occupancy = ti.field(ti.i32, shape=(2 * R ** 3 // 32 + 1,))  # A1
        vox = ti.field(ti.i8, shape=(R, R, R))
        vox.from_numpy(occ)
        rt._update_lods(vox, ti.Vector([0, 0, 0]))
        # occupancy bits of every LOD, through the reference's own query function
        bits = []
        for lod in range(rt.n_lods):
            r = R >> lod
            b = np.zeros((r, r, r), np.uint8)
            for x in range(r):
                for y in range(r):
                    for z in range(r):
                        b[x, y, z] = 1 if rt.query_occupancy(ti.Vector([x, y, z], ti.i32), np.int32(lod)) else 0
            bits.append(b.reshape(-1))
        out[name + "_grid"] = occ
        out[name + "_lodbits"] = np.concatenate(bits)

        inner = VoxelOctreeRaytracer.query_occupancy

        oob = [0]

        def checked(ipos, lod, _rt=rt, _R=R):
            r = _R >> int(lod)
            if int(ipos.data.min()) < 0 or int(ipos.data.max()) >= r:
                oob[0] = 1  # A3: the cell is outside the grid; read it as empty (robust-buffer behaviour)
                return False
            return inner(_rt, ipos, lod)

        rt.query_occupancy = checked
        rng = np.random.default_rng(R * 1000 + len(name))
        n = 300
        o = np.empty((n, 3), np.float32)
        d = unit(rng, n)
        third = n // 3
        # outside origins aimed at the box, origins inside the box, origins on / near cell boundaries
        o[:third] = (R / 2 + unit(rng, third) * R * rng.uniform(0.9, 2.0, (third, 1))).astype(np.float32)
        tgt = rng.uniform(0, R, (third, 3))
        dd = tgt - o[:third]
        d[:third] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(np.float32)
        o[third:2 * third] = rng.uniform(0, R, (third, 3)).astype(np.float32)
        o[2 * third:] = (np.floor(rng.uniform(0, R, (n - 2 * third, 3))) + rng.choice([0.0, 1e-6, 0.5, 0.999999], (n - 2 * third, 3))).astype(np.float32)
        t = np.zeros(n, np.float32)
        cell = np.zeros((n, 3), np.int32)
        nrm = np.zeros((n, 3), np.float32)
        iters = np.zeros(n, np.int32)
        flag = np.zeros(n, np.int32)
        for i in range(n):
            oob[0] = 0
            ht, hc, hn, hi = rt.raytrace(vec(o[i]), vec(d[i]), eps, inf)
            t[i], cell[i], nrm[i], iters[i], flag[i] = ht, hc.data, hn.data, hi, oob[0]
        out[name + "_o"], out[name + "_d"], out[name + "_t"] = o, d, t
        out[name + "_cell"], out[name + "_normal"], out[name + "_iters"], out[name + "_flag"] = cell, nrm, iters, flag
        print("raytrace %s: %d rays, %d hits, %d out-of-grid (A3)" % (name, n, int(np.isfinite(t).sum()), int(flag.sum())))
    np.savez_compressed(os.path.join(HERE, "ref_raytrace.npz"), **out)


# ------------------------------------------------------------------------------ math helpers
def section_math():
    import renderer.math_utils as mu

    rng = np.random.default_rng(7)
    out = {}
    n = 64
    v = unit(rng, n)
    bx, by = np.zeros((n, 3), F), np.zeros((n, 3), F)
    for i in range(n):
        x, y = mu.make_orthonormal_basis(vec(v[i]))
        bx[i], by[i] = x.data, y.data
    out["onb_n"], out["onb_x"], out["onb_y"] = v, bx, by
    # octahedral 2 x f16 round trip (math_utils.py:201-215)
    enc, dec = np.zeros((n, 2), np.float16), np.zeros((n, 3), F)
    for i in range(n):
        e = mu.encode_unit_vector_3x16(vec(v[i]))
        enc[i] = e.data
        dec[i] = mu.decode_unit_vector_3x16(e).data
    out["oct_v"], out["oct_enc"], out["oct_dec"] = v, enc, dec
    # hash3 (math_utils.py:217-229)
    hx = rng.integers(0, 2 ** 32, (n, 3), dtype=np.uint64).astype(np.uint32)
    out["hash3_in"] = hx
    out["hash3_out"] = np.array([mu.hash3(np.uint32(a), np.uint32(b), np.uint32(c)) for a, b, c in hx], np.uint32)
    # encode_material / decode (math_utils.py:231-247): mat id + albedo -> u32
    mid = rng.integers(0, 128, n).astype(np.int32)
    alb = rng.random((n, 3)).astype(F)
    out["encmat_id"], out["encmat_albedo"] = mid, alb
    out["encmat_out"] = np.array([mu.encode_material(np.int32(m), vec(a)) for m, a in zip(mid, alb)], np.uint32)
    # u8 colour conversions (math_utils.py:86-100)
    c = np.concatenate([rng.random((n - 4, 3)), [[0, 0, 0], [1, 1, 1], [1.5, -0.2, 0.5], [0.999, 0.001, 0.5]]]).astype(F)
    out["rgb_in"] = c
    out["rgb_u8"] = np.array([mu.rgb32f_to_rgb8(vec(x)).data for x in c], np.uint8)
    out["rgb_back"] = np.array([mu.rgb8_to_rgb32f(ti.Matrix(u, _noconv=True)).data for u in out["rgb_u8"]], F)
    # uchimura tonemap (math_utils.py:160-186) and luminance
    x = np.concatenate([rng.random((n, 3)) * 4.0, [[0, 0, 0], [0.22, 0.22, 0.22], [0.532, 0.1, 10.0]]]).astype(F)
    out["uchi_in"] = x
    out["uchi_out"] = np.array([mu.uchimura(vec(a)).data for a in x], F)
    out["lum_out"] = np.array([mu.luminance(vec(a)) for a in x], F)
    # cone pdf / samples with supplied random numbers (math_utils.py:44-63)
    u = rng.random((n, 2)).astype(F)
    cm = np.cos(rng.uniform(0.005, 0.3, n)).astype(F)
    samples = np.zeros((n, 3), F)
    hemi = np.zeros((n, 3), F)
    for i in range(n):
        q = list(u[i])
        ti.set_random_source(lambda name: q.pop(0))
        samples[i] = mu.sample_cone_oriented(cm[i], vec(v[i])).data
        q = list(u[i])
        hemi[i] = mu.sample_cosine_weighted_hemisphere(vec(v[i])).data
    ti.set_random_source(None)
    out["cone_u"], out["cone_cosmax"], out["cone_n"], out["cone_dir"], out["hemi_dir"] = u, cm, v, samples, hemi
    np.savez_compressed(os.path.join(HERE, "ref_math.npz"), **out)
    print("math: %d vectors per helper" % n)


# -------------------------------------------------------------------------------------- BSDF
def section_bsdf():
    from renderer.materials import MaterialList
    from renderer.math_utils import make_orthonormal_basis

    mats = MaterialList()
    bsdf = mats.bsdf
    rng = np.random.default_rng(11)
    ids = [0, 1, 2, 10, 11, 20, 21, 22, 30, 31, 32, 40, 41, 50, 51, 52, 53, 54, 80, 81, 82]
    per = 12
    n = len(ids) * per
    mat_id = np.repeat(np.array(ids, np.int32), per)
    albedo = rng.uniform(0.05, 1.0, (n, 3)).astype(F)
    nrm = unit(rng, n)
    # view / light directions in the upper hemisphere of n (plus a few below it)
    def hemi(k):
        w = unit(rng, n)
        s = np.sign((w * nrm).sum(1, keepdims=True))
        s[s == 0] = 1
        w = w * s
        flip = rng.random(n) < k
        w[flip] = -w[flip]
        return w.astype(F)

    v, l = hemi(0.0), hemi(0.1)
    u3 = rng.random((n, 3)).astype(F)
    ev_d, ev_s = np.zeros((n, 3), F), np.zeros((n, 3), F)
    pdf = np.zeros(n, F)
    sdir, sbrdf = np.zeros((n, 3), F), np.zeros((n, 3), F)
    spdf, slobe = np.zeros(n, F), np.zeros(n, np.int32)
    lw = np.zeros((n, 3), F)
    lobe_d, lobe_s, lobe_pdf = np.zeros((n, 3, 3), F), np.zeros((n, 3, 3), F), np.zeros((n, 3), F)
    table = np.zeros((128, 14), F)
    names = ["subsurface", "metallic", "specular", "specular_tint", "roughness", "anisotropic", "sheen", "sheen_tint", "clearcoat",
             "clearcoat_gloss", "ior_minus_one"]
    for m in range(128):
        s = mats.mat_list[m]
        table[m, :3] = s.base_col.data
        table[m, 3:] = [getattr(s, k) for k in names]
    for i in range(n):
        m = mats.mat_list[int(mat_id[i])]
        m.base_col = vec(albedo[i])
        N, V, L = vec(nrm[i]), vec(v[i]), vec(l[i])
        tang, bitang = make_orthonormal_basis(N)
        d, s = bsdf.disney_evaluate_split(m, V, N, L, tang, bitang)
        ev_d[i], ev_s[i] = d.data, s.data
        pdf[i] = bsdf.pdf_disney(m, V, N, L, tang, bitang)
        lw[i] = [float(x) for x in bsdf.disney_get_lobe_probabilities(m)]
        q = list(u3[i])
        ti.set_random_source(lambda name: q.pop(0))
        sd, sb, sp, sl = bsdf.sample_disney(m, V, N, tang, bitang)
        ti.set_random_source(None)
        sdir[i], sbrdf[i], spdf[i], slobe[i] = sd.data, sb.data, sp, sl
        for lobe in range(3):
            a, b = bsdf.disney_evaluate_lobewise_split(m, V, N, L, tang, bitang, np.int32(lobe))
            lobe_d[i, lobe], lobe_s[i, lobe] = a.data, b.data
            lobe_pdf[i, lobe] = bsdf.pdf_disney_lobewise(m, V, N, L, tang, bitang, np.int32(lobe))
    np.savez_compressed(os.path.join(HERE, "ref_bsdf.npz"), material_table=table, mat_id=mat_id, albedo=albedo, n=nrm, v=v, l=l, u3=u3,
                        eval_d=ev_d, eval_s=ev_s, pdf=pdf, lobe_w=lw, sample_dir=sdir, sample_brdf=sbrdf, sample_pdf=spdf, sample_lobe=slobe,
                        lobe_d=lobe_d, lobe_s=lobe_s, lobe_pdf=lobe_pdf)
    print("bsdf: %d probes over %d materials" % (n, len(ids)))


# ------------------------------------------------------------------- Renderer: hits + render()
def mix32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def sampler_rnd(pixel, sample, seed, dim):
    """The counter-based sampler shared by oracle and CUDA (DESIGN.md "Sampler")."""
    key = mix32((mix32(pixel ^ ((0x9E3779B9 * (sample + 1)) & 0xFFFFFFFF)) + seed) & 0xFFFFFFFF)
    h = mix32((key + 0x9E3779B9 * (dim + 1)) & 0xFFFFFFFF)
    return np.float32(h >> 8) * np.float32(1.0 / 16777216.0)


DIMS = {"sample_cone": (0, 2), "sample_disney": (2, 1), "sample_cosine_weighted_hemisphere": (3, 2), "GGX_VNDF_aniso": (3, 2),
        "sample_clearcoat": (3, 2), "sample_skybox": (5, 3)}


def render_scene(R, seed):
    rng = np.random.default_rng(seed)
    mat = np.zeros((R, R, R), np.int8)
    occ = rng.random((R, R, R)) < 0.08
    occ[R // 4: 3 * R // 4, : R // 3, R // 4: 3 * R // 4] |= rng.random((R // 2, R // 3, R // 2)) < 0.6
    ids = np.array([1, 1, 1, 11, 21, 32, 40, 50, 52, 54, 2], np.int8)
    mat[occ] = rng.choice(ids, int(occ.sum()))
    col = rng.integers(40, 250, (R, R, R, 3)).astype(np.uint8)
    return mat, col


def make_reference_renderer(W, H, R, mat, col, cfg):
    from renderer.pathtracer import Renderer
    from renderer.raytracer import VoxelOctreeRaytracer
    from renderer.voxel_world import VoxelWorld

    r = Renderer(dx=2.0 / R, image_res=(W, H), up=(0, 1, 0), voxel_edges=cfg["voxel_edges"], exposure=3)
    # the shipped Renderer hard-codes a 128^3 grid (pathtracer.py:80); same classes at R^3
    r.voxel_grid_res = R
    r.world = VoxelWorld(2.0 / R, R, cfg["voxel_edges"])
    r.voxel_raytracer = VoxelOctreeRaytracer(R)
    r.voxel_raytracer.occupancy = ti.field(ti.i32, shape=(2 * R ** 3 // 32 + 1,))  # A1
    r.world.voxel_material.arr[...] = mat
    r.world.voxel_color.arr[...] = col
    r.set_directional_light(cfg["light_dir"], cfg["light_cone"], cfg["light_color"])
    r.floor_height[None] = cfg["floor_height"]
    r.floor_color[None] = cfg["floor_color"]
    r.floor_material[None] = cfg["floor_material"]
    r.background_color[None] = cfg["background"]
    r.use_physical_atmosphere[None] = 0
    r.camera_is_moving[None] = 0
    r.render_scale[None] = 1.0
    r.world.update_data()
    r.voxel_raytracer._update_lods(r.world.voxel_material, ti.Vector(r.world.voxel_grid_offset))
    inner = VoxelOctreeRaytracer.query_occupancy

    def checked(ipos, lod, _rt=r.voxel_raytracer):  # A3
        n = R >> int(lod)
        if int(ipos.data.min()) < 0 or int(ipos.data.max()) >= n:
            return False
        return inner(_rt, ipos, lod)

    r.voxel_raytracer.query_occupancy = checked
    return r


def set_reference_camera(r, pos, view, proj):
    ti.set_random_source(lambda name: 0.5)
    r.set_camera_pos(*[float(x) for x in pos])
    r.set_view_mat(np.ascontiguousarray(view.T))  # the kernels transpose what GGUI hands them (pathtracer.py:262-281)
    r.set_proj_mat(np.ascontiguousarray(proj.T))
    ti.set_random_source(None)
    r.taa_jitter[None] = (0.0, 0.0)


def section_render():
    """Renderer.next_hit on every primary ray + sun shadow ray (hit buffer) and Renderer.render()
    (pathtracer.py:355-632) for a few samples per pixel, with ti.random() answering from the
    oracle's counter-based sampler: dimension 8*depth + {0,1 cone; 2 lobe; 3,4 direction; 5..7 sky
    jitter} of the path (pixel, sample), read from the calling frames. TAA jitter is zero."""
    sys.path.insert(0, ROOT)
    from renderer.math_utils import eps, inf
    from voxel_rt2_b200.camera import default_camera_matrices

    W, H, R, seed, n_samples = 32, 16, 32, 77, 4
    cfg = dict(voxel_edges=0.06, light_dir=(1.0, 1.0, 0.4), light_cone=0.06, light_color=(1.2, 1.1, 0.9), floor_height=-0.55,
               floor_color=(0.8, 0.75, 0.7), floor_material=1, background=(0.25, 0.35, 0.55))
    mat, col = render_scene(R, 5)
    r = make_reference_renderer(W, H, R, mat, col, cfg)
    pos, view, proj = default_camera_matrices(W, H, pos=(0.9, 0.8, 1.9))
    set_reference_camera(r, pos, view, proj)
    tex = r.world.voxel_color_texture
    out = dict(material=mat, color=col, cam_pos=pos, view=view, proj=proj, seed=np.int32(seed), W=np.int32(W), H=np.int32(H))
    for k, v in cfg.items():
        out["cfg_" + k] = np.asarray(v, np.float32)
    # ---- hit buffer: primary ray + shadow ray on the cone axis (the oracle's / CUDA's trace_primary)
    hit_t = np.zeros((H, W), np.float32)
    hit_n = np.zeros((H, W, 3), np.float32)
    hit_alb = np.zeros((H, W, 3), np.float32)
    hit_mat = np.zeros((H, W), np.int32)
    hit_light = np.zeros((H, W), np.int32)
    shadow = np.full((H, W), 3, np.int32)
    dirs = np.zeros((H, W, 3), np.float32)
    cam = r.camera_pos[None].copy()
    ldir = r.light_direction[None].copy()
    for v in range(H):
        for u in range(W):
            d = r.get_cast_dir(np.int32(u), np.int32(v))
            closest, normal, albedo, hl, iters, mid = r.next_hit(cam, d, inf, tex, shadow_ray=False)
            dirs[v, u], hit_t[v, u], hit_n[v, u], hit_alb[v, u], hit_mat[v, u], hit_light[v, u] = d.data, closest, normal.data, albedo.data, mid, hl
            if not hl and closest < inf:
                p = cam + closest * d + normal * eps
                if ldir.dot(normal) > 0:
                    dist = r.next_hit(p, ldir, inf, tex, shadow_ray=True)[0]
                    shadow[v, u] = 0 if dist >= inf else 1
                else:
                    shadow[v, u] = 2
    out.update(hit_dir=dirs, hit_t=hit_t, hit_normal=hit_n, hit_albedo=hit_alb, hit_mat=hit_mat, hit_light=hit_light, hit_shadow=shadow)
    print("hit buffer: %d voxel/floor hits of %d pixels" % (int(np.isfinite(hit_t).sum()), W * H))

    # ---- render(): per-pixel radiance of single samples
    state = {"sample": 0, "count": {}}

    def source(name, frame):
        f = frame
        while f is not None and f.f_code.co_name != "render":
            f = f.f_back
        if f is None or name not in DIMS:
            return 0.5  # set_proj_mat jitter, Reservoir.input_sample: not part of the non-ReSTIR pixel value
        u, v, depth = int(f.f_locals["u"]), int(f.f_locals["v"]), int(f.f_locals["depth"])
        base, n = DIMS[name]
        key = (u, v, depth, name)
        i = state["count"].get(key, 0)
        state["count"][key] = i + 1
        assert i < n, key
        return sampler_rnd(v * W + u, state["sample"], seed, 8 * depth + base + i)

    ti.set_random_source(source, with_frame=True)
    diff = np.zeros((n_samples, H, W, 3), np.float32)
    spec = np.zeros((n_samples, H, W, 3), np.float32)
    for s in range(n_samples):
        state["sample"], state["count"] = s, {}
        r.render(tex)
        diff[s] = np.transpose(r.color_buffer.arr, (1, 0, 2))
        spec[s] = np.transpose(r.color_buffer_specular.arr, (1, 0, 2))
        print("render sample %d: mean %.4f" % (s, float((diff[s] + spec[s]).mean())))
    ti.set_random_source(None)
    out.update(render_diffuse=diff, render_specular=spec)
    np.savez_compressed(os.path.join(HERE, "ref_render.npz"), **out)


def path_random_source(W, seed, state):
    """ti.random() of Renderer.render answered from the shared counter-based sampler (see section_render)."""
    def source(name, frame):
        f = frame
        while f is not None and f.f_code.co_name != "render":
            f = f.f_back
        if f is None or name not in DIMS:
            return 0.5  # set_proj_mat jitter, Reservoir.input_sample: not part of the non-ReSTIR pixel value
        u, v, depth = int(f.f_locals["u"]), int(f.f_locals["v"]), int(f.f_locals["depth"])
        base, n = DIMS[name]
        key = (u, v, depth, name)
        i = state["count"].get(key, 0)
        state["count"][key] = i + 1
        assert i < n, key
        return sampler_rnd(v * W + u, state["sample"], seed, 8 * depth + base + i)
    return source


def synthetic_sky_tables(S):
    """Smooth positive tables (the lookups, not the atmosphere model, are under test here)."""
    x, y = np.meshgrid((np.arange(S) + 0.5) / S, (np.arange(S) + 0.5) / S, indexing="ij")
    sc = np.stack([0.25 + 0.2 * np.sin(6.283 * x) * y, 0.3 + 0.25 * y, 0.45 + 0.3 * y * np.cos(6.283 * x) ** 2], -1)
    tr = np.stack([0.3 + 0.6 * y, 0.35 + 0.5 * y * (0.5 + 0.5 * np.cos(6.283 * x)), 0.2 + 0.7 * y ** 2], -1)
    return sc.astype(np.float32), tr.astype(np.float32)


def section_frame():
    """The static-camera frame loop of Scene.finish (scene.py:206-262) for 4 frames with the physical
    sky on (synthetic 16 x 16 tables): set_proj_mat / set_view_mat, accumulate() = render +
    temporal_filter_prepass + temporal_filter + temporal_filter_specular (pathtracer.py:1310-1319),
    fetch_image() = _render_to_image (:634-662), copy_prev_matrices. Stored: the accumulated HDR
    buffer and the tonemapped image after the last frame."""
    sys.path.insert(0, ROOT)
    from voxel_rt2_b200.camera import default_camera_matrices

    W, H, R, seed, n_frames, S = 32, 16, 32, 123, 4, 16
    cfg = dict(voxel_edges=0.06, light_dir=(0.5, 0.8, 0.6), light_cone=0.025, light_color=(1.3, 1.2337, 1.2181), floor_height=-0.4,
               floor_color=(1.0, 1.0, 1.0), floor_material=1, background=(0.0, 0.0, 0.0))
    mat, col = render_scene(R, 9)
    r = make_reference_renderer(W, H, R, mat, col, cfg)
    sc, tr = synthetic_sky_tables(S)
    r.use_physical_atmosphere[None] = 1
    r.atmos.skybox_res = ti.Vector([S, S])
    r.atmos.skybox_fres = ti.Vector([1.0 / S, 1.0 / S])
    r.atmos.skybox_scattering = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    r.atmos.skybox_transmittance = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    r.atmos.skybox_scattering.arr[...] = sc
    r.atmos.skybox_transmittance.arr[...] = tr
    r.color_buffer.oob_zero = r.color_buffer_specular.oob_zero = True  # bilinear_sample reads column W / row H with weight ~0 (pathtracer.py:1077-1090)
    pos, view, proj = default_camera_matrices(W, H, pos=(-0.7, 0.9, 1.8))
    state = {"sample": 0, "count": {}}
    r.set_max_samples(999999999.0)
    r.set_render_scale(1.0)
    r.set_camera_is_moving(False)
    r.reset_framebuffer()
    r.current_spp = 0  # reset_framebuffer leaves 1 (pathtracer.py:664-668); only _render_to_image's unused argument reads it
    for s in range(n_frames):
        set_reference_camera(r, pos, view, proj)
        state["sample"], state["count"] = s, {}
        ti.set_random_source(path_random_source(W, seed, state), with_frame=True)
        r.accumulate()
        ti.set_random_source(None)
        img = r.fetch_image()
        r.copy_prev_matrices()
        print("frame %d: hdr mean %.4f ldr mean %.4f" % (s, float(r.color_buffer.arr.mean()), float(img.arr[..., :3].mean())))
    out = dict(material=mat, color=col, cam_pos=pos, view=view, proj=proj, seed=np.int32(seed), W=np.int32(W), H=np.int32(H),
               sky_res=np.int32(S), sky_scatter=sc, sky_trans=tr, n_frames=np.int32(n_frames), exposure=np.float32(3.0),
               hdr=np.transpose(r.color_buffer.arr, (1, 0, 2)).copy(), ldr=np.transpose(img.arr, (1, 0, 2)).copy())
    for k, v in cfg.items():
        out["cfg_" + k] = np.asarray(v, np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_frame.npz"), **out)


# ----------------------------------------------------------------- voxel authoring (row a1)
def section_voxel():
    """Scene.round_idx (scene.py:131-137) and Renderer.set_voxel / get_voxel (pathtracer.py:1325-1334,
    math_utils.py:86-100): index rounding (ti.round: half away from zero), the i8 material cast and
    the u8 colour quantisation, on 96 hand-picked and random (index, material, colour) triples."""
    import scene as ref_scene  # /root/reference/scene.py; only the class is used, no window is opened

    R = 32
    cfg = dict(voxel_edges=0.06, light_dir=(1, 1, 1), light_cone=0.1, light_color=(0, 0, 0), floor_height=0.0, floor_color=(1, 1, 1),
               floor_material=1, background=(0, 0, 0))
    r = make_reference_renderer(32, 16, R, np.zeros((R, R, R), np.int8), np.zeros((R, R, R, 3), np.uint8), cfg)
    rng = np.random.default_rng(55)
    n = 96
    idx = rng.uniform(-R / 2 + 0.6, R / 2 - 1.6, (n, 3))
    idx[:12] = np.round(idx[:12]) + rng.choice([0.5, -0.5, 0.49999997, 0.0], (12, 3))   # ties and near-ties
    idx[12:20] = np.round(idx[12:20])                                                    # integer-valued
    idx = idx.astype(np.float32)
    mats = rng.choice(np.array([0, 1, 2, 11, 54, 82, 127, 128, 200, 255, -1]), n).astype(np.int32)
    cols = rng.uniform(-0.2, 1.2, (n, 3)).astype(np.float32)
    cols[:6] = [[0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125], [0.999999, 0.003921569, 0.00392], [254.5 / 255, 1 / 255, 2 / 255], [0.2, 0.4, 0.6]]
    rounded = np.zeros((n, 3), np.int32)
    got_mat = np.zeros(n, np.int32)
    got_col = np.zeros((n, 3), np.float32)
    for i in range(n):
        ri = ref_scene.Scene.round_idx(vec(idx[i]))
        rounded[i] = ri.data
        r.set_voxel(ri, np.int32(mats[i]), vec(cols[i]))
        m, c = r.get_voxel(ri)
        got_mat[i], got_col[i] = m, c.data
    np.savez_compressed(os.path.join(HERE, "ref_voxel.npz"), R=np.int32(R), idx=idx, mat=mats, color=cols, rounded=rounded, get_mat=got_mat,
                        get_color=got_col, material_field=r.world.voxel_material.arr.copy(), color_field=r.world.voxel_color.arr.copy())
    print("voxel: %d set/get round trips, %d distinct cells" % (n, len({tuple(x) for x in rounded})))


# ------------------------------------------------------------ moving-camera temporal path
class _SnapshotReads:
    """Field proxy: reads see the array as it was when the proxy was made, writes go to the field.
    Used for gbuff_depth_reflection during temporal_filter_prepass, which blurs that buffer in place
    while neighbouring threads still read it (pathtracer.py:1055,1066): the emulator would serialise
    the race in loop order; the snapshot is the "all reads before any write" outcome the oracle pins."""

    def __init__(self, field):
        self.f, self.snap = field, field.arr.copy()

    def __getitem__(self, k):
        return self.snap[self.f._key(k)]

    def __setitem__(self, k, v):
        self.f[k] = v


def section_moving():
    """Scene.finish's moving-camera loop (scene.py:214-262) for 4 frames of a camera translation:
    render at render_scale 0.5 with albedo demodulation, temporal_filter_prepass, temporal_filter and
    temporal_filter_specular with reprojection / Catmull-Rom history / depth + normal rejection
    (pathtracer.py:993-1303), copy_prev_matrices. Stored: color_buffer after every frame.
    Two upstream hazards are resolved the way the oracle / CUDA pin them (DESIGN.md "Moving-camera
    pins"), otherwise the comparison would measure the hazards, not the filters: (1) the prepass
    blurs gbuff_depth_reflection in place while neighbours read it — reads see a snapshot; (2) a
    glossy first bounce that escapes leaves an infinite reflection distance, whose NaN depth the 4x4
    blur spreads and which resets the specular history of every pixel it reaches — non-finite
    reflection depths are zeroed ("no reflection") after render()."""
    sys.path.insert(0, ROOT)
    from voxel_rt2_b200.camera import default_camera_matrices

    W, H, R, seed, n_frames, scale = 64, 32, 32, 321, 4, 0.5
    cfg = dict(voxel_edges=0.06, light_dir=(0.7, 0.9, 0.5), light_cone=0.05, light_color=(1.2, 1.1, 1.0), floor_height=-0.45,
               floor_color=(0.9, 0.9, 0.9), floor_material=1, background=(0.2, 0.3, 0.5))
    mat, col = render_scene(R, 12)
    r = make_reference_renderer(W, H, R, mat, col, cfg)
    for f in (r.color_buffer, r.color_buffer_specular):
        f.oob_zero = True
    r.set_max_samples(50.0)
    r.set_render_scale(scale)
    r.set_camera_is_moving(True)
    r.reset_framebuffer()
    state = {"sample": 0, "count": {}}
    cams, frames, bad_total = [], [], 0
    tex = r.world.voxel_color_texture
    for fidx in range(n_frames):
        pos, view, proj = default_camera_matrices(W, H, pos=(0.9 - 0.06 * fidx, 0.8 + 0.02 * fidx, 1.9 - 0.03 * fidx))
        set_reference_camera(r, pos, view, proj)
        if fidx == 0:
            r.copy_prev_matrices()  # the first moving frame reprojects into the same camera
        state["sample"], state["count"] = fidx, {}
        ti.set_random_source(path_random_source(W, seed, state), with_frame=True)
        r.render(tex)                                   # Renderer.accumulate (pathtracer.py:1310-1319), kernel by kernel
        ti.set_random_source(None)
        real = r.gbuff_depth_reflection
        n_bad = int((~np.isfinite(real.arr)).sum())
        real.arr[~np.isfinite(real.arr)] = 0.0          # see the docstring: non-finite reflection depth = no reflection
        bad_total += n_bad
        r.gbuff_depth_reflection = _SnapshotReads(real)
        r.temporal_filter_prepass()
        r.gbuff_depth_reflection = real
        r.temporal_filter()
        r.temporal_filter_specular()
        r.copy_prev_matrices()
        cams.append((pos, view, proj))
        frames.append(np.transpose(r.color_buffer.arr, (1, 0, 2)).copy())
        print("moving frame %d: mean %.4f, %d non-finite reflection depths zeroed" % (fidx, float(frames[-1][: H // 2, : W // 2].mean()), n_bad))
    out = dict(material=mat, color=col, seed=np.int32(seed), W=np.int32(W), H=np.int32(H), scale=np.float32(scale), max_accum=np.float32(50.0),
               cam_pos=np.array([c[0] for c in cams]), view=np.array([c[1] for c in cams]), proj=np.array([c[2] for c in cams]),
               frames=np.array(frames))
    for k, v in cfg.items():
        out["cfg_" + k] = np.asarray(v, np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_moving.npz"), **out)


# ------------------------------------------------ BASELINE config 1: example1 hit buffer
def section_example1():
    """BASELINE.json configs[0] at reduced resolution: the example1.py scene (shim seed 0, the committed
    tests/golden/example1_seed0.npz) in the Renderer exactly as shipped (128^3 grid, dx = 1/64, default
    camera, fov 50), 64 x 64 primary rays + the sun shadow ray on the cone axis, through the
    reference's own get_cast_dir / next_hit / _update_lods / _make_texture."""
    sys.path.insert(0, ROOT)
    from renderer.math_utils import eps, inf
    from renderer.pathtracer import Renderer
    from renderer.raytracer import VoxelOctreeRaytracer
    from voxel_rt2_b200.camera import default_camera_matrices

    z = np.load(os.path.join(HERE, "example1_seed0.npz"))
    W = H = 64
    r = Renderer(dx=1.0 / 64.0, image_res=(W, H), up=(0, 1, 0), voxel_edges=float(z["voxel_edges"]), exposure=float(z["exposure"]))
    R = r.voxel_grid_res
    r.voxel_raytracer.occupancy = ti.field(ti.i32, shape=(2 * R ** 3 // 32 + 1,))  # A1
    r.world.voxel_material.arr[...] = z["material"]
    r.world.voxel_color.arr[...] = z["color"]
    r.set_directional_light(tuple(float(x) for x in z["light_dir"]), float(z["light_noise"]), tuple(float(x) for x in z["light_color"]))
    r.floor_height[None] = float(z["floor_height"])
    r.floor_color[None] = tuple(float(x) for x in z["floor_color"])
    r.floor_material[None] = int(z["floor_material"])
    r.camera_is_moving[None] = 0
    r.render_scale[None] = 1.0
    r.prepare_data()  # world.update_data() + _update_lods, as Scene.finish calls it
    inner = VoxelOctreeRaytracer.query_occupancy

    def checked(ipos, lod, _rt=r.voxel_raytracer):  # A3
        n = R >> int(lod)
        if int(ipos.data.min()) < 0 or int(ipos.data.max()) >= n:
            return False
        return inner(_rt, ipos, lod)

    r.voxel_raytracer.query_occupancy = checked
    pos, view, proj = default_camera_matrices(W, H)
    set_reference_camera(r, pos, view, proj)
    tex = r.world.voxel_color_texture
    hit_t = np.zeros((H, W), np.float32)
    hit_n = np.zeros((H, W, 3), np.float32)
    hit_mat = np.zeros((H, W), np.int32)
    hit_light = np.zeros((H, W), np.int32)
    shadow = np.full((H, W), 3, np.int32)
    cam = r.camera_pos[None].copy()
    ldir = r.light_direction[None].copy()
    for v in range(H):
        for u in range(W):
            d = r.get_cast_dir(np.int32(u), np.int32(v))
            closest, normal, albedo, hl, iters, mid = r.next_hit(cam, d, inf, tex, shadow_ray=False)
            hit_t[v, u], hit_n[v, u], hit_mat[v, u], hit_light[v, u] = closest, normal.data, mid, hl
            if not hl and closest < inf:
                p = cam + closest * d + normal * eps
                if ldir.dot(normal) > 0:
                    shadow[v, u] = 0 if r.next_hit(p, ldir, inf, tex, shadow_ray=True)[0] >= inf else 1
                else:
                    shadow[v, u] = 2
    np.savez_compressed(os.path.join(HERE, "ref_example1_hits_64.npz"), W=np.int32(W), H=np.int32(H), cam_pos=pos, view=view, proj=proj,
                        hit_t=hit_t, hit_normal=hit_n, hit_mat=hit_mat, hit_light=hit_light, hit_shadow=shadow)
    print("example1: %d of %d pixels hit, %d voxel hits, shadow states %s" % (int(np.isfinite(hit_t).sum()), W * H,
                                                                              int((np.abs(hit_n[..., 1]) != 1).sum()), np.bincount(shadow.ravel())))


# ------------------------------------------------------------- ReSTIR reconnection shift
def section_shift():
    """Renderer.shift() (pathtracer.py:672-812), the reconnection shift of the ReSTIR-PT mode, on
    160 hand-built (destination vertex, source sample) pairs covering the four kinds of
    reconnection vertex (surface with continuation, last vertex, escape vertex, NEE-only) and all
    lobe codes. The zero-vector markers are set explicitly, so no NaN encodings are involved."""
    sys.path.insert(0, ROOT)
    import renderer.math_utils as mu
    from renderer.reservoir import Reservoir

    W, H, R = 32, 16, 32
    cfg = dict(voxel_edges=0.06, light_dir=(0.4, 0.8, 0.45), light_cone=0.05, light_color=(1.2, 1.1, 0.9), floor_height=-0.5,
               floor_color=(1.0, 1.0, 1.0), floor_material=1, background=(0.1, 0.1, 0.1))
    mat, col = render_scene(R, 3)
    r = make_reference_renderer(W, H, R, mat, col, cfg)
    cam = (0.7, 0.9, 1.6)
    r.set_camera_pos(*cam)
    rng = np.random.default_rng(2024)
    n = 160
    ids = np.array([1, 2, 11, 21, 32, 40, 50, 52, 54, 80, 82], np.int32)
    rows = np.zeros((n, 28), np.float32)
    out = np.zeros((n, 7), np.float32)
    for i in range(n):
        kind = i % 4  # 0 surface + continuation, 1 last vertex, 2 escape vertex, 3 surface, NEE invisible
        dst_pos = rng.uniform(-0.5, 0.5, 3)
        dst_n = unit(rng, 1)[0]
        src_pos = dst_pos + rng.normal(0, 0.05, 3)
        rc_pos = dst_pos + dst_n * rng.uniform(0.1, 0.6) + rng.normal(0, 0.2, 3)
        to_dst = dst_pos - rc_pos
        rc_n = unit(rng, 1)[0]
        if np.dot(rc_n, to_dst) < 0 and i % 8 < 6:  # mostly front-facing reconnections; some fail the N.L check
            rc_n = -rc_n
        inc = unit(rng, 1)[0]
        if np.dot(inc, rc_n) < 0:
            inc = -inc
        nee = np.asarray(cfg["light_dir"]) / np.linalg.norm(cfg["light_dir"]) + rng.normal(0, 0.01, 3)
        nee /= np.linalg.norm(nee)
        L = rng.uniform(0.0, 2.0, 3)
        if kind == 1:
            inc = np.zeros(3)
        if kind == 2:
            rc_n, inc, nee = np.zeros(3), np.zeros(3), np.zeros(3)
            rc_pos = unit(rng, 1)[0]  # a direction for escape vertices
            if np.dot(rc_pos, dst_n) < 0 and i % 8 < 6:
                rc_pos = -rc_pos
        if kind == 3 or (kind == 1 and i % 8 >= 4):
            nee = np.zeros(3)
        res = Reservoir()
        res.z.rc_pos, res.z.rc_normal, res.z.rc_incident_dir = vec(rc_pos), vec(rc_n), vec(inc)
        res.z.rc_incident_L, res.z.rc_NEE_dir = vec(L), vec(nee)
        rc_info = mu.encode_material(np.int32(rng.choice(ids)), vec(rng.uniform(0.1, 1.0, 3)))
        dst_info = mu.encode_material(np.int32(rng.choice(ids)), vec(rng.uniform(0.1, 1.0, 3)))
        res.z.rc_mat_info = rc_info
        res.z.lobes = np.int32(int(rng.integers(0, 3)) * 10 + int(rng.integers(0, 3)))
        res.z.cached_jacobian_term = np.float32(rng.uniform(0.05, 3.0))
        dst_mat, _ = mu.decode_material(r.mats.mat_list, dst_info)
        d, s, j = r.shift(vec(dst_pos), vec(dst_n), dst_mat, vec(src_pos), vec(dst_n), dst_mat, res)
        z = res.z
        rows[i, 0:3], rows[i, 3:6], rows[i, 6:9] = dst_pos, dst_n, src_pos
        rows[i, 9:12], rows[i, 12:15], rows[i, 15:18] = z.rc_pos.data, z.rc_normal.data, z.rc_incident_dir.data
        rows[i, 18:21], rows[i, 21:24] = z.rc_incident_L.data, z.rc_NEE_dir.data
        rows[i, 24], rows[i, 25] = z.cached_jacobian_term, float(z.lobes)
        rows[i, 26:28] = np.array([dst_info, rc_info], np.uint32).view(np.float32)
        out[i, 0:3], out[i, 3:6], out[i, 6] = d.data, s.data, j
    cfgo = {("cfg_" + k): np.asarray(v, np.float32) for k, v in cfg.items()}
    np.savez_compressed(os.path.join(HERE, "ref_shift.npz"), rows=rows, out=out, cam_pos=np.asarray(cam, np.float32), **cfgo)
    print("shift: %d probes, %d with non-zero Jacobian" % (n, int((out[:, 6] != 0).sum())))


# ------------------------------------------------------------------------- reservoir.py
def reservoir_probe_rows(n, seed):
    """Inputs of orc_reservoir_probe (oracle.cpp): two samples whose vectors are all non-zero, RIS
    weights (some zero, exercising the in_w > 0 branch), the random numbers of the three selection
    draws, a primary position and the neighbour's M."""
    rng = np.random.default_rng(seed)
    rows = np.zeros((n, 53), np.float32)
    for i in range(n):
        for k in (0, 21):
            rows[i, k:k + 3] = rng.uniform(0.0, 3.0, 3) * (0.0 if i % 16 == 15 and k == 21 else 1.0)  # F (black: finalize's p_hat < 1e-6)
            rows[i, k + 3:k + 6] = rng.uniform(-2.0, 2.0, 3)  # rc_pos
            rows[i, k + 6:k + 9], rows[i, k + 9:k + 12] = unit(rng, 1)[0], unit(rng, 1)[0]
            rows[i, k + 12:k + 15] = rng.uniform(0.0, 4.0, 3)
            rows[i, k + 15:k + 18] = unit(rng, 1)[0]
            rows[i, k + 18:k + 19] = np.array([rng.integers(0, 2 ** 32)], np.uint32).view(np.float32)
            rows[i, k + 19] = rng.uniform(0.01, 40.0)
            rows[i, k + 20] = float(int(rng.integers(0, 3)) * 10 + int(rng.integers(0, 3)))
        rows[i, 42:44] = rng.uniform(0.0, 2.0, 2) * (rng.random(2) > 0.15)
        rows[i, 44:46] = rng.random(2)
        rows[i, 46:49] = rng.uniform(-2.0, 2.0, 3)
        rows[i, 49] = rng.uniform(0.0, 2.0) * (rng.random() > 0.15)
        rows[i, 50] = rng.random()
        rows[i, 51], rows[i, 52] = float(rng.integers(1, 33)), rng.uniform(0.0, 5.0)
    return rows


def section_reservoir():
    """Reservoir.init / input_sample / update_cached_jacobian_term / merge / finalize_without_M /
    encode / decode (reservoir.py:41-141) on 192 reservoirs whose sample vectors are all non-zero
    (octahedral encodings of zero vectors, the part the upstream mode leaves undefined, are not
    exercised). ti.random() answers with the row's three selection numbers in call order."""
    sys.path.insert(0, ROOT)
    from renderer.reservoir import Reservoir, Sample

    rows = reservoir_probe_rows(192, 77)
    out = np.zeros((rows.shape[0], 28), np.float32)

    def sample_of(a):
        return Sample(F=vec(a[0:3]), rc_pos=vec(a[3:6]), rc_normal=vec(a[6:9]), rc_incident_dir=vec(a[9:12]), rc_incident_L=vec(a[12:15]),
                      rc_NEE_dir=vec(a[15:18]), rc_mat_info=a[18:19].view(np.uint32)[0], cached_jacobian_term=np.float32(a[19]),
                      lobes=np.int32(int(a[20])))

    for i, a in enumerate(rows):
        draws = [a[44], a[45], a[50]]
        ti.set_random_source(lambda name, q=draws: q.pop(0))
        A, B = sample_of(a), sample_of(a[21:])
        r = Reservoir()
        r.init()
        sel_a = r.input_sample(np.float32(a[42]), A)
        if a[42] <= 0:
            draws.pop(0)  # the draw sits inside the in_w > 0 branch
        sel_b = r.input_sample(np.float32(a[43]), B)
        if a[43] <= 0:
            draws.pop(0)
        r.update_cached_jacobian_term(vec(a[46:49]))
        other = Reservoir()
        other.init()
        other.z, other.M, other.weight = B, np.float32(a[51]), np.float32(a[52])
        sel_m = r.merge(other, np.float32(a[49]))
        r.finalize_without_M()
        out[i, 0:5] = float(sel_a), float(sel_b), float(sel_m), r.M, r.weight
        d = Reservoir()
        d.init()
        d.decode(r.encode())
        z = d.z
        out[i, 5:7] = d.M, d.weight
        out[i, 7:10], out[i, 10:13], out[i, 13:16] = z.F.data, z.rc_pos.data, z.rc_normal.data
        out[i, 16:19], out[i, 19:22], out[i, 22:25] = z.rc_incident_dir.data, z.rc_incident_L.data, z.rc_NEE_dir.data
        out[i, 25:26] = np.array([z.rc_mat_info], np.uint32).view(np.float32)
        out[i, 26], out[i, 27] = z.cached_jacobian_term, float(z.lobes)
    ti.set_random_source(None)
    np.savez_compressed(os.path.join(HERE, "ref_reservoir.npz"), rows=rows, out=out)
    print("reservoir: %d probes, selections %s, %d zero weights" % (rows.shape[0], out[:, 0:3].sum(0).astype(int).tolist(), int((out[:, 4] == 0).sum())))


# ------------------------------------------------------------------------- spatial_GRIS
def section_gris():
    """Renderer.spatial_GRIS(0, 24.0, 32, 1) (pathtracer.py:815-989, the call of accumulate() :1313)
    on hand-built buffers. The G-buffer is the reference's own: get_cast_dir + next_hit per pixel,
    then the writes of render() (:535-541). Every pixel's reservoir holds a sample whose vectors are
    all non-zero (a reconnection vertex with a continuation and a visible sun), packed with
    Reservoir.encode() as render() does (:607): the octahedral encodings of zero vectors, the part
    upstream leaves undefined, do not occur. The camera looks down so that no pixel sees the sky
    (the other hole). Out-of-image taps read zeros (Field.oob_zero) and fail the distance test.
    ti.random(): dimension 65 = radius shift, 66 + i = merge draw of tap i, 98 = canonical merge
    of the oracle's sampler, key (pixel, frame). The pass is run for every seventh pixel."""
    sys.path.insert(0, ROOT)
    import renderer.math_utils as mu
    from renderer.math_utils import eps, inf
    from renderer.reservoir import Reservoir, Sample
    from renderer.space_transformations import screen_to_view, view_to_screen, view_to_world, world_to_view
    from voxel_rt2_b200.camera import default_camera_matrices

    W, H, R, seed, frame, S = 48, 24, 32, 31, 5, 16
    cfg = dict(voxel_edges=0.06, light_dir=(0.4, 0.8, 0.45), light_cone=0.05, light_color=(1.2, 1.1, 0.9), floor_height=-0.5,
               floor_color=(0.9, 0.85, 0.8), floor_material=1, background=(0.1, 0.1, 0.1))
    mat, col = render_scene(R, 3)
    keep = np.random.default_rng(8).random(mat.shape) < 0.15  # mostly open floor: neighbouring pixels share a surface
    keep[R // 3: 2 * R // 3, : R // 4, R // 3: 2 * R // 3] = True
    mat = np.where(keep, mat, 0).astype(np.int8)
    r = make_reference_renderer(W, H, R, mat, col, cfg)
    sc, tr = synthetic_sky_tables(S)  # shift() looks the sun transmittance up for the reconnection vertex (:775-779)
    r.use_physical_atmosphere[None] = 1
    r.atmos.skybox_res = ti.Vector([S, S])
    r.atmos.skybox_fres = ti.Vector([1.0 / S, 1.0 / S])
    r.atmos.skybox_scattering = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    r.atmos.skybox_transmittance = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    r.atmos.skybox_scattering.arr[...] = sc
    r.atmos.skybox_transmittance.arr[...] = tr
    pos, view, proj = default_camera_matrices(W, H, pos=(0.5, 2.2, 1.1))
    set_reference_camera(r, pos, view, proj)
    r.current_frame = frame
    tex = r.world.voxel_color_texture
    cam = r.camera_pos[None].copy()
    rng = np.random.default_rng(4242)
    ids = np.array([1, 2, 11, 21, 32, 40, 50, 52, 54, 80, 82], np.int32)
    sun = np.asarray(cfg["light_dir"]) / np.linalg.norm(cfg["light_dir"])
    npx = W * H
    samples = np.zeros((npx, 23), np.float32)
    gbuf = np.zeros((npx, 7), np.float32)
    col_d = rng.uniform(0.0, 1.5, (npx, 3)).astype(np.float32)
    col_s = rng.uniform(0.0, 0.5, (npx, 3)).astype(np.float32)
    for f in (r.gbuff_normals, r.gbuff_depth, r.gbuff_mat_id, r.spatial_reservoirs):
        for g in (f.fields.values() if hasattr(f, "fields") else [f]):
            g.oob_zero = True
    for v in range(H):
        for u in range(W):
            i = v * W + u
            d = r.get_cast_dir(np.int32(u), np.int32(v))
            closest, normal, albedo, hl, iters, mid = r.next_hit(cam, d, inf, tex, shadow_ray=False)
            assert closest < inf, (u, v)
            hit_pos = cam + closest * d + normal * eps
            n_oct = mu.encode_unit_vector_3x16(normal)
            info = mu.encode_material(mid, albedo)
            depth = view_to_screen(world_to_view(hit_pos, r.view_mat[None]).xyz, r.proj_mat[None]).z
            r.gbuff_normals[u, v], r.gbuff_depth[u, v], r.gbuff_position[u, v], r.gbuff_mat_id[u, v] = n_oct, depth, hit_pos, info
            texcoord = (ti.Vector([np.int32(u), np.int32(v)]) + 0.5) * r.inv_image_res / r.render_scale[None]  # :823
            x1 = view_to_world(screen_to_view(texcoord, depth, r.proj_mat_inv[None]), r.view_mat_inv[None])  # :850-852
            gbuf[i, 0:3], gbuf[i, 3:5] = x1.data, n_oct.data.astype(np.float32)
            gbuf[i, 5:6] = np.array([info], np.uint32).view(np.float32)
            # the pixel's reservoir: a reconnection vertex above the surface, facing it
            p, n = np.asarray(hit_pos.data, np.float64), np.asarray(normal.data, np.float64)
            rc_pos = p + n * rng.uniform(0.5, 1.2) + rng.normal(0, 0.15, 3)
            rc_n = (p - rc_pos) / np.linalg.norm(p - rc_pos) - n + rng.normal(0, 0.3, 3)  # faces the pixel and, roughly, its neighbours
            rc_n /= np.linalg.norm(rc_n)
            if i % 8 == 7:
                rc_n = unit(rng, 1)[0]  # some reconnections fail the N.L checks
            inc = unit(rng, 1)[0]
            if np.dot(inc, rc_n) < 0:
                inc = -inc
            nee = sun + rng.normal(0, 0.01, 3)
            nee /= np.linalg.norm(nee)
            z = Sample(F=vec(rng.uniform(0.02, 1.5, 3)), rc_pos=vec(rc_pos), rc_normal=vec(rc_n), rc_incident_dir=vec(inc),
                       rc_incident_L=vec(rng.uniform(0.0, 2.0, 3)), rc_NEE_dir=vec(nee),
                       rc_mat_info=mu.encode_material(np.int32(rng.choice(ids)), vec(rng.uniform(0.1, 1.0, 3))),
                       cached_jacobian_term=np.float32(1.0), lobes=np.int32(int(rng.integers(0, 3)) * 10 + int(rng.integers(0, 3))))
            res = Reservoir()
            res.init()
            res.z, res.M, res.weight = z, np.float32(1.0), np.float32(rng.uniform(0.0, 3.0) * (rng.random() > 0.1))
            res.update_cached_jacobian_term(hit_pos)  # :553
            r.spatial_reservoirs[u, v, 0] = res.encode()
            r.color_buffer[u, v], r.color_buffer_specular[u, v] = vec(col_d[i]), vec(col_s[i])
            samples[i, 0:3], samples[i, 3:6], samples[i, 6:9] = res.z.F.data, res.z.rc_pos.data, res.z.rc_normal.data
            samples[i, 9:12], samples[i, 12:15], samples[i, 15:18] = res.z.rc_incident_dir.data, res.z.rc_incident_L.data, res.z.rc_NEE_dir.data
            samples[i, 18:19] = np.array([res.z.rc_mat_info], np.uint32).view(np.float32)
            samples[i, 19], samples[i, 20], samples[i, 21], samples[i, 22] = res.z.cached_jacobian_term, float(res.z.lobes), res.M, res.weight
    pixels = np.arange(1, npx, int(os.environ.get("GRIS_STEP", "7")), dtype=np.int32)
    r.gbuff_position._struct_for = lambda: iter([(np.int32(i % W), np.int32(i // W)) for i in pixels])
    state = {}

    def source(name, fr):
        f = fr
        while f is not None and f.f_code.co_name != "spatial_GRIS":
            f = f.f_back
        assert f is not None, name
        u, v = int(f.f_locals["u"]), int(f.f_locals["v"])
        if name == "spatial_GRIS":
            k = state[(u, v)] = state.get((u, v), 0) + 1
            assert k <= 2
            return 0.5 if k == 1 else sampler_rnd(v * W + u, frame, seed, 65)  # start_index (unused upstream), radius_shift
        assert name == "merge", name
        canonical = np.array_equal(fr.f_locals["in_r"].z.rc_pos.data, f.f_locals["center_reservoir"].z.rc_pos.data)  # rc_pos is unique per pixel
        dim = 98 if canonical else 66 + int(f.f_locals["i"])
        state[(u, v, dim)] = state.get((u, v, dim), 0) + 1
        assert state[(u, v, dim)] == 1, (u, v, dim)
        return sampler_rnd(v * W + u, frame, seed, dim)

    ti.set_random_source(source, with_frame=True)
    r.spatial_GRIS(0, 24.0, 32, 1, tex)
    ti.set_random_source(None)
    out_d = np.stack([r.color_buffer.arr[i % W, i // W] for i in pixels])
    out_s = np.stack([r.color_buffer_specular.arr[i % W, i // W] for i in pixels])
    merges = sum(1 for k in state if len(k) == 3 and k[2] != 98)
    print("gris: %d pixels, %d neighbour merges drawn, mean out %.4f (in %.4f)" % (len(pixels), merges, float((out_d + out_s).mean()),
                                                                                  float((col_d + col_s)[pixels].mean())))
    out = dict(material=mat, color=col, cam_pos=pos, view=view, proj=proj, seed=np.int32(seed), frame=np.int32(frame), W=np.int32(W),
               H=np.int32(H), sky_res=np.int32(S), sky_scatter=sc, sky_trans=tr, samples=samples, gbuf=gbuf, col_d=col_d, col_s=col_s, pixels=pixels, out_d=out_d, out_s=out_s)
    for k, v in cfg.items():
        out["cfg_" + k] = np.asarray(v, np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_gris.npz"), **out)


# ------------------------------------------------------------------------- render(), ReSTIR branch
def section_restir_render():
    """Renderer.render() with USE_RESTIR_PT = True (pathtracer.py:15,355-632): the branch that builds
    each pixel's input reservoir (reconnection vertex bookkeeping :405-520, the NEE-vs-BSDF RIS at
    the primary vertex :556-605) and writes the G-buffer and the canonical integrands. The module
    constant is switched on for this section; Reservoir.encode is intercepted, so what is stored is
    the reservoir BEFORE packing (zero-vector markers are exact zeros there, no 0/0 encodings).
    ti.random(): the path dimensions of section_render plus dimension 40 for input_sample's draw."""
    sys.path.insert(0, ROOT)
    import renderer.pathtracer as pt
    from renderer.reservoir import Reservoir, StorageReservoir
    from voxel_rt2_b200.camera import default_camera_matrices

    W, H, R, seed, n_samples, S = 32, 16, 32, 59, 2, 16
    cfg = dict(voxel_edges=0.06, light_dir=(1.0, 1.0, 0.4), light_cone=0.06, light_color=(1.2, 1.1, 0.9), floor_height=-0.55,
               floor_color=(0.8, 0.75, 0.7), floor_material=1, background=(0.25, 0.35, 0.55))
    mat, col = render_scene(R, 6)
    r = make_reference_renderer(W, H, R, mat, col, cfg)
    sc, tr = synthetic_sky_tables(S)
    r.use_physical_atmosphere[None] = 1
    r.atmos.skybox_res = ti.Vector([S, S])
    r.atmos.skybox_fres = ti.Vector([1.0 / S, 1.0 / S])
    r.atmos.skybox_scattering = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    r.atmos.skybox_transmittance = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    r.atmos.skybox_scattering.arr[...] = sc
    r.atmos.skybox_transmittance.arr[...] = tr
    pos, view, proj = default_camera_matrices(W, H, pos=(0.9, 0.8, 1.9))
    set_reference_camera(r, pos, view, proj)
    tex = r.world.voxel_color_texture
    npx = W * H
    state = {"sample": 0, "count": {}}
    path_source = path_random_source(W, seed, state)
    captured = {}

    def source(name, frame):
        if name == "input_sample":
            f = frame
            while f.f_code.co_name != "render":
                f = f.f_back
            return sampler_rnd(int(f.f_locals["v"]) * W + int(f.f_locals["u"]), state["sample"], seed, 40)
        return path_source(name, frame)

    def capture(self):
        f = sys._getframe(1)
        while f.f_code.co_name != "render":
            f = f.f_back
        z = self.z
        row = np.zeros(23, np.float32)
        row[0:3], row[3:6], row[6:9], row[9:12] = z.F.data, z.rc_pos.data, z.rc_normal.data, z.rc_incident_dir.data
        row[12:15], row[15:18] = z.rc_incident_L.data, z.rc_NEE_dir.data
        row[18:19] = np.array([z.rc_mat_info], np.uint32).view(np.float32)
        row[19], row[20], row[21], row[22] = z.cached_jacobian_term, float(z.lobes), self.M, self.weight
        captured[int(f.f_locals["v"]) * W + int(f.f_locals["u"])] = row
        return StorageReservoir()

    samples = np.zeros((n_samples, npx, 23), np.float32)
    gbuf = np.zeros((n_samples, npx, 7), np.float32)
    col_d = np.zeros((n_samples, npx, 3), np.float32)
    col_s = np.zeros((n_samples, npx, 3), np.float32)
    saved_encode, saved_flag = Reservoir.encode, pt.USE_RESTIR_PT
    Reservoir.encode, pt.USE_RESTIR_PT = capture, True
    ti.set_random_source(source, with_frame=True)
    try:
        with np.errstate(all="ignore"):
            for s in range(n_samples):
                state["sample"], state["count"] = s, {}
                captured.clear()
                r.render(tex)
                assert len(captured) == npx
                for i in range(npx):
                    u, v = i % W, i // W
                    samples[s, i] = captured[i]
                    gbuf[s, i, 0:3] = r.gbuff_position.arr[u, v]
                    gbuf[s, i, 3:5] = r.gbuff_normals.arr[u, v].astype(np.float32)
                    gbuf[s, i, 5:6] = np.array([r.gbuff_mat_id.arr[u, v]], np.uint32).view(np.float32)
                    col_d[s, i], col_s[s, i] = r.color_buffer.arr[u, v], r.color_buffer_specular.arr[u, v]
                esc = (np.abs(samples[s, :, 6:9]).sum(1) == 0).sum()
                print("restir render sample %d: %d escape rc vertices, %d NEE-visible, mean W %.3f" % (
                    s, int(esc), int((np.abs(samples[s, :, 15:18]).sum(1) > 0).sum()), float(np.nanmean(samples[s, :, 22]))))
    finally:
        ti.set_random_source(None)
        Reservoir.encode, pt.USE_RESTIR_PT = saved_encode, saved_flag
    out = dict(material=mat, color=col, cam_pos=pos, view=view, proj=proj, seed=np.int32(seed), W=np.int32(W), H=np.int32(H),
               sky_res=np.int32(S), sky_scatter=sc, sky_trans=tr, samples=samples, gbuf=gbuf, col_d=col_d, col_s=col_s)
    for k, v in cfg.items():
        out["cfg_" + k] = np.asarray(v, np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_restir_render.npz"), **out)


# ------------------------------------------------------------------------- atmos.py (sky)
def section_sky():
    """renderer/atmos.py: (1) 640 entries of the transmittance LUT from generate_transmittance_lut,
    (2) the whole precompute pipeline of Scene.finish (cloud ambient -> accumulate_clouds x passes ->
    compute_skybox, scene.py:243-253) on a 6 x 6 table, ti.random() answering from the oracle's
    per-texel counter sampler (key = (texel, pass), running counter) so tables compare per texel,
    (3) project_sky / unproject_sky / sample_skybox(_transmittance) lookups.
    The pipeline reads the LUT at arbitrary entries: after (1) the reference's LUT field is filled
    with the FULL table the reference's own generate_transmittance_lut produced through the emulator
    (tests/golden/ref_lut_full.npz, made by make_ref_lut.py in parallel worker processes; stored
    again in this fixture as lut_full). The oracle side of the comparison uses its own table."""
    sys.path.insert(0, ROOT)
    from oracle import binding as oracle
    from renderer.atmos import Atmos
    from voxel_rt2_b200.materials import material_table

    out = {}
    atm = Atmos()
    atm.load_textures()
    rng = np.random.default_rng(99)
    # (1) LUT subset: edges, horizon band, random interior
    sub = {(x, y) for x in (0, 1, 127, 128, 129, 254, 255) for y in (0, 1, 63, 126, 127)}
    sub |= {(int(x), int(y)) for x, y in zip(rng.integers(0, 256, 400), rng.integers(0, 128, 400))}
    sub |= {(int(x), int(y)) for x, y in zip(rng.integers(120, 140, 205), rng.integers(0, 128, 205))}
    sub = sorted(sub)
    atm.trans_LUT._struct_for = lambda: iter([(np.int32(x), np.int32(y)) for x, y in sub])
    atm.generate_transmittance_lut()
    del atm.trans_LUT._struct_for
    idx = np.array(sub, np.int32)
    out["lut_idx"] = idx
    out["lut_val"] = atm.trans_LUT.arr[idx[:, 0], idx[:, 1]].copy()
    print("sky: %d LUT entries" % len(sub))

    # (2) pipeline on an S x S table
    S, passes, seed = 6, 2, 5
    sun_dir, cone, sun_col = (0.6, 0.5, -0.62), 0.05, (1.3, 1.2337, 1.2181)
    tex = np.load(os.path.join(ROOT, "voxel_rt2_b200", "assets", "cloud_texture.npz"))["tex"]
    assert np.array_equal(tex, atm.cloud_tex.arr), "the packaged cloud texture differs from ti.tools.imread's decode"
    o = oracle.OracleRenderer(dx=2 / 16, image_res=(16, 16), grid_res=16, sky_res=S, cloud_passes=passes, seed=seed,
                              materials=material_table(), cloud_tex=tex)
    o.set_directional_light(sun_dir, cone, sun_col)
    o.set_use_physical_sky(True, True)
    o.prepare_data()
    lut = np.load(os.path.join(HERE, "ref_lut_full.npz"))["lut"]
    assert np.array_equal(lut[idx[:, 0], idx[:, 1]].view(np.uint16), out["lut_val"].view(np.uint16))  # same code, same entries
    out["lut_full"] = lut
    out["lut_oracle_diff"] = np.int32((o.get_trans_lut().view(np.uint16) != lut.view(np.uint16)).any(-1).sum())
    atm.trans_LUT.arr[...] = lut
    atm.skybox_res = ti.Vector([S, S])
    atm.skybox_fres = ti.Vector([1.0 / S, 1.0 / S])
    atm.skybox_scattering = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    atm.skybox_transmittance = ti.Vector.field(3, dtype=ti.f32, shape=(S, S))
    atm.use_clouds[None] = 1
    d = np.asarray(sun_dir, np.float64)
    light_dir = vec((d / np.sqrt((d * d).sum())).astype(np.float32))   # set_directional_light (pathtracer.py:138-144)
    light_col = vec(np.asarray(sun_col, np.float32)) * 3.0                # light_color * light_weight
    cosmax = np.float32(np.cos(cone * 0.5))
    state = {"pass": 0, "count": {}}

    def source(name, frame):
        f = frame
        while f is not None and f.f_code.co_name not in ("accumulate_clouds", "compute_skybox", "compute_cloud_ambient"):
            f = f.f_back
        k = f.f_code.co_name
        if k == "compute_cloud_ambient":
            texel, pas = 0xFFFFFFFF, 2000
        elif k == "accumulate_clouds":
            texel, pas = int(f.f_locals["u"]) * S + int(f.f_locals["v"]), state["pass"]
        else:
            off = f.f_locals["offset"]
            texel, pas = int(off[0]) * S + int(off[1]), 1000
        n = state["count"].get((texel, pas), 0)
        state["count"][(texel, pas)] = n + 1
        return sampler_rnd(texel, pas, seed, n)

    ti.set_random_source(source, with_frame=True)
    atm.compute_cloud_ambient(light_dir, light_col, cosmax)
    out["cloud_ambient"] = atm.cloud_ambient.arr.copy()
    for p in range(passes):
        state["pass"] = p
        atm.accumulate_clouds(light_dir, light_col, cosmax, passes)
    out["clouds_scatter"], out["clouds_trans"] = atm.skybox_scattering.arr.copy(), atm.skybox_transmittance.arr.copy()
    for sl in range(S):
        atm.compute_skybox(light_dir, light_col, cosmax, sl, S)
    ti.set_random_source(None)
    out["sky_scatter"], out["sky_trans"] = atm.skybox_scattering.arr.copy(), atm.skybox_transmittance.arr.copy()
    out.update(S=np.int32(S), passes=np.int32(passes), seed=np.int32(seed), sun_dir=np.asarray(sun_dir, np.float32), cone=np.float32(cone),
               sun_col=np.asarray(sun_col, np.float32))
    print("sky: %dx%d tables, mean scatter %.4f trans %.4f, %d draws" % (S, S, out["sky_scatter"].mean(), out["sky_trans"].mean(),
                                                                        sum(state["count"].values())))
    # (3) parameterisation and lookups on those tables
    n = 96
    dirs = unit(rng, n)
    dirs[:4] = [[0, 1, 0], [0, -1, 0], [1, 0, 0], [0.3, 1e-4, -0.95]]
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    uv = np.array([atm.project_sky(vec(x)).data for x in dirs], np.float32)
    uv_in = rng.random((n, 2)).astype(np.float32)
    back = np.array([atm.unproject_sky(vec(x)).data for x in uv_in], np.float32)
    tr = np.array([atm.sample_skybox_transmittance(vec(x)).data for x in dirs], np.float32)
    jit = rng.random((n, 3)).astype(np.float32)
    sc2, tr2 = np.zeros((n, 3), F), np.zeros((n, 3), F)
    for i in range(n):
        q = list(jit[i])
        ti.set_random_source(lambda name: q.pop(0))
        a, b = atm.sample_skybox(vec(dirs[i]))
        sc2[i], tr2[i] = a.data, b.data
    ti.set_random_source(None)
    out.update(dirs=dirs, project_uv=uv, unproject_in=uv_in, unproject_dir=back, lookup_trans=tr, lookup_jitter=jit, lookup_scatter_j=sc2,
               lookup_trans_j=tr2)
    np.savez_compressed(os.path.join(HERE, "ref_sky.npz"), **out)


SECTIONS = {"raytrace": section_raytrace, "math": section_math, "bsdf": section_bsdf, "render": section_render, "frame": section_frame,
            "shift": section_shift, "voxel": section_voxel, "moving": section_moving, "example1": section_example1, "sky": section_sky, "reservoir": section_reservoir, "gris": section_gris, "restir_render": section_restir_render}

if __name__ == "__main__":
    for s in (sys.argv[1:] or list(SECTIONS)):
        SECTIONS[s]()
