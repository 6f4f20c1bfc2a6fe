"""Known-answer and property tests that pin the CPU oracle (the reference ships no tests or
golden vectors; SURVEY.md §4 lists what has to be derived by hand from the source)."""
import math

import numpy as np
import pytest

import scenes


def _orc(oracle, R=128, res=(64, 64), **kw):
    from voxel_rt2_b200.materials import material_table

    return oracle.OracleRenderer(dx=2.0 / R, image_res=res, grid_res=R, sky_res=kw.pop("sky_res", 0), materials=material_table(), **kw)


def test_main_scene_ray_known_answer(oracle):
    """main.py:13 puts one voxel at index (0,0,0) => grid cell (64,64,64). A ray along -z from
    world (0.0078, 0.0078, 2) = voxel (64.4992, 64.4992, 192) enters the box at t=64, walks empty
    space and hits the +z face of cell (64,64,64) at t_voxel = 192 - 65 = 127."""
    o = _orc(oracle)
    o.set_voxels(*scenes.main_scene())
    org = np.array([[0.0078 * 64 + 64, 0.0078 * 64 + 64, 2 * 64 + 64]], np.float32)
    t, cell, n, it = o.raytrace(org, np.array([[0, 0, -1]], np.float32))
    assert t[0] == 127.0
    assert cell[0].tolist() == [64, 64, 64]
    assert (n[0] + 0.0).tolist() == [0.0, 0.0, 1.0]
    assert 1 <= it[0] <= 64
    # a ray that passes one voxel to the side misses
    t2, *_ = o.raytrace(org + np.array([[1.0, 0, 0]], np.float32), np.array([[0, 0, -1]], np.float32))
    assert np.isinf(t2[0])


def test_ray_inside_occupied_voxel_hits_at_eps_with_initial_normal(oracle):
    """a ray born inside an occupied voxel hits at t = 1e-6 with the normal of the
    axis of largest |p - R/2| (raytracer.py:88-101)."""
    o = _orc(oracle, R=32)
    mat, col = scenes.empty(32)
    mat[20, 10, 12] = 1
    o.set_voxels(mat, col)
    t, cell, n, it = o.raytrace(np.array([[20.5, 10.5, 12.5]], np.float32), np.array([[0.6, 0.0, 0.8]], np.float32))
    assert t[0] == np.float32(1e-6) and it[0] == 0
    assert cell[0].tolist() == [20, 10, 12]
    # |p - 16| = (4.5, 5.5, 3.5) -> y axis; flipped against the ray only if d.n > 0 (d.y = 0 here)
    assert np.abs(n[0]).tolist() == [0.0, 1.0, 0.0]


def test_occupancy_pyramid_is_or_of_children(oracle):
    R = 32
    o = _orc(oracle, R=R)
    mat, col = scenes.random_grid(R, 0.01, 4)
    mat[5, 6, 7] = -3  # negative material = empty
    o.set_voxels(mat, col)
    occ0 = mat > 0
    for x, y, z in [(5, 6, 7), (0, 0, 0), (31, 31, 31), (12, 3, 30)]:
        assert o.occupancy(x, y, z, 0) == int(occ0[x, y, z])
    cur = occ0
    for lod in range(1, 5):
        r = R >> lod
        cur = cur.reshape(r, 2, r, 2, r, 2).any(axis=(1, 3, 5))
        rng = np.random.default_rng(lod)
        for x, y, z in rng.integers(0, r, (64, 3)):
            assert o.occupancy(x, y, z, lod) == int(cur[x, y, z])
    assert o.occupancy(0, 0, 0, 5) == -1  # top LOD is log2(R)-1


def _brute_force_first_hit(occ, o, d, R, tmax=400.0):
    """Independent check: Amanatides-Woo LOD-0 walk in float64."""
    o = np.asarray(o, np.float64)
    d = np.asarray(d, np.float64)
    # enter the box
    t0, t1 = -np.inf, np.inf
    for i in range(3):
        if d[i] != 0:
            a, b = (0 - o[i]) / d[i], (R - o[i]) / d[i]
            t0, t1 = max(t0, min(a, b)), min(t1, max(a, b))
    if t0 > t1 or t1 < 0:
        return None
    t = max(t0, 0.0) + 1e-9
    p = o + d * t
    c = np.clip(np.floor(p), 0, R - 1).astype(int)
    while True:
        if occ[c[0], c[1], c[2]]:
            return tuple(c)
        tn = np.full(3, np.inf)
        for i in range(3):
            if d[i] > 0:
                tn[i] = (c[i] + 1 - o[i]) / d[i]
            elif d[i] < 0:
                tn[i] = (c[i] - o[i]) / d[i]
        i = int(np.argmin(tn))
        c[i] += 1 if d[i] > 0 else -1
        if c[i] < 0 or c[i] >= R:
            return None


def test_hierarchical_dda_equals_brute_force_first_hit(oracle):
    """Hit voxel of the hierarchical DDA == first occupied cell along the ray, for rays in
    general position (ties and grazing rays are excluded by the random directions; SURVEY
    Appendix C measured ~1e-5 disagreement from float rounding, so allow 0.2 %)."""
    R = 32
    o = _orc(oracle, R=R)
    mat, col = scenes.random_grid(R, 0.03, 21)
    o.set_voxels(mat, col)
    occ = mat > 0
    rng = np.random.default_rng(0)
    n = 1500
    org = rng.uniform(-8, R + 8, (n, 3)).astype(np.float32)
    tgt = rng.uniform(4, R - 4, (n, 3)).astype(np.float32)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    t, cell, nrm, it = o.raytrace(org, d)
    bad = 0
    hits = 0
    for k in range(n):
        ref = _brute_force_first_hit(occ, org[k], d[k], R)
        got = tuple(cell[k]) if np.isfinite(t[k]) else None
        hits += ref is not None
        if ref != got:
            bad += 1
    assert hits > 0.5 * n
    assert bad <= 0.002 * n, "%d of %d rays disagree with the brute-force walk" % (bad, n)
    # face normals are axis aligned and oppose the ray
    fin = np.isfinite(t)
    assert np.all(np.abs(nrm[fin]).sum(axis=1) >= 1)
    assert np.all((nrm[fin] * d[fin]).sum(axis=1) <= 0)


def test_floor_hit_known_answer(oracle):
    """Floor plane y = h: t = (h - o.y)/d.y, normal +y, accepted only within radius 10 of the
    diagonal."""
    o = _orc(oracle, R=32, res=(8, 4))
    o.set_voxels(*scenes.empty(32))
    o.set_floor(-0.05, (0.2, 0.4, 0.6), 1)
    o.set_camera_pos(0.0, 1.0, 0.0)
    o.set_look_at(0.0, 0.0, -1e-3)
    o.prepare_data()
    h = o.trace_primary()
    c = h[2, 4]  # a pixel near the image centre looks (almost) straight down
    assert (c["flags"] & 255) == 1
    assert abs(c["t"] - 1.05) < 0.02
    assert c["normal"].tolist() == [0.0, 1.0, 0.0]
    assert c["cell"].tolist() == [-1, -1, -1]


def test_material_table_rows(oracle):
    from voxel_rt2_b200.materials import material_table

    t = material_table()
    assert t.shape == (128, 14) and t.dtype == np.float32
    default = [1, 1, 1, 0, 0, 0.04, 0, 0.9, 0, 0, 0, 0, 0, 0]  # materials.py:50-63
    for mid in (0, 1, 2, 3, 127):
        assert np.allclose(t[mid], default)
    assert np.allclose(t[50, 3:9], [0, 1, 0.8, 0, 0.4, 0])         # rough metal
    assert np.allclose(t[54, [4, 5, 7, 11, 12]], [0.7, 0.8, 0.3, 0.7, 0.9])  # car paint
    assert np.allclose(t[82, [3, 5, 9, 10]], [0.95, 0.0, 0.9, 0.4])  # cloth


def test_material_table_matches_reference_csv():
    import os

    p = "/root/reference/default_material_set.csv"
    if not os.path.exists(p):
        pytest.skip("reference tree not mounted")
    from voxel_rt2_b200.materials import load_csv, material_table

    assert np.array_equal(load_csv(p), material_table())


def _probe(oracle, mat_ids, v, n, l, u3, albedo=None):
    o = _orc(oracle, R=8, res=(8, 4))
    k = len(mat_ids)
    albedo = np.tile(np.array([[0.8, 0.6, 0.4]], np.float32), (k, 1)) if albedo is None else albedo
    return o.bsdf_probe(np.asarray(mat_ids, np.int32), albedo, v, n, l, u3)


def test_lobe_probabilities_and_sampler_choice(oracle):
    """bsdf.py:351-363: diffuse = (1-metallic)*clamp(1-specular,0.4,0.9), spec = 1-diffuse,
    clearcoat = 0.7*clearcoat, normalised. The lobe id returned by sample_disney follows the
    thresholds u <= d, u <= d+s."""
    ids = [1, 50, 21, 54]
    exp = []
    from voxel_rt2_b200.materials import material_table

    t = material_table()
    for m in ids:
        met, spec, cc = t[m, 4], t[m, 5], t[m, 11]
        d = (1 - met) * min(max(1 - spec, 0.4), 0.9)
        s = 1 - d
        c = 0.7 * cc
        w = d + s + c
        exp.append((d / w, s / w, c / w))
    v = np.tile(np.array([[0.3, 0.8, 0.52]], np.float32), (len(ids), 1))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    n = np.tile(np.array([[0, 1, 0]], np.float32), (len(ids), 1))
    l = np.tile(np.array([[-0.2, 0.9, 0.39]], np.float32), (len(ids), 1))
    l /= np.linalg.norm(l, axis=1, keepdims=True)
    for u, want in ((0.0, None), (0.999999, None)):
        u3 = np.tile(np.array([[u, 0.3, 0.6]], np.float32), (len(ids), 1))
        out = _probe(oracle, ids, v, n, l, u3)
        for k, (d, s, c) in enumerate(exp):
            lobe = int(out[k, 11])
            expect = 0 if u <= d else (1 if u <= d + s else 2)
            assert lobe == expect
    # metal: diffuse weight 0 -> u=0.0 is still "<= 0" -> diffuse lobe is chosen with pdf 0 -> pdf=1 fixup? no: pdf = 0*...
    # (kept as the reference behaves; only checked for consistency above)


def test_diffuse_material_eval_known_answer(oracle):
    """Default material, normal incidence: F_L = F_V = 0 so f_d = albedo/pi * (1-metallic);
    pdf_disney's diffuse part = cos/pi * 0.9."""
    v = np.array([[0, 1, 0]], np.float32)
    out = _probe(oracle, [1], v, v, v, np.array([[0.5, 0.5, 0.5]], np.float32))
    assert np.allclose(out[0, 0:3], np.array([0.8, 0.6, 0.4]) / math.pi, rtol=1e-5)
    assert out[0, 6] >= 0.9 / math.pi  # + the specular lobe's share


def test_sampled_directions_are_unit_and_pdf_positive(oracle):
    rng = np.random.default_rng(3)
    k = 512
    ids = rng.choice([1, 10, 11, 20, 21, 22, 30, 32, 40, 50, 51, 53, 54, 80, 82], k)
    n = np.tile(np.array([[0, 0, 1]], np.float32), (k, 1))
    v = rng.normal(size=(k, 3)).astype(np.float32)
    v[:, 2] = np.abs(v[:, 2]) + 0.05
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    l = rng.normal(size=(k, 3)).astype(np.float32)
    l[:, 2] = np.abs(l[:, 2]) + 0.05
    l /= np.linalg.norm(l, axis=1, keepdims=True)
    u3 = rng.random((k, 3)).astype(np.float32)
    out = _probe(oracle, ids, v, n, l, u3)
    assert np.isfinite(out).all()
    nrm = np.linalg.norm(out[:, 7:10], axis=1)
    not_cc = out[:, 11] != 2  # the clear-coat sampler clamps sin and cos separately (bsdf.py:205-206): not unit
    assert np.allclose(nrm[not_cc], 1.0, atol=1e-4)
    assert np.all(np.abs(nrm - 1.0) < 0.05)
    assert (out[:, 10] > 0).all() and (out[:, 6] > 0).all()
    assert (out[:, 0:6] >= 0).all()


def test_cone_pdf_and_power_heuristic_formulas(oracle):
    # math_utils.py:61-63 and pathtracer.py:349-353, through the radiance of a 1-voxel scene is
    # overkill; check the closed forms the kernels rely on
    cosmax = math.cos(0.025 * 0.5)
    pdf = 1.0 / (2 * math.pi * (1 - cosmax))
    assert pdf > 1e3
    a, b = 2.0, 1.0
    assert abs(a * a / max(a * a + b * b, 1e-4) - 0.8) < 1e-12


def test_rng_is_uniform_and_deterministic(oracle):
    lib = oracle.load()
    xs = np.array([lib.orc_rnd(p, s, 7, d) for p in range(50) for s in range(10) for d in range(8)])
    assert xs.min() >= 0.0 and xs.max() < 1.0
    assert abs(xs.mean() - 0.5) < 0.02 and abs(xs.var() - 1 / 12) < 0.01
    assert lib.orc_rnd(3, 4, 5, 6) == lib.orc_rnd(3, 4, 5, 6)
    assert lib.orc_rnd(3, 4, 5, 6) != lib.orc_rnd(3, 4, 5, 7)


def test_f16_conversion_matches_numpy(oracle):
    lib = oracle.load()
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.normal(size=300), rng.normal(size=100) * 1e-6, rng.normal(size=100) * 7e4,
                           [0.0, -0.0, 1.0, 65504.0, 65520.0, 1e-8, 6.1e-5, 5.96e-8]]).astype(np.float32)
    for x in vals:
        want = np.float16(x)
        got = np.uint16(lib.orc_f32_to_f16(float(x))).view(np.float16)
        assert got == want or (np.isinf(got) and np.isinf(want)), (x, got, want)
        back = lib.orc_f16_to_f32(int(np.float16(x).view(np.uint16)))
        assert back == np.float32(np.float16(x)) or np.isinf(back)


def test_sky_projection_round_trip(oracle):
    """project_sky(unproject_sky(uv)) == uv (atmos.py:428-455)."""
    lib = oracle.load()
    import ctypes as C

    S = 256
    rng = np.random.default_rng(2)
    uv = rng.uniform(0.02, 0.98, (400, 2)).astype(np.float32)
    d = np.empty((400, 3), np.float32)
    back = np.empty((400, 2), np.float32)
    fp = C.POINTER(C.c_float)
    lib.orc_unproject_sky(400, S, uv.ctypes.data_as(fp), d.ctypes.data_as(fp))
    lib.orc_project_sky(400, S, d.ctypes.data_as(fp), back.ctypes.data_as(fp))
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    assert np.abs(back - uv).max() < 2e-3  # sqrt parametrisation is ill-conditioned near the horizon
    # zenith maps to the top edge, horizon to the middle row
    z = np.array([[0.0, 1.0, 1e-4], [1.0, 0.0, 0.0]], np.float32)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    out = np.empty((2, 2), np.float32)
    lib.orc_project_sky(2, S, z.ctypes.data_as(fp), out.ctypes.data_as(fp))
    assert out[0, 1] > 0.97 and abs(out[1, 1] - 0.5) < 1e-3


def test_transmittance_lut_physics(oracle):
    """atmos.py:462-498: looking up from sea level is nearly transparent, blue is attenuated
    more than red (Rayleigh), and transmittance increases with altitude."""
    o = _orc(oracle, R=8, res=(8, 4), sky_res=8)
    lut = o.get_trans_lut().astype(np.float32)
    up_sea = lut[255, 0]
    assert 0.5 < up_sea[2] < up_sea[0] < 1.0
    assert (lut[255, 100] >= lut[255, 0] - 1e-3).all()
    horiz = lut[128, 0]
    assert (horiz < up_sea).all()


def test_emissive_primary_hit_returns_quantised_albedo(oracle):
    """a primary ray that hits a material-2 voxel returns emission = albedo
    re-quantised through encode_material (u32(albedo*255)/255)."""
    R = 32
    o = _orc(oracle, R=R, res=(8, 4), jitter=False)
    mat, col = scenes.empty(R)
    mat[:, :, 10:14] = 2
    col[:, :, 10:14] = (200, 100, 50)
    o.set_voxels(mat, col)
    o.set_floor(-1e5, (1, 1, 1))
    o.set_use_physical_sky(False)
    o.prepare_data()
    o.accumulate(2)
    hdr = o.fetch_hdr()
    px = hdr[2, 4, :3]
    # voxel_edges = 0.06 darkens edges only; the centre pixel sees a face interior
    want = np.floor(np.array([200, 100, 50], np.float32) / 255.0 * 255.0) / 255.0
    assert np.allclose(px, want, atol=1.0 / 255.0)


def test_white_sky_furnace_energy_bound(oracle):
    """A diffuse floor under a uniform background and a black sun: the estimator must stay
    below the incoming radiance (no energy gain) and be well above zero."""
    R = 32
    o = _orc(oracle, R=R, res=(16, 8))
    o.set_voxels(*scenes.empty(R))
    o.set_floor(-0.5, (1.0, 1.0, 1.0), 1)
    o.set_background_color((1.0, 1.0, 1.0))
    o.set_camera_pos(0.0, 0.5, 0.0)
    o.set_look_at(0.0, -1.0, -0.3)
    o.prepare_data()
    o.accumulate(256)
    hdr = o.fetch_hdr()[..., :3]
    m = hdr.mean()
    assert 0.5 < m < 1.05, m


def test_tonemap_known_points(oracle):
    """math_utils.py:160-186 + pathtracer.py:636-661: black stays black, the image centre has no
    vignette (uv = (0.5, 0.5) only at even sizes), large values saturate to 1."""
    o = _orc(oracle, R=8, res=(8, 4), exposure=1.0)
    hdr = np.zeros((4, 8, 4), np.float32)
    hdr[2, 4, :3] = 0.22          # pixel (i=4, j=2): uv = (0.5, 0.5) -> darken = 1
    hdr[0, 0, :3] = 1000.0
    ldr = o.tonemap(hdr)
    assert np.all(ldr[1, 1, :3] == 0.0) and np.all(ldr[..., 3] == 1.0)
    # uchimura(0.22) with m = 0.22: w0 = 0, linear section: L = m + a(x - m) = 0.22 -> ^(1/2.2)
    assert np.allclose(ldr[2, 4, :3], 0.22 ** (1 / 2.2), atol=2e-4)
    assert np.allclose(ldr[0, 0, :3], 1.0, atol=1e-3)


def test_oracle_temporal_reservoir_reuse_is_stable_and_helps(oracle):
    """The oracle's temporal reservoir reuse (oracle.cpp temporal_reuse_pixel; no upstream counterpart) on the
    example3 fixture at 64 x 48: the chain accumulates confidence (M grows to the cap), the accumulated image stays
    within 12 % of a path-traced mean and 6 % of the spatial-only mean, the per-frame error drops against the spatial pass alone, and
    reset_framebuffer drops the history (the next frame equals the first frame of a fresh chain)."""
    import os

    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example3_seed0.npz"))

    def mk(seed):
        o = oracle.OracleRenderer(dx=1.0 / 64, image_res=(64, 48), grid_res=128, sky_res=0, seed=seed, voxel_edges=0.0, exposure=30.0,
                                  materials=material_table())
        o.set_voxels(z["material"], z["color"])
        o.set_floor(0.0, (1.0, 1.0, 1.0))
        o.set_directional_light((1, 1, 1), 0.1, (0.0, 0.0, 0.0))
        o.prepare_data()
        return o

    ref = mk(1)
    ref.accumulate(1024)
    m = ref.fetch_hdr()[..., :3]
    geo = (ref.trace_primary()["flags"] & 255) > 0
    res = {}
    for temporal in (False, True):
        o = mk(5)
        o.set_restir_temporal(temporal)
        prev, errs, first = np.zeros_like(m), [], None
        for k in range(10):
            o.accumulate_restir(1)
            cur = o.fetch_hdr()[..., :3] * (k + 1)
            if k == 0:
                first = cur.copy()
            errs.append(np.abs((cur - prev) - m)[geo].mean() / m[geo].mean())
            prev = cur
        res[temporal] = (np.mean(errs[2:]), (cur / 10)[geo].mean() / m[geo].mean(), first, o)
    assert res[True][0] < 0.97 * res[False][0], (res[True][0], res[False][0])
    # both estimators clamp W to 50 and radiance to 300 (upstream), which costs energy on this high-variance scene: measured
    # over 3 seeds x 40 frames the spatial pass alone reaches 0.966 of the path-traced mean, temporal + spatial 0.944
    assert abs(res[True][1] - 1.0) < 0.12 and abs(res[True][1] / res[False][1] - 1.0) < 0.06
    raw = res[True][3].get_reservoirs()
    M = raw[..., 0:2].copy().view(np.float16)[..., 0].astype(np.float32)
    assert M.max() >= 20.0 and M[geo].mean() > 4.0
    o = res[True][3]
    o.reset_framebuffer()
    o.accumulate_restir(1)
    assert np.allclose(o.fetch_hdr()[..., :3], res[True][2], rtol=0, atol=0)  # the reset dropped the history
