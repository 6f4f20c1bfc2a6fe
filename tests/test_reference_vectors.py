"""The CPU oracle against golden vectors produced by the REFERENCE'S OWN SOURCE
(/root/reference/renderer/*.py executed through oracle/ti_emu, the float32 Taichi emulator; see
tests/golden/make_ref_vectors.py). These are the pins that tie oracle/ to upstream code rather than
to our reading of it. /root/reference is not needed at test time."""
import os

import numpy as np
import pytest

import scenes

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _orc(oracle, R, **kw):
    from voxel_rt2_b200.materials import material_table

    return oracle.OracleRenderer(dx=2.0 / R, image_res=(16, 16), grid_res=R, sky_res=kw.pop("sky_res", 0), materials=material_table(), **kw)


@pytest.mark.parametrize("grid", ["rand16_15", "rand32_40", "struct32"])
def test_traversal_matches_reference_raytrace_bit_for_bit(oracle, grid):
    """raytracer.py:72-155 run on 300 rays per grid (outside, inside and on-boundary origins): the
    oracle's hit distance (raw f32 bits), cell, normal and iteration count are identical."""
    z = np.load(os.path.join(G, "ref_raytrace.npz"))
    occ = z[grid + "_grid"]
    R = occ.shape[0]
    o = _orc(oracle, R)
    col = np.full(occ.shape + (3,), 200, np.uint8)
    o.set_voxels(occ, col)
    o.prepare_data()
    # occupancy pyramid (raytracer.py:46-70), every LOD, via the reference's query_occupancy
    ref_bits = z[grid + "_lodbits"]
    got, lod = [], 0
    while (R >> lod) >= 2:
        r = R >> lod
        got.append(np.array([[[o.occupancy(x, y, zz, lod) for zz in range(r)] for y in range(r)] for x in range(r)], np.uint8).reshape(-1))
        lod += 1
    got = np.concatenate(got)
    assert got.shape == ref_bits.shape and np.array_equal(got, ref_bits)
    t, cell, nrm, it = o.raytrace(z[grid + "_o"], z[grid + "_d"])
    rt, rcell, rn, rit, flag = z[grid + "_t"], z[grid + "_cell"], z[grid + "_normal"], z[grid + "_iters"], z[grid + "_flag"]
    ok = flag == 0
    assert ok.sum() >= 200
    assert np.array_equal(t[ok].view(np.uint32), rt[ok].view(np.uint32))
    hit = ok & np.isfinite(rt)
    assert hit.sum() >= 50
    assert np.array_equal(cell[hit], rcell[hit])
    assert np.array_equal(nrm[hit] + 0.0, rn[hit] + 0.0)
    assert np.array_equal(it[hit], rit[hit])
    # rays that step out of the grid before `hit_distance > far` fires — the reference
    # (reading the out-of-range cell as empty) and the oracle's pin both report a miss
    assert np.isinf(rt[~ok]).all() and np.isinf(t[~ok]).all()


def _close(a, b, rtol, atol=0.0):
    """|a - b| <= atol + rtol |b| elementwise; a NaN matches a NaN (the reference's zenith / nadir
    azimuth is 0/0 in project_sky, and so is the oracle's)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (np.abs(a - b) <= atol + rtol * np.abs(b)) | (np.isnan(a) & np.isnan(b))


def test_math_helpers_match_reference(oracle):
    """renderer/math_utils.py helpers: orthonormal basis, cone and cosine-hemisphere samplers fed the
    same random numbers, Uchimura tonemap, luminance, material packing, octahedral coding, hash3."""
    z = np.load(os.path.join(G, "ref_math.npz"))
    lib = oracle.load()
    onb = oracle.math_probe(0, z["onb_n"], out_per=6)
    assert _close(onb[:, :3], z["onb_x"], 1e-6, 1e-7).all() and _close(onb[:, 3:], z["onb_y"], 1e-6, 1e-7).all()
    b = np.concatenate([z["cone_cosmax"][:, None], z["cone_u"]], axis=1)
    assert _close(oracle.math_probe(1, z["cone_n"], b), z["cone_dir"], 2e-6, 2e-7).all()
    assert _close(oracle.math_probe(2, z["cone_n"], b), z["hemi_dir"], 2e-6, 2e-7).all()
    assert _close(oracle.math_probe(3, z["uchi_in"]), z["uchi_out"], 2e-6, 1e-7).all()
    assert _close(oracle.math_probe(4, z["uchi_in"], out_per=1), z["lum_out"], 1e-6, 1e-7).all()
    enc = oracle.math_probe(5, z["encmat_albedo"], z["encmat_id"].astype(np.float32), out_per=1).view(np.uint32)
    assert np.array_equal(enc, z["encmat_out"])
    for (x, y, w), h in zip(z["hash3_in"], z["hash3_out"]):
        assert lib.orc_hash3(int(x), int(y), int(w)) == int(h)
    # octahedral 2 x f16 coding of unit vectors (math_utils.py:201-215)
    import ctypes as C

    v = np.ascontiguousarray(z["oct_v"], np.float32)
    enc = np.zeros((len(v), 2), np.float32)
    dec = np.zeros((len(v), 3), np.float32)
    fp = C.POINTER(C.c_float)
    lib.orc_oct_round_trip(len(v), v.ctypes.data_as(fp), enc.ctypes.data_as(fp), dec.ctypes.data_as(fp))
    assert np.array_equal(enc.astype(np.float16).view(np.uint16), z["oct_enc"].view(np.uint16))
    assert _close(dec, z["oct_dec"], 1e-6, 1e-7).all()


def test_host_colour_quantisation_matches_reference():
    """rgb32f_to_rgb8 / rgb8_to_rgb32f (math_utils.py:86-100) vs the colour conversion of the host's
    Scene.set_voxel / get_voxel."""
    from voxel_rt2_b200.scene import _f32, _u8

    z = np.load(os.path.join(G, "ref_math.npz"))
    for c, u, back in zip(z["rgb_in"], z["rgb_u8"], z["rgb_back"]):
        assert [_u8(float(x)) for x in c] == [int(x) for x in u]
        assert [np.float32(_f32(int(x) / 255.0)) for x in u] == [x for x in back]


def test_material_table_matches_reference_material_list():
    """MaterialList (materials.py:47-112 + default_material_set.csv) read back field by field."""
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_bsdf.npz"))
    assert np.array_equal(np.asarray(material_table(), np.float32).reshape(128, 14), z["material_table"])


def test_disney_bsdf_matches_reference(oracle):
    """bsdf.py evaluated by the reference source for 252 (material, albedo, v, n, l, u) probes over all
    21 material rows: disney_evaluate_split, pdf_disney, lobe probabilities, the lobe-wise variants and
    sample_disney (direction, brdf, pdf, lobe) with the same three random numbers."""
    z = np.load(os.path.join(G, "ref_bsdf.npz"))
    o = _orc(oracle, 16)
    out = o.bsdf_probe(z["mat_id"], z["albedo"], z["v"], z["n"], z["l"], z["u3"])
    rt, at = 2e-6, 1e-7  # measured 1.5e-7: libm vs numpy float32 pow / sqrt / log differ by an ulp
    assert _close(out[:, 0:3], z["eval_d"], rt, at).all()
    assert _close(out[:, 3:6], z["eval_s"], rt, at).all()
    assert _close(out[:, 6], z["pdf"], rt, at).all()
    assert np.array_equal(out[:, 11].astype(np.int32), z["sample_lobe"])
    assert _close(out[:, 7:10], z["sample_dir"], 1e-5, 2e-6).all()
    good = np.isfinite(z["sample_pdf"]) & (np.abs(z["sample_pdf"]) < 1e6)
    assert good.sum() > 200
    assert _close(out[good, 10], z["sample_pdf"][good], 1e-4, 1e-6).all()  # sharp GGX lobes amplify the ulp of the sampled half vector
    assert _close(out[good, 12:15], z["sample_brdf"][good], 1e-4, 1e-6).all()
    lw_out, lw = o.bsdf_lobewise_probe(z["mat_id"], z["albedo"], z["v"], z["n"], z["l"])
    assert _close(lw, z["lobe_w"], 1e-6, 1e-7).all()
    assert _close(lw_out[:, :, 0:3], z["lobe_d"], rt, at).all()
    assert _close(lw_out[:, :, 3:6], z["lobe_s"], rt, at).all()
    assert _close(lw_out[:, :, 6], z["lobe_pdf"], rt, at).all()


def test_hit_buffer_matches_reference_next_hit_bit_for_bit(oracle):
    """Renderer.get_cast_dir + next_hit (+ the sun shadow ray) executed by the reference source for
    every pixel of a 32x16 view of a 32^3 scene (voxels of 10 materials incl. emissive, floor):
    f32 bits of t, normal, material id, emissive flag and shadow state are identical."""
    from util import reference_hit_fields, renderer_from_reference_fixture
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_render.npz"))
    o = renderer_from_reference_fixture(oracle.OracleRenderer, z, materials=material_table())
    o.prepare_data()
    h = reference_hit_fields(o.trace_primary())
    assert np.array_equal(h["t"].view(np.uint32), z["hit_t"].view(np.uint32))
    hit = np.isfinite(z["hit_t"])
    assert hit.sum() > 400
    assert np.array_equal(h["normal"][hit], z["hit_normal"][hit] + 0.0)
    assert np.array_equal(h["mat"][hit], z["hit_mat"][hit])
    assert np.array_equal(h["light"][hit], z["hit_light"][hit])
    assert np.array_equal(h["shadow"], z["hit_shadow"])
    assert len(np.unique(z["hit_shadow"])) >= 3 and len(np.unique(z["hit_mat"][hit])) >= 8


def test_path_estimator_matches_reference_render_per_pixel(oracle):
    """Renderer.render() (pathtracer.py:355-632) run by the reference source, sample by sample, with
    ti.random() answering from the shared counter-based sampler: the oracle's per-pixel value of
    each of 4 samples agrees to 1e-4 relative on EVERY pixel (measured 3e-5 worst case) — NEE,
    MIS, primary-vertex bookkeeping, emissive voxels, clamping, all lobes."""
    from util import reference_radiance, renderer_from_reference_fixture
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_render.npz"))
    o = renderer_from_reference_fixture(oracle.OracleRenderer, z, materials=material_table())
    o.prepare_data()
    for s in range(z["render_diffuse"].shape[0]):
        o.reset_framebuffer()
        o.sample_offset = s
        o.accumulate(1)
        a, b = o.fetch_hdr()[..., :3], reference_radiance(z, s)
        assert b.mean() > 0.1
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
        assert err.max() < 1e-4, "sample %d: worst pixel differs by %.3e" % (s, err.max())


def test_static_frame_loop_matches_reference_accumulate_and_fetch_image(oracle):
    """Scene.finish's static-camera loop run by the reference source for 4 frames with the physical
    sky on: accumulate() (render + the three temporal-filter kernels) and fetch_image() (vignette,
    exposure, Uchimura, gamma). The oracle's accumulated HDR mean and tonemapped image agree per
    pixel to 2e-5 (measured 1.4e-6 / 5e-7; the reference blends a running mean with mix(), the
    oracle sums)."""
    from util import renderer_from_reference_fixture
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_frame.npz"))
    o = renderer_from_reference_fixture(oracle.OracleRenderer, z, materials=material_table(), exposure=float(z["exposure"]))
    o.set_use_physical_sky(True, False)
    o.set_sky_tables(z["sky_scatter"], z["sky_trans"])
    o.prepare_data()
    o.accumulate(int(z["n_frames"]))
    hdr, ldr = o.fetch_hdr(), o.fetch_image()
    assert (hdr[..., 3] == int(z["n_frames"])).all()
    err = np.abs(hdr[..., :3] - z["hdr"]) / np.maximum(np.abs(z["hdr"]), 1e-3)
    assert err.max() < 2e-5, "HDR worst pixel %.3e" % err.max()
    assert np.abs(ldr[..., :3] - z["ldr"][..., :3]).max() < 2e-5
    assert (z["ldr"][..., 3] == 1.0).all() and (ldr[..., 3] == 1.0).all()
    assert z["ldr"][..., :3].std() > 0.05


def test_sky_precompute_matches_reference_atmos(oracle):
    """renderer/atmos.py run by the reference source: the full transmittance LUT (32 768 entries), then cloud
    ambient -> accumulate_clouds x 2 -> compute_skybox on a 6 x 6 table with ti.random() answering
    from the per-texel counter sampler, and the run-time lookups on those tables."""
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_sky.npz"))
    S = int(z["S"])
    tex = np.load(os.path.join(os.path.dirname(G), "..", "voxel_rt2_b200", "assets", "cloud_texture.npz"))["tex"]
    o = oracle.OracleRenderer(dx=2 / 16, image_res=(16, 16), grid_res=16, sky_res=S, cloud_passes=int(z["passes"]), seed=int(z["seed"]),
                              materials=material_table(), cloud_tex=tex)
    o.set_directional_light(z["sun_dir"], float(z["cone"]), z["sun_col"])
    o.set_use_physical_sky(True, True)
    o.prepare_data()
    lut = o.get_trans_lut().astype(np.float32)
    idx, ref = z["lut_idx"], z["lut_val"].astype(np.float32)
    got = lut[idx[:, 0], idx[:, 1]]
    # f16 storage: allow one f16 ulp (2^-10 relative) where exp() differs in the last float32 bit
    assert (np.abs(got - ref) <= 1.0e-3 * np.abs(ref) + 1e-7).all()
    assert (got == ref).mean() > 0.98  # measured: all 627 identical
    # the WHOLE 256 x 128 table against the reference's own generate_transmittance_lut (ref_lut_full.npz,
    # tests/golden/make_ref_lut.py): measured 32 763 of 32 768 entries bit-identical, the other 5 one f16 ulp off
    full = np.load(os.path.join(G, "ref_lut_full.npz"))["lut"]
    assert np.array_equal(full.view(np.uint16), z["lut_full"].view(np.uint16))  # the table the reference pipeline run was given
    mine = o.get_trans_lut()
    differ = (mine.view(np.uint16) != full.view(np.uint16)).any(-1)
    assert differ.sum() <= 8, differ.sum()
    assert (np.abs(mine.astype(np.float32) - full.astype(np.float32)) <= 2.0 ** -10 * np.maximum(np.abs(full.astype(np.float32)), 2.0 ** -4)).all()
    rt = 1e-4  # measured: scattering 3.5e-5, transmittance 7e-7, cloud ambient 8e-6
    assert _close(o.get_cloud_ambient(), z["cloud_ambient"], rt, 1e-7).all()
    sc, tr = o.get_sky_tables()
    assert _close(sc, z["sky_scatter"], rt, 1e-6).all(), np.abs(sc - z["sky_scatter"]).max()
    assert _close(tr, z["sky_trans"], rt, 1e-6).all()
    assert sc.mean() > 1e-3 and 0.0 < tr.mean() < 1.0
    # parameterisation and lookups
    assert _close(oracle.project_sky(z["dirs"], S), z["project_uv"], 1e-5, 2e-6).all()
    assert _close(oracle.unproject_sky(z["unproject_in"], S), z["unproject_dir"], 1e-5, 2e-6).all()
    assert _close(o.sample_sky_trans(z["dirs"]), z["lookup_trans"], 1e-4, 1e-6).all()
    s2, t2 = o.sample_skybox(z["dirs"], z["lookup_jitter"])
    assert _close(s2, z["lookup_scatter_j"], 1e-4, 1e-6).all() and _close(t2, z["lookup_trans_j"], 1e-4, 1e-6).all()


def test_restir_shift_matches_reference(oracle):
    """Renderer.shift() (pathtracer.py:672-812) evaluated by the reference source on 160 reconnections
    (surface with continuation / last vertex / escape vertex / NEE-invisible, all lobe codes, some
    failing the N.L checks): the oracle's shifted integrand (diffuse, specular) and Jacobian agree."""
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_shift.npz"))
    o = oracle.OracleRenderer(dx=2.0 / 32, image_res=(32, 16), grid_res=32, sky_res=0, materials=material_table())
    o.set_voxels(*scenes.empty(32))
    o.set_directional_light(z["cfg_light_dir"], float(z["cfg_light_cone"]), z["cfg_light_color"])
    o.set_camera_pos(*[float(x) for x in z["cam_pos"]])
    got, ref = o.shift_probe(z["rows"]), z["out"]
    assert (ref[:, 6] != 0).sum() >= 100 and (ref[:, 6] == 0).sum() >= 10
    assert _close(got[:, 6], ref[:, 6], 2e-6, 1e-9).all()
    live = ref[:, 6] != 0  # a zero Jacobian multiplies every use of the integrand (the CUDA path skips it)
    assert _close(got[live, :6], ref[live, :6], 1e-4, 1e-7).all(), np.abs(got[live, :6] - ref[live, :6]).max()
    assert np.abs(ref[live, :6]).max() > 0.01


def test_restir_reservoir_bookkeeping_and_packing_match_reference(oracle):
    """Reservoir.init / input_sample / update_cached_jacobian_term / merge / finalize_without_M and the
    56-byte record's encode -> decode (reservoir.py:41-141), run by the reference source on 192
    reservoirs with non-zero sample vectors: selections, M, W and every decoded field (f16 M / W /
    Jacobian term, 8-bit and f16 octahedral directions, i8 lobes) are BIT-identical. Rows in which
    nothing was selected keep the all-zero init sample: there upstream decodes the octahedral
    encoding of (0,0,0) (a 0/0) into a spurious (0,0,-1) normal and NaN incident direction, the
    hazard the oracle's flag bits close; the oracle returns the zero vectors the algorithm tests for."""
    z = np.load(os.path.join(G, "ref_reservoir.npz"))
    got, ref = oracle.reservoir_probe(z["rows"]), z["out"]
    gb, rb = got.view(np.uint32), ref.view(np.uint32)
    assert (gb[:, :7] == rb[:, :7]).all()  # selections, M, weight, decoded M / W
    assert ref[:, 0].sum() > 100 and ref[:, 1].sum() > 50 and ref[:, 2].sum() > 30 and (ref[:, 4] == 0).sum() >= 5
    empty = ref[:, :3].sum(1) == 0
    assert 1 <= empty.sum() <= 8
    assert (gb[~empty] == rb[~empty]).all(), np.argwhere(gb[~empty] != rb[~empty])[:5]
    assert np.isnan(ref[empty, 16:19]).all() and (ref[empty, 15] == -1).all()  # the upstream hazard, as recorded
    assert (got[empty, 13:19] == 0).all() and (got[empty, 22:25] == 0).all()
    keep = [c for c in range(7, 28) if c not in (*range(13, 19), *range(22, 25))]
    assert (gb[empty][:, keep] == rb[empty][:, keep]).all()


def test_restir_spatial_gris_matches_reference(oracle):
    """Renderer.spatial_GRIS(0, 24.0, 32, 1) (pathtracer.py:815-989) run by the reference source on
    hand-built buffers: the reference's own G-buffer of a 48 x 24 view (no sky pixels), one packed
    reservoir per pixel whose sample vectors are all non-zero (so the zero-vector encodings upstream
    leaves undefined do not occur), random canonical integrands; ti.random() answered from the shared
    sampler (dimension 65 radius shift, 66 + i tap merges, 98 canonical merge). Pins the tap spiral
    (hash3 seed, golden angle), the similarity rejection, both shift() directions, the pairwise MIS
    weights, the merge order, the visibility ray and the finalisation, with the physical-sky lookup
    of shift() on: the colour buffers after the pass agree on every processed pixel."""
    from util import renderer_from_reference_fixture
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_gris.npz"))
    o = renderer_from_reference_fixture(oracle.OracleRenderer, z, materials=material_table())
    o.set_use_physical_sky(True, False)
    o.set_sky_tables(z["sky_scatter"], z["sky_trans"])
    px = z["pixels"]
    got = o.gris_probe(int(z["frame"]), z["samples"], z["gbuf"], z["col_d"], z["col_s"], px)
    ref = np.concatenate([z["out_d"], z["out_s"]], 1)
    assert len(px) >= 150 and np.isfinite(ref).all()
    # the pass did resample: many pixels end on a neighbour's shifted sample, not on W x their own integrand
    own = np.concatenate([z["col_d"][px], z["col_s"][px]], 1)
    ratio = ref / own
    resampled = (ratio.max(1) - ratio.min(1)) > 1e-3 * np.abs(ratio).max(1)
    assert resampled.sum() >= 20 and (~resampled).sum() >= 20, resampled.sum()
    assert _close(got, ref, 1e-5, 1e-7).all(), (np.abs(got - ref) / (np.abs(ref) + 1e-6)).max()  # measured: identical bits


def test_restir_render_branch_matches_reference(oracle):
    """Renderer.render() with USE_RESTIR_PT = True (pathtracer.py:355-632) run by the reference
    source, Reservoir.encode intercepted so the fixture holds every pixel's reservoir BEFORE packing
    (the zero-vector markers are exact zeros there): reconnection vertex, incident direction /
    radiance, NEE direction, lobes, cached Jacobian term, the NEE-vs-BSDF RIS at the primary vertex
    (M, W, which sample was kept), the G-buffer and the canonical integrands agree per pixel for two
    samples (measured 7.5e-6 worst). Sky pixels: upstream stores the octahedral encoding of a zero
    normal (NaN, recorded in the fixture); the oracle stores zeros and a flag."""
    from util import renderer_from_reference_fixture
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_restir_render.npz"))
    o = renderer_from_reference_fixture(oracle.OracleRenderer, z, materials=material_table())
    o.set_use_physical_sky(True, False)
    o.set_sky_tables(z["sky_scatter"], z["sky_trans"])
    o.prepare_data()
    for s in range(z["samples"].shape[0]):
        sm, gb, cd, cs = o.restir_render_probe(s)
        ref, g = z["samples"][s], z["gbuf"][s]
        # exact parts: material info bits, lobes, M, zero-vector markers, geometric normals / sun directions
        assert (sm[:, 18].view(np.uint32) == ref[:, 18].view(np.uint32)).all()
        assert (sm[:, 20:22] == ref[:, 20:22]).all()
        for k in (6, 9, 15):
            assert ((np.abs(sm[:, k:k + 3]).sum(1) == 0) == (np.abs(ref[:, k:k + 3]).sum(1) == 0)).all(), k
        assert (sm[:, 6:9] == ref[:, 6:9]).all() and (sm[:, 15:18] == ref[:, 15:18]).all()
        # the cached Jacobian term of an escape vertex divides by |N| = 0 on both sides
        assert (np.isfinite(sm) == np.isfinite(ref)).all()
        fin = np.isfinite(ref)
        cols = [c for c in range(23) if c != 18]
        assert _close(sm[:, cols][fin[:, cols]], ref[:, cols][fin[:, cols]], 5e-5, 1e-6).all()
        sky = gb[:, 6] == 1
        assert (sky == (np.abs(g[:, :3]).sum(1) == 0)).all() and 20 < sky.sum() < 200
        assert np.isnan(g[sky, 3:5]).all() and (gb[sky, 3:5] == 0).all()
        assert (gb[:, :3] == g[:, :3]).all() and (gb[~sky, 3:5] == g[~sky, 3:5]).all()
        assert (gb[:, 5].view(np.uint32) == g[:, 5].view(np.uint32)).all()
        assert _close(cd, z["col_d"][s], 5e-5, 1e-6).all() and _close(cs, z["col_s"][s], 5e-5, 1e-6).all()
        # coverage: escape and surface reconnection vertices, terminated paths, both outcomes of the RIS
        surf = np.abs(ref[:, 6:9]).sum(1) > 0
        assert surf.sum() > 80 and (surf & (np.abs(ref[:, 9:12]).sum(1) == 0)).sum() > 5 and (np.abs(ref[:, 15:18]).sum(1) > 0).sum() > 30
        nee_kept = (~sky) & (ref[:, 20] == 99)  # LOBE_ALL * 10 + LOBE_ALL (bsdf.py:20) marks the light sample
        assert nee_kept.sum() > 20 and ((~sky) & ~nee_kept).sum() > 100


def test_config1_example1_hit_buffer_matches_reference(oracle):
    """BASELINE config 1 at 64 x 64: the example1.py scene in the Renderer exactly as shipped (128^3
    grid built by the reference's _update_lods / _make_texture), primary ray + sun shadow ray per
    pixel through the reference's get_cast_dir / next_hit — the oracle's hit buffer is bit-identical."""
    from util import assert_hits_equal_reference, example1_renderer
    from voxel_rt2_b200.materials import material_table

    o, h = example1_renderer(oracle.OracleRenderer, materials=material_table())
    hit = assert_hits_equal_reference(o.trace_primary(), h)
    assert hit.sum() > 1000 and (h["hit_mat"][hit] == 2).any() and (np.abs(h["hit_normal"][hit][:, 1]) != 1).sum() > 300


def test_moving_camera_path_matches_reference_filters(oracle):
    """Scene.finish's moving-camera loop run by the reference source for 4 frames of a camera
    translation (render at render_scale 0.5 with albedo demodulation, temporal_filter_prepass,
    temporal_filter, temporal_filter_specular, copy_prev_matrices; pathtracer.py:993-1303), with the two
    upstream hazards resolved as DESIGN.md "Moving-camera pins" states (snapshot reads in the in-place
    blur, non-finite reflection depths = no reflection). The oracle's colour buffer agrees per pixel:
    measured worst 1.5e-6 / 3.4e-7 / 7.6e-5 / 4.7e-3 over the four frames (the last: a 5 % depth /
    0.642 normal rejection decision flipping on float rounding in a handful of pixels)."""
    from util import renderer_from_reference_fixture
    from voxel_rt2_b200.materials import material_table

    z = np.load(os.path.join(G, "ref_moving.npz"))
    W, H = int(z["W"]), int(z["H"])
    zz = dict(z)
    zz["cam_pos"], zz["view"], zz["proj"] = z["cam_pos"][0], z["view"][0], z["proj"][0]
    o = renderer_from_reference_fixture(oracle.OracleRenderer, zz, materials=material_table())
    o.prepare_data()
    for f in range(z["frames"].shape[0]):
        o.set_view_proj(z["cam_pos"][f], z["view"][f], z["proj"][f])
        o.accumulate_moving(float(z["scale"]), float(z["max_accum"]))
        a = o.fetch_hdr_moving()[::2, ::2, :3][: H // 2, : W // 2]
        b = z["frames"][f][: H // 2, : W // 2]
        err = np.abs(a - b).max(-1) / np.maximum(np.abs(b).max(-1), 1e-3)
        assert b.mean() > 0.1
        assert (err < 1e-4).mean() >= 0.98 and err.max() < 2e-2, "frame %d: %.4f within 1e-4, worst %.3e" % (f, (err < 1e-4).mean(), err.max())
        if f < 2:
            assert err.max() < 1e-5


def test_voxel_authoring_matches_reference_set_get_voxel(monkeypatch):
    """Scene.round_idx + Renderer.set_voxel / get_voxel run by the reference source on 96 (index,
    material, colour) triples (ties of ti.round, materials >= 128 wrapping in i8, colours outside
    [0,1]): the host Scene of the product produces the same cells, material bytes, colour bytes and
    read-back values."""
    monkeypatch.setenv("VRT_GRID", "32")
    import voxel_rt2_b200.scene as S

    z = np.load(os.path.join(G, "ref_voxel.npz"))
    R = int(z["R"])
    sc = S.Scene(renderer_factory=lambda **kw: None)
    assert sc.grid_res == R
    for i in range(len(z["idx"])):
        idx = [float(x) for x in z["idx"][i]]
        assert list(S.Scene.round_idx(idx)) == [int(x) for x in z["rounded"][i]]
        sc.set_voxel(idx, int(z["mat"][i]), tuple(float(x) for x in z["color"][i]))
        m, c = sc.get_voxel(idx)
        assert int(m) == int(z["get_mat"][i])
        assert [np.float32(x) for x in c] == [x for x in z["get_color"][i]]
    assert np.array_equal(sc.voxel_material, z["material_field"]) and np.array_equal(sc.voxel_color, z["color_field"])
    assert (z["material_field"] < 0).any() and (z["get_mat"] == 127).any()
