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
