"""Full transmittance LUT (256 x 128 x f16x3) computed by the REFERENCE'S OWN SOURCE
(/root/reference/renderer/atmos.py:462-498, generate_transmittance_lut) through the float32 Taichi
emulator in oracle/ti_emu. One emulator run over all 32 768 entries takes ~45 min on one core, so
the entries are split over worker processes (each restricts the kernel's struct-for to its slice,
exactly as make_ref_vectors.section_sky does for its subset).

    python tests/golden/make_ref_lut.py [n_workers]

Output: tests/golden/ref_lut_full.npz (committed; 196 KB of binary16). make_ref_vectors.section_sky
feeds this table — not the oracle's — to the reference's cloud / skybox pipeline, and
tests/test_reference_vectors.py holds the oracle's whole LUT to it.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def work(args):
    k, n = args
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ti_emu"))
    sys.path.insert(0, REF)
    os.chdir(REF)
    import taichi as ti  # noqa: F401  (the emulator)
    from renderer.atmos import Atmos

    atm = Atmos()
    cells = [(x, y) for x in range(256) for y in range(128)][k::n]
    atm.trans_LUT._struct_for = lambda: iter([(np.int32(x), np.int32(y)) for x, y in cells])
    atm.generate_transmittance_lut()
    idx = np.array(cells, np.int32)
    return idx, atm.trans_LUT.arr[idx[:, 0], idx[:, 1]].copy()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, (os.cpu_count() or 2) - 1)
    with mp.get_context("spawn").Pool(n) as pool:
        parts = pool.map(work, [(k, n) for k in range(n)])
    lut = None
    for idx, val in parts:
        if lut is None:
            lut = np.zeros((256, 128) + val.shape[1:], val.dtype)
        lut[idx[:, 0], idx[:, 1]] = val
    np.savez_compressed(os.path.join(HERE, "ref_lut_full.npz"), lut=lut)
    print("ref_lut_full: dtype %s shape %s mean %.5f" % (lut.dtype, lut.shape, float(lut.astype(np.float32).mean())))


if __name__ == "__main__":
    main()
