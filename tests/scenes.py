"""Seeded NumPy scene builders shared by the tests, bench.py and smoke().

They produce the host-side voxel arrays of the Scene API: material int8 [R,R,R] and colour
uint8 [R,R,R,3], index = ijk + R/2 (voxel_world.py:14-18)."""
import numpy as np


def empty(R):
    return np.zeros((R, R, R), np.int8), np.zeros((R, R, R, 3), np.uint8)


def main_scene(R=128):
    """main.py:13 — one emissive voxel at index (0,0,0)."""
    mat, col = empty(R)
    mat[R // 2, R // 2, R // 2] = 2
    col[R // 2, R // 2, R // 2] = (229, 25, 25)  # u8(0.9*255), u8(0.1*255)
    return mat, col


def random_grid(R, occupancy=0.5, seed=1234, materials=(1,)):
    """BASELINE.json config 3: dense random grid, material 1, colours 32..223."""
    rng = np.random.default_rng(seed)
    occ = rng.random((R, R, R)) < occupancy
    col = rng.integers(32, 224, (R, R, R, 3), dtype=np.uint8)
    if len(materials) == 1:
        mat = np.where(occ, materials[0], 0).astype(np.int8)
    else:
        pick = rng.integers(0, len(materials), (R, R, R))
        mat = np.where(occ, np.asarray(materials, np.int8)[pick], 0).astype(np.int8)
    return mat, col


def city(R=128, seed=0, n=50):
    """NumPy port of the *generator logic* of example1.py:9-24 (emissive rim, random towers with
    light caps) driven by a seeded NumPy RNG instead of ti.random()."""
    rng = np.random.default_rng(seed)
    mat, col = empty(R)
    h = R // 2

    def put(i, j, k, m, c):
        mat[i + h, j + h, k + h] = m
        col[i + h, j + h, k + h] = np.clip(np.asarray(c) * 255, 0, 255).astype(np.uint8)

    for i in range(n):
        for j in range(n):
            if min(i, j) == 0 or max(i, j) == n - 1:
                put(i, 0, j, 2, (0.9, 0.1, 0.1))
            else:
                put(i, 0, j, 1, (0.9, 0.1, 0.1))
                if rng.random() < 0.04:
                    height = min(int(rng.random() * 20), R // 2 - 1)
                    for k in range(1, height):
                        put(i, k, j, 1, (0.0, 0.5, 0.9))
                    if height:
                        put(i, height, j, 2, (1, 1, 1))
    return mat, col


def material_zoo(R=64, seed=3):
    """Slabs of every CSV material + lights, to exercise all three BSDF lobes."""
    ids = [1, 10, 11, 20, 21, 22, 30, 31, 32, 40, 41, 50, 51, 52, 53, 54, 80, 81, 82]
    rng = np.random.default_rng(seed)
    mat, col = empty(R)
    for x in range(R):
        for z in range(R):
            m = ids[((x // 4) + (z // 4) * 5) % len(ids)]
            hgt = R // 4 + int(4 * np.sin(x * 0.4) + 4 * np.cos(z * 0.3))
            mat[x, :hgt, z] = m
            col[x, :hgt, z] = rng.integers(60, 250, 3)
    # a few emissive blocks and pillars
    for _ in range(12):
        x, z = rng.integers(4, R - 4, 2)
        top = R // 4 + 10 + int(rng.integers(0, 10))
        mat[x:x + 2, R // 4:top, z:z + 2] = 54
        col[x:x + 2, R // 4:top, z:z + 2] = rng.integers(60, 250, 3)
        mat[x:x + 2, top, z:z + 2] = 2
        col[x:x + 2, top, z:z + 2] = (255, 240, 200)
    return mat, col
