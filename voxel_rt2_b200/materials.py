"""Material table of the reference (renderer/materials.py:47-112 + default_material_set.csv).

128 rows x 14 float32: base colour rgb, subsurface, metallic, specular, specular_tint,
roughness, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss, ior_minus_one
(struct order of renderer/bsdf.py:26-37). Every id not listed uses the default rough
dielectric (materials.py:50-63); the listed ids carry the values of the reference's CSV."""
import numpy as np

_DEFAULT = (1.0, 1.0, 1.0, 0.0, 0.0, 0.04, 0.0, 0.9, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)

# id: (subsurface, metallic, specular, specular_tint, roughness, anisotropic, sheen, sheen_tint,
#      clearcoat, clearcoat_gloss)  — base colour is white and ior_minus_one is 0 in every row
_OVERRIDES = {
    10: (0, 0, 0.3, 0, 0.6, 0, 0, 0, 0, 0),          # rough concrete
    11: (0, 0, 0.3, 0, 0.2, 0, 0, 0, 0, 0),          # smooth concrete
    20: (0.9, 0, 0.5, 0.2, 0.04, 0, 0, 0, 0, 0),     # silicate
    21: (0.5, 0, 0.35, 0, 0.22, 0, 0, 0, 0.6, 0.8),  # smooth ceramic
    22: (0.5, 0, 0.35, 0, 0.8, 0, 0, 0, 0.6, 0.8),   # rough ceramic
    30: (0.3, 0, 0.2, 0, 0.6, 0, 0.4, 0.5, 0, 0),    # bark
    31: (0.3, 0, 0.5, 0, 0.5, 0, 0.4, 0, 0, 0),      # wood plank
    32: (0.3, 0, 0.5, 0, 0.5, 0, 0.4, 0, 0.6, 0.7),  # gloss coated wood plank
    40: (0.4, 0, 0.3, 0.9, 0.2, 0, 0, 0, 0, 0),      # smooth plastic
    41: (0.4, 0, 0.3, 0.9, 0.6, 0, 0, 0, 0, 0),      # rough plastic
    50: (0, 1, 0.8, 0, 0.4, 0, 0, 0, 0, 0),          # rough metal
    51: (0, 1, 0.8, 0, 0.11, 0, 0, 0, 0, 0),         # smooth metal
    52: (0, 1, 0.8, 0, 1.0, 0, 0, 0, 0, 0),          # "mirror"
    53: (0, 1, 0.8, 0, 0.4, 0.8, 0, 0, 0, 0),        # brushed metal
    54: (0, 0.7, 0.8, 0, 0.3, 0, 0, 0, 0.7, 0.9),    # car paint
    80: (0.9, 0, 0.04, 0, 0.8, 0, 0, 0, 0, 0),       # plant
    81: (0.9, 0, 0.3, 0, 0.4, 0, 0, 0, 0, 0),        # light skin
    82: (0.95, 0, 0.0, 0, 0.4, 0, 0.9, 0.4, 0, 0),   # cloth
}


def material_table():
    """float32 [128, 14] table in the C-ABI layout of vrt_set_materials."""
    t = np.tile(np.asarray(_DEFAULT, np.float32), (128, 1))
    for mid, vals in _OVERRIDES.items():
        t[mid, 3:13] = np.asarray(vals, np.float32)
    return np.ascontiguousarray(t)


def load_csv(path):
    """Same table from a CSV in the reference's format (id, 14 values per row, one header line)."""
    t = np.tile(np.asarray(_DEFAULT, np.float32), (128, 1))
    with open(path, newline="") as f:
        rows = [r for r in f.read().splitlines()[1:] if r.strip()]
    for r in rows:
        v = [float(x) for x in r.split(",")]
        t[int(v[0])] = np.asarray(v[1:15], np.float32)
    return np.ascontiguousarray(t)
