"""Helpers shared by the parity tests."""
import numpy as np


def make_pair(vrt, oracle, *, image_res, grid_res, dx=None, sky_res=0, cloud_passes=2, seed=7, jitter=True, max_depth=4,
              voxel_edges=0.06, exposure=3.0):
    """A CUDA Renderer and an OracleRenderer with identical configuration."""
    from voxel_rt2_b200.materials import material_table
    import os

    dx = dx if dx is not None else 2.0 / grid_res
    kw = dict(dx=dx, image_res=image_res, voxel_edges=voxel_edges, exposure=exposure, grid_res=grid_res, max_depth=max_depth,
              sky_res=sky_res, cloud_passes=cloud_passes, seed=seed, jitter=jitter)
    g = vrt.Renderer(**kw)
    tex = np.load(os.path.join(os.path.dirname(vrt.__file__), "assets", "cloud_texture.npz"))["tex"]
    o = oracle.OracleRenderer(materials=material_table(), cloud_tex=tex, **kw)
    return g, o


def apply_both(objs, name, *args, **kw):
    for x in objs:
        getattr(x, name)(*args, **kw)


def rel_rmse(a, b):
    """sqrt(mean((a-b)^2)) / mean(b) on the rgb channels."""
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-12))
