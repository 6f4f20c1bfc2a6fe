"""The C-ABI library loads and exports every symbol include/voxelrt.h declares (no compute calls
without a GPU), the ctypes structs match the header's layout, and the product path fails loudly
— never falls back — when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "voxelrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vrt_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vrt):
    from voxel_rt2_b200 import _cabi, build

    build.build()
    lib = _cabi.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libvoxelrt.so does not export %s" % n
    assert sorted(_cabi.EXPORTS) == names


def test_struct_layouts(vrt):
    from voxel_rt2_b200 import _cabi
    from voxel_rt2_b200.renderer import HIT_DTYPE

    assert C.sizeof(_cabi.vrt_hit) == 32 == HIT_DTYPE.itemsize
    assert C.sizeof(_cabi.vrt_config) == 16 * 4
    assert C.sizeof(_cabi.vrt_stats) == 8 * 8 + 10 * 4  # 9 x 4-byte fields + tail padding to 8


def test_bad_arguments_are_rejected_without_touching_a_device(vrt):
    from voxel_rt2_b200 import _cabi

    lib = _cabi.load()
    h = C.c_void_p()
    cfg = _cabi.vrt_config(width=100, height=100, grid_res=128, voxel_dx=1 / 64, voxel_edges=0.06, exposure=3, max_depth=4, sky_res=0,
                           cloud_passes=1, device=0, seed=0, jitter_mode=1)
    assert lib.vrt_create(C.byref(cfg), C.byref(h)) == -2  # width not a multiple of 8
    assert b"multiple" in lib.vrt_last_error(None)
    cfg.width, cfg.height, cfg.grid_res = 64, 64, 100
    assert lib.vrt_create(C.byref(cfg), C.byref(h)) == -2  # grid not a power of two
    assert lib.vrt_create(None, C.byref(h)) == -2
    assert lib.vrt_accumulate(None, 0, 1, 1, 0) == -2


def test_no_cpu_fallback(vrt):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        vrt.Renderer(image_res=(64, 64), grid_res=32, sky_res=0)


def test_product_package_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "voxel_rt2_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dp, f)
    txt = open(os.path.join(ROOT, "scene.py")).read()
    assert "oracle" not in txt
