"""ctypes binding of libvoxelrt.so (include/voxelrt.h). Thin: no logic, no fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VRT_LIB selects an alternative build of the same library (kernel tuning variants); never a fallback
LIB_PATH = os.environ.get("VRT_LIB") or os.path.join(HERE, "libvoxelrt.so")

EXPORTS = [
    "vrt_create", "vrt_destroy", "vrt_last_error", "vrt_set_stream", "vrt_upload_voxels", "vrt_set_camera",
    "vrt_set_light", "vrt_set_floor", "vrt_set_background", "vrt_set_sky", "vrt_set_materials",
    "vrt_set_cloud_texture", "vrt_prepare", "vrt_set_sky_shard", "vrt_sky_tables_device_ptr", "vrt_sky_tables_complete", "vrt_sky_tables_pending", "vrt_get_sky_tables", "vrt_set_sky_tables", "vrt_set_sky_format", "vrt_get_trans_lut",
    "vrt_trace_primary", "vrt_accumulate", "vrt_accumulate_restir", "vrt_set_restir_temporal", "vrt_accumulate_moving", "vrt_get_reservoirs", "vrt_spatial_gris", "vrt_set_tile_shard", "vrt_set_row_shard", "vrt_set_row_range", "vrt_reset", "vrt_get_accum", "vrt_set_accum", "vrt_accum_device_ptr",
    "vrt_fetch_hdr", "vrt_fetch_ldr", "vrt_fetch_ldr_async", "vrt_fetch_wait", "vrt_accum_ipc_handle", "vrt_open_peer_accum", "vrt_close_peer_accum", "vrt_fetch_ldr_merged", "vrt_set_accum_slot", "vrt_out_ipc_handle", "vrt_merge_slice", "vrt_copy_ldr_async", "vrt_stream_wait_copy", "vrt_resolve_ldr_device", "vrt_get_stats", "vrt_synchronize",
]


class vrt_config(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("grid_res", C.c_int32), ("voxel_dx", C.c_float),
        ("voxel_edges", C.c_float), ("exposure", C.c_float), ("max_depth", C.c_int32), ("sky_res", C.c_int32),
        ("cloud_passes", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint32), ("jitter_mode", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class vrt_hit(C.Structure):
    _fields_ = [("t", C.c_float), ("cell", C.c_int32 * 3), ("normal", C.c_float * 3), ("flags", C.c_uint32)]


class vrt_stats(C.Structure):
    _fields_ = [
        ("paths", C.c_uint64), ("rays", C.c_uint64), ("steps", C.c_uint64), ("queries", C.c_uint64),
        ("hits", C.c_uint64), ("sky_escapes", C.c_uint64), ("nee_visible", C.c_uint64), ("vertices", C.c_uint64),
        ("last_render_ms", C.c_float), ("last_resolve_ms", C.c_float), ("sky_precompute_ms", C.c_float),
        ("kernel_launches", C.c_uint32), ("last_gris_ms", C.c_float), ("render_ms_sum", C.c_float), ("render_launches", C.c_uint32),
        ("launches_total", C.c_uint32), ("last_temporal_ms", C.c_float),
    ]


_lib = None


def load():
    """Load libvoxelrt.so; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libvoxelrt.so is missing (%s). Build it with `python -m voxel_rt2_b200.build`; "
            "there is no CPU fallback for the rendering path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    fp = C.POINTER(C.c_float)
    lib.vrt_create.argtypes = [C.POINTER(vrt_config), C.POINTER(P)]
    lib.vrt_destroy.argtypes = [P]
    lib.vrt_destroy.restype = None
    lib.vrt_last_error.argtypes = [P]
    lib.vrt_last_error.restype = C.c_char_p
    lib.vrt_set_stream.argtypes = [P, P]
    lib.vrt_upload_voxels.argtypes = [P, P, P]
    lib.vrt_set_camera.argtypes = [P, fp, fp, fp]
    lib.vrt_set_light.argtypes = [P, fp, C.c_float, fp]
    lib.vrt_set_floor.argtypes = [P, C.c_float, fp, C.c_int32]
    lib.vrt_set_background.argtypes = [P, fp]
    lib.vrt_set_sky.argtypes = [P, C.c_int32, C.c_int32]
    lib.vrt_set_materials.argtypes = [P, fp]
    lib.vrt_set_cloud_texture.argtypes = [P, P]
    lib.vrt_prepare.argtypes = [P]
    lib.vrt_set_sky_shard.argtypes = [P, C.c_int32, C.c_int32]
    lib.vrt_sky_tables_device_ptr.argtypes = [P, C.POINTER(P), C.POINTER(P), C.POINTER(C.c_uint64)]
    lib.vrt_sky_tables_complete.argtypes = [P]
    lib.vrt_sky_tables_pending.argtypes = [P]
    lib.vrt_get_sky_tables.argtypes = [P, fp, fp]
    lib.vrt_set_sky_tables.argtypes = [P, fp, fp]
    lib.vrt_set_sky_format.argtypes = [P, C.c_int32]
    lib.vrt_fetch_ldr_async.argtypes = [P, fp]
    lib.vrt_fetch_wait.argtypes = [P]
    lib.vrt_get_trans_lut.argtypes = [P, P]
    lib.vrt_trace_primary.argtypes = [P, P]
    lib.vrt_accumulate.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.vrt_set_restir_temporal.argtypes = [P, C.c_int32]
    lib.vrt_accumulate_restir.argtypes = [P, C.c_int32, C.c_int32, C.c_int32]
    lib.vrt_accumulate_moving.argtypes = [P, C.c_int32, C.c_float, C.c_float]
    lib.vrt_get_reservoirs.argtypes = [P, P]
    lib.vrt_spatial_gris.argtypes = [P, C.c_int32, P, P, P, P, P]
    lib.vrt_set_tile_shard.argtypes = [P, C.c_int32, C.c_int32]
    lib.vrt_set_row_shard.argtypes = [P, C.c_int32, C.c_int32]
    lib.vrt_set_row_range.argtypes = [P, C.c_int32, C.c_int32]
    lib.vrt_reset.argtypes = [P]
    lib.vrt_get_accum.argtypes = [P, fp]
    lib.vrt_set_accum.argtypes = [P, fp]
    lib.vrt_accum_device_ptr.argtypes = [P, C.POINTER(P), C.POINTER(C.c_uint64)]
    lib.vrt_accum_ipc_handle.argtypes = [P, P]
    lib.vrt_open_peer_accum.argtypes = [P, P, C.POINTER(P)]
    lib.vrt_close_peer_accum.argtypes = [P, P]
    lib.vrt_fetch_ldr_merged.argtypes = [P, C.POINTER(P), C.c_int32, fp]
    lib.vrt_set_accum_slot.argtypes = [P, C.c_int32]
    lib.vrt_out_ipc_handle.argtypes = [P, P]
    lib.vrt_merge_slice.argtypes = [P, C.POINTER(P), C.c_int32, C.c_int32, C.c_int32, P, C.c_int32]
    lib.vrt_copy_ldr_async.argtypes = [P, fp]
    lib.vrt_stream_wait_copy.argtypes = [P]
    lib.vrt_fetch_hdr.argtypes = [P, fp]
    lib.vrt_fetch_ldr.argtypes = [P, fp]
    lib.vrt_resolve_ldr_device.argtypes = [P, C.POINTER(P)]
    lib.vrt_get_stats.argtypes = [P, C.POINTER(vrt_stats)]
    lib.vrt_synchronize.argtypes = [P]
    for name in EXPORTS:
        if name not in ("vrt_destroy", "vrt_last_error"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
