/*
 * libvoxelrt — C-ABI of the B200-native rendering hot path for voxel-rt2.
 *
 * The reference has no FFI: its hot path sits behind the Python object `Scene.renderer`
 * (renderer/pathtracer.py `class Renderer`) plus Taichi fields poked directly from scene.py.
 * Each entry point below names the reference interface it replaces (file:line under the
 * reference tree). The Python host (`voxel_rt2_b200/renderer.py`) mirrors `Renderer`'s method
 * names on top of these calls, and `scene.py` keeps the reference's Scene API unchanged.
 *
 * Conventions: plain pointers and sizes only; every host pointer is caller-allocated and
 * borrowed for the duration of the call; every function returns 0 on success or a negative
 * vrt_status; the message for the last failure is vrt_last_error(ctx) (ctx may be NULL for
 * creation failures). No C++ exceptions cross the boundary. A context is driven by one host
 * thread; several contexts may coexist (one per GPU / per process).
 */
#ifndef VOXELRT_H
#define VOXELRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vrt_status {
  VRT_OK = 0,
  VRT_ERR_CUDA = -1,
  VRT_ERR_BAD_ARG = -2,
  VRT_ERR_NOT_PREPARED = -3,
  VRT_ERR_NO_DEVICE = -4,
  VRT_ERR_OOM = -5
} vrt_status;

typedef struct vrt_ctx vrt_ctx;

/* Construction parameters. Replaces the constants baked into the reference:
 * SCREEN_RES / VOXEL_DX (scene.py:11-12), voxel_grid_res = 128 (pathtracer.py:83),
 * MAX_RAY_DEPTH = 4 (pathtracer.py:17), skybox_res = 3840 (atmos.py:66-67),
 * Renderer(dx, image_res, up, voxel_edges, exposure) (pathtracer.py:28). */
typedef struct vrt_config {
  int32_t width;        /* multiple of 8 */
  int32_t height;       /* multiple of 4 */
  int32_t grid_res;     /* R: power of two, 8..512 */
  float voxel_dx;       /* world size of one voxel (1/64 at R=128) */
  float voxel_edges;    /* Scene(voxel_edges=0.06) */
  float exposure;       /* Scene(exposure=3) */
  int32_t max_depth;    /* 1..8, reference 4 */
  int32_t sky_res;      /* sky table resolution, reference 3840 */
  int32_t cloud_passes; /* reference 32 (scene.py:199) */
  int32_t device;       /* CUDA device ordinal */
  uint32_t seed;
  int32_t jitter_mode;  /* 0 = no TAA jitter, 1 = per-sample Halton(2,3) jitter (pathtracer.py:264-265) */
  int32_t reserved[4];
} vrt_config;

/* One record of the primary-hit dump (north_star: "voxel id, face normal, hit t"). 32 bytes. */
typedef struct vrt_hit {
  float t;          /* world-space distance along the primary ray; +inf on miss */
  int32_t cell[3];  /* grid cell [0,R)^3 for voxel hits, (-1,-1,-1) otherwise */
  float normal[3];  /* raytracer.py:152-155 normal (components in {-1,0,1}; floor: (0,+-1,0)) */
  uint32_t flags;   /* bits 0-7 kind (0 miss, 1 floor, 2 voxel); 8-15 shadow (0 lit, 1 occluded,
                       2 back-facing: no ray, 3 n/a); 16-23 material id; 24-31 is_light */
} vrt_hit;

/* Counters of the last vrt_accumulate with stats enabled, device times, launch counts.
 * vrt_get_stats waits for the asynchronous launches still in flight before it answers. */
typedef struct vrt_stats {
  uint64_t paths, rays, steps, queries, hits, sky_escapes, nee_visible, vertices;
  float last_render_ms;   /* device time of the path kernel(s) in the last vrt_accumulate */
  float last_resolve_ms;  /* device time of the last tonemap pass */
  float sky_precompute_ms;
  uint32_t kernel_launches; /* kernels launched by the last vrt_accumulate */
  float last_gris_ms;       /* device time of the spatial resampling kernel(s), ReSTIR mode */
  float render_ms_sum;      /* device time of all vrt_accumulate path-kernel launches since the previous vrt_get_stats */
  uint32_t render_launches; /* ... and how many launches that sum covers */
  uint32_t launches_total;  /* kernels launched by the library since the previous vrt_get_stats */
  float last_temporal_ms;   /* device time of the temporal reservoir reuse kernel(s), ReSTIR mode */
} vrt_stats;

/* Renderer.__init__ (pathtracer.py:28-136) */
int vrt_create(const vrt_config* cfg, vrt_ctx** out);
void vrt_destroy(vrt_ctx* ctx);
const char* vrt_last_error(const vrt_ctx* ctx);

/* Use a caller-owned CUDA stream (cudaStream_t as void*) for all work; NULL = library stream. */
int vrt_set_stream(vrt_ctx* ctx, void* cuda_stream);

/* Renderer.set_voxel/get_voxel storage (pathtracer.py:1325-1334, voxel_world.py:6-25):
 * material int8 [R][R][R] and colour uint8 [R][R][R][3], C order, index = ijk + R/2. */
int vrt_upload_voxels(vrt_ctx* ctx, const int8_t* material, const uint8_t* rgb);

/* scene.py:233-237 -> Renderer.set_proj_mat / set_view_mat / set_camera_pos
 * (pathtracer.py:246-281). Row-major 4x4 view and projection (GL convention). */
int vrt_set_camera(vrt_ctx* ctx, const float pos[3], const float view[16], const float proj[16]);

/* Renderer.set_directional_light (pathtracer.py:139-144): direction is normalised here,
 * cos_theta_max = cos(cone_angle/2), radiance = 3 * rgb. */
int vrt_set_light(vrt_ctx* ctx, const float direction[3], float cone_angle, const float rgb[3]);

/* Scene.set_floor (scene.py:148-151) */
int vrt_set_floor(vrt_ctx* ctx, float height, const float rgb[3], int32_t material);
/* Scene.set_background_color (scene.py:156-157) */
int vrt_set_background(vrt_ctx* ctx, const float rgb[3]);
/* Scene.set_use_physical_sky / set_use_clouds (scene.py:159-169) */
int vrt_set_sky(vrt_ctx* ctx, int32_t physical, int32_t clouds);
/* MaterialList (materials.py:47-112): 128 rows x 14 floats (base_col rgb + 11 scalars). */
int vrt_set_materials(vrt_ctx* ctx, const float* table128x14);
/* Atmos.load_textures (atmos.py:85-90): uint8 [256][256][3] indexed [x][y] (imread layout). */
int vrt_set_cloud_texture(vrt_ctx* ctx, const uint8_t* tex256x256x3);

/* Renderer.prepare_data (pathtracer.py:314-323) + the sky start-up frames of Scene.finish
 * (scene.py:243-253): bricks + occupancy mips + colour SoA; if the physical sky is on,
 * transmittance LUT, cloud passes and the two sky tables. */
int vrt_prepare(vrt_ctx* ctx);

/* Sharded sky precompute (one process per GPU): after vrt_set_sky_shard(rank, n) vrt_prepare computes
 * only rows [rank, rank + 1) * sky_res / n of the two tables (compute_skybox is already sliced by rows upstream,
 * atmos.py:159-189) and leaves them incomplete; the caller gathers the other ranks' slices into the device buffers
 * returned by vrt_sky_tables_device_ptr (float4 texels [x][y], one all-gather per table) and then calls
 * vrt_sky_tables_complete. Rendering calls fail with VRT_ERR_NOT_PREPARED in between. sky_res must be a multiple of n. */
int vrt_set_sky_shard(vrt_ctx* ctx, int32_t rank, int32_t n);
int vrt_sky_tables_device_ptr(vrt_ctx* ctx, void** scattering, void** transmittance, uint64_t* bytes_each);
int vrt_sky_tables_complete(vrt_ctx* ctx);
/* 1 while a sharded precompute awaits its gather (between vrt_prepare and vrt_sky_tables_complete), else 0. */
int vrt_sky_tables_pending(vrt_ctx* ctx);

/* Sky tables as float3 [sky_res][sky_res] (atmos.py:68-69). get: copy out after vrt_prepare;
 * set: install externally computed tables (skips the precompute in vrt_prepare). */
int vrt_get_sky_tables(vrt_ctx* ctx, float* scattering, float* transmittance);
int vrt_set_sky_tables(vrt_ctx* ctx, const float* scattering, const float* transmittance);
/* Compact sky tables (no counterpart upstream, whose tables are f32 only,
 * atmos.py:68-69): format 0 (default) keeps the two float tables; format 1 packs both into ONE table
 * of 16-byte texels {scattering rgb, transmittance rgb} as binary16 (236 MB instead of 2 x 236 MB at
 * 3840^2; relative texel error <= 2^-11). Applies to the static-camera path kernel (vrt_accumulate);
 * the ReSTIR and moving-camera modes and vrt_get_sky_tables keep reading the float tables. */
int vrt_set_sky_format(vrt_ctx* ctx, int32_t format);
/* Atmos.trans_LUT (atmos.py:63): binary16 bits [256][128][3] */
int vrt_get_trans_lut(vrt_ctx* ctx, uint16_t* lut);

/* Primary-hit dump (config 1): next_hit for the camera ray (pathtracer.py:218-244) and the
 * sun shadow ray along the cone axis (pathtracer.py:435-450). out has width*height records,
 * row-major with v (y) slowest: out[v*width + u]. */
int vrt_trace_primary(vrt_ctx* ctx, vrt_hit* out);

/* Renderer.accumulate (pathtracer.py:1310-1319) for sample indices first, first+stride, ...
 * (n_samples of them). stats != 0 also fills the counters (and then waits for the launch). With
 * stats == 0 the call is asynchronous: it enqueues the batch on the context's stream and returns;
 * the fetch / stats / synchronize calls wait as needed. */
int vrt_accumulate(vrt_ctx* ctx, int32_t first_sample, int32_t n_samples, int32_t stride, int32_t stats);

/* Renderer.accumulate with USE_RESTIR_PT = True (pathtracer.py:15,1310-1319): per frame, render
 * writes one packed reservoir + G-buffer record per pixel (pathtracer.py:535-607) and
 * spatial_GRIS(0, 24.0, 32, 1) (pathtracer.py:815-989) resamples 32 neighbours; the frame is then
 * accumulated like a path-traced one. n_frames frames with sample indices first, first+stride...
 * Asynchronous like vrt_accumulate: the frames are enqueued on the context's stream; vrt_get_stats reads the
 * per-phase device times back (and waits for them). */
int vrt_accumulate_restir(vrt_ctx* ctx, int32_t first_sample, int32_t n_frames, int32_t stride);
/* Temporal reservoir reuse for vrt_accumulate_restir (BASELINE.json configs[3]: "temporal+spatial resampling per
 * frame"). No upstream counterpart: the reference allocates two reservoir slots per pixel (pathtracer.py:108-109) and
 * writes the second (:989) but never reads it. With enable != 0 every frame resamples, between render and
 * spatial_GRIS, the pixel's canonical reservoir with its own reservoir of the previous frame kept in that second
 * slot (same shift / pairwise-MIS / merge primitives as spatial_GRIS with one tap; see k_temporal in
 * csrc/vrt_restir.cu). Static camera: vrt_reset, a changed camera, light or scene drop the history. Default 0
 * (the reference's behaviour: spatial pass only). */
int vrt_set_restir_temporal(vrt_ctx* ctx, int32_t enable);
/* Renderer.accumulate with camera_is_moving = 1 (scene.py:214-228, pathtracer.py:146-150): one
 * frame rendered at render_scale (reference 0.5) with albedo-demodulated diffuse, then
 * temporal_filter_prepass / temporal_filter / temporal_filter_specular (pathtracer.py:1020-1303)
 * reprojecting the history of the previous moving frame with the previous camera matrices, then
 * copy_prev_matrices (pathtracer.py:284-287). max_accum = set_max_samples (reference 50 while
 * moving). vrt_fetch_hdr / vrt_fetch_ldr afterwards return this path's colour buffer, nearest
 * up-sampled as _render_to_image does; vrt_reset (reset_framebuffer) drops the history. */
int vrt_accumulate_moving(vrt_ctx* ctx, int32_t sample, float render_scale, float max_accum);
/* Packed reservoirs of the last ReSTIR frame: 56 bytes per pixel, row-major (reservoir.py:8-19
 * field order; byte 55 carries the escape / last-vertex / NEE-visible flags). */
int vrt_get_reservoirs(vrt_ctx* ctx, void* out56_per_pixel);
/* Renderer.spatial_GRIS(0, 24.0, 32, 1) (pathtracer.py:815-989, the call of accumulate() :1313) on
 * caller-supplied per-pixel buffers instead of the ones the path kernel wrote: packed reservoirs
 * (56 B each, the layout of vrt_get_reservoirs; spatial_reservoirs :108), gpos = float4 (primary
 * position, sky flag; gbuff_position :116), gattr = 2 x u32 (octahedral f16x2 primary normal,
 * packed material info; gbuff_normals / gbuff_mat_id :112-113), col_d / col_s = float4 canonical
 * integrands (color_buffer / color_buffer_specular :39-40). frame = current_frame. The pass adds
 * its result to the accumulation buffer like a rendered frame. Host pointers, row-major; the parity
 * tests use it to hold the resampling kernel to reference-derived vectors. */
int vrt_spatial_gris(vrt_ctx* ctx, int32_t frame, const void* reservoirs56, const float* gpos4, const uint32_t* gattr2, const float* col_d4,
                     const float* col_s4);

/* Only pixels in 8x4 tiles with tile_id % n == rank are rendered (tile sharding). Default 0,1. */
int vrt_set_tile_shard(vrt_ctx* ctx, int32_t rank, int32_t n);

/* Contiguous strips instead of interleaved tiles: this context owns the tile rows
 * [rank, rank + 1) * (height / 4) / n. vrt_accumulate renders the own rows only; vrt_accumulate_restir renders the
 * reservoirs / G-buffer (and runs the per-pixel temporal pass) for the own rows plus a halo of 24 pixels on either side
 * — spatial_GRIS's max_radius (pathtracer.py:1313) — and resamples the own rows, so ONE reservoir chain is spread over
 * the GPUs and the merged frame equals the unsharded one. n = 1 switches the mode off. */
int vrt_set_row_shard(vrt_ctx* ctx, int32_t rank, int32_t n);
/* The same with an explicit strip: tile rows [first_row, first_row + n_rows) (a tile row = 4 pixel rows). Lets the host
 * balance the strips by cost — sky rows are nearly free, geometry rows are not (parallel.shard_rows). */
int vrt_set_row_range(vrt_ctx* ctx, int32_t first_row, int32_t n_rows);

/* Renderer.reset_framebuffer (pathtracer.py:664-668). Deferred: if a full-frame vrt_accumulate follows, its
 * kernel overwrites the buffer (no memset, no read-modify-write); any other use clears it first. */
int vrt_reset(vrt_ctx* ctx);

/* Two accumulation slots, each with its own image buffer (no upstream counterpart; the reference has one
 * GPU): every call that reads or writes "the accumulation buffer" / "the image buffer" uses the current slot. With one process per GPU, batch k+1 renders into
 * one slot while the partial sums of batch k in the other are still being merged by the peers. */
int vrt_set_accum_slot(vrt_ctx* ctx, int32_t slot);

/* Accumulation checkpoint (upstream keeps its history buffers on the device only, pathtracer.py:39-44):
 * copy the float4 sums (rgb sums, w = samples) of the current slot to / from host memory, so a long render can be
 * saved and resumed (Scene.finish: VRT_CHECKPOINT=file.npz). */
int vrt_get_accum(vrt_ctx* ctx, float* rgba_sums);
int vrt_set_accum(vrt_ctx* ctx, const float* rgba_sums);

/* Device pointer of the float4 [height][width] accumulation buffer (rgb sums, w = samples),
 * for device-side collectives (NCCL all-reduce through torch.distributed). */
int vrt_accum_device_ptr(vrt_ctx* ctx, void** ptr, uint64_t* bytes);

/* Multi-GPU merge without a separate collective (one process per GPU on one NVLink box): every
 * rank exports its accumulation buffer (cudaIpcMemHandle_t, 64 bytes), the displaying rank opens
 * the peers' buffers and ONE kernel reads all partial sums over NVLink peer mappings, adds them to
 * its own and applies _render_to_image (pathtracer.py:634-662) — reduce + tonemap fused. The caller
 * orders the ranks (a barrier after the peers' vrt_accumulate and another before they touch the
 * buffer again). ldr_rgba may be NULL (result stays on the device, see vrt_resolve_ldr_device). */
int vrt_accum_ipc_handle(vrt_ctx* ctx, void* handle64);
int vrt_open_peer_accum(vrt_ctx* ctx, const void* handle64, void** peer_ptr);
int vrt_close_peer_accum(vrt_ctx* ctx, void* peer_ptr);
int vrt_fetch_ldr_merged(vrt_ctx* ctx, const void* const* peer_ptrs, int32_t n_peers, float* ldr_rgba);
/* The same merge spread over the ranks (reduce-scatter + tonemap in ONE kernel): this rank sums its own and
 * the peers' partial sums (current slot) for pixels [first_pixel, first_pixel + n_pixels) and writes the
 * tonemapped pixels (_render_to_image, pathtracer.py:634-662) to ldr_dst — a device pointer, typically the
 * peer mapping of the displaying rank's image buffer (vrt_out_ipc_handle + vrt_open_peer_accum), or NULL
 * for this context's own image buffer. write_sums != 0 also stores the merged float4 sums back into this
 * rank's slice of its accumulation buffer. Asynchronous on the context's stream; the caller orders the
 * ranks (all partial sums complete before, nobody overwrites them until every rank has merged).
 * vrt_copy_ldr_async then moves the displaying rank's image buffer to pinned host memory as
 * vrt_fetch_ldr_async does, without running the tonemap pass again. */
int vrt_out_ipc_handle(vrt_ctx* ctx, void* handle64);
int vrt_merge_slice(vrt_ctx* ctx, const void* const* peer_ptrs, int32_t n_peers, int32_t first_pixel, int32_t n_pixels, void* ldr_dst,
                    int32_t write_sums);
int vrt_copy_ldr_async(vrt_ctx* ctx, float* rgba_pinned);
/* Device-side ordering for that protocol: work enqueued on the context's stream after this call runs after
 * the pipelined image copy in flight (the displaying rank calls it before it lets the peers write the next
 * frame into the image buffer being copied). Does not block the host. */
int vrt_stream_wait_copy(vrt_ctx* ctx);

/* Renderer.color_buffer after accumulate: mean linear radiance, float4 [height][width]. */
int vrt_fetch_hdr(vrt_ctx* ctx, float* rgba);
/* Renderer.fetch_image / _render_to_image (pathtracer.py:634-662,1321-1323) */
int vrt_fetch_ldr(vrt_ctx* ctx, float* rgba);
/* Pipelined fetch_image for a frame loop (no upstream counterpart: the reference blits its texture on
 * the device): the tonemap pass runs on the context's stream, the device-to-host copy on a separate
 * copy-engine stream, and the call returns without waiting for it, so the copy of frame k overlaps the
 * rendering of frame k+1. rgba_pinned must be page-locked host memory that stays valid until
 * vrt_fetch_wait (or the next vrt_fetch_* call, which waits first) returns. */
int vrt_fetch_ldr_async(vrt_ctx* ctx, float* rgba_pinned);
int vrt_fetch_wait(vrt_ctx* ctx);
/* Same pass, result left on the device (no copy): returns the device pointer. */
int vrt_resolve_ldr_device(vrt_ctx* ctx, void** ptr);

int vrt_get_stats(vrt_ctx* ctx, vrt_stats* out);
int vrt_synchronize(vrt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VOXELRT_H */
