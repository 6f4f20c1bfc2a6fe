// C-ABI of libvoxelrt (see include/voxelrt.h for the reference interfaces each call replaces).
// Host side only: context, device memory, parameter marshalling, launches. No CPU fallback: if
// there is no CUDA device, vrt_create fails with VRT_ERR_NO_DEVICE.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vrt_internal.h"

struct vrt_ctx {
  vrt_config cfg;
  int device = 0, sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;

  // voxel world
  int8_t* d_mat = nullptr;
  uint8_t* d_rgb = nullptr;
  unsigned long long* d_bricks = nullptr;
  uint32_t* d_color = nullptr;
  uint32_t* d_upper = nullptr;
  uint32_t upper_off[8] = {0};
  int n_lods = 0, upper_words = 0;
  bool voxels_uploaded = false, prepared = false;

  // uniforms
  float cam_pos[3] = {0.4f, 0.5f, 2.0f};
  float inv_proj[16], inv_view[16];
  bool camera_set = false;
  float light_dir[3], light_cos_max, light_color[3] = {0, 0, 0};
  float floor_height = 0.0f, floor_color[3] = {1, 1, 1};
  int floor_material = 1;
  float background[3] = {0, 0, 0};
  int use_sky = 0, use_clouds = 0;

  // tables
  float4* d_mats = nullptr;
  float4* d_sky_scatter = nullptr;
  float4* d_sky_trans = nullptr;
  uint4* d_sky_packed = nullptr;  // format 1 (vrt_set_sky_format), built from the float tables
  int sky_format = 0;
  bool sky_packed_valid = false;
  __half* d_trans_lut = nullptr;
  uint8_t* d_cloud_tex = nullptr;
  float* d_cloud_ambient = nullptr;
  bool sky_valid = false, cloud_tex_set = false, lut_valid = false;
  int sky_shard_rank = 0, sky_shard_n = 1;  // vrt_set_sky_shard: vrt_prepare computes rows [rank, rank + 1) * sky_res / n only
  bool sky_partial = false;                  // ... and the tables are complete only after vrt_sky_tables_complete

  // frame buffers. Two accumulation slots (the second allocated on first use, vrt_set_accum_slot): while the
  // partial sums of batch k are being merged across GPUs, batch k+1 renders into the other slot.
  float4* d_accum = nullptr;  // == accum_slot[cur_slot]
  float4* accum_slot[2] = {nullptr, nullptr};
  bool zero_pending[2] = {false, false};  // reset_framebuffer deferred: the next full-frame batch overwrites instead of adding
  int cur_slot = 0;
  float4* d_out = nullptr;  // resolve target (hdr or ldr) == out_slot[cur_slot]
  float4* out_slot[2] = {nullptr, nullptr};
  vrt_hit* d_hits = nullptr;
  float2* d_jitter = nullptr;
  int jitter_cap = 0;
  unsigned int* d_work = nullptr;
  unsigned long long* d_stats = nullptr;
  RestirBuffers rb{};  // allocated on first ReSTIR frame
  bool restir_temporal = false, hist_valid = false;  // vrt_set_restir_temporal; history of the previous ReSTIR frame usable
  // moving-camera temporal path (allocated on the first moving frame)
  struct {
    float4 *col_d = nullptr, *col_s = nullptr, *hd[2] = {nullptr, nullptr}, *hs[2] = {nullptr, nullptr}, *out = nullptr, *full = nullptr;
    float *depth[2] = {nullptr, nullptr}, *refl = nullptr, *refl_blur = nullptr, *hsd[2] = {nullptr, nullptr};
    uint2* attr[2] = {nullptr, nullptr};
    float prev_view[16], prev_proj[16];
    int cur = 0;
    bool has_prev = false, active = false;
    float scale = 1.0f;
  } mv;
  float view[16], proj[16];

  int tile_rank = 0, tile_n = 1;
  // vrt_set_row_shard: this context owns the tile rows [strip_row0, strip_row1) of the frame (contiguous strips, the
  // partition that lets the ReSTIR passes run on a shard: their 24-pixel neighbourhood becomes a 6-tile-row halo)
  int strip_n = 1, strip_row0 = 0, strip_row1 = 0;
  vrt_stats stats;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // vrt_accumulate is asynchronous: the device time of each launch is bracketed by an event pair from this ring
  // and read back lazily (vrt_get_stats, or when the ring wraps onto a pair that has not been read yet)
  static constexpr int EV_RING = 32;
  cudaEvent_t ring_a[EV_RING] = {nullptr}, ring_b[EV_RING] = {nullptr};
  bool ring_pending[EV_RING] = {false};
  int ring_head = 0;
  std::vector<cudaEvent_t> frame_events;  // ReSTIR mode: 4 events per frame of a call
  int restir_pending_frames = 0;           // ... whose times have not been read back yet (vrt_accumulate_restir is asynchronous)
  // pipelined image fetch (vrt_fetch_ldr_async): copy engine stream + events, created on first use
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_resolved = nullptr, ev_copied = nullptr;
  bool copy_pending = false;
};

static thread_local std::string g_create_err;

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      char buf_[512];                                                                  \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = buf_;                                                                 \
      return e_ == cudaErrorMemoryAllocation ? VRT_ERR_OOM : VRT_ERR_CUDA;              \
    }                                                                                  \
  } while (0)

#define REQUIRE(cond, msg)      \
  do {                          \
    if (!(cond)) {              \
      ctx->err = msg;           \
      return VRT_ERR_BAD_ARG;   \
    }                           \
  } while (0)

static bool invert4(const double m[16], double inv[16]) {
  double a[4][8];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) a[i][j] = m[i * 4 + j], a[i][4 + j] = i == j ? 1.0 : 0.0;
  for (int c = 0; c < 4; c++) {
    int p = c;
    for (int r = c + 1; r < 4; r++)
      if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
    if (std::fabs(a[p][c]) < 1e-300) return false;
    if (p != c)
      for (int j = 0; j < 8; j++) std::swap(a[p][j], a[c][j]);
    double d = a[c][c];
    for (int j = 0; j < 8; j++) a[c][j] /= d;
    for (int r = 0; r < 4; r++)
      if (r != c) {
        double f = a[r][c];
        for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) inv[i * 4 + j] = a[i][4 + j];
  return true;
}

static double halton(uint32_t i, uint32_t b) {
  double f = 1.0, r = 0.0;
  while (i > 0) {
    f /= (double)b;
    r += f * (double)(i % b);
    i /= b;
  }
  return r;
}

// One 20-float device row per material from the 14 reference parameters (bsdf.py:26-37 order:
// base rgb, subsurface, metallic, specular, specular_tint, roughness, anisotropic, sheen,
// sheen_tint, clearcoat, clearcoat_gloss, ior_minus_one) + the per-material constants the
// reference recomputes at every vertex (bsdf.py:92-95, :113-118, :351-363), in float32.
static void pack_material(const float* p, float* r) {
  const float pi = 3.14159265358979323846f;
  const float subsurface = p[3], metallic = p[4], specular = p[5], specular_tint = p[6], roughness = p[7], anisotropic = p[8];
  const float sheen = p[9], sheen_tint = p[10], clearcoat = p[11], clearcoat_gloss = p[12];
  float dw = (1.0f - metallic) * fmaxf(0.4f, fminf(0.9f, 1.0f - specular));
  float sw = 1.0f - dw;
  float cw = clearcoat * 0.7f;
  const float w_sum = dw + sw + cw;
  dw /= w_sum, sw /= w_sum, cw /= w_sum;
  const float aspect = sqrtf(1.0f - 0.9f * anisotropic);
  const float ax = fmaxf(roughness * roughness / aspect, 1e-3f), ay = fmaxf(roughness * roughness * aspect, 1e-3f);
  const float cc_alpha = 0.1f * (1.0f - clearcoat_gloss) + 0.001f * clearcoat_gloss;
  const float a2 = cc_alpha * cc_alpha;
  const float cc_norm = cc_alpha >= 1.0f ? 1.0f / pi : (a2 - 1.0f) / (pi * logf(a2));
  const float row[20] = {p[0], p[1], p[2], subsurface, metallic, specular, specular_tint, roughness, sheen, sheen_tint,
                         clearcoat, cc_alpha, dw, sw, cw, cc_norm, ax, ay, 1.0f / (pi * ax * ay), 0.0f};
  memcpy(r, row, sizeof row);
}

static void default_materials(float* rows) {  // materials.py:50-63
  const float d[14] = {1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.04f, 0.0f, 0.9f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  for (int i = 0; i < 128; i++) pack_material(d, rows + i * 20);
}

static bool is_sharded(const vrt_ctx* ctx) { return ctx->tile_n != 1 || ctx->strip_n > 1; }

static int n_local_tiles(const vrt_ctx* ctx) {
  if (ctx->strip_n > 1) return (ctx->strip_row1 - ctx->strip_row0) * (ctx->cfg.width / 8);
  int total = (ctx->cfg.width / 8) * (ctx->cfg.height / 4);
  if (ctx->tile_rank >= total) return 0;
  return (total - ctx->tile_rank + ctx->tile_n - 1) / ctx->tile_n;
}

static void fill_params(const vrt_ctx* ctx, Params& P) {
  memset(&P, 0, sizeof P);
  const vrt_config& c = ctx->cfg;
  P.bricks = ctx->d_bricks, P.upper = ctx->d_upper, P.color = ctx->d_color;
  P.R = c.grid_res, P.n_lods = ctx->n_lods, P.brick_res = c.grid_res / 4, P.upper_words = ctx->upper_words;
  for (int i = 0; i < 8; i++) P.upper_off[i] = ctx->upper_off[i];
  P.voxel_size = c.voxel_dx;
  P.voxel_inv_size = (float)(1.0 / (double)c.voxel_dx);  // voxel_world.py:11
  P.voxel_edges = c.voxel_edges;
  P.grid_half = (float)(c.grid_res / 2);
  P.floor_height = ctx->floor_height;
  P.floor_color = f3{ctx->floor_color[0], ctx->floor_color[1], ctx->floor_color[2]};
  P.floor_material = ctx->floor_material;
  P.light_dir = f3{ctx->light_dir[0], ctx->light_dir[1], ctx->light_dir[2]};
  P.light_cos_max = ctx->light_cos_max;
  P.light_color = f3{ctx->light_color[0], ctx->light_color[1], ctx->light_color[2]};
  P.light_weight = 3.0f;  // pathtracer.py:144
  P.background = f3{ctx->background[0], ctx->background[1], ctx->background[2]};
  P.use_sky = ctx->use_sky && ctx->sky_valid;
  P.cam_pos = f3{ctx->cam_pos[0], ctx->cam_pos[1], ctx->cam_pos[2]};
  memcpy(P.inv_proj, ctx->inv_proj, sizeof P.inv_proj);
  memcpy(P.inv_view, ctx->inv_view, sizeof P.inv_view);
  memcpy(P.view, ctx->view, sizeof P.view);
  memcpy(P.proj, ctx->proj, sizeof P.proj);
  P.W = c.width, P.H = c.height;
  P.inv_w = 1.0f / (float)c.width, P.inv_h = 1.0f / (float)c.height;
  P.sky_scatter = ctx->d_sky_scatter, P.sky_trans = ctx->d_sky_trans, P.sky_res = c.sky_res;
  P.sky_packed = (ctx->sky_format == 1 && ctx->sky_packed_valid) ? ctx->d_sky_packed : nullptr;
  P.mats = ctx->d_mats;
  P.accum = ctx->d_accum;
  P.seed = c.seed;
  P.max_depth = c.max_depth;
  P.tile_rank = ctx->tile_rank, P.tile_n = ctx->tile_n;
  P.tiles_x = c.width / 8;
  P.n_tiles = n_local_tiles(ctx);
  if (ctx->strip_n > 1) P.tile_rank = ctx->strip_row0 * P.tiles_x, P.tile_n = 1;  // a contiguous tile range
  P.work_counter = ctx->d_work;
  P.stats = nullptr;
  P.jitter = ctx->d_jitter;
}

// The deferred reset_framebuffer of the current slot becomes a real memset: needed before anything
// other than a full-frame path-tracing batch touches the buffer.
static int flush_pending_zero(vrt_ctx* ctx) {
  if (ctx->zero_pending[ctx->cur_slot]) {
    CK(cudaMemsetAsync(ctx->d_accum, 0, (size_t)ctx->cfg.width * ctx->cfg.height * sizeof(float4), ctx->stream));
    ctx->zero_pending[ctx->cur_slot] = false;
  }
  return VRT_OK;
}

// Read back the event pairs of finished asynchronous launches (blocking on the ones still in flight).
static int drain_ring_slot(vrt_ctx* ctx, int i) {
  if (!ctx->ring_pending[i]) return VRT_OK;
  CK(cudaEventSynchronize(ctx->ring_b[i]));
  float ms = 0.0f;
  CK(cudaEventElapsedTime(&ms, ctx->ring_a[i], ctx->ring_b[i]));
  ctx->ring_pending[i] = false;
  ctx->stats.last_render_ms = ms;
  ctx->stats.render_ms_sum += ms;
  ctx->stats.render_launches += 1;
  return VRT_OK;
}
static int drain_ring(vrt_ctx* ctx) {
  for (int k = 0; k < vrt_ctx::EV_RING; k++)  // oldest first, so last_render_ms ends up being the newest launch
    if (int rc = drain_ring_slot(ctx, (ctx->ring_head + k) % vrt_ctx::EV_RING)) return rc;
  return VRT_OK;
}

extern "C" {

const char* vrt_last_error(const vrt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

void vrt_destroy(vrt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->d_mat), cudaFree(ctx->d_rgb), cudaFree(ctx->d_bricks), cudaFree(ctx->d_color), cudaFree(ctx->d_upper);
  cudaFree(ctx->d_mats), cudaFree(ctx->d_sky_scatter), cudaFree(ctx->d_sky_trans), cudaFree(ctx->d_sky_packed), cudaFree(ctx->d_trans_lut);
  cudaFree(ctx->d_cloud_tex), cudaFree(ctx->d_cloud_ambient), cudaFree(ctx->accum_slot[0]), cudaFree(ctx->accum_slot[1]), cudaFree(ctx->out_slot[0]), cudaFree(ctx->out_slot[1]), cudaFree(ctx->d_hits);
  cudaFree(ctx->d_jitter), cudaFree(ctx->d_work), cudaFree(ctx->d_stats);
  cudaFree(ctx->mv.col_d), cudaFree(ctx->mv.col_s), cudaFree(ctx->mv.out), cudaFree(ctx->mv.full), cudaFree(ctx->mv.refl), cudaFree(ctx->mv.refl_blur);
  for (int k = 0; k < 2; k++) cudaFree(ctx->mv.hd[k]), cudaFree(ctx->mv.hs[k]), cudaFree(ctx->mv.hsd[k]), cudaFree(ctx->mv.depth[k]), cudaFree(ctx->mv.attr[k]);
  cudaFree(ctx->rb.reservoirs), cudaFree(ctx->rb.gpos), cudaFree(ctx->rb.gattr), cudaFree(ctx->rb.col_d), cudaFree(ctx->rb.col_s), cudaFree(ctx->rb.rc_skyT);
  cudaFree(ctx->rb.hist_res), cudaFree(ctx->rb.hist_gpos), cudaFree(ctx->rb.hist_gattr), cudaFree(ctx->rb.hist_skyT);
  if (ctx->copy_pending) cudaEventSynchronize(ctx->ev_copied);
  if (ctx->ev_resolved) cudaEventDestroy(ctx->ev_resolved);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int i = 0; i < vrt_ctx::EV_RING; i++) {
    if (ctx->ring_a[i]) cudaEventDestroy(ctx->ring_a[i]);
    if (ctx->ring_b[i]) cudaEventDestroy(ctx->ring_b[i]);
  }
  for (cudaEvent_t e : ctx->frame_events) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int vrt_create(const vrt_config* cfg, vrt_ctx** out) {
  if (!cfg || !out) {
    g_create_err = "vrt_create: null argument";
    return VRT_ERR_BAD_ARG;
  }
  *out = nullptr;
  const int R = cfg->grid_res;
  if (cfg->width <= 0 || cfg->height <= 0 || cfg->width % 8 || cfg->height % 4) {
    g_create_err = "vrt_create: width must be a positive multiple of 8 and height of 4";
    return VRT_ERR_BAD_ARG;
  }
  if (R < 8 || R > 512 || (R & (R - 1))) {
    g_create_err = "vrt_create: grid_res must be a power of two in [8, 512]";
    return VRT_ERR_BAD_ARG;
  }
  if (cfg->max_depth < 1 || cfg->max_depth > 8 || !(cfg->voxel_dx > 0.0f) || cfg->sky_res < 0 || cfg->sky_res > 8192 ||
      cfg->cloud_passes < 0) {
    g_create_err = "vrt_create: bad max_depth / voxel_dx / sky_res / cloud_passes";
    return VRT_ERR_BAD_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("vrt_create: no CUDA device (") + cudaGetErrorString(e) + "); libvoxelrt has no CPU path";
    return VRT_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_create_err = "vrt_create: device ordinal out of range";
    return VRT_ERR_BAD_ARG;
  }
  vrt_ctx* ctx = new vrt_ctx();
  ctx->cfg = *cfg;
  ctx->device = cfg->device;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  auto fail = [&](int code) {
    g_create_err = ctx->err;
    vrt_destroy(ctx);
    return code;
  };
#define CKC(call)                                                                         \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ctx->err = std::string(#call " failed: ") + cudaGetErrorString(e_);                 \
      return fail(e_ == cudaErrorMemoryAllocation ? VRT_ERR_OOM : VRT_ERR_CUDA);           \
    }                                                                                     \
  } while (0)
  CKC(cudaSetDevice(ctx->device));
  CKC(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
  CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  CKC(cudaEventCreate(&ctx->ev0));
  CKC(cudaEventCreate(&ctx->ev1));
  for (int i = 0; i < vrt_ctx::EV_RING; i++) {
    CKC(cudaEventCreate(&ctx->ring_a[i]));
    CKC(cudaEventCreate(&ctx->ring_b[i]));
  }

  const size_t nvox = (size_t)R * R * R;
  ctx->n_lods = 0;
  while ((1 << ctx->n_lods) < R) ctx->n_lods++;
  ctx->upper_words = 0;
  for (int l = 3; l < ctx->n_lods; l++) {
    size_t r = (size_t)(R >> l);
    ctx->upper_off[l - 3] = (uint32_t)ctx->upper_words;
    ctx->upper_words += (int)((r * r * r + 31) / 32);
  }
  CKC(cudaMalloc(&ctx->d_mat, nvox));
  CKC(cudaMalloc(&ctx->d_rgb, nvox * 3));
  CKC(cudaMalloc(&ctx->d_bricks, nvox / 64 * sizeof(unsigned long long)));
  CKC(cudaMalloc(&ctx->d_color, nvox * 4));
  CKC(cudaMalloc(&ctx->d_upper, (size_t)(ctx->upper_words > 0 ? ctx->upper_words : 1) * 4));
  CKC(cudaMalloc(&ctx->d_mats, 128 * 20 * sizeof(float)));
  const size_t npx = (size_t)cfg->width * cfg->height;
  CKC(cudaMalloc(&ctx->accum_slot[0], npx * sizeof(float4)));
  ctx->d_accum = ctx->accum_slot[0];
  CKC(cudaMalloc(&ctx->out_slot[0], npx * sizeof(float4)));
  ctx->d_out = ctx->out_slot[0];
  CKC(cudaMemsetAsync(ctx->d_accum, 0, npx * sizeof(float4), ctx->stream));
  CKC(cudaMalloc(&ctx->d_work, sizeof(unsigned int)));
  ctx->jitter_cap = 64;
  CKC(cudaMalloc(&ctx->d_jitter, (size_t)ctx->jitter_cap * sizeof(float2)));
  CKC(cudaMalloc(&ctx->d_stats, 8 * sizeof(unsigned long long)));
  CKC(cudaMalloc(&ctx->d_cloud_ambient, 3 * sizeof(float)));
  CKC(cudaMalloc(&ctx->d_cloud_tex, 256 * 256 * 3));
  CKC(cudaMemsetAsync(ctx->d_cloud_tex, 0, 256 * 256 * 3, ctx->stream));
  if (cfg->sky_res > 0) {
    const size_t ns = (size_t)cfg->sky_res * cfg->sky_res;
    CKC(cudaMalloc(&ctx->d_sky_scatter, ns * sizeof(float4)));
    CKC(cudaMalloc(&ctx->d_sky_trans, ns * sizeof(float4)));
    CKC(cudaMalloc(&ctx->d_trans_lut, 256 * 128 * 3 * sizeof(__half)));
  }
  {
    std::vector<float> t(128 * 20);
    default_materials(t.data());
    CKC(cudaMemcpyAsync(ctx->d_mats, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
  }
  // defaults: light ((1,1,1), 0.1, black) scene.py:127; identity-free camera must be set by the caller
  const float dir[3] = {1, 1, 1}, rgb[3] = {0, 0, 0};
  vrt_set_light(ctx, dir, 0.1f, rgb);
#undef CKC
  *out = ctx;
  return VRT_OK;
}

int vrt_set_stream(vrt_ctx* ctx, void* s) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) {
    cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
  }
  if (s) {
    ctx->stream = (cudaStream_t)s;
  } else {
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return VRT_OK;
}

int vrt_upload_voxels(vrt_ctx* ctx, const int8_t* material, const uint8_t* rgb) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(material && rgb, "vrt_upload_voxels: null pointer");
  CK(cudaSetDevice(ctx->device));
  const size_t nvox = (size_t)ctx->cfg.grid_res * ctx->cfg.grid_res * ctx->cfg.grid_res;
  CK(cudaMemcpyAsync(ctx->d_mat, material, nvox, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_rgb, rgb, nvox * 3, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->voxels_uploaded = true;
  ctx->prepared = false;
  ctx->hist_valid = false;
  return VRT_OK;
}

int vrt_set_camera(vrt_ctx* ctx, const float pos[3], const float view[16], const float proj[16]) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(pos && view && proj, "vrt_set_camera: null pointer");
  double v[16], p[16], vi[16], pi[16];
  for (int i = 0; i < 16; i++) v[i] = view[i], p[i] = proj[i];
  REQUIRE(invert4(v, vi) && invert4(p, pi), "vrt_set_camera: singular matrix");
  // temporal reservoir reuse is defined for a static camera: a different camera drops the history
  if (!ctx->camera_set || memcmp(ctx->view, view, sizeof ctx->view) || memcmp(ctx->proj, proj, sizeof ctx->proj) || memcmp(ctx->cam_pos, pos, sizeof ctx->cam_pos))
    ctx->hist_valid = false;
  for (int i = 0; i < 16; i++) ctx->inv_view[i] = (float)vi[i], ctx->inv_proj[i] = (float)pi[i];
  memcpy(ctx->cam_pos, pos, sizeof ctx->cam_pos);
  memcpy(ctx->view, view, sizeof ctx->view);
  memcpy(ctx->proj, proj, sizeof ctx->proj);
  ctx->camera_set = true;
  return VRT_OK;
}

int vrt_set_light(vrt_ctx* ctx, const float direction[3], float cone_angle, const float rgb[3]) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(direction && rgb, "vrt_set_light: null pointer");
  double x = direction[0], y = direction[1], z = direction[2];
  double n = std::sqrt(x * x + y * y + z * z);
  REQUIRE(n > 0.0, "vrt_set_light: zero direction");
  ctx->light_dir[0] = (float)(x / n), ctx->light_dir[1] = (float)(y / n), ctx->light_dir[2] = (float)(z / n);
  ctx->light_cos_max = (float)std::cos((double)cone_angle * 0.5);
  memcpy(ctx->light_color, rgb, sizeof ctx->light_color);
  ctx->sky_valid = false;  // the sky tables depend on the sun
  ctx->sky_packed_valid = false;
  ctx->hist_valid = false;
  return VRT_OK;
}

int vrt_set_floor(vrt_ctx* ctx, float height, const float rgb[3], int32_t material) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgb, "vrt_set_floor: null pointer");
  ctx->floor_height = height;
  memcpy(ctx->floor_color, rgb, sizeof ctx->floor_color);
  ctx->floor_material = material;
  return VRT_OK;
}

int vrt_set_background(vrt_ctx* ctx, const float rgb[3]) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgb, "vrt_set_background: null pointer");
  memcpy(ctx->background, rgb, sizeof ctx->background);
  return VRT_OK;
}

int vrt_set_sky(vrt_ctx* ctx, int32_t physical, int32_t clouds) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(!physical || ctx->cfg.sky_res > 0, "vrt_set_sky: context was created with sky_res = 0");
  if ((clouds != 0) != (ctx->use_clouds != 0)) ctx->sky_valid = false;
  ctx->use_sky = physical ? 1 : 0;
  ctx->use_clouds = clouds ? 1 : 0;
  return VRT_OK;
}

int vrt_set_materials(vrt_ctx* ctx, const float* t) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(t, "vrt_set_materials: null pointer");
  CK(cudaSetDevice(ctx->device));
  std::vector<float> rows(128 * 20, 0.0f);
  for (int i = 0; i < 128; i++) pack_material(t + i * 14, rows.data() + i * 20);
  CK(cudaMemcpyAsync(ctx->d_mats, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}

int vrt_set_cloud_texture(vrt_ctx* ctx, const uint8_t* tex) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(tex, "vrt_set_cloud_texture: null pointer");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->d_cloud_tex, tex, 256 * 256 * 3, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->cloud_tex_set = true;
  ctx->sky_valid = false;
  return VRT_OK;
}

static int sync_packed_sky(vrt_ctx* ctx);

int vrt_prepare(vrt_ctx* ctx) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  if (!ctx->voxels_uploaded) {
    ctx->err = "vrt_prepare: vrt_upload_voxels has not been called";
    return VRT_ERR_NOT_PREPARED;
  }
  CK(cudaSetDevice(ctx->device));
  CK(vrt_launch_build(ctx->d_mat, ctx->d_rgb, ctx->cfg.grid_res, ctx->d_bricks, ctx->d_color, ctx->d_upper, ctx->upper_off, ctx->n_lods,
                      ctx->upper_words, ctx->stream));
  if (ctx->use_sky && !ctx->sky_valid) {
    SkyBuild B;
    B.S = ctx->cfg.sky_res;
    B.scatter = ctx->d_sky_scatter, B.trans = ctx->d_sky_trans, B.trans_lut = ctx->d_trans_lut;
    B.cloud_tex = ctx->d_cloud_tex, B.cloud_ambient = ctx->d_cloud_ambient;
    B.sun_dir = f3{ctx->light_dir[0], ctx->light_dir[1], ctx->light_dir[2]};
    // sun_col = light_color * light_weight (pathtracer.py:320,326,329)
    B.sun_col = f3{ctx->light_color[0] * 3.0f, ctx->light_color[1] * 3.0f, ctx->light_color[2] * 3.0f};
    B.cosmax = ctx->light_cos_max;
    B.use_clouds = ctx->use_clouds;
    B.cloud_passes = ctx->cfg.cloud_passes > 0 ? ctx->cfg.cloud_passes : 1;
    B.seed = ctx->cfg.seed;
    const int rows = ctx->cfg.sky_res / ctx->sky_shard_n;
    B.first_texel = ctx->sky_shard_rank * rows * ctx->cfg.sky_res;
    B.n_texels = rows * ctx->cfg.sky_res;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    CK(vrt_launch_sky_precompute(B, ctx->stream));
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->stats.sky_precompute_ms, ctx->ev0, ctx->ev1));
    ctx->lut_valid = true;
    ctx->sky_packed_valid = false;
    if (ctx->sky_shard_n > 1) {
      // only this rank's rows are filled: the caller gathers the other ranks' slices into the buffers
      // (vrt_sky_tables_device_ptr) and then calls vrt_sky_tables_complete
      ctx->sky_partial = true;
      ctx->prepared = true;
      return VRT_OK;
    }
    ctx->sky_valid = true;
  }
  if (int rc = sync_packed_sky(ctx)) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->prepared = true;
  return VRT_OK;
}

int vrt_set_sky_shard(vrt_ctx* ctx, int32_t rank, int32_t n) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(n >= 1 && rank >= 0 && rank < n, "vrt_set_sky_shard: need 0 <= rank < n");
  REQUIRE(n == 1 || (ctx->cfg.sky_res > 0 && ctx->cfg.sky_res % n == 0), "vrt_set_sky_shard: sky_res must be a multiple of n");
  ctx->sky_shard_rank = rank, ctx->sky_shard_n = n;
  return VRT_OK;
}

int vrt_sky_tables_device_ptr(vrt_ctx* ctx, void** scattering, void** transmittance, uint64_t* bytes_each) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(scattering && transmittance, "vrt_sky_tables_device_ptr: null pointer");
  REQUIRE(ctx->cfg.sky_res > 0, "vrt_sky_tables_device_ptr: context was created with sky_res = 0");
  *scattering = ctx->d_sky_scatter, *transmittance = ctx->d_sky_trans;
  if (bytes_each) *bytes_each = (uint64_t)ctx->cfg.sky_res * ctx->cfg.sky_res * sizeof(float4);
  return VRT_OK;
}

int vrt_sky_tables_pending(vrt_ctx* ctx) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  return ctx->sky_partial ? 1 : 0;
}

int vrt_sky_tables_complete(vrt_ctx* ctx) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  if (!ctx->sky_partial) {
    ctx->err = "vrt_sky_tables_complete: no sharded sky precompute is pending";
    return VRT_ERR_NOT_PREPARED;
  }
  ctx->sky_partial = false;
  ctx->sky_valid = true;
  ctx->sky_packed_valid = false;
  if (int rc = sync_packed_sky(ctx)) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}

int vrt_get_sky_tables(vrt_ctx* ctx, float* scattering, float* transmittance) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(scattering && transmittance, "vrt_get_sky_tables: null pointer");
  if (!ctx->sky_valid) {
    ctx->err = "vrt_get_sky_tables: sky tables have not been computed";
    return VRT_ERR_NOT_PREPARED;
  }
  CK(cudaSetDevice(ctx->device));
  const size_t ns = (size_t)ctx->cfg.sky_res * ctx->cfg.sky_res;
  std::vector<float4> tmp(ns);
  for (int k = 0; k < 2; k++) {
    CK(cudaMemcpyAsync(tmp.data(), k == 0 ? ctx->d_sky_scatter : ctx->d_sky_trans, ns * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float* dst = k == 0 ? scattering : transmittance;
    for (size_t i = 0; i < ns; i++) dst[3 * i] = tmp[i].x, dst[3 * i + 1] = tmp[i].y, dst[3 * i + 2] = tmp[i].z;
  }
  return VRT_OK;
}

int vrt_set_sky_tables(vrt_ctx* ctx, const float* scattering, const float* transmittance) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(scattering && transmittance, "vrt_set_sky_tables: null pointer");
  REQUIRE(ctx->cfg.sky_res > 0, "vrt_set_sky_tables: context was created with sky_res = 0");
  CK(cudaSetDevice(ctx->device));
  const size_t ns = (size_t)ctx->cfg.sky_res * ctx->cfg.sky_res;
  std::vector<float4> tmp(ns);
  for (int k = 0; k < 2; k++) {
    const float* src = k == 0 ? scattering : transmittance;
    for (size_t i = 0; i < ns; i++) tmp[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.0f);
    CK(cudaMemcpyAsync(k == 0 ? ctx->d_sky_scatter : ctx->d_sky_trans, tmp.data(), ns * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->sky_valid = true;
  ctx->sky_partial = false;
  ctx->sky_packed_valid = false;
  return sync_packed_sky(ctx);
}

// (Re)build the packed table whenever the float tables changed and format 1 is selected.
static int sync_packed_sky(vrt_ctx* ctx) {
  if (ctx->sky_format != 1 || !ctx->sky_valid || ctx->sky_packed_valid) return VRT_OK;
  CK(cudaSetDevice(ctx->device));
  const size_t ns = (size_t)ctx->cfg.sky_res * ctx->cfg.sky_res;
  if (!ctx->d_sky_packed) CK(cudaMalloc(&ctx->d_sky_packed, ns * sizeof(uint4)));
  CK(vrt_launch_pack_sky(ctx->d_sky_scatter, ctx->d_sky_trans, ctx->d_sky_packed, ns, ctx->stream));
  ctx->sky_packed_valid = true;
  return VRT_OK;
}

int vrt_set_sky_format(vrt_ctx* ctx, int32_t format) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(format == 0 || format == 1, "vrt_set_sky_format: format must be 0 (float tables) or 1 (packed binary16)");
  REQUIRE(format == 0 || ctx->cfg.sky_res > 0, "vrt_set_sky_format: context was created with sky_res = 0");
  ctx->sky_format = format;
  return sync_packed_sky(ctx);
}

int vrt_get_trans_lut(vrt_ctx* ctx, uint16_t* lut) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(lut, "vrt_get_trans_lut: null pointer");
  if (!ctx->d_trans_lut || !ctx->lut_valid) {
    ctx->err = "vrt_get_trans_lut: sky has not been computed";
    return VRT_ERR_NOT_PREPARED;
  }
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(lut, ctx->d_trans_lut, 256 * 128 * 3 * sizeof(__half), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}

static int check_ready(vrt_ctx* ctx, const char* who) {
  if (!ctx->prepared || !ctx->camera_set) {
    ctx->err = std::string(who) + ": call vrt_prepare and vrt_set_camera first";
    return VRT_ERR_NOT_PREPARED;
  }
  if (ctx->use_sky && !ctx->sky_valid) {
    ctx->err = std::string(who) + ": physical sky enabled but tables are stale; call vrt_prepare";
    return VRT_ERR_NOT_PREPARED;
  }
  return VRT_OK;
}

int vrt_trace_primary(vrt_ctx* ctx, vrt_hit* out) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(out, "vrt_trace_primary: null pointer");
  int rc = check_ready(ctx, "vrt_trace_primary");
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t npx = (size_t)ctx->cfg.width * ctx->cfg.height;
  if (!ctx->d_hits) CK(cudaMalloc(&ctx->d_hits, npx * sizeof(vrt_hit)));
  Params P;
  fill_params(ctx, P);
  P.tile_rank = 0, P.tile_n = 1;
  P.n_tiles = (ctx->cfg.width / 8) * (ctx->cfg.height / 4);
  CK(vrt_launch_primary(P, ctx->d_hits, ctx->stream));
  CK(cudaMemcpyAsync(out, ctx->d_hits, npx * sizeof(vrt_hit), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}

int vrt_accumulate(vrt_ctx* ctx, int32_t first_sample, int32_t n_samples, int32_t stride, int32_t stats) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(n_samples > 0 && stride > 0 && first_sample >= 0, "vrt_accumulate: bad sample range");
  int rc = check_ready(ctx, "vrt_accumulate");
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  if (n_samples > ctx->jitter_cap) {
    CK(cudaStreamSynchronize(ctx->stream));  // a launch in flight may still read the old buffer
    if (ctx->d_jitter) cudaFree(ctx->d_jitter);
    ctx->d_jitter = nullptr;
    ctx->jitter_cap = 0;
    CK(cudaMalloc(&ctx->d_jitter, (size_t)n_samples * sizeof(float2)));
    ctx->jitter_cap = n_samples;
  }
  ctx->mv.active = false;
  Params P;
  fill_params(ctx, P);
  P.first_sample = first_sample, P.n_samples = n_samples, P.stride = stride;
  P.stats = stats ? ctx->d_stats : nullptr;
  if (P.n_tiles <= 0) return VRT_OK;
  // reset_framebuffer + full-frame batch: the kernel writes every texel, so the memset (and the read) is skipped
  if (ctx->zero_pending[ctx->cur_slot] && !is_sharded(ctx)) {
    P.accum_overwrite = 1;
    ctx->zero_pending[ctx->cur_slot] = false;
  } else if (int rcz = flush_pending_zero(ctx)) {
    return rcz;
  }
  // pathtracer.py:264-265: taa_jitter = (rand*2-1) * inv_image_res, one value per frame. Ours is a Halton(2,3)
  // point per sample index (stratified pixel footprint), computed on the device (k_jitter also clears the tile
  // queue counter): nothing on the host outlives the call, so it returns without synchronising.
  CK(vrt_launch_jitter(ctx->d_jitter, first_sample, stride, n_samples, ctx->cfg.width, ctx->cfg.height, ctx->cfg.jitter_mode, ctx->d_work, ctx->stream));
  if (stats) CK(cudaMemsetAsync(ctx->d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
  const int slot = ctx->ring_head;
  if (int rcd = drain_ring_slot(ctx, slot)) return rcd;
  CK(cudaEventRecord(ctx->ring_a[slot], ctx->stream));
  CK(vrt_launch_path(P, stats != 0, ctx->sm_count, ctx->stream, nullptr));
  CK(cudaEventRecord(ctx->ring_b[slot], ctx->stream));
  ctx->ring_pending[slot] = true;
  ctx->ring_head = (slot + 1) % vrt_ctx::EV_RING;
  ctx->stats.kernel_launches = 2;  // k_jitter, k_path
  ctx->stats.launches_total += 2;
  if (stats) {
    unsigned long long h[8];
    CK(cudaMemcpyAsync(h, ctx->d_stats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.paths = h[0], ctx->stats.rays = h[1], ctx->stats.steps = h[2], ctx->stats.queries = h[3], ctx->stats.hits = h[4];
    ctx->stats.sky_escapes = h[5], ctx->stats.nee_visible = h[6], ctx->stats.vertices = h[7];
  }
  return VRT_OK;
}

static int ensure_restir_buffers(vrt_ctx* ctx) {
  const size_t npx = (size_t)ctx->cfg.width * ctx->cfg.height;
  if (!ctx->rb.reservoirs) {
    CK(cudaMalloc(&ctx->rb.reservoirs, npx * 56));
    CK(cudaMalloc(&ctx->rb.gpos, npx * sizeof(float4)));
    CK(cudaMalloc(&ctx->rb.gattr, npx * sizeof(uint2)));
    CK(cudaMalloc(&ctx->rb.col_d, npx * sizeof(float4)));
    CK(cudaMalloc(&ctx->rb.col_s, npx * sizeof(float4)));
    CK(cudaMalloc(&ctx->rb.rc_skyT, npx * sizeof(float4)));
  }
  if (ctx->restir_temporal && !ctx->rb.hist_res) {
    CK(cudaMalloc(&ctx->rb.hist_res, npx * 56));
    CK(cudaMalloc(&ctx->rb.hist_gpos, npx * sizeof(float4)));
    CK(cudaMalloc(&ctx->rb.hist_gattr, npx * sizeof(uint2)));
    CK(cudaMalloc(&ctx->rb.hist_skyT, npx * sizeof(float4)));
    ctx->hist_valid = false;
  }
  ctx->rb.temporal = ctx->restir_temporal ? 1 : 0;
  return VRT_OK;
}

int vrt_accumulate_restir(vrt_ctx* ctx, int32_t first_sample, int32_t n_frames, int32_t stride) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(n_frames > 0 && stride > 0 && first_sample >= 0, "vrt_accumulate_restir: bad sample range");
  REQUIRE(ctx->tile_n == 1, "vrt_accumulate_restir: interleaved tile sharding cannot carry the 24-pixel neighbourhood; use vrt_set_row_shard or sample sharding");
  int rc = check_ready(ctx, "vrt_accumulate_restir");
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  rc = ensure_restir_buffers(ctx);
  if (rc) return rc;
  if (int rcz = flush_pending_zero(ctx)) return rcz;  // the resampling pass adds its frame to the buffer
  ctx->mv.active = false;
  while (ctx->frame_events.size() < 4u * (size_t)n_frames) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    ctx->frame_events.push_back(e);
  }
  for (int k = 0; k < n_frames; k++) {
    const uint32_t s = (uint32_t)(first_sample + k * stride);
    CK(vrt_launch_jitter(ctx->d_jitter, (int)s, 1, 1, ctx->cfg.width, ctx->cfg.height, ctx->cfg.jitter_mode, ctx->d_work, ctx->stream));
    Params P;
    fill_params(ctx, P);
    P.first_sample = (int)s, P.n_samples = 1, P.stride = 1;
    // Row shard: the reservoirs / G-buffer (and the per-pixel temporal chain) are produced for the own rows plus a halo of
    // 6 tile rows = 24 pixels on either side (spatial_GRIS max_radius, pathtracer.py:1313), the spatial pass runs on the
    // own rows only. Every halo pixel is computed exactly as its owner computes it, so the merged frame equals the
    // unsharded one.
    Params PX = P;
    if (ctx->strip_n > 1) {
      const int rows = ctx->cfg.height / 4;
      const int r0 = ctx->strip_row0 - 6 > 0 ? ctx->strip_row0 - 6 : 0, r1 = ctx->strip_row1 + 6 < rows ? ctx->strip_row1 + 6 : rows;
      PX.tile_rank = r0 * P.tiles_x, PX.n_tiles = (r1 - r0) * P.tiles_x;
    }
    cudaEvent_t* ev = &ctx->frame_events[4 * (size_t)k];
    CK(cudaEventRecord(ev[0], ctx->stream));
    CK(vrt_launch_path_restir(PX, ctx->rb, ctx->sm_count, ctx->stream));
    CK(cudaEventRecord(ev[1], ctx->stream));
    if (ctx->restir_temporal) {  // temporal reuse of the previous frame's reservoirs (k_rc_sky + k_temporal), in place
      CK(vrt_launch_temporal(PX, ctx->rb, s, ctx->hist_valid ? 1 : 0, ctx->stream));
      ctx->hist_valid = true;
    } else {
      CK(vrt_launch_rc_sky(PX, ctx->rb, ctx->stream));
    }
    CK(cudaEventRecord(ev[2], ctx->stream));
    CK(vrt_launch_gris(P, ctx->rb, s, ctx->stream));
    CK(cudaEventRecord(ev[3], ctx->stream));
  }
  // asynchronous like vrt_accumulate: the per-phase device times are read back by vrt_get_stats
  ctx->restir_pending_frames = n_frames;
  ctx->stats.kernel_launches = (ctx->restir_temporal ? 5u : 4u) * (uint32_t)n_frames;  // k_jitter, k_path, k_rc_sky, [k_temporal,] k_gris
  ctx->stats.launches_total += ctx->stats.kernel_launches;
  return VRT_OK;
}

static int drain_restir_events(vrt_ctx* ctx) {
  if (ctx->restir_pending_frames <= 0) return VRT_OK;
  const int n_frames = ctx->restir_pending_frames;
  ctx->restir_pending_frames = 0;
  CK(cudaEventSynchronize(ctx->frame_events[4 * (size_t)(n_frames - 1) + 3]));
  float render_ms = 0.0f, gris_ms = 0.0f, temporal_ms = 0.0f;
  for (int k = 0; k < n_frames; k++) {
    float a = 0.0f, b = 0.0f, c = 0.0f;
    const cudaEvent_t* ev = &ctx->frame_events[4 * (size_t)k];
    CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
    CK(cudaEventElapsedTime(&c, ev[1], ev[2]));
    CK(cudaEventElapsedTime(&b, ev[2], ev[3]));
    render_ms += a, gris_ms += b, temporal_ms += c;
  }
  ctx->stats.last_render_ms = render_ms;
  ctx->stats.last_gris_ms = gris_ms;
  ctx->stats.last_temporal_ms = ctx->restir_temporal ? temporal_ms : 0.0f;
  return VRT_OK;
}

int vrt_set_restir_temporal(vrt_ctx* ctx, int32_t enable) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  ctx->restir_temporal = enable != 0;
  ctx->hist_valid = false;
  return VRT_OK;
}

int vrt_spatial_gris(vrt_ctx* ctx, int32_t frame, const void* reservoirs, const float* gpos, const uint32_t* gattr, const float* col_d,
                     const float* col_s) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(frame >= 0 && reservoirs && gpos && gattr && col_d && col_s, "vrt_spatial_gris: bad arguments");
  REQUIRE(!is_sharded(ctx), "vrt_spatial_gris: sharding is not supported here (taps cross tiles)");
  int rc = check_ready(ctx, "vrt_spatial_gris");
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  rc = ensure_restir_buffers(ctx);
  if (rc) return rc;
  if (int rcz = flush_pending_zero(ctx)) return rcz;
  ctx->mv.active = false;
  const size_t npx = (size_t)ctx->cfg.width * ctx->cfg.height;
  CK(cudaMemcpyAsync(ctx->rb.reservoirs, reservoirs, npx * 56, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->rb.gpos, gpos, npx * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->rb.gattr, gattr, npx * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->rb.col_d, col_d, npx * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->rb.col_s, col_s, npx * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  Params P;
  fill_params(ctx, P);
  P.first_sample = frame, P.n_samples = 1, P.stride = 1;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(vrt_launch_rc_sky(P, ctx->rb, ctx->stream));
  CK(vrt_launch_gris(P, ctx->rb, (uint32_t)frame, ctx->stream));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the host buffers may be freed on return
  float ms = 0.0f;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->stats.last_render_ms = 0.0f;
  ctx->stats.last_gris_ms = ms;
  ctx->stats.kernel_launches = 2u;  // k_rc_sky, k_gris
  ctx->stats.launches_total += 2;
  return VRT_OK;
}

int vrt_accumulate_moving(vrt_ctx* ctx, int32_t sample, float render_scale, float max_accum) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(sample >= 0 && render_scale > 0.0f && render_scale <= 1.0f && max_accum >= 1.0f, "vrt_accumulate_moving: bad arguments");
  REQUIRE(!is_sharded(ctx), "vrt_accumulate_moving: sharding is not supported (history taps cross tiles)");
  int rc = check_ready(ctx, "vrt_accumulate_moving");
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const size_t npx = (size_t)ctx->cfg.width * ctx->cfg.height;
  auto& m = ctx->mv;
  if (!m.col_d) {
    CK(cudaMalloc(&m.col_d, npx * sizeof(float4)));
    CK(cudaMalloc(&m.col_s, npx * sizeof(float4)));
    CK(cudaMalloc(&m.out, npx * sizeof(float4)));
    CK(cudaMalloc(&m.full, npx * sizeof(float4)));
    CK(cudaMalloc(&m.refl, npx * sizeof(float)));
    CK(cudaMalloc(&m.refl_blur, npx * sizeof(float)));
    for (int k = 0; k < 2; k++) {
      CK(cudaMalloc(&m.hd[k], npx * sizeof(float4)));
      CK(cudaMalloc(&m.hs[k], npx * sizeof(float4)));
      CK(cudaMalloc(&m.hsd[k], npx * sizeof(float)));
      CK(cudaMalloc(&m.depth[k], npx * sizeof(float)));
      CK(cudaMalloc(&m.attr[k], npx * sizeof(uint2)));
    }
    m.has_prev = false;
  }
  if (!m.has_prev) {
    // reset_framebuffer semantics (pathtracer.py:664-668): empty history, previous matrices = current
    CK(cudaMemsetAsync(m.col_d, 0, npx * sizeof(float4), ctx->stream));
    CK(cudaMemsetAsync(m.col_s, 0, npx * sizeof(float4), ctx->stream));
    CK(cudaMemsetAsync(m.out, 0, npx * sizeof(float4), ctx->stream));
    CK(cudaMemsetAsync(m.refl, 0, npx * sizeof(float), ctx->stream));
    for (int k = 0; k < 2; k++) {
      CK(cudaMemsetAsync(m.hd[k], 0, npx * sizeof(float4), ctx->stream));
      CK(cudaMemsetAsync(m.hs[k], 0, npx * sizeof(float4), ctx->stream));
      CK(cudaMemsetAsync(m.hsd[k], 0, npx * sizeof(float), ctx->stream));
      CK(cudaMemsetAsync(m.depth[k], 0, npx * sizeof(float), ctx->stream));
      CK(cudaMemsetAsync(m.attr[k], 0, npx * sizeof(uint2), ctx->stream));
    }
    memcpy(m.prev_view, ctx->view, sizeof m.prev_view);
    memcpy(m.prev_proj, ctx->proj, sizeof m.prev_proj);
    m.has_prev = true;
    m.cur = 0;
  }
  m.scale = render_scale;
  m.active = true;
  const int cur = m.cur, prev = cur ^ 1;
  CK(cudaMemsetAsync(ctx->d_work, 0, sizeof(unsigned int), ctx->stream));
  Params P;
  fill_params(ctx, P);
  P.first_sample = sample, P.n_samples = 1, P.stride = 1;
  MovingOut MO{m.col_d, m.col_s, m.depth[cur], m.attr[cur], m.refl, render_scale};
  MovingFrame F;
  F.col_d = m.col_d, F.col_s = m.col_s, F.depth = m.depth[cur], F.attr = m.attr[cur], F.refl = m.refl, F.refl_blur = m.refl_blur;
  F.depth_prev = m.depth[prev], F.attr_prev = m.attr[prev], F.hd_prev = m.hd[prev], F.hs_prev = m.hs[prev], F.hsd_prev = m.hsd[prev];
  F.hd = m.hd[cur], F.hs = m.hs[cur], F.hsd = m.hsd[cur], F.out = m.out;
  memcpy(F.prev_view, m.prev_view, sizeof F.prev_view);
  memcpy(F.prev_proj, m.prev_proj, sizeof F.prev_proj);
  F.scale = render_scale;
  // pixels the filters skip keep their history (the reference copies slot 1 -> slot 0 for every
  // pixel, :1297-1303): start this frame's slot from the previous one
  CK(cudaMemcpyAsync(m.hd[cur], m.hd[prev], npx * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(m.hs[cur], m.hs[prev], npx * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(m.hsd[cur], m.hsd[prev], npx * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(m.depth[cur], m.depth[prev], npx * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(m.attr[cur], m.attr[prev], npx * sizeof(uint2), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(vrt_launch_path_moving(P, MO, ctx->sm_count, ctx->stream));
  CK(vrt_launch_moving_filters(P, F, max_accum, ctx->stream));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventElapsedTime(&ctx->stats.last_render_ms, ctx->ev0, ctx->ev1));
  ctx->stats.kernel_launches = 3;
  ctx->stats.launches_total += 3;
  // copy_prev_matrices (pathtracer.py:284-287) and the slot swap
  memcpy(m.prev_view, ctx->view, sizeof m.prev_view);
  memcpy(m.prev_proj, ctx->proj, sizeof m.prev_proj);
  m.cur = prev;
  return VRT_OK;
}

int vrt_get_reservoirs(vrt_ctx* ctx, void* out) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(out, "vrt_get_reservoirs: null pointer");
  if (!ctx->rb.reservoirs) {
    ctx->err = "vrt_get_reservoirs: no ReSTIR frame has been rendered";
    return VRT_ERR_NOT_PREPARED;
  }
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->rb.reservoirs, (size_t)ctx->cfg.width * ctx->cfg.height * 56, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}

// Accumulation checkpoint: the float4 sums + sample counts of the current slot, to / from host memory.
int vrt_get_accum(vrt_ctx* ctx, float* rgba_sums) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgba_sums, "vrt_get_accum: null pointer");
  CK(cudaSetDevice(ctx->device));
  if (int rc = flush_pending_zero(ctx)) return rc;
  CK(cudaMemcpyAsync(rgba_sums, ctx->d_accum, (size_t)ctx->cfg.width * ctx->cfg.height * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}
int vrt_set_accum(vrt_ctx* ctx, const float* rgba_sums) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgba_sums, "vrt_set_accum: null pointer");
  CK(cudaSetDevice(ctx->device));
  ctx->zero_pending[ctx->cur_slot] = false;
  ctx->mv.active = false;
  CK(cudaMemcpyAsync(ctx->d_accum, rgba_sums, (size_t)ctx->cfg.width * ctx->cfg.height * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}

static int flush_all_pending_zero(vrt_ctx* ctx) {
  CK(cudaSetDevice(ctx->device));
  const int keep = ctx->cur_slot;
  for (int k = 0; k < 2; k++)
    if (ctx->accum_slot[k]) {
      ctx->cur_slot = k, ctx->d_accum = ctx->accum_slot[k];
      if (int rcz = flush_pending_zero(ctx)) return rcz;
    }
  ctx->cur_slot = keep, ctx->d_accum = ctx->accum_slot[keep];
  return VRT_OK;
}

int vrt_set_row_shard(vrt_ctx* ctx, int32_t rank, int32_t n) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  const int rows = ctx->cfg.height / 4;
  REQUIRE(n >= 1 && rank >= 0 && rank < n && n <= rows, "vrt_set_row_shard: need 0 <= rank < n <= height / 4");
  REQUIRE(n == 1 || ctx->tile_n == 1, "vrt_set_row_shard: interleaved tile sharding is already active");
  if (n != 1)
    if (int rc = flush_all_pending_zero(ctx)) return rc;  // a strip covers only its rows: pending resets must really clear
  ctx->strip_n = n;
  ctx->strip_row0 = (int)((long long)rank * rows / n), ctx->strip_row1 = (int)((long long)(rank + 1) * rows / n);
  ctx->hist_valid = false;
  return VRT_OK;
}

int vrt_set_row_range(vrt_ctx* ctx, int32_t first_row, int32_t n_rows) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  const int rows = ctx->cfg.height / 4;
  REQUIRE(first_row >= 0 && n_rows >= 0 && first_row + n_rows <= rows, "vrt_set_row_range: tile rows outside the frame");
  REQUIRE(ctx->tile_n == 1, "vrt_set_row_range: interleaved tile sharding is already active");
  if (int rc = flush_all_pending_zero(ctx)) return rc;
  ctx->strip_n = 2;  // any value > 1: "a strip is active"
  ctx->strip_row0 = first_row, ctx->strip_row1 = first_row + n_rows;
  ctx->hist_valid = false;
  return VRT_OK;
}

int vrt_set_tile_shard(vrt_ctx* ctx, int32_t rank, int32_t n) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(n >= 1 && rank >= 0 && rank < n, "vrt_set_tile_shard: need 0 <= rank < n");
  REQUIRE(n == 1 || ctx->strip_n <= 1, "vrt_set_tile_shard: row sharding is already active");
  if (n != 1) {  // a sharded batch covers only its own tiles: pending resets must really clear the buffers
    CK(cudaSetDevice(ctx->device));
    const int keep = ctx->cur_slot;
    for (int k = 0; k < 2; k++)
      if (ctx->accum_slot[k]) {
        ctx->cur_slot = k, ctx->d_accum = ctx->accum_slot[k];
        if (int rcz = flush_pending_zero(ctx)) return rcz;
      }
    ctx->cur_slot = keep, ctx->d_accum = ctx->accum_slot[keep];
  }
  ctx->tile_rank = rank, ctx->tile_n = n;
  return VRT_OK;
}

int vrt_reset(vrt_ctx* ctx) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  CK(cudaSetDevice(ctx->device));
  // deferred: a full-frame path-tracing batch that follows overwrites the buffer (no memset, no read-modify-write);
  // every other consumer turns the flag into the memset first (flush_pending_zero)
  ctx->zero_pending[ctx->cur_slot] = true;
  if (is_sharded(ctx))
    if (int rcz = flush_pending_zero(ctx)) return rcz;
  ctx->mv.has_prev = false;
  ctx->mv.active = false;
  ctx->hist_valid = false;
  return VRT_OK;
}

int vrt_accum_device_ptr(vrt_ctx* ctx, void** ptr, uint64_t* bytes) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(ptr, "vrt_accum_device_ptr: null pointer");
  CK(cudaSetDevice(ctx->device));
  if (int rcz = flush_pending_zero(ctx)) return rcz;  // the caller is about to use the memory directly
  *ptr = ctx->d_accum;
  if (bytes) *bytes = (uint64_t)ctx->cfg.width * ctx->cfg.height * sizeof(float4);
  return VRT_OK;
}

static int wait_pending_copy(vrt_ctx* ctx) {
  if (ctx->copy_pending) {
    CK(cudaEventSynchronize(ctx->ev_copied));
    ctx->copy_pending = false;
  }
  return VRT_OK;
}

static int resolve(vrt_ctx* ctx, bool ldr, float* host) {
  CK(cudaSetDevice(ctx->device));
  if (int rc = wait_pending_copy(ctx)) return rc;  // d_out may still be the source of a pipelined copy
  if (int rc = flush_pending_zero(ctx)) return rc;
  const int W = ctx->cfg.width, H = ctx->cfg.height;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  const float4* src = ctx->d_accum;
  if (ctx->mv.active) {  // the last frame came from the moving-camera path: show its colour buffer
    CK(vrt_launch_moving_upsample(ctx->mv.out, ctx->mv.full, W, H, ctx->mv.scale, ctx->stream));
    src = ctx->mv.full;
  }
  CK(vrt_launch_resolve(src, ldr ? nullptr : ctx->d_out, ldr ? ctx->d_out : nullptr, W, H, ctx->cfg.exposure, ctx->stream));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (host) CK(cudaMemcpyAsync(host, ctx->d_out, (size_t)W * H * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventElapsedTime(&ctx->stats.last_resolve_ms, ctx->ev0, ctx->ev1));
  return VRT_OK;
}

int vrt_fetch_hdr(vrt_ctx* ctx, float* rgba) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgba, "vrt_fetch_hdr: null pointer");
  return resolve(ctx, false, rgba);
}
int vrt_fetch_ldr(vrt_ctx* ctx, float* rgba) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgba, "vrt_fetch_ldr: null pointer");
  return resolve(ctx, true, rgba);
}
// Pipelined variant of vrt_fetch_ldr: the tonemap pass runs on the context's stream, the device-to-host
// copy on a separate copy-engine stream, and the call returns without waiting for it, so the copy of
// frame k overlaps the rendering of frame k+1. The image is complete after vrt_fetch_wait (or the next
// vrt_fetch_* call, which waits first: the device image buffer is single).
int vrt_fetch_ldr_async(vrt_ctx* ctx, float* rgba_pinned) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgba_pinned, "vrt_fetch_ldr_async: null pointer");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_resolved, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
  }
  if (int rc = wait_pending_copy(ctx)) return rc;
  if (int rc = flush_pending_zero(ctx)) return rc;
  const int W = ctx->cfg.width, H = ctx->cfg.height;
  const float4* src = ctx->d_accum;
  if (ctx->mv.active) {
    CK(vrt_launch_moving_upsample(ctx->mv.out, ctx->mv.full, W, H, ctx->mv.scale, ctx->stream));
    src = ctx->mv.full;
  }
  CK(vrt_launch_resolve(src, nullptr, ctx->d_out, W, H, ctx->cfg.exposure, ctx->stream));
  CK(cudaEventRecord(ctx->ev_resolved, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_resolved, 0));
  CK(cudaMemcpyAsync(rgba_pinned, ctx->d_out, (size_t)W * H * sizeof(float4), cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
  ctx->copy_pending = true;
  return VRT_OK;
}
int vrt_fetch_wait(vrt_ctx* ctx) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  CK(cudaSetDevice(ctx->device));
  return wait_pending_copy(ctx);
}
int vrt_resolve_ldr_device(vrt_ctx* ctx, void** ptr) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(ptr, "vrt_resolve_ldr_device: null pointer");
  int rc = resolve(ctx, true, nullptr);
  *ptr = ctx->d_out;
  return rc;
}

int vrt_accum_ipc_handle(vrt_ctx* ctx, void* handle64) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(handle64, "vrt_accum_ipc_handle: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  if (int rcz = flush_pending_zero(ctx)) return rcz;
  CK(cudaIpcGetMemHandle(&h, ctx->d_accum));
  memcpy(handle64, &h, sizeof h);
  return VRT_OK;
}

int vrt_open_peer_accum(vrt_ctx* ctx, const void* handle64, void** peer_ptr) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(handle64 && peer_ptr, "vrt_open_peer_accum: null pointer");
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof h);
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return VRT_OK;
}

int vrt_close_peer_accum(vrt_ctx* ctx, void* peer_ptr) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (peer_ptr) CK(cudaIpcCloseMemHandle(peer_ptr));
  return VRT_OK;
}

int vrt_fetch_ldr_merged(vrt_ctx* ctx, const void* const* peer_ptrs, int32_t n_peers, float* ldr_rgba) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(n_peers >= 0 && n_peers <= 8 && (n_peers == 0 || peer_ptrs), "vrt_fetch_ldr_merged: 0..8 peers");
  REQUIRE(!ctx->mv.active, "vrt_fetch_ldr_merged: the last frame came from the moving-camera path, which does not shard");
  CK(cudaSetDevice(ctx->device));
  if (int rc = wait_pending_copy(ctx)) return rc;  // d_out may still be the source of a pipelined copy
  if (int rc = flush_pending_zero(ctx)) return rc;
  const int W = ctx->cfg.width, H = ctx->cfg.height;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(vrt_launch_resolve_merged(ctx->d_accum, reinterpret_cast<const float4* const*>(peer_ptrs), n_peers, ctx->d_out, W, H, ctx->cfg.exposure, ctx->stream));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (ldr_rgba) CK(cudaMemcpyAsync(ldr_rgba, ctx->d_out, (size_t)W * H * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventElapsedTime(&ctx->stats.last_resolve_ms, ctx->ev0, ctx->ev1));
  return VRT_OK;
}

// ---- double-buffered accumulation and the fused multi-GPU merge (reduce-scatter + tonemap over peer memory)
int vrt_set_accum_slot(vrt_ctx* ctx, int32_t slot) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(slot == 0 || slot == 1, "vrt_set_accum_slot: slot must be 0 or 1");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->accum_slot[slot]) {
    const size_t bytes = (size_t)ctx->cfg.width * ctx->cfg.height * sizeof(float4);
    CK(cudaMalloc(&ctx->accum_slot[slot], bytes));
    CK(cudaMemsetAsync(ctx->accum_slot[slot], 0, bytes, ctx->stream));
    CK(cudaMalloc(&ctx->out_slot[slot], bytes));
  }
  ctx->cur_slot = slot;
  ctx->d_accum = ctx->accum_slot[slot];
  ctx->d_out = ctx->out_slot[slot];
  return VRT_OK;
}

int vrt_out_ipc_handle(vrt_ctx* ctx, void* handle64) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(handle64, "vrt_out_ipc_handle: null pointer");
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->d_out));
  memcpy(handle64, &h, sizeof h);
  return VRT_OK;
}

int vrt_merge_slice(vrt_ctx* ctx, const void* const* peer_ptrs, int32_t n_peers, int32_t first_pixel, int32_t n_pixels, void* ldr_dst,
                    int32_t write_sums) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  const int npx = ctx->cfg.width * ctx->cfg.height;
  REQUIRE(n_peers >= 0 && n_peers <= 8 && (n_peers == 0 || peer_ptrs), "vrt_merge_slice: 0..8 peers");
  REQUIRE(first_pixel >= 0 && n_pixels >= 0 && first_pixel + n_pixels <= npx, "vrt_merge_slice: pixel range outside the frame");
  REQUIRE(!ctx->mv.active, "vrt_merge_slice: the last frame came from the moving-camera path, which does not shard");
  CK(cudaSetDevice(ctx->device));
  if (int rc = flush_pending_zero(ctx)) return rc;
  if (!ldr_dst)
    if (int rc = wait_pending_copy(ctx)) return rc;  // own image buffer: it may still be the source of a pipelined copy
  float4* ldr = ldr_dst ? reinterpret_cast<float4*>(ldr_dst) : ctx->d_out;
  CK(vrt_launch_merge_slice(ctx->d_accum, reinterpret_cast<const float4* const*>(peer_ptrs), n_peers, write_sums ? ctx->d_accum : nullptr, ldr,
                            first_pixel, n_pixels, ctx->cfg.width, ctx->cfg.height, ctx->cfg.exposure, ctx->stream));
  ctx->stats.launches_total += 1;
  return VRT_OK;
}

int vrt_copy_ldr_async(vrt_ctx* ctx, float* rgba_pinned) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(rgba_pinned, "vrt_copy_ldr_async: null pointer");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_resolved, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
  }
  if (int rc = wait_pending_copy(ctx)) return rc;
  CK(cudaEventRecord(ctx->ev_resolved, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_resolved, 0));
  CK(cudaMemcpyAsync(rgba_pinned, ctx->d_out, (size_t)ctx->cfg.width * ctx->cfg.height * sizeof(float4), cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
  ctx->copy_pending = true;
  return VRT_OK;
}

int vrt_stream_wait_copy(vrt_ctx* ctx) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->copy_pending) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));
  return VRT_OK;
}

int vrt_get_stats(vrt_ctx* ctx, vrt_stats* out) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  REQUIRE(out, "vrt_get_stats: null pointer");
  CK(cudaSetDevice(ctx->device));
  if (int rc = drain_ring(ctx)) return rc;  // waits for the asynchronous launches still in flight
  if (int rc = drain_restir_events(ctx)) return rc;
  *out = ctx->stats;
  ctx->stats.render_ms_sum = 0.0f, ctx->stats.render_launches = 0, ctx->stats.launches_total = 0;  // "since the last query"
  return VRT_OK;
}

int vrt_synchronize(vrt_ctx* ctx) {
  if (!ctx) return VRT_ERR_BAD_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return VRT_OK;
}

}  // extern "C"
