// k_path_pool — EXPERIMENTAL shared-memory wavefront variant of the path kernel (VRT_KERNEL=pool).
// Parity-green (same tests as k_path) but 25-40 % slower than the per-lane kernel in round 1:
// 16.8 vs 11 active threads per instruction, but issue utilisation drops from 63 % to 43 % (five
// CTA barriers per iteration with 2 CTAs/SM, phase tails, sky / accumulation loads exposed at 25 %
// occupancy). Kept as the starting point for a barrier-free queue version; see
// profiles/r01_pool_experiment.md.
//
// Replaces Renderer.render + the static-camera temporal filters (renderer/pathtracer.py:355-632,
// :1185-1303) like k_path in vrt_render.cu, with the same estimator and sampler, but schedules
// the work as a WAVEFRONT INSIDE SHARED MEMORY instead of one path per lane:
//
//   * each CTA (256 threads, 2 CTAs per SM) owns a pool of 640 path states in shared memory
//     (31 words per path, structure-of-arrays);
//   * one pool iteration = five phases separated by __syncthreads():
//       T  trace      every live path has exactly one pending ray (segment or sun shadow ray);
//                     warps pull rays from the pool with a persistent "while-while" DDA loop and
//                     refill idle lanes in batches, so short and long rays do not share a warp's
//                     fate (warp vote + popc ranks for the refill, smem atomic for the cursor);
//       C  classify   one thread per slot: floor test, voxel colour fetch, hit / miss / emissive,
//                     NEE direction; builds compact sky / shade / retire lists (ballot compaction);
//       K  sky        full warps over the escaped paths: two bilinear sky-table lookups;
//       S  shade      full warps over the surface vertices: Disney eval + pdf, MIS, BSDF sample;
//       F  retire     full warps over finished paths: pixel sample value, next sample of the pixel
//                     or the single accumulation read-modify-write; free slots take new 8x4 tiles.
//   * a pixel stays in its slot for all `spp` samples of the launch, so accumulation order is
//     deterministic and there is one 32-byte RMW per pixel per launch.
//
// ncu on the per-lane kernel showed 10-11 active threads per warp instruction (DDA trip-count
// variance, shade run by ~45 % of the lanes, sky lookups by ~20 %); the phases above run each of
// those pieces over compacted lists instead (profiles/r01*_k_path*.md).
#include "vrt_bsdf.cuh"
#include "vrt_internal.h"
#include "vrt_sky.cuh"
#include "vrt_trace.cuh"

namespace {

#define POOL_THREADS 256
#define POOL_SLOTS 640
#define POOL_WORDS 31
#define REFILL_BELOW 20  // refill a warp's idle lanes when fewer than this many are tracing

#define RADIANCE_CLAMP 300.0f
HD f3 firefly_filter(f3 v) { return clamp3(v, 0.0f, RADIANCE_CLAMP); }
HD float power_heuristic(float a, float b) {
  float a_sqr = a * a;
  return __fdividef(a_sqr, fmaxf(a_sqr + b * b, 1e-4f));
}
HD uint32_t encode_material(int mat_id, f3 albedo) {
  return (uint32_t)mat_id | ((uint32_t)(albedo.x * 255.0f) << 8) | ((uint32_t)(albedo.y * 255.0f) << 16) | ((uint32_t)(albedo.z * 255.0f) << 24);
}
HD bool bad3(f3 c) { return isbad(c.x) || isbad(c.y) || isbad(c.z) || c.x < 0.0f || c.y < 0.0f || c.z < 0.0f; }

// field indices of the SoA pool (word arrays of POOL_SLOTS entries)
enum {
  F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ,        // pending ray (world space)
  F_TR, F_TG, F_TB,                          // throughput
  F_CR, F_CG, F_CB,                          // contrib
  F_NDR, F_NDG, F_NDB, F_NSR, F_NSG, F_NSB,  // first-vertex NEE, diffuse / specular
  F_AR, F_AG, F_AB,                          // per-pixel accumulator of this launch
  F_INVPDF, F_PIX, F_MISC, F_PMINFO,
  F_HIT_T,     // T: voxel-space t of the pending ray (inf = miss)
  F_HIT_CELL,  // T: cell x | y<<10 | z<<20 | in-grid flag << 30 (the face normal goes into F_MISC)
  F_HIT_COL,   // C: RGBA8 colour word of the surface (kept while the shadow ray flies)
  F_VX, F_VY, F_VZ,  // view vector at the surface (= -segment direction)
};
// F_MISC bits: 0-3 depth, 4 state (0 segment, 1 shadow), 5-6 first lobe, 7 visible, 8-9 kind (1 floor, 2 voxel),
//              10 edge, 11-16 normal code, 17-31 sample counter s_i
#define MISC_DEPTH(m) ((m) & 15u)
#define MISC_STATE(m) (((m) >> 4) & 1u)
#define MISC_LOBE(m) (((m) >> 5) & 3u)
#define MISC_SI(m) ((m) >> 17)

HD uint32_t normal_code(float nx, float ny, float nz) {  // components in {-1, 0, 1} -> 2 bits each
  return (uint32_t)((int)nx + 1) | ((uint32_t)((int)ny + 1) << 2) | ((uint32_t)((int)nz + 1) << 4);
}
HD f3 normal_decode(uint32_t c) { return f3{(float)((int)(c & 3u) - 1), (float)((int)((c >> 2) & 3u) - 1), (float)((int)((c >> 4) & 3u) - 1)}; }

struct Pool {
  uint32_t* w;  // POOL_WORDS x POOL_SLOTS
  HD float& f(int field, int slot) const { return reinterpret_cast<float*>(w)[field * POOL_SLOTS + slot]; }
  HD uint32_t& u(int field, int slot) const { return w[field * POOL_SLOTS + slot]; }
  HD int& i(int field, int slot) const { return reinterpret_cast<int*>(w)[field * POOL_SLOTS + slot]; }
  HD f3 get3(int field, int slot) const { return f3{f(field, slot), f(field + 1, slot), f(field + 2, slot)}; }
  HD void set3(int field, int slot, f3 v) const { f(field, slot) = v.x, f(field + 1, slot) = v.y, f(field + 2, slot) = v.z; }
};

struct Lists {
  unsigned short* sky;
  unsigned short* shade;
  unsigned short* retire;
  unsigned short* freel;
  int* counts;  // [0] sky [1] shade [2] retire [3] free [4] trace cursor [5] live [6] tile base [7] tiles taken
};

// Append `slot` to a list for every lane with `pred` (warp-aggregated shared-memory atomic).
HD void push_list(unsigned short* list, int* count, bool pred, int slot) {
  const unsigned m = __ballot_sync(0xffffffffu, pred);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)slot;
}

// (Re)start the path of `slot` for sample s_i of its pixel (pathtracer.py:331-347).
HD void start_path(const Params& P, const Pool& S, int slot, int pix, uint32_t s_i) {
  const int u = pix & 0xffff, v = pix >> 16;
  const float2 j = P.jitter[s_i];
  const f3 d = get_cast_dir(P, (float)u, (float)v, j.x, j.y);
  S.set3(F_OX, slot, P.cam_pos);
  S.set3(F_DX, slot, d);
  S.set3(F_TR, slot, mk3(1.0f));
  S.set3(F_CR, slot, mk3(0.0f));
  S.set3(F_NDR, slot, mk3(0.0f));
  S.set3(F_NSR, slot, mk3(0.0f));
  S.f(F_INVPDF, slot) = 1.0f;
  S.u(F_PMINFO, slot) = 0u;
  S.u(F_MISC, slot) = s_i << 17;  // depth 0, segment ray, lobe 0
}

template <bool STATS>
__global__ void __launch_bounds__(POOL_THREADS, 2) k_path_pool(const __grid_constant__ Params P, int upper_in_smem, int fixed_words, int upper_words_smem) {
  extern __shared__ uint32_t smem[];
  // ---- static staging: materials, UNORM8 table, upper pyramid
  float4* s_mats = reinterpret_cast<float4*>(smem);
  for (int i = threadIdx.x; i < 128 * MAT_ROW_F4; i += blockDim.x) s_mats[i] = P.mats[i];
  float* unorm8 = reinterpret_cast<float*>(smem + 128 * MAT_ROW_F4 * 4);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) unorm8[i] = xdiv((float)i, 255.0f);
  const uint32_t* upper = P.upper;
  if (upper_in_smem) {
    uint32_t* s_upper = smem + fixed_words;
    for (int i = threadIdx.x; i < P.upper_words; i += blockDim.x) s_upper[i] = P.upper[i];
    upper = s_upper;
  }
  uint32_t* dyn = smem + fixed_words + upper_words_smem;
  Pool S{dyn};
  Lists L;
  L.sky = reinterpret_cast<unsigned short*>(dyn + POOL_WORDS * POOL_SLOTS);
  L.shade = L.sky + POOL_SLOTS;
  L.retire = L.shade + POOL_SLOTS;
  L.freel = L.retire + POOL_SLOTS;
  L.counts = reinterpret_cast<int*>(L.freel + POOL_SLOTS);
  for (int s = threadIdx.x; s < POOL_SLOTS; s += blockDim.x) S.i(F_PIX, s) = -1;
  if (threadIdx.x < 8) L.counts[threadIdx.x] = 0;
  __syncthreads();

  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  f3 sun_bx, sun_by;
  make_orthonormal_basis(P.light_dir, sun_bx, sun_by);
  const float light_pdf_axis = cone_sample_pdf(P.light_cos_max, 1.0f);
  const f3 sun_rad = P.light_weight * P.light_color;
  const float sky_fres = 1.0f / (float)P.sky_res;
  const float Rf = (float)P.R;
  const int top_lod = P.n_lods - 1;
  uint32_t c_rays = 0, c_steps = 0, c_queries = 0, c_hits = 0, c_escapes = 0, c_nee = 0, c_vertices = 0, c_paths = 0;
  bool work_left = true;  // uniform per CTA (derived from shared state after a barrier)

  for (;;) {
    // =========================================================== R: free slots take new tiles
    if (tid < 4) L.counts[tid] = 0;
    if (tid == 4) L.counts[4] = 0;
    __syncthreads();
    for (int base = 0; base < POOL_SLOTS; base += POOL_THREADS) {
      const int s = base + tid;
      const bool is_free = s < POOL_SLOTS && S.i(F_PIX, s) < 0;
      push_list(L.freel, &L.counts[3], is_free, s);
    }
    __syncthreads();
    const int n_free = L.counts[3];
    if (tid == 0) {
      int want = work_left ? n_free / 32 : 0, got = 0, base = 0;
      if (want > 0) {
        base = (int)atomicAdd(P.work_counter, (unsigned)want);
        got = min(want, max(P.n_tiles - base, 0));
      }
      L.counts[6] = base, L.counts[7] = got;
    }
    __syncthreads();
    const int tiles_got = L.counts[7], tile_base = L.counts[6];
    if (work_left && tiles_got < n_free / 32) work_left = false;  // the global queue ran dry
    for (int j = tid; j < tiles_got * 32; j += POOL_THREADS) {
      const int slot = L.freel[j];
      const int tile = P.tile_rank + P.tile_n * (tile_base + (j >> 5));
      const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
      const int pix = (tx * 8 + (j & 7)) | ((ty * 4 + ((j >> 3) & 3)) << 16);
      S.i(F_PIX, slot) = pix;
      S.set3(F_AR, slot, mk3(0.0f));
      start_path(P, S, slot, pix, 0u);
      if (STATS) c_paths++;
    }
    __syncthreads();
    if (n_free - tiles_got * 32 == POOL_SLOTS) break;  // nothing live and nothing left to fetch

    // =========================================================== T: trace all pending rays
    {
      int slot = -1;
      bool exhausted = false;
      // per-ray traversal registers (raytracer.py:72-155, same op order as vrt_trace.cuh::raytrace)
      f3 o = mk3(0.0f), d = mk3(0.0f);
      float ivx = 0.0f, ivy = 0.0f, ivz = 0.0f, hit_distance = 0.0f, far = 0.0f, nx = 0.0f, ny = 0.0f, nz = 0.0f;
      int px = 0, py = 0, pz = 0, lod = 0, last_b = -1, iters = 0;
      unsigned long long w = 0ull;
      bool shadow = false;
      for (;;) {
        unsigned act = __ballot_sync(FULL, slot >= 0);
        if (!exhausted && __popc(act) < REFILL_BELOW) {
          const unsigned idle = ~act;
          const int n = __popc(idle);
          int base = 0;
          if (lane == 0) base = atomicAdd(&L.counts[4], n);
          base = __shfl_sync(FULL, base, 0);
          if (slot < 0) {
            const int s = base + __popc(idle & lt_mask);
            if (s < POOL_SLOTS && S.i(F_PIX, s) >= 0) {
              // ---- ray setup: world -> voxel space, box test, start cell
              slot = s;
              shadow = MISC_STATE(S.u(F_MISC, s)) != 0u;
              const f3 pos = S.get3(F_OX, s);
              d = S.get3(F_DX, s);
              o = f3{xsub(xmul(P.voxel_inv_size, pos.x), -P.grid_half), xsub(xmul(P.voxel_inv_size, pos.y), -P.grid_half),
                     xsub(xmul(P.voxel_inv_size, pos.z), -P.grid_half)};
              if (STATS) c_rays++;
              float near_int = -VRT_INF, far_int = VRT_INF;
              const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
              for (int i = 0; i < 3; i++) {
                if (dd[i] != 0.0f) {
                  float i1 = xdiv(xsub(0.0f, oo[i]), dd[i]);
                  float i2 = xdiv(xsub(Rf, oo[i]), dd[i]);
                  far_int = fminf(fmaxf(i1, i2), far_int);
                  near_int = fmaxf(fminf(i1, i2), near_int);
                }
              }
              if (!(near_int <= far_int && VRT_EPS < far_int && near_int < VRT_INF)) {
                S.f(F_HIT_T, s) = VRT_INF;  // the ray never enters the grid
                slot = -1;
              } else {
                hit_distance = fmaxf(near_int, VRT_EPS);
                const float t0 = xadd(hit_distance, VRT_EPS);
                const float ipx = xadd(o.x, xmul(d.x, t0)), ipy = xadd(o.y, xmul(d.y, t0)), ipz = xadd(o.z, xmul(d.z, t0));
                px = (int)clampf(floorf(ipx), 0.0f, Rf - 1.0f);
                py = (int)clampf(floorf(ipy), 0.0f, Rf - 1.0f);
                pz = (int)clampf(floorf(ipz), 0.0f, Rf - 1.0f);
                ivx = __frcp_rn(fabsf(d.x)), ivy = __frcp_rn(fabsf(d.y)), ivz = __frcp_rn(fabsf(d.z));
                far = xsub(fminf(VRT_INF, far_int), VRT_EPS);
                const float ax = fabsf(xsub(ipx, xmul(Rf, 0.5f))), ay = fabsf(xsub(ipy, xmul(Rf, 0.5f))), az = fabsf(xsub(ipz, xmul(Rf, 0.5f)));
                const float m = fmaxf(fmaxf(ax, ay), az);
                nx = (m == ax) ? 1.0f : 0.0f, ny = (m == ay) ? 1.0f : 0.0f, nz = (m == az) ? 1.0f : 0.0f;
                lod = 0, last_b = -1, iters = 0;
              }
            }
          }
          if (base + n >= POOL_SLOTS) exhausted = true;
          act = __ballot_sync(FULL, slot >= 0);
        }
        if (act == 0u) {
          if (exhausted) break;
          continue;
        }
        if (slot >= 0) {
          // ---- one iteration of the hierarchical DDA
          bool finished = false;
          float t_out = VRT_INF;
          if (iters >= 512) {
            finished = true, t_out = hit_distance;  // iteration cap reports a hit (pinned)
          } else if (hit_distance > far) {
            finished = true;
          } else if ((unsigned)px >= (unsigned)P.R || (unsigned)py >= (unsigned)P.R || (unsigned)pz >= (unsigned)P.R) {
            finished = true;  // stepped outside the grid (pinned)
          } else {
            bool occ = true;
            while (lod >= 3) {
              occ = upper_bit(P, upper, px >> lod, py >> lod, pz >> lod, lod);
              if (STATS) c_queries++;
              if (!occ) break;
              lod--;
            }
            if (occ) {
              const int b = ((pz >> 2) * P.brick_res + (py >> 2)) * P.brick_res + (px >> 2);
              if (b != last_b) {
                w = __ldg(P.bricks + b);
                last_b = b;
              }
              if (lod == 2) {
                if (STATS) c_queries++;
                if (w == 0ull)
                  occ = false;
                else
                  lod = 1;
              }
              if (occ && lod == 1) {
                if (STATS) c_queries++;
                const int sh = ((px >> 1) & 1) * 2 + ((py >> 1) & 1) * 8 + ((pz >> 1) & 1) * 32;
                if ((w & (0x0000000000330033ull << sh)) == 0ull)
                  occ = false;
                else
                  lod = 0;
              }
              if (occ && lod == 0) {
                if (STATS) c_queries++;
                occ = (w >> ((pz & 3) * 16 + (py & 3) * 4 + (px & 3))) & 1ull;
              }
            }
            if (occ) {
              finished = true, t_out = hit_distance;
            } else {
              const float cell_size = (float)(1 << lod);
              const float bx = xmul((float)(px >> lod), cell_size), by = xmul((float)(py >> lod), cell_size), bz = xmul((float)(pz >> lod), cell_size);
              const float fx = xsub(xadd(o.x, xmul(d.x, hit_distance)), bx);
              const float fy = xsub(xadd(o.y, xmul(d.y, hit_distance)), by);
              const float fz = xsub(xadd(o.z, xmul(d.z, hit_distance)), bz);
              float tx = xmul(d.x > 0.0f ? xsub(cell_size, fx) : fx, ivx);
              float ty = xmul(d.y > 0.0f ? xsub(cell_size, fy) : fy, ivy);
              float tz = xmul(d.z > 0.0f ? xsub(cell_size, fz) : fz, ivz);
              if (d.x == 0.0f) tx = VRT_INF;
              if (d.y == 0.0f) ty = VRT_INF;
              if (d.z == 0.0f) tz = VRT_INF;
              const float min_t = fminf(fminf(tx, ty), tz);
              const float ex = clampf(floorf(xadd(fx, xmul(min_t, d.x))), 0.0f, cell_size - 1.0f);
              const float ey = clampf(floorf(xadd(fy, xmul(min_t, d.y))), 0.0f, cell_size - 1.0f);
              const float ez = clampf(floorf(xadd(fz, xmul(min_t, d.z))), 0.0f, cell_size - 1.0f);
              hit_distance = xadd(hit_distance, min_t);
              nx = (tx == min_t ? 1.0f : 0.0f) * signf(d.x);
              ny = (ty == min_t ? 1.0f : 0.0f) * signf(d.y);
              nz = (tz == min_t ? 1.0f : 0.0f) * signf(d.z);
              px = (int)(bx + ex + nx);
              py = (int)(by + ey + ny);
              pz = (int)(bz + ez + nz);
              lod = min(top_lod, lod + 1);
              iters++;
              if (STATS) c_steps++;
            }
          }
          if (finished) {
            S.f(F_HIT_T, slot) = t_out;
            if (!shadow && t_out < VRT_INF) {
              if (xadd(xadd(xmul(d.x, nx), xmul(d.y, ny)), xmul(d.z, nz)) > 0.0f) nx = -nx, ny = -ny, nz = -nz;
              S.u(F_HIT_CELL, slot) = ((uint32_t)px & 1023u) | (((uint32_t)py & 1023u) << 10) | (((uint32_t)pz & 1023u) << 20) |
                                      ((uint32_t)((unsigned)px < (unsigned)P.R && (unsigned)py < (unsigned)P.R && (unsigned)pz < (unsigned)P.R) << 30);
              uint32_t m = S.u(F_MISC, slot);
              S.u(F_MISC, slot) = (m & ~(63u << 11)) | (normal_code(nx, ny, nz) << 11);
            }
            slot = -1;
          }
        }
      }
    }
    __syncthreads();

    // =========================================================== C: classify the traced rays
    for (int base = 0; base < POOL_SLOTS; base += POOL_THREADS) {
      const int s = base + tid;
      bool to_sky = false, to_shade = false, to_retire = false;
      if (s < POOL_SLOTS && S.i(F_PIX, s) >= 0) {
        uint32_t misc = S.u(F_MISC, s);
        const f3 pos = S.get3(F_OX, s), d = S.get3(F_DX, s);
        const float tv = S.f(F_HIT_T, s);
        // floor plane first, then the grid (pathtracer.py:173-216)
        float closest = VRT_INF;
        int kind = 0;
        float fny = 1.0f;
        {
          const float dist = xdiv(xsub(P.floor_height, pos.y), d.y);
          if (dist > VRT_EPS && dist < closest) {
            const float hx = xadd(pos.x, xmul(d.x, dist)), hy = xadd(pos.y, xmul(d.y, dist)), hz = xadd(pos.z, xmul(d.z, dist));
            const float dn = xadd(xadd(xmul(hx, 0.0f), xmul(hy, 1.0f)), xmul(hz, 0.0f));
            const float ax = xsub(hx, dn), ay = xsub(hy, dn), az = xsub(hz, dn);
            if (xsqrt(xadd(xadd(xmul(ax, ax), xmul(ay, ay)), xmul(az, az))) < 10.0f) {
              closest = dist, kind = 1;
              if (xadd(xadd(xmul(0.0f, d.x), xmul(1.0f, d.y)), xmul(0.0f, d.z)) > 0.0f) fny = -1.0f;
            }
          }
        }
        const float tw = xmul(tv, P.voxel_size);
        if (tw < closest) closest = tw, kind = 2;
        if (MISC_STATE(misc) == 1u) {
          // sun shadow ray came back (pathtracer.py:444-449)
          misc = (misc & ~((1u << 4) | (1u << 7))) | ((closest >= VRT_INF ? 1u : 0u) << 7);
          S.u(F_MISC, s) = misc;
          to_shade = true;
        } else if (closest == VRT_INF) {
          to_sky = true;
        } else {
          // surface attributes (voxel_world.py:34-56)
          f3 n, albedo;
          int mat, hit_light;
          uint32_t colw = 0u, edge = 0u;
          if (kind == 1) {
            n = f3{0.0f, fny, 0.0f};
            albedo = P.floor_color, mat = P.floor_material, hit_light = P.floor_material == 2;
          } else {
            const uint32_t cellw = S.u(F_HIT_CELL, s);
            const int cx = cellw & 1023u, cy = (cellw >> 10) & 1023u, cz = (cellw >> 20) & 1023u;
            n = normal_decode((misc >> 11) & 63u);
            const f3 eye{xsub(xmul(P.voxel_inv_size, pos.x), -P.grid_half), xsub(xmul(P.voxel_inv_size, pos.y), -P.grid_half),
                         xsub(xmul(P.voxel_inv_size, pos.z), -P.grid_half)};
            const float uvx = clampf(xsub(xadd(eye.x, xmul(tv, d.x)), (float)cx), 0.0f, 1.0f);
            const float uvy = clampf(xsub(xadd(eye.y, xmul(tv, d.y)), (float)cy), 0.0f, 1.0f);
            const float uvz = clampf(xsub(xadd(eye.z, xmul(tv, d.z)), (float)cz), 0.0f, 1.0f);
            const float bnd = P.voxel_edges, hib = xsub(1.0f, P.voxel_edges);
            const int count = (uvx < bnd || uvx > hib) + (uvy < bnd || uvy > hib) + (uvz < bnd || uvz > hib);
            edge = count >= 2 ? 1u : 0u;
            if ((cellw >> 30) & 1u) {
              const int b = ((cz >> 2) * P.brick_res + (cy >> 2)) * P.brick_res + (cx >> 2);
              colw = __ldg(P.color + (size_t)b * 64 + ((cz & 3) * 16 + (cy & 3) * 4 + (cx & 3)));
              if (STATS) c_hits++;
            }
            const float k = xsub(1.0f, xmul(0.9f, edge ? 1.0f : 0.0f));
            albedo = (cellw >> 30) & 1u ? f3{xmul(unorm8[colw & 255u], k), xmul(unorm8[(colw >> 8) & 255u], k), xmul(unorm8[(colw >> 16) & 255u], k)} : mk3(0.0f);
            mat = (int)(colw >> 24), hit_light = mat == 2;
          }
          const uint32_t depth = MISC_DEPTH(misc);
          if (hit_light) {
            // emissive voxel / floor terminates the path (pathtracer.py:519-525)
            if (depth > 0u) S.set3(F_CR, s, S.get3(F_CR, s) + S.get3(F_TR, s) * albedo);
            if (depth == 0u) S.u(F_PMINFO, s) = encode_material(mat, albedo);
            to_retire = true;
          } else {
            if (STATS) c_vertices++;
            const f3 npos = (pos + closest * d) + n * VRT_EPS;
            S.set3(F_OX, s, npos);
            S.set3(F_VX, s, -d);
            S.u(F_HIT_COL, s) = kind == 1 ? 0u : colw;
            misc = (misc & ~((3u << 8) | (1u << 10) | (63u << 11))) | ((uint32_t)kind << 8) | (edge << 10) | (normal_code(n.x, n.y, n.z) << 11);
            const uint32_t base_dim = 8u * depth;
            const int pix = S.i(F_PIX, s);
            const uint32_t key = path_key((uint32_t)((pix >> 16) * P.W + (pix & 0xffff)), (uint32_t)(P.first_sample + (int)MISC_SI(misc) * P.stride), P.seed);
            const f3 light_dir = sample_cone_oriented(P.light_cos_max, P.light_dir, sun_bx, sun_by, rnd(key, base_dim + 0), rnd(key, base_dim + 1));
            if (dot(light_dir, n) > 0.0f) {
              S.set3(F_DX, s, light_dir);
              misc |= 1u << 4;  // trace the shadow ray next
            } else {
              misc &= ~(1u << 7);  // not visible
              to_shade = true;
            }
            S.u(F_MISC, s) = misc;
          }
        }
      }
      push_list(L.sky, &L.counts[0], to_sky, s);
      push_list(L.shade, &L.counts[1], to_shade, s);
      push_list(L.retire, &L.counts[2], to_retire || to_sky, s);  // escaped paths retire after phase K
    }
    __syncthreads();

    // =========================================================== K: escaped paths (pathtracer.py:499-511)
    const int n_sky = L.counts[0], n_shade = L.counts[1];
    for (int j = tid; j < n_sky; j += POOL_THREADS) {
      const int s = L.sky[j];
      const uint32_t misc = S.u(F_MISC, s);
      const f3 d = S.get3(F_DX, s);
      const float hit_sun = dot(P.light_dir, d) >= P.light_cos_max ? 1.0f : 0.0f;
      f3 sky_scattering = P.background, sky_T = mk3(1.0f);
      if (P.use_sky) {
        const int pix = S.i(F_PIX, s);
        const uint32_t key = path_key((uint32_t)((pix >> 16) * P.W + (pix & 0xffff)), (uint32_t)(P.first_sample + (int)MISC_SI(misc) * P.stride), P.seed);
        const uint32_t base_dim = 8u * MISC_DEPTH(misc);
        const f3 dj = normalize(d + f3{rnd(key, base_dim + 5), rnd(key, base_dim + 6), rnd(key, base_dim + 7)} * 0.0015f);
        const SkyTap t = sky_tap(P.sky_res, project_sky(dj, sky_fres));
        sky_scattering = sky_fetch(P.sky_scatter, t);
        sky_T = sky_fetch(P.sky_trans, t);
        if (STATS) c_escapes++;
      }
      const f3 sky_emission = firefly_filter(sky_scattering + sky_T * sun_rad * hit_sun);
      S.set3(F_CR, s, S.get3(F_CR, s) + S.get3(F_TR, s) * sky_emission);
    }

    // =========================================================== S: shade (NEE contribution + BSDF sample)
    for (int base = 0; base < n_shade; base += POOL_THREADS) {
      const int j = base + tid;
      bool done = false;
      int s = 0;
      if (j < n_shade) {
        s = L.shade[j];
        uint32_t misc = S.u(F_MISC, s);
        const uint32_t depth = MISC_DEPTH(misc);
        const bool visible = (misc >> 7) & 1u;
        const int kind = (misc >> 8) & 3u;
        const f3 n = normal_decode((misc >> 11) & 63u);
        const uint32_t colw = S.u(F_HIT_COL, s);
        Mat m;
        if (kind == 1) {
          m = load_mat(s_mats, P.floor_material);
          m.base_col = P.floor_color;
        } else {
          m = load_mat(s_mats, (int)(colw >> 24));
          const float k = xsub(1.0f, xmul(0.9f, (misc >> 10) & 1u ? 1.0f : 0.0f));
          m.base_col = f3{xmul(unorm8[colw & 255u], k), xmul(unorm8[(colw >> 8) & 255u], k), xmul(unorm8[(colw >> 16) & 255u], k)};
        }
        const f3 view = S.get3(F_VX, s);
        const f3 light_dir = S.get3(F_DX, s);
        f3 thr = S.get3(F_TR, s);
        const int pix = S.i(F_PIX, s);
        const uint32_t key = path_key((uint32_t)((pix >> 16) * P.W + (pix & 0xffff)), (uint32_t)(P.first_sample + (int)MISC_SI(misc) * P.stride), P.seed);
        const uint32_t base_dim = 8u * depth;
        f3 tang, bitang;
        make_orthonormal_basis(n, tang, bitang);
        if (visible) {
          f3 bd, bs;
          float lpdf;
          eval_and_pdf(m, view, n, light_dir, tang, bitang, bd, bs, lpdf);
          const float mis = power_heuristic(light_pdf_axis, lpdf);
          f3 sky_T = mk3(1.0f);
          if (P.use_sky) {
            const SkyTap t = sky_tap(P.sky_res, project_sky(light_dir, sky_fres));
            sky_T = sky_fetch(P.sky_trans, t);
            if (STATS) c_nee++;
          }
          const float ndl = dot(light_dir, n);
          const f3 lr = sky_T * sun_rad * ndl;
          if (depth == 0u) {
            S.set3(F_NDR, s, firefly_filter(thr * (bd * lr)) * mis);
            S.set3(F_NSR, s, firefly_filter(thr * (bs * lr)) * mis);
          } else {
            S.set3(F_CR, s, S.get3(F_CR, s) + firefly_filter(thr * ((mis * (bd + bs)) * lr)));
          }
        }
        f3 brdf;
        float pdf;
        int lobe;
        const f3 nd = sample_disney(m, view, n, tang, bitang, rnd(key, base_dim + 2), rnd(key, base_dim + 3), rnd(key, base_dim + 4), brdf, pdf, lobe);
        f3 bounce_weight = brdf * saturate(dot(nd, n));
        if (depth == 0u) {
          S.f(F_INVPDF, s) = frcp(pdf);
          misc = (misc & ~(3u << 5)) | ((uint32_t)lobe << 5);
        } else {
          bounce_weight = bounce_weight * frcp(pdf);
          const float bsdf_sample_light_pdf = cone_sample_pdf(P.light_cos_max, dot(P.light_dir, nd));
          bounce_weight *= power_heuristic(pdf, (visible ? 1.0f : 0.0f) * bsdf_sample_light_pdf);
        }
        thr *= bounce_weight;
        S.set3(F_TR, s, thr);
        S.set3(F_DX, s, nd);
        misc = (misc & ~(15u | (1u << 4))) | (depth + 1u);  // next segment ray
        S.u(F_MISC, s) = misc;
        const bool dead = thr.x == 0.0f && thr.y == 0.0f && thr.z == 0.0f;
        done = (int)(depth + 1u) >= P.max_depth || dead;
      }
      push_list(L.retire, &L.counts[2], done, s);
    }
    __syncthreads();

    // =========================================================== F: retire finished paths
    const int n_retire = L.counts[2];
    for (int j = tid; j < n_retire; j += POOL_THREADS) {
      const int s = L.retire[j];
      const uint32_t misc = S.u(F_MISC, s);
      const uint32_t pm_info = S.u(F_PMINFO, s);
      f3 emission = mk3(0.0f);
      if ((pm_info & 255u) == 2u) emission = f3{unorm8[(pm_info >> 8) & 255u], unorm8[(pm_info >> 16) & 255u], unorm8[(pm_info >> 24) & 255u]};
      f3 diffuse = S.get3(F_NDR, s), specular = S.get3(F_NSR, s);
      const f3 contrib = S.get3(F_CR, s);
      const float invpdf = S.f(F_INVPDF, s);
      const uint32_t lobe = MISC_LOBE(misc);
      if (lobe == LOBE_DIFFUSE) diffuse += contrib * invpdf + emission;
      if (lobe == LOBE_SPEC_REFL) specular += contrib * invpdf;
      if (bad3(diffuse)) diffuse = mk3(0.0f);
      if (bad3(specular)) specular = mk3(0.0f);
      const f3 acc = S.get3(F_AR, s) + (diffuse + specular);
      const uint32_t s_i = MISC_SI(misc) + 1u;
      const int pix = S.i(F_PIX, s);
      if ((int)s_i < P.n_samples) {
        S.set3(F_AR, s, acc);
        start_path(P, S, s, pix, s_i);
        if (STATS) c_paths++;
      } else {
        float4* dst = P.accum + (size_t)(pix >> 16) * P.W + (pix & 0xffff);
        float4 a = *dst;
        a.x += acc.x, a.y += acc.y, a.z += acc.z, a.w += (float)P.n_samples;
        *dst = a;
        S.i(F_PIX, s) = -1;
      }
    }
    __syncthreads();
  }

  if (STATS) {
    unsigned long long vals[8] = {c_paths, c_rays, c_steps, c_queries, c_hits, c_escapes, c_nee, c_vertices};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      unsigned long long x = vals[i];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
      if (lane == 0 && x) atomicAdd(P.stats + i, x);
    }
  }
}

}  // namespace

cudaError_t vrt_launch_path_pool(const Params& P, bool stats, int sm_count, cudaStream_t st, int* blocks_out) {
  const int fixed_words = 128 * MAT_ROW_F4 * 4 + 256;
  const size_t pool_bytes = (size_t)POOL_WORDS * POOL_SLOTS * 4 + 4 * POOL_SLOTS * sizeof(unsigned short) + 8 * sizeof(int);
  // the upper pyramid joins the pool in shared memory when two CTAs still fit on an SM
  const size_t budget = 110 * 1024;
  int upper_in_smem = ((size_t)fixed_words * 4 + (size_t)P.upper_words * 4 + pool_bytes <= budget) ? 1 : 0;
  const int upper_words_smem = upper_in_smem ? ((P.upper_words + 3) & ~3) : 0;
  const size_t sm = (size_t)fixed_words * 4 + (size_t)upper_words_smem * 4 + pool_bytes;
  cudaError_t e;
  if (stats)
    e = cudaFuncSetAttribute(k_path_pool<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  else
    e = cudaFuncSetAttribute(k_path_pool<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  if (stats)
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path_pool<true>, POOL_THREADS, sm);
  else
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path_pool<false>, POOL_THREADS, sm);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  int blocks = sm_count * per_sm;
  const int max_useful = (P.n_tiles * 32 + POOL_SLOTS - 1) / POOL_SLOTS;
  if (blocks > max_useful) blocks = max_useful > 0 ? max_useful : 1;
  if (blocks_out) *blocks_out = blocks;
  if (stats)
    k_path_pool<true><<<blocks, POOL_THREADS, sm, st>>>(P, upper_in_smem, fixed_words, upper_words_smem);
  else
    k_path_pool<false><<<blocks, POOL_THREADS, sm, st>>>(P, upper_in_smem, fixed_words, upper_words_smem);
  return cudaGetLastError();
}
