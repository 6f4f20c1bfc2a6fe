// k_wave — warp-private wavefront form of the fused bounce / shade / NEE path kernel (same estimator,
// same sampler dimensions and the same arithmetic as k_path in vrt_render.cu: renderer/pathtracer.py:355-632,
// non-ReSTIR, static camera; traversal renderer/raytracer.py:72-155 op for op).
//
// Why: in k_path one lane owns one path, so every stage of an outer iteration (retire, camera ray, ray
// set-up, DDA loop, classify, sky lookup, shade) runs with whatever lanes happen to need it — 14 of 32
// on the dense 256^3 scene (profiles/r01k_k_path_final.md): the DDA loop at 8-10 lanes (trip-count
// variance), shade at 11, path start at 12. Here a WARP owns NS = 64 path slots whose state lives in
// shared memory (structure of arrays, one 32-bit word array per field), twice as many as it has lanes,
// and the path is cut into stages at every point where lanes used to diverge:
//
//   GEN     retire the finished path of a slot, start the next sample / fetch the next pixel, camera ray
//   SETUP   floor plane, slab test, DDA initialisation, FIRST occupancy query of the ray
//   STEP    one DDA step + the next occupancy query (a ray that needs k steps passes here k times)
//   CLASS   surface colour fetch, classification (escape / emissive / surface), sun sample -> shadow ray
//   SKY     sky-table projection + bilinear footprint (escaped segments and visible sun samples)
//   SHADE   Disney eval + pdf for the sun sample, BSDF sampling, MIS, throughput update
//
// Each stage has a 64-bit ready mask (a warp-uniform register pair, no atomics, no barriers: the pool
// is private to the warp). Per iteration the warp picks the stage with the most ready slots, hands the
// first 32 of them to its lanes (ballot + popc ranking through a 32-byte shared scratch), loads the words
// that stage needs, runs it with (nearly) all lanes active and routes every slot to its next stage with
// warp OR-reductions. The DDA trip-count variance disappears as a source of idle lanes because a ray
// re-enters STEP as a new work item; shade only ever sees slots that need shading.
// Samples of a pixel stay in one slot and are retired in order, so the per-pixel sums are formed in the
// same order as in k_path.
#include "vrt_internal.h"
#include "vrt_kshared.cuh"
#include "vrt_sky.cuh"
#include "vrt_trace.cuh"

#ifndef VRT_WAVE_SLOTS
#define VRT_WAVE_SLOTS 64
#endif
#ifndef VRT_WAVE_WARPS
#define VRT_WAVE_WARPS 8  // warps per CTA: the staged tables (16 KB) are shared by more pools
#endif
#ifndef VRT_WAVE_MIN_BLOCKS
#define VRT_WAVE_MIN_BLOCKS 2
#endif
#ifndef VRT_WAVE_STEPS
#define VRT_WAVE_STEPS 1  // DDA steps per STEP pass
#endif

namespace {

constexpr int NS = VRT_WAVE_SLOTS;
static_assert(NS == 64 || NS == 32, "the ready masks are 64-bit");

// slot words
enum {
  W_PIX = 0,    // u | v << 16, or -1: the slot has no pixel
  W_META,       // s_i:16 | depth:3 | f_lobe:2 | shadow:1 | visible:1 | mat:8 | has_path (bit 31)
  W_KEY,
  W_PMI,        // packed primary material / albedo
  W_POS,        // 3: ray origin (world)
  W_DIR = W_POS + 3,   // 3: ray direction
  W_THR = W_DIR + 3,   // 3
  W_CON = W_THR + 3,   // 3: contrib
  W_ACC = W_CON + 3,   // 3: per-pixel sum over the samples of this launch
  W_FND = W_ACC + 3,   // 3: primary-vertex sun sample, diffuse
  W_FNS = W_FND + 3,   // 3: ... specular
  W_INVPDF = W_FNS + 3,
  W_ALB,               // 3: surface albedo (stash while the shadow ray is in flight)
  W_VIEW = W_ALB + 3,  // 3: -d of the segment that hit the surface; SKY request of an escape: jittered direction
  W_T = W_VIEW + 3,    // DDA: distance so far / result t (voxel units, +inf on miss)
  W_CELL,              // DDA: (px+1) | (py+1) << 10 | (pz+1) << 20
  W_NRM,               // DDA: normal codes (6 bits) | lod << 6 | iters << 10 | floor_hit << 20 | floor_neg << 21 | kind << 22
  W_FAR,
  W_FLOOR,             // floor-plane distance of the current ray (valid when floor_hit)
  W_IV,                // 3: 1 / |d|
  W_SKYT = W_IV + 3,   // 3: sun transmittance of a visible sun sample
  W_SN = W_SKYT + 3,   // surface normal codes of the stash
  NW
};

enum { ST_GEN = 0, ST_SETUP, ST_STEP, ST_CLASS, ST_SKY, ST_SHADE, ST_NONE };

__device__ __forceinline__ uint32_t ncode(float x) { return (x != 0.0f ? 1u : 0u) | (__float_as_uint(x) >> 31 << 1); }
__device__ __forceinline__ float ndecode(uint32_t c) { return __uint_as_float(((c & 1u) ? 0x3f800000u : 0u) | ((c & 2u) << 30)); }
__device__ __forceinline__ uint32_t npack(float x, float y, float z) { return ncode(x) | (ncode(y) << 2) | (ncode(z) << 4); }

struct Dda {
  float t;
  int px, py, pz, lod, iters;
  float nx, ny, nz;
};

// One occupancy query of raytrace (raytracer.py:103-118): 0 = empty cell, the ray must step; 1 = hit (also the
// 512-iteration cap); 2 = miss (past the far plane or outside the grid).
__device__ __forceinline__ int wave_head(const Params& P, const uint32_t* __restrict__ upper, Dda& s, float far) {
  if (s.iters >= 512) return 1;
  if (s.t > far) return 2;
  if ((unsigned)s.px >= (unsigned)P.R || (unsigned)s.py >= (unsigned)P.R || (unsigned)s.pz >= (unsigned)P.R) return 2;
  bool occ = true;
  int lod = s.lod;
  while (lod >= 3) {
    occ = upper_bit(P, upper, s.px >> lod, s.py >> lod, s.pz >> lod, lod);
    if (!occ) break;
    lod--;
  }
  if (occ) {
    const int b = ((s.pz >> 2) * P.brick_res + (s.py >> 2)) * P.brick_res + (s.px >> 2);
    const unsigned long long w = __ldg(P.bricks + b);
    const int sh1 = ((s.px >> 1) & 1) * 2 + ((s.py >> 1) & 1) * 8 + ((s.pz >> 1) & 1) * 32;
    const bool e2 = w == 0ull;
    const bool e1 = (w & (0x0000000000330033ull << sh1)) == 0ull;
    const bool e0 = ((w >> ((s.pz & 3) * 16 + (s.py & 3) * 4 + (s.px & 3))) & 1ull) == 0ull;
    const int first_empty = e2 ? 2 : (e1 ? 1 : 0);
    lod = min(lod, first_empty);
    occ = !e0;
  }
  s.lod = lod;
  return occ ? 1 : 0;
}

// Step to the exit face of the empty LOD-`lod` cell (raytracer.py:124-147), same float ops as vrt_trace.cuh.
__device__ __forceinline__ void wave_step(const Params& P, Dda& s, f3 o, f3 d, float ivx, float ivy, float ivz) {
  const int lod = s.lod;
  const float cell_size = (float)(1 << lod);
  const int cmask = -(1 << lod);
  const float bx = (float)(s.px & cmask), by = (float)(s.py & cmask), bz = (float)(s.pz & cmask);
  const float fx = xsub(xadd(o.x, xmul(d.x, s.t)), bx);
  const float fy = xsub(xadd(o.y, xmul(d.y, s.t)), by);
  const float fz = xsub(xadd(o.z, xmul(d.z, s.t)), bz);
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
  s.t = xadd(s.t, min_t);
  s.nx = __uint_as_float((tx == min_t ? 0x3f800000u : 0u) | (d.x < 0.0f ? 0x80000000u : 0u));
  s.ny = __uint_as_float((ty == min_t ? 0x3f800000u : 0u) | (d.y < 0.0f ? 0x80000000u : 0u));
  s.nz = __uint_as_float((tz == min_t ? 0x3f800000u : 0u) | (d.z < 0.0f ? 0x80000000u : 0u));
  s.px = (int)(bx + ex + s.nx);
  s.py = (int)(by + ey + s.ny);
  s.pz = (int)(bz + ez + s.nz);
  s.lod = min(P.n_lods - 1, lod + 1);
  s.iters++;
}

__device__ __forceinline__ f3 world_to_voxel(const Params& P, f3 pos) {  // pathtracer.py:165-167
  return f3{xsub(xmul(P.voxel_inv_size, pos.x), -P.grid_half), xsub(xmul(P.voxel_inv_size, pos.y), -P.grid_half),
            xsub(xmul(P.voxel_inv_size, pos.z), -P.grid_half)};
}

}  // namespace

template <bool SKY16>
__global__ void __launch_bounds__(VRT_WAVE_WARPS * 32, VRT_WAVE_MIN_BLOCKS) k_wave(const __grid_constant__ Params P, int upper_in_smem, int fixed_words) {
  extern __shared__ uint32_t smem[];
  const uint32_t* upper = stage_shared(P, smem, upper_in_smem);
  const float4* s_mats = reinterpret_cast<const float4*>(smem);
  const float* unorm8 = reinterpret_cast<const float*>(smem + SMEM_MAT_WORDS);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t* pool = smem + fixed_words + wid * (NS * NW + 8);
  volatile uint8_t* sel = reinterpret_cast<volatile uint8_t*>(pool + NS * NW);

#define LDU(w) pool[(w) * NS + slot]
#define LDF(w) __uint_as_float(pool[(w) * NS + slot])
#define STU(w, v) pool[(w) * NS + slot] = (v)
#define STF(w, v) pool[(w) * NS + slot] = __float_as_uint(v)
#define LD3(w) f3{LDF(w), LDF((w) + 1), LDF((w) + 2)}
#define ST3(w, v)        \
  do {                   \
    const f3 v_ = (v);   \
    STF((w), v_.x);      \
    STF((w) + 1, v_.y);  \
    STF((w) + 2, v_.z);  \
  } while (0)

  // per-launch constants
  f3 sun_bx, sun_by;
  make_orthonormal_basis(P.light_dir, sun_bx, sun_by);
  const float light_pdf_axis = cone_sample_pdf(P.light_cos_max, 1.0f);
  const f3 sun_rad = P.light_weight * P.light_color;
  const float sky_fres = 1.0f / (float)P.sky_res;

  for (int s = lane; s < NS; s += 32) {
    pool[W_PIX * NS + s] = 0xffffffffu;
    pool[W_META * NS + s] = 0u;
  }
  __syncwarp();

  // ready masks (warp-uniform)
  unsigned long long m_gen = NS == 64 ? ~0ull : 0xffffffffull, m_setup = 0ull, m_step = 0ull, m_class = 0ull, m_sky = 0ull, m_shade = 0ull;
  int chunk_base = 0, chunk_rem = 0;
  bool queue_empty = false;

  for (;;) {
    // ---- pick the stage with the most ready slots (ties: the later stage, which frees slots sooner)
    const int c_gen = __popcll(m_gen);
    int stage = ST_NONE, best = 0;
    {
      const int c[6] = {c_gen, __popcll(m_setup), __popcll(m_step), __popcll(m_class), __popcll(m_sky), __popcll(m_shade)};
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int v = min(c[k], 32);
        if (v >= best && v > 0) best = v, stage = k;
      }
    }
    if (stage == ST_NONE) break;
    const unsigned long long M = stage == ST_GEN ? m_gen : stage == ST_SETUP ? m_setup : stage == ST_STEP ? m_step : stage == ST_CLASS ? m_class
                                 : stage == ST_SKY ? m_sky : m_shade;
    // ---- hand the first (up to) 32 ready slots to the lanes
    const unsigned lo = (unsigned)M, hi = (unsigned)(M >> 32);
    const int n_lo = __popc(lo);
    const bool my_lo = (lo >> lane) & 1u, my_hi = (hi >> lane) & 1u;
    const int r_lo = __popc(lo & lt_mask), r_hi = n_lo + __popc(hi & lt_mask);
    if (my_lo) sel[r_lo] = (uint8_t)lane;
    const bool take_hi = my_hi && r_hi < 32;
    if (take_hi) sel[r_hi] = (uint8_t)(32 + lane);
    const unsigned hi_taken = __ballot_sync(FULL, take_hi);
    __syncwarp();
    const int n = min(32, n_lo + __popc(hi));
    const bool act = lane < n;
    const int slot = act ? (int)sel[lane] : 0;
    const unsigned long long taken = (unsigned long long)lo | ((unsigned long long)hi_taken << 32);
    __syncwarp();
    int next = ST_NONE;

    if (stage == ST_GEN) {
      m_gen &= ~taken;
      // ---- retire the finished path of the slot (pathtracer.py:609-619, NaN scrub :1068-1075), then either the
      // next sample of its pixel or the single read-modify-write of the accumulation texel
      int pix = -1;
      uint32_t meta = 0u;
      f3 acc = mk3(0.0f);
      bool restart = false;
      if (act) {
        pix = (int)LDU(W_PIX);
        meta = LDU(W_META);
        if (pix >= 0) {
          acc = LD3(W_ACC);
          const uint32_t pm_info = LDU(W_PMI);
          const int f_lobe = (meta >> 19) & 3;
          const float f_invpdf = LDF(W_INVPDF);
          const f3 contrib = LD3(W_CON);
          f3 emission = mk3(0.0f);
          if ((pm_info & 255u) == 2u)
            emission = f3{unorm8[(pm_info >> 8) & 255u], unorm8[(pm_info >> 16) & 255u], unorm8[(pm_info >> 24) & 255u]};
          f3 diffuse = LD3(W_FND), specular = LD3(W_FNS);
          if (f_lobe == LOBE_DIFFUSE) diffuse += contrib * f_invpdf + emission;
          if (f_lobe == LOBE_SPEC_REFL) specular += contrib * f_invpdf;
          if (bad3(diffuse)) diffuse = mk3(0.0f);
          if (bad3(specular)) specular = mk3(0.0f);
          acc += diffuse + specular;
          const int s_i = (int)(meta & 0xffffu) + 1;
          if (s_i < P.n_samples) {
            meta = (uint32_t)s_i;
            restart = true;
          } else {
            const size_t pidx = (size_t)(pix >> 16) * P.W + (pix & 0xffff);
            float4* dst = P.accum + pidx;
            float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (!P.accum_overwrite) a = *dst;
            a.x += acc.x, a.y += acc.y, a.z += acc.z, a.w += (float)P.n_samples;
            *dst = a;
            pix = -1;
          }
        }
      }
      // ---- slots without a pixel take the next one from the warp's tile queue (one tile per global atomic)
      for (;;) {
        const unsigned need = __ballot_sync(FULL, act && pix == -1);
        if (need == 0u) break;
        if (chunk_rem == 0) {
          unsigned c = 0xffffffffu;
          if (!queue_empty) {
            if (lane == 0) c = atomicAdd(P.work_counter, 1u);
            c = __shfl_sync(FULL, c, 0);
          }
          if (c >= (unsigned)P.n_tiles) {
            queue_empty = true;
            if (act && pix == -1) pix = -2;  // the slot retires
            break;
          }
          chunk_base = (int)c * 32;
          chunk_rem = 32;
        }
        const int rank = __popc(need & lt_mask);
        if (act && pix == -1 && rank < chunk_rem) {
          const int item = chunk_base + (32 - chunk_rem) + rank;
          const int tile = P.tile_rank + P.tile_n * (item >> 5);
          const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
          pix = (tx * 8 + (item & 7)) | ((ty * 4 + ((item >> 3) & 3)) << 16);
          meta = 0u;
          acc = mk3(0.0f);
          restart = true;
        }
        chunk_rem -= min(__popc(need), chunk_rem);
      }
      if (act) {
        if (restart) {
          const int u = pix & 0xffff, v = pix >> 16;
          const int s_i = (int)(meta & 0xffffu);
          const uint32_t sample = (uint32_t)(P.first_sample + s_i * P.stride);
          const float2 j = P.jitter[s_i];
          STU(W_PIX, (uint32_t)pix);
          STU(W_META, (uint32_t)s_i | 0x80000000u);  // depth 0, lobe 0, segment ray
          STU(W_KEY, path_key((uint32_t)(v * P.W + u), sample, P.seed));
          STU(W_PMI, 0u);
          ST3(W_POS, P.cam_pos);
          ST3(W_DIR, get_cast_dir(P, (float)u, (float)v, j.x, j.y));
          ST3(W_THR, mk3(1.0f));
          ST3(W_CON, mk3(0.0f));
          ST3(W_ACC, acc);
          ST3(W_FND, mk3(0.0f));
          ST3(W_FNS, mk3(0.0f));
          STF(W_INVPDF, 1.0f);
          next = ST_SETUP;
        } else {
          STU(W_PIX, 0xffffffffu);  // retired: no stage owns the slot any more
        }
      }
    } else if (stage == ST_SETUP || stage == ST_STEP) {
      if (stage == ST_SETUP) m_setup &= ~taken; else m_step &= ~taken;
      if (act) {
        const f3 pos = LD3(W_POS), d = LD3(W_DIR);
        const f3 eye = world_to_voxel(P, pos);
        Dda s;
        float far, ivx, ivy, ivz;
        uint32_t fl = 0u;  // floor_hit | floor_neg << 1
        int res;
        if (stage == ST_SETUP) {
          // floor: dist = (h - p.y)/d.y ; accept if eps < dist and |(x-y, 0, z-y)| < 10   
          {
            const float dist = xdiv(xsub(P.floor_height, pos.y), d.y);
            if (dist > VRT_EPS) {
              const float hx = xadd(pos.x, xmul(d.x, dist)), hy = xadd(pos.y, xmul(d.y, dist)), hz = xadd(pos.z, xmul(d.z, dist));
              const float dn = xadd(xadd(xmul(hx, 0.0f), xmul(hy, 1.0f)), xmul(hz, 0.0f));
              const float ax = xsub(hx, dn), ay = xsub(hy, dn), az = xsub(hz, dn);
              const float len = xsqrt(xadd(xadd(xmul(ax, ax), xmul(ay, ay)), xmul(az, az)));
              if (len < 10.0f) {
                fl = 1u | ((xadd(xadd(xmul(0.0f, d.x), xmul(1.0f, d.y)), xmul(0.0f, d.z)) > 0.0f) ? 2u : 0u);
                STF(W_FLOOR, dist);
              }
            }
          }
          // ray_aabb_intersection against [0,R]^3 (math_utils.py:103-123), then the DDA start (raytracer.py:84-101)
          const float Rf = (float)P.R;
          float near_int = -VRT_INF, far_int = VRT_INF;
          slab_axis(eye.x, d.x, Rf, near_int, far_int);
          slab_axis(eye.y, d.y, Rf, near_int, far_int);
          slab_axis(eye.z, d.z, Rf, near_int, far_int);
          s.t = VRT_INF, s.px = s.py = s.pz = -1, s.lod = 0, s.iters = 0, s.nx = s.ny = s.nz = 0.0f;
          far = 0.0f, ivx = ivy = ivz = 0.0f;
          if (!(near_int <= far_int && VRT_EPS < far_int && near_int < VRT_INF)) {
            res = 2;  // the ray misses the grid
          } else {
            s.t = fmaxf(near_int, VRT_EPS);
            const float t0 = xadd(s.t, VRT_EPS);
            const float ipx = xadd(eye.x, xmul(d.x, t0)), ipy = xadd(eye.y, xmul(d.y, t0)), ipz = xadd(eye.z, xmul(d.z, t0));
            s.px = (int)clampf(floorf(ipx), 0.0f, Rf - 1.0f);
            s.py = (int)clampf(floorf(ipy), 0.0f, Rf - 1.0f);
            s.pz = (int)clampf(floorf(ipz), 0.0f, Rf - 1.0f);
            ivx = __frcp_rn(fabsf(d.x)), ivy = __frcp_rn(fabsf(d.y)), ivz = __frcp_rn(fabsf(d.z));
            far = xsub(fminf(VRT_INF, far_int), VRT_EPS);
            const float ax = fabsf(xsub(ipx, xmul(Rf, 0.5f))), ay = fabsf(xsub(ipy, xmul(Rf, 0.5f))), az = fabsf(xsub(ipz, xmul(Rf, 0.5f)));
            const float m = fmaxf(fmaxf(ax, ay), az);
            s.nx = (m == ax) ? 1.0f : 0.0f, s.ny = (m == ay) ? 1.0f : 0.0f, s.nz = (m == az) ? 1.0f : 0.0f;
            res = wave_head(P, upper, s, far);
            if (res == 0) {
              STF(W_FAR, far);
              STF(W_IV, ivx), STF(W_IV + 1, ivy), STF(W_IV + 2, ivz);
            }
          }
        } else {
          const uint32_t cw = LDU(W_CELL), nw = LDU(W_NRM);
          s.t = LDF(W_T);
          s.px = (int)(cw & 1023u) - 1, s.py = (int)((cw >> 10) & 1023u) - 1, s.pz = (int)((cw >> 20) & 1023u) - 1;
          s.lod = (int)((nw >> 6) & 15u), s.iters = (int)((nw >> 10) & 1023u);
          s.nx = ndecode(nw & 3u), s.ny = ndecode((nw >> 2) & 3u), s.nz = ndecode((nw >> 4) & 3u);
          fl = (nw >> 20) & 3u;
          far = LDF(W_FAR);
          ivx = LDF(W_IV), ivy = LDF(W_IV + 1), ivz = LDF(W_IV + 2);
          res = 0;
#pragma unroll 1
          for (int k = 0; k < VRT_WAVE_STEPS && res == 0; k++) {
            wave_step(P, s, eye, d, ivx, ivy, ivz);
            res = wave_head(P, upper, s, far);
          }
        }
        if (res == 0) {
          STF(W_T, s.t);
          STU(W_CELL, (uint32_t)(s.px + 1) | ((uint32_t)(s.py + 1) << 10) | ((uint32_t)(s.pz + 1) << 20));
          STU(W_NRM, npack(s.nx, s.ny, s.nz) | ((uint32_t)s.lod << 6) | ((uint32_t)s.iters << 10) | (fl << 20));
          next = ST_STEP;
        } else {
          // the ray is decided: raytrace tail (normal against the ray, raytracer.py:152-153) + next_hit (pathtracer.py:218-244)
          const float rt = res == 1 ? s.t : VRT_INF;
          if (xadd(xadd(xmul(d.x, s.nx), xmul(d.y, s.ny)), xmul(d.z, s.nz)) > 0.0f) s.nx = -s.nx, s.ny = -s.ny, s.nz = -s.nz;
          const float floor_t = (fl & 1u) ? LDF(W_FLOOR) : VRT_INF;
          const float tw = xmul(rt, P.voxel_size);
          uint32_t kind = (fl & 1u) ? 1u : 0u;
          if (tw < floor_t) kind = 2u;
          const uint32_t meta = LDU(W_META);
          if (meta & (1u << 21)) {
            // sun shadow ray: visible iff nothing was hit; its transmittance comes from the sky tables
            const bool visible = kind == 0u;
            STU(W_META, (meta & ~(3u << 21)) | (visible ? (1u << 22) : 0u));
            next = (visible && P.use_sky) ? ST_SKY : ST_SHADE;
          } else {
            STF(W_T, rt);
            STU(W_CELL, (uint32_t)(s.px + 1) | ((uint32_t)(s.py + 1) << 10) | ((uint32_t)(s.pz + 1) << 20));
            STU(W_NRM, npack(s.nx, s.ny, s.nz) | (fl << 20) | (kind << 22));
            next = ST_CLASS;
          }
        }
      }
    } else if (stage == ST_CLASS) {
      m_class &= ~taken;
      if (act) {
        const f3 pos = LD3(W_POS), d = LD3(W_DIR);
        const uint32_t nw = LDU(W_NRM);
        const uint32_t kind = (nw >> 22) & 3u;
        uint32_t meta = LDU(W_META);
        const int depth = (meta >> 16) & 7;
        const uint32_t key = LDU(W_KEY);
        const uint32_t base = 8u * (uint32_t)depth;
        // ---- surface attributes (voxel_world.py:34-56 / floor)
        float closest = VRT_INF;
        f3 n = mk3(0.0f), albedo = mk3(1.0f);
        int mat = 0, hit_light = 0;
        if (kind == 1u) {
          closest = LDF(W_FLOOR);
          n = f3{0.0f, (nw & (1u << 21)) ? -1.0f : 1.0f, 0.0f};
          albedo = P.floor_color, hit_light = P.floor_material == 2, mat = P.floor_material;
        } else if (kind == 2u) {
          const float rt = LDF(W_T);
          const uint32_t cw = LDU(W_CELL);
          const int cx = (int)(cw & 1023u) - 1, cy = (int)((cw >> 10) & 1023u) - 1, cz = (int)((cw >> 20) & 1023u) - 1;
          closest = xmul(rt, P.voxel_size);
          const f3 eye = world_to_voxel(P, pos);
          const float uvx = clampf(xsub(xadd(eye.x, xmul(rt, d.x)), (float)cx), 0.0f, 1.0f);
          const float uvy = clampf(xsub(xadd(eye.y, xmul(rt, d.y)), (float)cy), 0.0f, 1.0f);
          const float uvz = clampf(xsub(xadd(eye.z, xmul(rt, d.z)), (float)cz), 0.0f, 1.0f);
          const float bnd = P.voxel_edges, hib = xsub(1.0f, P.voxel_edges);
          const int count = (uvx < bnd || uvx > hib) + (uvy < bnd || uvy > hib) + (uvz < bnd || uvz > hib);
          const float f = count >= 2 ? 1.0f : 0.0f;
          f3 col = mk3(0.0f);
          if ((unsigned)cx < (unsigned)P.R && (unsigned)cy < (unsigned)P.R && (unsigned)cz < (unsigned)P.R) {
            const int b = ((cz >> 2) * P.brick_res + (cy >> 2)) * P.brick_res + (cx >> 2);
            uint32_t c;
            asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(c) : "l"(P.color + (size_t)b * 64 + ((cz & 3) * 16 + (cy & 3) * 4 + (cx & 3))));
            col = f3{unorm8[c & 255u], unorm8[(c >> 8) & 255u], unorm8[(c >> 16) & 255u]};
            mat = (int)(c >> 24);
          }
          const float k = xsub(1.0f, xmul(0.9f, f));
          albedo = f3{xmul(col.x, k), xmul(col.y, k), xmul(col.z, k)};
          hit_light = mat == 2;
          n = f3{ndecode(nw & 3u), ndecode((nw >> 2) & 3u), ndecode((nw >> 4) & 3u)};
        }
        // ---- classify (pathtracer.py:499-525)
        if (closest == VRT_INF) {
          if (P.use_sky) {
            ST3(W_VIEW, normalize(d + f3{rnd(key, base + 5), rnd(key, base + 6), rnd(key, base + 7)} * 0.0015f));
            next = ST_SKY;
          } else {
            const float hit_sun = dot(P.light_dir, d) >= P.light_cos_max ? 1.0f : 0.0f;
            const f3 sky_emission = firefly_filter(P.background + mk3(1.0f) * sun_rad * hit_sun);
            f3 contrib = LD3(W_CON);
            contrib += LD3(W_THR) * sky_emission;
            ST3(W_CON, contrib);
            next = ST_GEN;
          }
        } else if (hit_light) {
          if (depth > 0) {
            f3 contrib = LD3(W_CON);
            contrib += LD3(W_THR) * albedo;
            ST3(W_CON, contrib);
          }
          if (depth == 0) STU(W_PMI, encode_material(mat, albedo));
          next = ST_GEN;
        } else {
          ST3(W_ALB, albedo);
          ST3(W_VIEW, -d);
          STU(W_SN, npack(n.x, n.y, n.z));
          ST3(W_POS, (pos + closest * d) + n * VRT_EPS);
          meta = (meta & ~(0xffu << 23)) | (((uint32_t)mat & 255u) << 23);
          const f3 light_dir = sample_cone_oriented(P.light_cos_max, P.light_dir, sun_bx, sun_by, rnd(key, base + 0), rnd(key, base + 1));
          if (dot(light_dir, n) > 0.0f) {
            ST3(W_DIR, light_dir);
            STU(W_META, (meta & ~(1u << 22)) | (1u << 21));  // shadow ray in flight
            next = ST_SETUP;
          } else {
            STU(W_META, meta & ~(3u << 21));  // not visible, no ray
            next = ST_SHADE;
          }
        }
      }
    } else if (stage == ST_SKY) {
      m_sky &= ~taken;
      if (act) {
        // one projection + bilinear footprint for both users of the tables: escaped segments (scattering +
        // transmittance, atmos.py:94-115) and visible sun samples (transmittance only, atmos.py:117-131)
        const uint32_t meta = LDU(W_META);
        const bool shadow = (meta >> 22) & 1u;  // a visible sun sample asked; otherwise an escaped segment
        const f3 d = LD3(W_DIR);
        const f3 sky_dir = shadow ? d : LD3(W_VIEW);
        const SkyTap t = sky_tap(P.sky_res, project_sky(sky_dir, sky_fres));
        f3 sky_T, sc = mk3(0.0f);
        if (SKY16) {
          sky_fetch_packed(P.sky_packed, t, sc, sky_T);
        } else {
          sky_T = sky_fetch(P.sky_trans, t);
          if (!shadow) sc = sky_fetch(P.sky_scatter, t);
        }
        if (shadow) {
          ST3(W_SKYT, sky_T);
          next = ST_SHADE;
        } else {
          const float hit_sun = dot(P.light_dir, d) >= P.light_cos_max ? 1.0f : 0.0f;
          const f3 sky_emission = firefly_filter(sc + sky_T * sun_rad * hit_sun);
          f3 contrib = LD3(W_CON);
          contrib += LD3(W_THR) * sky_emission;
          ST3(W_CON, contrib);
          next = ST_GEN;
        }
      }
    } else {  // ST_SHADE
      m_shade &= ~taken;
      if (act) {
        // NEE contribution (if the sun is visible) + BSDF sample (pathtracer.py:427-497, :556-579)
        uint32_t meta = LDU(W_META);
        int depth = (meta >> 16) & 7;
        const bool vis = (meta >> 22) & 1u;
        const float visible = vis ? 1.0f : 0.0f;
        const uint32_t key = LDU(W_KEY);
        const uint32_t base = 8u * (uint32_t)depth;
        const uint32_t sn = LDU(W_SN);
        const f3 s_n{ndecode(sn & 3u), ndecode((sn >> 2) & 3u), ndecode((sn >> 4) & 3u)};
        const f3 s_view = LD3(W_VIEW);
        f3 thr = LD3(W_THR);
        Mat m = load_mat(s_mats, (int)((meta >> 23) & 255u));
        m.base_col = LD3(W_ALB);
        f3 tang, bitang;
        make_orthonormal_basis(s_n, tang, bitang);
        if (vis) {
          const f3 light_dir = LD3(W_DIR);
          f3 sky_T = mk3(1.0f);
          if (P.use_sky) sky_T = LD3(W_SKYT);
          f3 bd, bs;
          float lpdf;
          eval_and_pdf(m, s_view, s_n, light_dir, tang, bitang, bd, bs, lpdf);
          const float mis = power_heuristic(light_pdf_axis, lpdf);
          const float ndl = dot(light_dir, s_n);
          const f3 lr = sky_T * sun_rad * ndl;
          if (depth == 0) {
            ST3(W_FND, firefly_filter(thr * (bd * lr)) * mis);
            ST3(W_FNS, firefly_filter(thr * (bs * lr)) * mis);
          } else {
            f3 contrib = LD3(W_CON);
            contrib += firefly_filter(thr * ((mis * (bd + bs)) * lr));
            ST3(W_CON, contrib);
          }
        }
        f3 brdf;
        float pdf;
        int lobe;
        const f3 nd = sample_disney(m, s_view, s_n, tang, bitang, rnd(key, base + 2), rnd(key, base + 3), rnd(key, base + 4), brdf, pdf, lobe);
        f3 bounce_weight = brdf * saturate(dot(nd, s_n));
        if (depth == 0) {
          STF(W_INVPDF, frcp(pdf));
          meta = (meta & ~(3u << 19)) | ((uint32_t)lobe << 19);
        } else {
          bounce_weight = bounce_weight * frcp(pdf);
          const float bsdf_sample_light_pdf = cone_sample_pdf(P.light_cos_max, dot(P.light_dir, nd));
          bounce_weight *= power_heuristic(pdf, visible * bsdf_sample_light_pdf);
        }
        thr *= bounce_weight;
        depth++;
        ST3(W_THR, thr);
        ST3(W_DIR, nd);
        STU(W_META, (meta & ~((7u << 16) | (3u << 21))) | ((uint32_t)depth << 16));  // segment ray, visibility cleared
        // The reference keeps tracing zero-throughput paths; they add exact zeros, so stop here.
        const bool dead = thr.x == 0.0f && thr.y == 0.0f && thr.z == 0.0f;
        next = (depth >= P.max_depth || dead) ? ST_GEN : ST_SETUP;
      }
    }
    // ---- route every processed slot to its next stage
    {
      const unsigned long long bit = 1ull << slot;
      const unsigned blo = (unsigned)bit, bhi = (unsigned)(bit >> 32);
#define ROUTE(ST, MASK)                                                            \
  do {                                                                             \
    const bool q = act && next == (ST);                                            \
    const unsigned a_ = __reduce_or_sync(FULL, q ? blo : 0u);                      \
    const unsigned b_ = NS == 64 ? __reduce_or_sync(FULL, q ? bhi : 0u) : 0u;      \
    MASK |= (unsigned long long)a_ | ((unsigned long long)b_ << 32);               \
  } while (0)
      if (stage == ST_GEN) {
        ROUTE(ST_SETUP, m_setup);
      } else if (stage == ST_SETUP || stage == ST_STEP) {
        ROUTE(ST_STEP, m_step);
        ROUTE(ST_CLASS, m_class);
        ROUTE(ST_SKY, m_sky);
        ROUTE(ST_SHADE, m_shade);
      } else if (stage == ST_CLASS) {
        ROUTE(ST_SETUP, m_setup);
        ROUTE(ST_SHADE, m_shade);
        ROUTE(ST_SKY, m_sky);
        ROUTE(ST_GEN, m_gen);
      } else if (stage == ST_SKY) {
        ROUTE(ST_SHADE, m_shade);
        ROUTE(ST_GEN, m_gen);
      } else {
        ROUTE(ST_SETUP, m_setup);
        ROUTE(ST_GEN, m_gen);
      }
#undef ROUTE
    }
    __syncwarp();
  }
#undef LDU
#undef LDF
#undef STU
#undef STF
#undef LD3
#undef ST3
}

static size_t wave_smem_bytes(const Params& P, int* upper_in_smem, int* fixed_words) {
  int fw = SMEM_FIXED_WORDS;
  const size_t pools = (size_t)VRT_WAVE_WARPS * (NS * NW + 8) * 4;
  *upper_in_smem = ((size_t)(fw + P.upper_words) * 4 + pools <= 110 * 1024) ? 1 : 0;
  if (*upper_in_smem) fw += P.upper_words;
  fw = (fw + 3) & ~3;
  *fixed_words = fw;
  return (size_t)fw * 4 + pools;
}

template <bool SKY16>
static cudaError_t launch_wave_t(const Params& P, int sm_count, cudaStream_t st) {
  int uis, fw;
  const size_t sm = wave_smem_bytes(P, &uis, &fw);
  static bool attr_set[2] = {false, false};
  if (!attr_set[SKY16 ? 1 : 0]) {
    cudaError_t e = cudaFuncSetAttribute(k_wave<SKY16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[SKY16 ? 1 : 0] = true;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wave<SKY16>, VRT_WAVE_WARPS * 32, sm);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  int blocks = sm_count * per_sm;  // persistent: one wave, a multiple of the SM count
  const int max_useful = (P.n_tiles + 2 * VRT_WAVE_WARPS - 1) / (2 * VRT_WAVE_WARPS);  // a pool holds two tiles
  if (blocks > max_useful) blocks = max_useful > 0 ? max_useful : 1;
  k_wave<SKY16><<<blocks, VRT_WAVE_WARPS * 32, sm, st>>>(P, uis, fw);
  return cudaGetLastError();
}

cudaError_t vrt_launch_wave(const Params& P, int sm_count, cudaStream_t st) {
  if (P.sky_packed && P.use_sky) return launch_wave_t<true>(P, sm_count, st);
  return launch_wave_t<false>(P, sm_count, st);
}
