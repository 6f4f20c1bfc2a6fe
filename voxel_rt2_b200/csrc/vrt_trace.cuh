// Hierarchical 3-D DDA over 4^3 occupancy bricks + an upper bit pyramid.
//
// Replaces renderer/raytracer.py:17-155 (linearize_index / query_occupancy / raytrace),
// renderer/math_utils.py:103-123 (ray_aabb_intersection), renderer/voxel_world.py:27-56
// (voxel_surface_color) and renderer/pathtracer.py:152-244 (floor plane, next_hit).
//
// Layout (ours, not the reference's): LOD 0, 1 and 2 of a 4x4x4 block live in ONE 64-bit brick
// word (bit = (z&3)*16 + (y&3)*4 + (x&3)); LOD1 is a 2^3 sub-mask test, LOD2 is word != 0, so a
// single 8-byte load answers up to three of the reference's occupancy queries and the last word
// is kept in a register while the ray stays inside the brick. LOD >= 3 are plain bit arrays
// (<= 37 KB at 512^3) that the render kernels stage in shared memory.
//
// The stepping sequence (descend while occupied, step to the exit face of the empty LOD-l cell,
// ascend one LOD) and every float op are the reference's, with no FMA contraction, so hits are
// bit-identical to the CPU oracle.
#pragma once
#include "vrt_common.cuh"

struct RayHit {
  float t;  // voxel units, +inf on miss
  int cx, cy, cz;
  float nx, ny, nz;
  int iters;
};

struct TraceCounters {
  uint32_t rays, steps, queries;
};

HD bool upper_bit(const Params& P, const uint32_t* __restrict__ upper, int x, int y, int z, int lod) {
  int r = P.R >> lod;
  uint32_t idx = (uint32_t)((z * r + y) * r + x);
  return (upper[P.upper_off[lod - 3] + (idx >> 5)] >> (idx & 31)) & 1u;
}

// One axis of ray_aabb_intersection against [0,R] (math_utils.py:103-123; an axis with d == 0 is
// skipped, as in the reference). Out of line: three calls share one copy of the two IEEE divisions
// (measured +0.8 %; sharing the hotter helpers this way — every division, the sky fetches — was
// measured 4 % slower: a call on the hot path costs more fetch redirects than the footprint saves).
struct SlabRange {
  float near_int, far_int;
};
static __device__ __noinline__ SlabRange slab_axis_nl(float o, float d, float Rf, float near_int, float far_int) {
  if (d != 0.0f) {
    float i1 = __fdiv_rn(xsub(0.0f, o), d);
    float i2 = __fdiv_rn(xsub(Rf, o), d);
    far_int = fminf(fmaxf(i1, i2), far_int);
    near_int = fmaxf(fminf(i1, i2), near_int);
  }
  return SlabRange{near_int, far_int};
}
HD void slab_axis(float o, float d, float Rf, float& near_int, float& far_int) {
  const SlabRange r = slab_axis_nl(o, d, Rf, near_int, far_int);
  near_int = r.near_int, far_int = r.far_int;
}

template <bool STATS>
HD RayHit raytrace(const Params& P, const uint32_t* __restrict__ upper, f3 o, f3 d, TraceCounters* tc) {
  RayHit h;
  h.t = VRT_INF;
  h.cx = h.cy = h.cz = -1;
  h.nx = h.ny = h.nz = 0.0f;
  h.iters = 0;
  const float Rf = (float)P.R;
  if (STATS) tc->rays++;

  // ray_aabb_intersection against [0,R]^3 (axes with d == 0 are skipped, as in the reference).
  // (Skipping the three entry-side divisions for origins inside the box is exact — for d > 0 the exit plane is
  // always x = R, the entry plane x = 0 — but was measured 6 % slower in round 1 and 5 % slower in round 2 with
  // the far side computed for every lane and the near side under a branch: profiles/r02d_ab_inside_f16.log.)
  float near_int = -VRT_INF, far_int = VRT_INF;
  slab_axis(o.x, d.x, Rf, near_int, far_int);
  slab_axis(o.y, d.y, Rf, near_int, far_int);
  slab_axis(o.z, d.z, Rf, near_int, far_int);
  if (!(near_int <= far_int && VRT_EPS < far_int && near_int < VRT_INF)) {
    return h;
  }

  float hit_distance = fmaxf(near_int, VRT_EPS);
  const float t0 = xadd(hit_distance, VRT_EPS);
  const float ipx = xadd(o.x, xmul(d.x, t0)), ipy = xadd(o.y, xmul(d.y, t0)), ipz = xadd(o.z, xmul(d.z, t0));
  int px = (int)clampf(floorf(ipx), 0.0f, Rf - 1.0f);
  int py = (int)clampf(floorf(ipy), 0.0f, Rf - 1.0f);
  int pz = (int)clampf(floorf(ipz), 0.0f, Rf - 1.0f);
  const float ivx = __frcp_rn(fabsf(d.x)), ivy = __frcp_rn(fabsf(d.y)), ivz = __frcp_rn(fabsf(d.z));  // == 1.0f / |d|, correctly rounded
  int lod = 0;
  const float far = xsub(fminf(VRT_INF, far_int), VRT_EPS);
  {
    float ax = fabsf(xsub(ipx, xmul(Rf, 0.5f))), ay = fabsf(xsub(ipy, xmul(Rf, 0.5f))), az = fabsf(xsub(ipz, xmul(Rf, 0.5f)));
    float m = fmaxf(fmaxf(ax, ay), az);
    h.nx = (m == ax) ? 1.0f : 0.0f;
    h.ny = (m == ay) ? 1.0f : 0.0f;
    h.nz = (m == az) ? 1.0f : 0.0f;
  }
  const int top_lod = P.n_lods - 1;
  int last_b = -1;
  unsigned long long w = 0ull;
  int iters = 0;
  while (iters < 512) {
    if (hit_distance > far) {
      hit_distance = VRT_INF;
      break;
    }
    if ((unsigned)px >= (unsigned)P.R || (unsigned)py >= (unsigned)P.R || (unsigned)pz >= (unsigned)P.R) {
      hit_distance = VRT_INF;  // stepped outside the grid: miss (pinned)
      break;
    }
    // --- descend while occupied (raytracer.py:110-118)
    bool occ = true;
    while (lod >= 3) {
      occ = upper_bit(P, upper, px >> lod, py >> lod, pz >> lod, lod);
      if (STATS) tc->queries++;
      if (!occ) break;
      lod--;
    }
    if (occ) {
      int b = ((pz >> 2) * P.brick_res + (py >> 2)) * P.brick_res + (px >> 2);
      if (b != last_b) {
        w = __ldg(P.bricks + b);
        last_b = b;
      }
      // LOD 2 / 1 / 0 all come from the one brick word: e2 => e1 => e0 (an empty 4^3 block has empty
      // 2^3 sub-blocks and voxels), so "descend from `lod` while occupied" is the first empty level
      // at or below `lod`: min(lod, coarsest empty level); a hit iff the voxel bit is set.
      const int sh1 = ((px >> 1) & 1) * 2 + ((py >> 1) & 1) * 8 + ((pz >> 1) & 1) * 32;
      const bool e2 = w == 0ull;
      const bool e1 = (w & (0x0000000000330033ull << sh1)) == 0ull;
      const bool e0 = ((w >> ((pz & 3) * 16 + (py & 3) * 4 + (px & 3))) & 1ull) == 0ull;
      const int first_empty = e2 ? 2 : (e1 ? 1 : 0);
      if (STATS) tc->queries += e0 ? lod - min(lod, first_empty) + 1 : lod + 1;
      lod = min(lod, first_empty);
      occ = !e0;
    }
    if (occ) break;
    // --- step to the exit face of the empty LOD-`lod` cell (raytracer.py:124-147)
    const float cell_size = (float)(1 << lod);
    // (px >> lod) * 2^lod == px with the low `lod` bits cleared: same float, no multiply
    const int cmask = -(1 << lod);
    const float bx = (float)(px & cmask), by = (float)(py & cmask), bz = (float)(pz & cmask);
    const float fx = xsub(xadd(o.x, xmul(d.x, hit_distance)), bx);
    const float fy = xsub(xadd(o.y, xmul(d.y, hit_distance)), by);
    const float fz = xsub(xadd(o.z, xmul(d.z, hit_distance)), bz);
    float tx = xmul(d.x > 0.0f ? xsub(cell_size, fx) : fx, ivx);
    float ty = xmul(d.y > 0.0f ? xsub(cell_size, fy) : fy, ivy);
    float tz = xmul(d.z > 0.0f ? xsub(cell_size, fz) : fz, ivz);
    if (d.x == 0.0f) tx = VRT_INF;  // pinned, SURVEY A4 (also covers origins outside the slab of that axis)
    if (d.y == 0.0f) ty = VRT_INF;
    if (d.z == 0.0f) tz = VRT_INF;
    const float min_t = fminf(fminf(tx, ty), tz);
    const float ex = clampf(floorf(xadd(fx, xmul(min_t, d.x))), 0.0f, cell_size - 1.0f);
    const float ey = clampf(floorf(xadd(fy, xmul(min_t, d.y))), 0.0f, cell_size - 1.0f);
    const float ez = clampf(floorf(xadd(fz, xmul(min_t, d.z))), 0.0f, cell_size - 1.0f);
    hit_distance = xadd(hit_distance, min_t);
    // (t == min_t ? 1 : 0) * sign(d): for d != 0 the product is 1 or 0 carrying d's sign bit; an
    // axis with d == +-0 has t = +inf, never the minimum, and the reference product is +0.
    h.nx = __uint_as_float((tx == min_t ? 0x3f800000u : 0u) | (d.x < 0.0f ? 0x80000000u : 0u));
    h.ny = __uint_as_float((ty == min_t ? 0x3f800000u : 0u) | (d.y < 0.0f ? 0x80000000u : 0u));
    h.nz = __uint_as_float((tz == min_t ? 0x3f800000u : 0u) | (d.z < 0.0f ? 0x80000000u : 0u));
    px = (int)(bx + ex + h.nx);
    py = (int)(by + ey + h.ny);
    pz = (int)(bz + ez + h.nz);
    lod = min(top_lod, lod + 1);
    iters++;
    if (STATS) tc->steps++;
  }
  h.t = hit_distance;
  h.cx = px, h.cy = py, h.cz = pz;
  h.iters = iters;
  // flip the normal against the ray (raytracer.py:152-153)
  if (xadd(xadd(xmul(d.x, h.nx), xmul(d.y, h.ny)), xmul(d.z, h.nz)) > 0.0f) {
    h.nx = -h.nx, h.ny = -h.ny, h.nz = -h.nz;
  }
  return h;
}

// Result of next_hit. Surface attributes are only filled for non-shadow rays.
struct Hit {
  float closest;  // world units, +inf on miss
  float nx, ny, nz;
  f3 albedo;
  int mat_id;
  int hit_light;
  int kind;  // 0 miss, 1 floor, 2 voxel
  int cx, cy, cz;
};

// pathtracer.py:218-244: floor plane first (:173-190), then the voxel grid (:192-216).
// `shadow` is a run-time flag so that segment rays and shadow rays of different lanes share one
// instance of the traversal loop (the path kernel traces both kinds in the same iteration).
template <bool STATS>
HD Hit next_hit(const Params& P, const uint32_t* __restrict__ upper, const float* __restrict__ unorm8, f3 pos, f3 d, const bool SHADOW,
                TraceCounters* tc, uint32_t* n_hits) {
  Hit h;
  h.closest = VRT_INF;
  h.nx = h.ny = h.nz = 0.0f;
  h.albedo = mk3(1.0f);
  h.mat_id = 0;
  h.hit_light = 0;
  h.kind = 0;
  h.cx = h.cy = h.cz = -1;
  // floor: dist = (h - p.y)/d.y ; accept if eps < dist and |(x-y, 0, z-y)| < 10   
  {
    float dist = xdiv(xsub(P.floor_height, pos.y), d.y);
    if (dist > VRT_EPS && dist < h.closest) {
      float hx = xadd(pos.x, xmul(d.x, dist)), hy = xadd(pos.y, xmul(d.y, dist)), hz = xadd(pos.z, xmul(d.z, dist));
      float dn = xadd(xadd(xmul(hx, 0.0f), xmul(hy, 1.0f)), xmul(hz, 0.0f));
      float ax = xsub(hx, dn), ay = xsub(hy, dn), az = xsub(hz, dn);
      float len = xsqrt(xadd(xadd(xmul(ax, ax), xmul(ay, ay)), xmul(az, az)));
      if (len < 10.0f) {
        h.closest = dist;
        h.nx = 0.0f, h.ny = 1.0f, h.nz = 0.0f;
        if (xadd(xadd(xmul(0.0f, d.x), xmul(1.0f, d.y)), xmul(0.0f, d.z)) > 0.0f) h.ny = -1.0f;
        if (!SHADOW) {
          h.albedo = P.floor_color;
          h.hit_light = P.floor_material == 2;
          h.mat_id = P.floor_material;
        }
        h.kind = 1;
      }
    }
  }
  // world -> voxel space: inv_size * p - offset, offset = -R/2  (pathtracer.py:165-167)
  f3 eye{xsub(xmul(P.voxel_inv_size, pos.x), -P.grid_half), xsub(xmul(P.voxel_inv_size, pos.y), -P.grid_half),
         xsub(xmul(P.voxel_inv_size, pos.z), -P.grid_half)};
  RayHit r = raytrace<STATS>(P, upper, eye, d, tc);
  float tw = xmul(r.t, P.voxel_size);
  if (tw < h.closest) {
    h.closest = tw;
    h.kind = 2;
    h.cx = r.cx, h.cy = r.cy, h.cz = r.cz;
    if (!SHADOW) {
      // voxel_world.py:34-56
      float uvx = clampf(xsub(xadd(eye.x, xmul(r.t, d.x)), (float)r.cx), 0.0f, 1.0f);
      float uvy = clampf(xsub(xadd(eye.y, xmul(r.t, d.y)), (float)r.cy), 0.0f, 1.0f);
      float uvz = clampf(xsub(xadd(eye.z, xmul(r.t, d.z)), (float)r.cz), 0.0f, 1.0f);
      const float bnd = P.voxel_edges, hib = xsub(1.0f, P.voxel_edges);
      int count = (uvx < bnd || uvx > hib) + (uvy < bnd || uvy > hib) + (uvz < bnd || uvz > hib);
      float f = count >= 2 ? 1.0f : 0.0f;
      f3 col = mk3(0.0f);
      int mat = 0;
      if ((unsigned)r.cx < (unsigned)P.R && (unsigned)r.cy < (unsigned)P.R && (unsigned)r.cz < (unsigned)P.R) {
        int b = ((r.cz >> 2) * P.brick_res + (r.cy >> 2)) * P.brick_res + (r.cx >> 2);
        // one texel per surface hit out of 8-64 MB: bypass L1 allocation so the occupancy bricks stay resident
        uint32_t c;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(c) : "l"(P.color + (size_t)b * 64 + ((r.cz & 3) * 16 + (r.cy & 3) * 4 + (r.cx & 3))));
        col = f3{unorm8[c & 255u], unorm8[(c >> 8) & 255u], unorm8[(c >> 16) & 255u]};  // k / 255.0f, correctly rounded
        mat = (int)(c >> 24);
        if (STATS) (*n_hits)++;
      }
      float k = xsub(1.0f, xmul(0.9f, f));
      h.albedo = f3{xmul(col.x, k), xmul(col.y, k), xmul(col.z, k)};
      h.mat_id = mat;
      h.hit_light = mat == 2;
      h.nx = r.nx, h.ny = r.ny, h.nz = r.nz;
    }
  }
  return h;
}

// pathtracer.py:293-312 + space_transformations.py:14-30, render_scale = 1, static camera.
HD f3 get_cast_dir(const Params& P, float u, float v, float jx, float jy) {
  float tx = xadd(xmul(xadd(u, 0.5f), P.inv_w), xmul(jx, 0.5f));
  float ty = xadd(xmul(xadd(v, 0.5f), P.inv_h), xmul(jy, 0.5f));
  const float px = xsub(xmul(tx, 2.0f), 1.0f), py = xsub(xmul(ty, 2.0f), 1.0f);
  // inv_proj @ (px, py, 1, 1): the two products by 1.0f are exact and dropped (same bits)
  float q[4];
#pragma unroll
  for (int i = 0; i < 4; i++)
    q[i] = xadd(xadd(xadd(xmul(P.inv_proj[i * 4 + 0], px), xmul(P.inv_proj[i * 4 + 1], py)), P.inv_proj[i * 4 + 2]), P.inv_proj[i * 4 + 3]);
  f3 dv = xnormalize(f3{xdiv(q[0], q[3]), xdiv(q[1], q[3]), xdiv(q[2], q[3])});
  // inv_view @ (dv, 0): the product by 0 contributes +-0 and is dropped
  float w[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
    w[i] = xadd(xadd(xmul(P.inv_view[i * 4 + 0], dv.x), xmul(P.inv_view[i * 4 + 1], dv.y)), xmul(P.inv_view[i * 4 + 2], dv.z));
  return f3{w[0], w[1], w[2]};
}

// Moving camera (pathtracer.py:307-309): texcoord / render_scale, no TAA jitter.
HD f3 get_cast_dir_scaled(const Params& P, float u, float v, float scale) {
  const float tx = xdiv(xmul(xadd(u, 0.5f), P.inv_w), scale);
  const float ty = xdiv(xmul(xadd(v, 0.5f), P.inv_h), scale);
  const float px = xsub(xmul(tx, 2.0f), 1.0f), py = xsub(xmul(ty, 2.0f), 1.0f);
  float q[4];
#pragma unroll
  for (int i = 0; i < 4; i++)
    q[i] = xadd(xadd(xadd(xmul(P.inv_proj[i * 4 + 0], px), xmul(P.inv_proj[i * 4 + 1], py)), P.inv_proj[i * 4 + 2]), P.inv_proj[i * 4 + 3]);
  f3 dv = xnormalize(f3{xdiv(q[0], q[3]), xdiv(q[1], q[3]), xdiv(q[2], q[3])});
  float w[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
    w[i] = xadd(xadd(xmul(P.inv_view[i * 4 + 0], dv.x), xmul(P.inv_view[i * 4 + 1], dv.y)), xmul(P.inv_view[i * 4 + 2], dv.z));
  return f3{w[0], w[1], w[2]};
}

// renderer/space_transformations.py:6-34 (row-major matrices, M @ v)
HD void mat4_mul(const float* m, float x, float y, float z, float w, float q[4]) {
#pragma unroll
  for (int i = 0; i < 4; i++) q[i] = ((m[i * 4 + 0] * x + m[i * 4 + 1] * y) + m[i * 4 + 2] * z) + m[i * 4 + 3] * w;
}
HD float linearize_depth(const Params& P, float depth) { return 1.0f / ((depth * 2.0f - 1.0f) * P.inv_proj[3 * 4 + 2] + P.inv_proj[3 * 4 + 3]); }
HD float delinearize_depth(const Params& P, float lindepth) { return ((-lindepth * P.proj[2 * 4 + 2] + P.proj[2 * 4 + 3]) / -lindepth) * -0.5f + 0.5f; }
HD f3 screen_to_view(const Params& P, float ux, float uy, float depth) {
  float q[4];
  mat4_mul(P.inv_proj, ux * 2.0f - 1.0f, uy * 2.0f - 1.0f, depth * 2.0f - 1.0f, 1.0f, q);
  return f3{q[0] / q[3], q[1] / q[3], q[2] / q[3]};
}
HD f3 view_to_world(const Params& P, f3 v) {
  float q[4];
  mat4_mul(P.inv_view, v.x, v.y, v.z, 1.0f, q);
  return f3{q[0], q[1], q[2]};
}
HD float view_to_screen_z(const Params& P, f3 world_pos) {  // view_to_screen(world_to_view(p)).z
  float a[4], q[4];
  mat4_mul(P.view, world_pos.x, world_pos.y, world_pos.z, 1.0f, a);
  mat4_mul(P.proj, a[0], a[1], a[2], 1.0f, q);
  return q[2] / q[3] * 0.5f + 0.5f;
}
