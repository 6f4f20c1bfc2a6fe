// ORACLE — TEST INFRASTRUCTURE ONLY. See ovec.h header for the rules.
// CPU restatement of the traversal half of voxel-rt2's hot path:
//   renderer/raytracer.py:7-155   occupancy pyramid + hierarchical 3-D DDA
//   renderer/math_utils.py:103-123 ray_aabb_intersection
//   renderer/voxel_world.py:27-56  inside_grid / voxel_surface_color
//   renderer/pathtracer.py:152-244 floor plane, world<->voxel, next_hit
//   renderer/pathtracer.py:293-312 + space_transformations.py:14-30  camera rays
// Storage layout is the oracle's own (one bit array per LOD); SURVEY.md Appendix A1 explains
// why the reference's linearised offsets are not reproduced (they overrun the field) — the
// traversal result depends only on the bits: bit(l, cell) = OR of its 2^3 children.
#pragma once
#include <vector>

#include "ovec.h"

namespace orc {

struct Counters {  // per-path algorithmic counters
  uint64_t rays = 0, steps = 0, Q = 0, H = 0, E = 0, N = 0, paths = 0, vertices = 0;
  void add(const Counters& o) {
    rays += o.rays, steps += o.steps, Q += o.Q, H += o.H, E += o.E, N += o.N, paths += o.paths,
        vertices += o.vertices;
  }
};

struct Scene {
  // voxel store (voxel_world.py:6-25). Arrays are indexed [x][y][z] (C order, x slowest) with
  // x = i + R/2 — the NumPy layout of the host-side Scene.
  int R = 0, n_lods = 0;
  float voxel_size = 0, voxel_inv_size = 0, voxel_edges = 0;
  std::vector<int8_t> material;
  std::vector<uint8_t> color;  // 3 per voxel
  std::vector<std::vector<uint32_t>> occ;  // occ[lod] bit array, index z*r*r + y*r + x (raytracer.py:35-37)

  // floor / light / background uniforms (pathtracer.py:49-69,91-93,139-144)
  float floor_height = 0;
  V3 floor_color{1, 1, 1};
  int floor_material = 1;
  V3 light_dir{0.57735026f, 0.57735026f, 0.57735026f};
  float light_cos_max = 1.0f;
  V3 light_color{0, 0, 0};
  float light_weight = 3.0f;  // pathtracer.py:144
  V3 background{0, 0, 0};
  int use_physical_sky = 0;

  // camera (inverse matrices prepared on the host in float64)
  int W = 0, H = 0;
  V3 cam_pos{0, 0, 0};
  float inv_proj[16], inv_view[16], view[16], proj[16];
  float jitter[2] = {0, 0};

  // sky tables (atmos.py:66-69): [x][y] y fastest, 3 floats per texel
  int sky_res = 0;
  const float* sky_scatter = nullptr;
  const float* sky_trans = nullptr;

  inline bool occ_bit(int x, int y, int z, int lod) const {
    int r = R >> lod;
    uint32_t idx = (uint32_t)((z * r + y) * r + x);
    return (occ[lod][idx >> 5] >> (idx & 31)) & 1u;
  }
};

// raytracer.py:46-70 (_update_lods): LOD0 bit = material > 0; LOD l = OR of children.
static inline void build_occupancy(Scene& s) {
  int R = s.R;
  s.n_lods = 0;
  while ((1 << s.n_lods) < R) s.n_lods++;  // int(log2(R)), raytracer.py:9
  s.occ.assign(s.n_lods, {});
  for (int l = 0; l < s.n_lods; l++) {
    size_t r = (size_t)(R >> l);
    s.occ[l].assign((r * r * r + 31) / 32, 0u);
  }
  for (int x = 0; x < R; x++)
    for (int y = 0; y < R; y++)
      for (int z = 0; z < R; z++)
        if (s.material[((size_t)x * R + y) * R + z] > 0) {
          uint32_t idx = (uint32_t)((z * R + y) * R + x);
          s.occ[0][idx >> 5] |= 1u << (idx & 31);
        }
  for (int l = 1; l < s.n_lods; l++) {
    int r = R >> l;
    for (int z = 0; z < r; z++)
      for (int y = 0; y < r; y++)
        for (int x = 0; x < r; x++) {
          bool any = false;
          for (int c = 0; c < 8 && !any; c++)
            any = s.occ_bit(2 * x + (c & 1), 2 * y + ((c >> 1) & 1), 2 * z + (c >> 2), l - 1);
          if (any) {
            uint32_t idx = (uint32_t)((z * r + y) * r + x);
            s.occ[l][idx >> 5] |= 1u << (idx & 31);
          }
        }
  }
}

// math_utils.py:103-123. NB the d[i]==0 "outside the slab" test writes `intersect` but the final
// line overwrites it, so axis-parallel rays are never rejected by that test (kept as is).
static inline bool ray_aabb(float bmin, float bmax, V3 o, V3 d, float& near_int, float& far_int) {
  near_int = -kInf;
  far_int = kInf;
  for (int i = 0; i < 3; i++) {
    if (d[i] == 0.0f) continue;
    float i1 = (bmin - o[i]) / d[i];
    float i2 = (bmax - o[i]) / d[i];
    float new_far = fmaxf_(i1, i2);
    float new_near = fminf_(i1, i2);
    far_int = fminf_(new_far, far_int);
    near_int = fmaxf_(new_near, near_int);
  }
  return near_int <= far_int;
}

struct RayHit {
  float t;  // voxel units, inf on miss
  I3 cell;  // LOD0 cell, (-1,-1,-1) if the ray never entered the box
  V3 normal;
  int iters;
};

// raytracer.py:72-155, op for op in float32. Pinned undefined behaviour:
//  A3 a stepped-to cell outside [0,R)^3 is a miss;  A4 d[i]==0 => t[i]=+inf;
//  A5 the 512-iteration cap returns the current finite distance as a "hit".
static inline RayHit raytrace(const Scene& s, V3 origin, V3 direction, float ray_min_t, float ray_max_t,
                              Counters* cnt) {
  RayHit h;
  h.t = kInf;
  h.cell = I3{-1, -1, -1};
  h.normal = V3{0, 0, 0};
  h.iters = 0;
  const float Rf = (float)s.R;
  float bbox_near, bbox_far;
  bool hit_box = ray_aabb(0.0f, Rf, origin, direction, bbox_near, bbox_far);
  if (cnt) cnt->rays++;
  if (hit_box && ray_min_t < bbox_far && ray_max_t > bbox_near) {
    float hit_distance = fmaxf_(bbox_near, ray_min_t);
    V3 initial_p = origin + direction * (hit_distance + kEps);
    I3 ipos_lod0{(int)clampf(std::floor(initial_p.x), 0.0f, Rf - 1.0f),
                 (int)clampf(std::floor(initial_p.y), 0.0f, Rf - 1.0f),
                 (int)clampf(std::floor(initial_p.z), 0.0f, Rf - 1.0f)};
    V3 inv_dir{1.0f / std::fabs(direction.x), 1.0f / std::fabs(direction.y), 1.0f / std::fabs(direction.z)};
    int current_lod = 0;
    float far = fminf_(ray_max_t, bbox_far) - kEps;

    V3 initial_dist{std::fabs(initial_p.x - Rf * 0.5f), std::fabs(initial_p.y - Rf * 0.5f),
                    std::fabs(initial_p.z - Rf * 0.5f)};
    float max_dist = fmaxf_(fmaxf_(initial_dist.x, initial_dist.y), initial_dist.z);
    V3 hit_normal{max_dist == initial_dist.x ? 1.0f : 0.0f, max_dist == initial_dist.y ? 1.0f : 0.0f,
                  max_dist == initial_dist.z ? 1.0f : 0.0f};
    int iters = 0;
    while (iters < 512) {
      if (hit_distance > far) {
        hit_distance = kInf;
        break;
      }
      // A3: cell left the grid -> miss (checked before any query)
      if (ipos_lod0.x < 0 || ipos_lod0.y < 0 || ipos_lod0.z < 0 || ipos_lod0.x >= s.R || ipos_lod0.y >= s.R ||
          ipos_lod0.z >= s.R) {
        hit_distance = kInf;
        break;
      }
      I3 ipos{0, 0, 0};
      bool sample = false;
      while (true) {
        ipos = I3{ipos_lod0.x >> current_lod, ipos_lod0.y >> current_lod, ipos_lod0.z >> current_lod};
        sample = s.occ_bit(ipos.x, ipos.y, ipos.z, current_lod);
        if (cnt) cnt->Q++;
        if (sample && current_lod > 0)
          current_lod -= 1;
        else
          break;
      }
      if (sample) break;

      float cell_size = (float)(1 << current_lod);
      V3 cell_base{(float)ipos.x * cell_size, (float)ipos.y * cell_size, (float)ipos.z * cell_size};
      V3 voxel_pos = origin + direction * hit_distance;
      V3 frac_pos = voxel_pos - cell_base;
      V3 dist = frac_pos;
      if (direction.x > 0.0f) dist.x = cell_size - frac_pos.x;
      if (direction.y > 0.0f) dist.y = cell_size - frac_pos.y;
      if (direction.z > 0.0f) dist.z = cell_size - frac_pos.z;
      V3 t = dist * inv_dir;
      if (direction.x == 0.0f) t.x = kInf;  // A4
      if (direction.y == 0.0f) t.y = kInf;
      if (direction.z == 0.0f) t.z = kInf;
      float min_t = fminf_(fminf_(t.x, t.y), t.z);
      V3 adv = frac_pos + min_t * direction;
      V3 edge_frac_pos{clampf(std::floor(adv.x), 0.0f, cell_size - 1.0f),
                       clampf(std::floor(adv.y), 0.0f, cell_size - 1.0f),
                       clampf(std::floor(adv.z), 0.0f, cell_size - 1.0f)};
      hit_distance += min_t;
      hit_normal = V3{(t.x == min_t ? 1.0f : 0.0f) * signf(direction.x), (t.y == min_t ? 1.0f : 0.0f) * signf(direction.y),
                      (t.z == min_t ? 1.0f : 0.0f) * signf(direction.z)};
      V3 np = cell_base + edge_frac_pos + hit_normal;
      ipos_lod0 = I3{(int)np.x, (int)np.y, (int)np.z};
      current_lod = current_lod + 1 < s.n_lods - 1 ? current_lod + 1 : s.n_lods - 1;
      iters += 1;
      if (cnt) cnt->steps++;
    }
    h.t = hit_distance;
    h.cell = ipos_lod0;
    h.normal = hit_normal;
    h.iters = iters;
  }
  if (dot(direction, h.normal) > 0.0f) h.normal = -h.normal;
  return h;
}

struct Hit {
  float closest;  // world units, inf on miss
  V3 normal;
  V3 albedo;
  int hit_light;
  int mat_id;
  int iters;
  I3 cell;   // voxel cell in grid coordinates [0,R) when kind==2 else (-1,-1,-1)
  int kind;  // 0 miss, 1 floor, 2 voxel   (oracle bookkeeping for the hit-buffer dump)
};

// pathtracer.py:165-167
static inline V3 world_to_voxel(const Scene& s, V3 p) {
  float off = (float)(-(s.R / 2));
  return V3{s.voxel_inv_size * p.x - off, s.voxel_inv_size * p.y - off, s.voxel_inv_size * p.z - off};
}

// voxel_world.py:34-56 — colour texel (rgb = u8/255, a = material/255 -> int(a*255)) + edge darkening.
static inline void voxel_surface_color(const Scene& s, I3 cell, V3 uv, V3& color, int& is_light, int& mat) {
  float boundary = s.voxel_edges;
  int count = 0;
  for (int i = 0; i < 3; i++)
    if (uv[i] < boundary || uv[i] > 1.0f - boundary) count++;
  float f = count >= 2 ? 1.0f : 0.0f;
  color = V3{0, 0, 0};
  mat = 0;
  is_light = 0;
  if (cell.x >= 0 && cell.y >= 0 && cell.z >= 0 && cell.x < s.R && cell.y < s.R && cell.z < s.R) {  // inside_grid
    size_t idx = ((size_t)cell.x * s.R + cell.y) * s.R + cell.z;
    color = V3{(float)s.color[idx * 3 + 0] / 255.0f, (float)s.color[idx * 3 + 1] / 255.0f,
               (float)s.color[idx * 3 + 2] / 255.0f};
    int8_t m = s.material[idx];
    // _make_texture stores f32(material)/255 into a UNORM8 channel: negatives clamp to 0.
    float a = m > 0 ? (float)m / 255.0f : 0.0f;
    mat = (int)(a * 255.0f);
    if (mat == 2) is_light = 1;
  }
  color = color * (1.0f - 0.9f * f);
}

// pathtracer.py:218-244 (next_hit) = _trace_sdf (:173-190) then _trace_voxel (:192-216).
// The "selected voxel" highlight (:235-242) is dead: cast_voxel_hit is never set.
static inline Hit next_hit(const Scene& s, V3 pos, V3 d, float max_dist, bool shadow_ray, Counters* cnt) {
  Hit h;
  h.closest = max_dist;
  h.normal = V3{0, 0, 0};
  h.albedo = V3{1, 1, 1};
  h.hit_light = 0;
  h.mat_id = 0;
  h.cell = I3{-1, -1, -1};
  h.kind = 0;
  // floor plane (A8)
  float ray_march_dist = (s.floor_height - pos.y) / d.y;
  if (ray_march_dist > kEps && ray_march_dist < h.closest) {
    V3 hit_pos = pos + d * ray_march_dist;
    V3 sdf_normal{0.0f, 1.0f, 0.0f};
    float dn = dot(hit_pos, sdf_normal);
    if (length(hit_pos - dn) < 10.0f) {
      h.closest = ray_march_dist;
      h.normal = sdf_normal;
      if (dot(h.normal, d) > 0.0f) h.normal = -h.normal;
      h.albedo = s.floor_color;
      h.hit_light = s.floor_material == 2;
      h.mat_id = s.floor_material;
      h.kind = 1;
    }
  }
  // voxel grid
  V3 eye = world_to_voxel(s, pos);
  RayHit r = raytrace(s, eye, d, kEps, kInf, cnt);
  h.iters = r.iters;
  if (r.t * s.voxel_size < h.closest) {
    h.closest = r.t * s.voxel_size;
    h.kind = 2;
    h.cell = r.cell;
    if (!shadow_ray) {
      V3 p = eye + r.t * d;
      V3 uv{clampf(p.x - (float)r.cell.x, 0.0f, 1.0f), clampf(p.y - (float)r.cell.y, 0.0f, 1.0f),
            clampf(p.z - (float)r.cell.z, 0.0f, 1.0f)};
      voxel_surface_color(s, r.cell, uv, h.albedo, h.hit_light, h.mat_id);
      h.normal = r.normal;
      if (cnt) cnt->H++;
    }
  }
  return h;
}

// row-major 4x4 * vec4
static inline void mat4_mul(const float* m, const float v[4], float out[4]) {
  for (int i = 0; i < 4; i++) out[i] = ((m[i * 4 + 0] * v[0] + m[i * 4 + 1] * v[1]) + m[i * 4 + 2] * v[2]) + m[i * 4 + 3] * v[3];
}

// pathtracer.py:293-312 get_cast_dir; space_transformations.py:14-30. Static camera:
// render_scale = 1 and texcoord += 0.5 * taa_jitter; moving camera: texcoord / render_scale, no
// jitter (pathtracer.py:307-309).
static inline V3 get_cast_dir(const Scene& s, float u, float v, float render_scale = 1.0f, bool moving = false) {
  float tx = (u + 0.5f) * (1.0f / (float)s.W);
  float ty = (v + 0.5f) * (1.0f / (float)s.H);
  if (moving) {
    tx = tx / render_scale;
    ty = ty / render_scale;
  } else {
    tx = tx + s.jitter[0] * 0.5f;
    ty = ty + s.jitter[1] * 0.5f;
  }
  float pos[4] = {tx * 2.0f - 1.0f, ty * 2.0f - 1.0f, 1.0f * 2.0f - 1.0f, 1.0f};
  float q[4];
  mat4_mul(s.inv_proj, pos, q);
  V3 dv = normalize(V3{q[0] / q[3], q[1] / q[3], q[2] / q[3]});
  float dv4[4] = {dv.x, dv.y, dv.z, 0.0f};
  float w[4];
  mat4_mul(s.inv_view, dv4, w);
  return V3{w[0], w[1], w[2]};
}

// space_transformations.py:6-34 (row-major matrices, M @ v)
static inline float linearize_depth(float depth, const float* inv_proj) { return 1.0f / ((depth * 2.0f - 1.0f) * inv_proj[3 * 4 + 2] + inv_proj[3 * 4 + 3]); }
static inline float delinearize_depth(float lindepth, const float* proj) {
  return ((-lindepth * proj[2 * 4 + 2] + proj[2 * 4 + 3]) / -lindepth) * -0.5f + 0.5f;
}
static inline V3 screen_to_view(float ux, float uy, float depth, const float* inv_proj) {
  float p[4] = {ux * 2.0f - 1.0f, uy * 2.0f - 1.0f, depth * 2.0f - 1.0f, 1.0f}, q[4];
  mat4_mul(inv_proj, p, q);
  return V3{q[0] / q[3], q[1] / q[3], q[2] / q[3]};
}
static inline V3 view_to_screen(V3 vp, const float* proj) {
  float p[4] = {vp.x, vp.y, vp.z, 1.0f}, q[4];
  mat4_mul(proj, p, q);
  return V3{q[0] / q[3] * 0.5f + 0.5f, q[1] / q[3] * 0.5f + 0.5f, q[2] / q[3] * 0.5f + 0.5f};
}
static inline V3 xform_point(const float* m, V3 p, float w) {  // view_to_world / world_to_view
  float a[4] = {p.x, p.y, p.z, w}, q[4];
  mat4_mul(m, a, q);
  return V3{q[0], q[1], q[2]};
}

}  // namespace orc
