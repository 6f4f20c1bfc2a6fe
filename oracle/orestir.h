// ORACLE — TEST INFRASTRUCTURE ONLY. See ovec.h header for the rules.
// CPU restatement of voxel-rt2's experimental ReSTIR-PT mode (USE_RESTIR_PT, pathtracer.py:15):
//   renderer/reservoir.py:8-141           Sample / Reservoir / StorageReservoir packing
//   renderer/math_utils.py:201-215,250-263 octahedral f16 and arbitrary-bit u32 packing
//   renderer/pathtracer.py:672-812         shift() (reconnection shift + Jacobian)
//   renderer/pathtracer.py:815-989         spatial_GRIS (32-tap golden-angle spiral, pairwise MIS)
//
// PARITY: the ReSTIR branch of render() (reservoir before packing: 7.5e-6), shift() (Jacobian
// identical), the reservoir bookkeeping / packing and spatial_GRIS (both BIT-IDENTICAL) are pinned
// against the reference source run through oracle/ti_emu (tests/golden/ref_{restir_render,shift,
// reservoir,gris}.npz), the last three on inputs with non-zero sample vectors and no sky pixels.
// UNPINNED: only what the upstream mode (compiled out, USE_RESTIR_PT = False) leaves
// undefined — octahedral encodings of zero vectors = 0/0, out-of-image taps, sky pixels processed
// with NaN normals. This file states the INTENDED algorithm with those holes pinned:
//   * the "zero vector" markers of a Sample (escape vertex / last vertex / NEE invisible) travel
//     through the packed reservoir as explicit flag bits (the spare byte of the 56-byte record);
//   * primary-sky pixels pass their own sample through (the is_vec_zero(center_x1) branch);
//   * primary positions come from gbuff_position (the commented alternative at :851,:904);
//   * taps outside the image or on sky pixels are skipped;
//   * the visibility ray is traced only if the resampling selected something.
#pragma once
#include "obsdf.h"
#include "osky.h"
#include "otrace.h"

namespace orc {

struct Sample {  // reservoir.py:23-39
  V3 F{0, 0, 0};
  V3 rc_pos{0, 0, 0};
  V3 rc_normal{0, 0, 0};        // zero => rc vertex is an escape vertex
  V3 rc_incident_dir{0, 0, 0};  // zero => path terminated at the rc vertex
  V3 rc_incident_L{0, 0, 0};
  V3 rc_NEE_dir{0, 0, 0};  // zero => NEE invisible
  uint32_t rc_mat_info = 0;
  float cached_jacobian_term = 1.0f;
  int lobes = 0;
};
struct Reservoir {  // reservoir.py:41-100
  Sample z;
  float M = 0.0f, weight = 0.0f;
  void update_cached_jacobian_term(V3 x1) {  // :59-62
    V3 dir = z.rc_pos - x1;
    z.cached_jacobian_term = dot(dir, dir) / std::fabs(dot(normalize(dir), z.rc_normal));
  }
  bool input_sample(float in_w, const Sample& in_z, float u, bool force_add = false) {  // :64-74
    M += 1.0f;
    bool selected = false;
    if (in_w > 0.0f) {
      weight += in_w;
      selected = (u * weight <= in_w) || force_add;
      if (selected) z = in_z;
    }
    return selected;
  }
  bool merge(const Reservoir& in_r, float in_w, float u, bool force_add = false) {  // :76-86
    M += in_r.M;
    bool selected = false;
    if (in_w > 0.0f) {
      weight += in_w;
      selected = (u * weight <= in_w) || force_add;
      if (selected) z = in_r.z;
    }
    return selected;
  }
  void finalize_without_M() {  // :96-102
    float p_hat = luminance(z.F);
    weight = p_hat < 1e-6f ? 0.0f : weight / p_hat;
  }
};

// 56-byte packed record (reservoir.py:8-19) + flags in the spare byte.
struct StorageReservoir {
  uint16_t M, W;  // f16
  float F[3];
  float rc_pos[3];
  uint32_t rc_normal_and_NEE_dir;  // 4 x 8-bit octahedral
  uint16_t rc_incident_dir[2];     // octahedral f16
  float rc_incident_L[3];
  uint32_t rc_mat_info;
  uint16_t cached_jacobian_term;  // f16
  int8_t lobes;
  uint8_t flags;  // bit0 escape vertex, bit1 last vertex, bit2 NEE visible
};
static_assert(sizeof(StorageReservoir) == 56, "packed reservoir must be 56 bytes");

// math_utils.py:201-207 (returns the two values after the f16 cast)
static inline void encode_unit_vector_3x16(V3 v, float& ex, float& ey) {
  float s = std::fabs(v.x) + std::fabs(v.y) + std::fabs(v.z);
  float x = v.x / s, y = v.y / s;
  float ox, oy;
  if (v.z <= 0.0f) {
    ox = (1.0f - std::fabs(y)) * (x >= 0.0f ? 1.0f : -1.0f);
    oy = (1.0f - std::fabs(x)) * (y >= 0.0f ? 1.0f : -1.0f);
  } else {
    ox = x, oy = y;
  }
  ex = round_f16(ox * 0.5f + 0.5f);
  ey = round_f16(oy * 0.5f + 0.5f);
}
// math_utils.py:209-215
static inline V3 decode_unit_vector_3x16(float ax, float ay) {
  float ex = ax * 2.0f - 1.0f, ey = ay * 2.0f - 1.0f;
  V3 v{ex, ey, 1.0f - std::fabs(ex) - std::fabs(ey)};
  float t = fmaxf_(-v.z, 0.0f);
  v.x += v.x >= 0.0f ? -t : t;
  v.y += v.y >= 0.0f ? -t : t;
  return normalize(v);
}
static inline uint32_t unorm8(float x) { return (uint32_t)(x * 255.0f + 0.5f); }  // math_utils.py:250-256, size 8

static inline StorageReservoir encode_reservoir(const Reservoir& r) {  // reservoir.py:104-122
  StorageReservoir e;
  std::memset(&e, 0, sizeof e);
  e.M = f32_to_f16_bits(r.M);
  e.W = f32_to_f16_bits(r.weight);
  e.F[0] = r.z.F.x, e.F[1] = r.z.F.y, e.F[2] = r.z.F.z;
  e.rc_pos[0] = r.z.rc_pos.x, e.rc_pos[1] = r.z.rc_pos.y, e.rc_pos[2] = r.z.rc_pos.z;
  const bool escape = is_vec_zero(r.z.rc_normal), last = is_vec_zero(r.z.rc_incident_dir), nee = !is_vec_zero(r.z.rc_NEE_dir);
  float nx = 0, ny = 0, lx = 0, ly = 0, ix = 0, iy = 0;
  if (!escape) encode_unit_vector_3x16(r.z.rc_normal, nx, ny);
  if (nee) encode_unit_vector_3x16(r.z.rc_NEE_dir, lx, ly);
  if (!last) encode_unit_vector_3x16(r.z.rc_incident_dir, ix, iy);
  e.rc_normal_and_NEE_dir = unorm8(nx) | (unorm8(ny) << 8) | (unorm8(lx) << 16) | (unorm8(ly) << 24);
  e.rc_incident_dir[0] = f32_to_f16_bits(ix), e.rc_incident_dir[1] = f32_to_f16_bits(iy);
  e.rc_incident_L[0] = r.z.rc_incident_L.x, e.rc_incident_L[1] = r.z.rc_incident_L.y, e.rc_incident_L[2] = r.z.rc_incident_L.z;
  e.rc_mat_info = r.z.rc_mat_info;
  e.cached_jacobian_term = f32_to_f16_bits(r.z.cached_jacobian_term);
  e.lobes = (int8_t)r.z.lobes;
  e.flags = (uint8_t)((escape ? 1 : 0) | (last ? 2 : 0) | (nee ? 4 : 0));
  return e;
}
static inline Reservoir decode_reservoir(const StorageReservoir& e) {  // reservoir.py:124-141
  Reservoir r;
  r.M = f16_bits_to_f32(e.M);
  r.weight = f16_bits_to_f32(e.W);
  r.z.F = V3{e.F[0], e.F[1], e.F[2]};
  r.z.rc_pos = V3{e.rc_pos[0], e.rc_pos[1], e.rc_pos[2]};
  const uint32_t p = e.rc_normal_and_NEE_dir;
  r.z.rc_normal = (e.flags & 1) ? V3{0, 0, 0} : decode_unit_vector_3x16((float)(p & 255u) / 255.0f, (float)((p >> 8) & 255u) / 255.0f);
  r.z.rc_NEE_dir = (e.flags & 4) ? decode_unit_vector_3x16((float)((p >> 16) & 255u) / 255.0f, (float)((p >> 24) & 255u) / 255.0f) : V3{0, 0, 0};
  r.z.rc_incident_dir =
      (e.flags & 2) ? V3{0, 0, 0} : decode_unit_vector_3x16(f16_bits_to_f32(e.rc_incident_dir[0]), f16_bits_to_f32(e.rc_incident_dir[1]));
  r.z.rc_incident_L = V3{e.rc_incident_L[0], e.rc_incident_L[1], e.rc_incident_L[2]};
  r.z.rc_mat_info = e.rc_mat_info;
  r.z.cached_jacobian_term = f16_bits_to_f32(e.cached_jacobian_term);
  r.z.lobes = (int)e.lobes;
  return r;
}

// math_utils.py:217-229
static inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t z) {
  x += x >> 11;
  x ^= x << 7;
  x += y;
  x ^= x << 3;
  x += z ^ (x >> 14);
  x ^= x << 6;
  x += x >> 15;
  x ^= x << 5;
  x += x >> 12;
  x ^= x << 9;
  return x;
}

struct GBufferPx {  // pathtracer.py:535-540 (position instead of depth, see header)
  V3 position{0, 0, 0};
  float n_oct[2] = {0, 0};  // octahedral f16 values of the primary normal
  uint32_t mat_info = 0;
  uint8_t sky = 1;
};

}  // namespace orc
