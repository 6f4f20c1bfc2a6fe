// ReSTIR-PT storage shared by the path kernel (producer) and the spatial resampling kernel.
// Replaces renderer/reservoir.py:8-141 (Sample / Reservoir / StorageReservoir) and the packing
// helpers renderer/math_utils.py:201-215,250-263. The 56-byte record keeps the reference's field
// precision (f16 M / W / Jacobian term, 4 x 8-bit octahedral normal + NEE direction, 2 x f16
// octahedral incident direction); the "zero vector" markers travel as flag bits in the spare
// byte (ReSTIR pins).
#pragma once
#include <cuda_fp16.h>

#include "vrt_common.cuh"

struct RSample {
  f3 F, rc_pos, rc_normal, rc_incident_dir, rc_incident_L, rc_NEE_dir;
  uint32_t rc_mat_info;
  float cached_jacobian_term;
  int lobes;
};
struct RReservoir {
  RSample z;
  float M, weight;
};

HD bool is_vec_zero(f3 v) { return dot(v, v) < 1e-7f; }
HD float h16(float x) { return __half2float(__float2half_rn(x)); }
HD uint32_t h16bits(float x) { return (uint32_t)__half_as_ushort(__float2half_rn(x)); }
HD float h16val(uint32_t b) { return __half2float(__ushort_as_half((unsigned short)(b & 0xffffu))); }

HD void rinit(RReservoir& r) {
  r.z.F = r.z.rc_pos = r.z.rc_normal = r.z.rc_incident_dir = r.z.rc_incident_L = r.z.rc_NEE_dir = mk3(0.0f);
  r.z.rc_mat_info = 0u, r.z.cached_jacobian_term = 1.0f, r.z.lobes = 0;
  r.M = 0.0f, r.weight = 0.0f;
}
// math_utils.py:201-207 (IEEE division: the 8-bit / f16 quantisation amplifies nothing, but the
// record is compared byte-wise with the oracle)
HD void encode_unit_vector_3x16(f3 v, float& ex, float& ey) {
  float s = __fadd_rn(__fadd_rn(fabsf(v.x), fabsf(v.y)), fabsf(v.z));
  float x = __fdiv_rn(v.x, s), y = __fdiv_rn(v.y, s);
  float ox, oy;
  if (v.z <= 0.0f) {
    ox = __fmul_rn(__fsub_rn(1.0f, fabsf(y)), x >= 0.0f ? 1.0f : -1.0f);
    oy = __fmul_rn(__fsub_rn(1.0f, fabsf(x)), y >= 0.0f ? 1.0f : -1.0f);
  } else {
    ox = x, oy = y;
  }
  ex = h16(__fadd_rn(__fmul_rn(ox, 0.5f), 0.5f));
  ey = h16(__fadd_rn(__fmul_rn(oy, 0.5f), 0.5f));
}
// math_utils.py:209-215
HD f3 decode_unit_vector_3x16(float ax, float ay) {
  float ex = ax * 2.0f - 1.0f, ey = ay * 2.0f - 1.0f;
  f3 v{ex, ey, 1.0f - fabsf(ex) - fabsf(ey)};
  float t = fmaxf(-v.z, 0.0f);
  v.x += v.x >= 0.0f ? -t : t;
  v.y += v.y >= 0.0f ? -t : t;
  return normalize(v);
}
HD uint32_t unorm8(float x) { return (uint32_t)__fadd_rn(__fmul_rn(x, 255.0f), 0.5f); }

// 14 words: [0] M|W<<16  [1-3] F  [4-6] rc_pos  [7] normal+NEE oct8x4  [8] incident oct f16x2
// [9-11] rc_incident_L  [12] rc_mat_info  [13] jacobian f16 | lobes<<16 | flags<<24
HD void encode_reservoir(const RReservoir& r, uint32_t w[14]) {
  const bool escape = is_vec_zero(r.z.rc_normal), last = is_vec_zero(r.z.rc_incident_dir), nee = !is_vec_zero(r.z.rc_NEE_dir);
  float nx = 0.0f, ny = 0.0f, lx = 0.0f, ly = 0.0f, ix = 0.0f, iy = 0.0f;
  if (!escape) encode_unit_vector_3x16(r.z.rc_normal, nx, ny);
  if (nee) encode_unit_vector_3x16(r.z.rc_NEE_dir, lx, ly);
  if (!last) encode_unit_vector_3x16(r.z.rc_incident_dir, ix, iy);
  w[0] = h16bits(r.M) | (h16bits(r.weight) << 16);
  w[1] = __float_as_uint(r.z.F.x), w[2] = __float_as_uint(r.z.F.y), w[3] = __float_as_uint(r.z.F.z);
  w[4] = __float_as_uint(r.z.rc_pos.x), w[5] = __float_as_uint(r.z.rc_pos.y), w[6] = __float_as_uint(r.z.rc_pos.z);
  w[7] = unorm8(nx) | (unorm8(ny) << 8) | (unorm8(lx) << 16) | (unorm8(ly) << 24);
  w[8] = h16bits(ix) | (h16bits(iy) << 16);
  w[9] = __float_as_uint(r.z.rc_incident_L.x), w[10] = __float_as_uint(r.z.rc_incident_L.y), w[11] = __float_as_uint(r.z.rc_incident_L.z);
  w[12] = r.z.rc_mat_info;
  const uint32_t flags = (escape ? 1u : 0u) | (last ? 2u : 0u) | (nee ? 4u : 0u);
  w[13] = h16bits(r.z.cached_jacobian_term) | (((uint32_t)r.z.lobes & 255u) << 16) | (flags << 24);
}
HD void decode_reservoir(const uint32_t w[14], const float* __restrict__ unorm8_lut, RReservoir& r) {
  r.M = h16val(w[0]);
  r.weight = h16val(w[0] >> 16);
  r.z.F = f3{__uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3])};
  r.z.rc_pos = f3{__uint_as_float(w[4]), __uint_as_float(w[5]), __uint_as_float(w[6])};
  const uint32_t flags = w[13] >> 24, p = w[7];
  r.z.rc_normal = (flags & 1u) ? mk3(0.0f) : decode_unit_vector_3x16(unorm8_lut[p & 255u], unorm8_lut[(p >> 8) & 255u]);
  r.z.rc_NEE_dir = (flags & 4u) ? decode_unit_vector_3x16(unorm8_lut[(p >> 16) & 255u], unorm8_lut[(p >> 24) & 255u]) : mk3(0.0f);
  r.z.rc_incident_dir = (flags & 2u) ? mk3(0.0f) : decode_unit_vector_3x16(h16val(w[8]), h16val(w[8] >> 16));
  r.z.rc_incident_L = f3{__uint_as_float(w[9]), __uint_as_float(w[10]), __uint_as_float(w[11])};
  r.z.rc_mat_info = w[12];
  r.z.cached_jacobian_term = h16val(w[13]);
  r.z.lobes = (int)(signed char)((w[13] >> 16) & 255u);
}
// reservoir.py:59-62
HD float jacobian_term(f3 rc_pos, f3 rc_normal, f3 x1) {
  f3 dir = rc_pos - x1;
  return fdiv(dot(dir, dir), fabsf(dot(normalize(dir), rc_normal)));
}

// Per-pixel G-buffer of the ReSTIR mode (pathtracer.py:112-125): float4 (position, sky flag) and
// uint2 (octahedral f16x2 primary normal, packed material + albedo).
struct RestirBuffers {
  uint2* reservoirs;  // 7 x uint2 per pixel
  float4* gpos;
  uint2* gattr;
  float4* col_d;
  float4* col_s;
  float4* rc_skyT;  // sun transmittance at the reservoir's reconnection vertex (k_rc_sky), read by k_gris
  // temporal reuse (vrt_set_restir_temporal): the second reservoir slot of pathtracer.py:108-109 — the previous
  // frame's temporally resampled reservoir of every pixel with its G-buffer record and sun transmittance
  uint2* hist_res;
  float4* hist_gpos;
  uint2* hist_gattr;
  float4* hist_skyT;
  int temporal;  // 1: spatial_GRIS takes the neighbour's own target value from its stored integrand
};

// Outputs of the moving-camera variant of the path kernel (pathtracer.py:535-546,628-632).
struct MovingOut {
  float4* col_d;  // albedo-demodulated diffuse of the frame
  float4* col_s;
  float* depth;   // gbuff_depth (NDC z in [0,1])
  uint2* attr;    // octahedral f16x2 normal, packed material + albedo
  float* refl;    // gbuff_depth_reflection (linear), 0 = none
  float scale;    // render_scale
};

// One frame of the moving-camera temporal filters: current G-buffer / colours, previous frame's
// G-buffer and history slot (read), this frame's history slot and colour (written).
struct MovingFrame {
  float4* col_d;
  float4* col_s;
  const float* depth;
  const uint2* attr;
  const float* refl;
  float* refl_blur;
  const float* depth_prev;
  const uint2* attr_prev;
  const float4* hd_prev;
  const float4* hs_prev;
  const float* hsd_prev;
  float4* hd;
  float4* hs;
  float* hsd;
  float4* out;
  float prev_view[16], prev_proj[16];
  float scale;
};
