// Shared device-side definitions for libvoxelrt (sm_100a).
// Everything here is written for the B200 path; there is no host fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VRT_INF __int_as_float(0x7f800000)
#define VRT_EPS 1e-6f  // renderer/math_utils.py:5
#define VRT_PI 3.14159265358979323846f

struct f3 {
  float x, y, z;
};
struct f2 {
  float x, y;
};

#define HD __device__ __forceinline__

HD f3 mk3(float a, float b, float c) { return f3{a, b, c}; }
HD f3 mk3(float a) { return f3{a, a, a}; }
HD f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
HD f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
HD f3 operator*(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
HD f3 operator/(f3 a, f3 b) { return f3{__fdividef(a.x, b.x), __fdividef(a.y, b.y), __fdividef(a.z, b.z)}; }
HD f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
HD f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
// Approximate (1-2 ulp, flush-to-zero) SFU ops for the shading code; never used by the traversal.
HD float frcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
HD float fdiv(float a, float b) { return __fdividef(a, b); }
HD float fsqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
HD float frsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
HD f3 operator/(f3 a, float s) {
  float r = frcp(s);
  return f3{a.x * r, a.y * r, a.z * r};
}
HD f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
HD f3& operator+=(f3& a, f3 b) {
  a = a + b;
  return a;
}
HD f3& operator*=(f3& a, f3 b) {
  a = a * b;
  return a;
}
HD f3& operator*=(f3& a, float s) {
  a = a * s;
  return a;
}
HD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
HD f3 cross(f3 a, f3 b) { return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
HD float length(f3 a) { return fsqrt(dot(a, a)); }
HD f3 normalize(f3 a) {
  float inv = frsqrt(dot(a, a));
  return inv * a;
}
HD float clampf(float x, float lo, float hi) { return fmaxf(lo, fminf(hi, x)); }
HD f3 clamp3(f3 v, float lo, float hi) { return f3{clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)}; }
HD float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
HD f3 saturate3(f3 v) { return f3{saturate(v.x), saturate(v.y), saturate(v.z)}; }
HD float sqr(float x) { return x * x; }
HD float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
HD f3 mix3(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
HD float fractf(float x) { return x - floorf(x); }
// sign(x) in {-1, 0, +1}: copy the sign bit onto 1.0f, zero (and NaN) map to 0 like the comparisons did
HD float signf(float x) {
  const float s = __uint_as_float((__float_as_uint(x) & 0x80000000u) | 0x3f800000u);
  return (x > 0.0f || x < 0.0f) ? s : 0.0f;
}
HD f3 reflect(f3 i, f3 n) { return i - (2.0f * dot(i, n)) * n; }
HD f3 exp3(f3 a) { return f3{expf(a.x), expf(a.y), expf(a.z)}; }
HD float luminance(f3 c) { return dot(f3{0.2125f, 0.7154f, 0.0721f}, c); }
HD bool isbad(float x) { return isnan(x) || isinf(x); }
HD float pow5(float x) {
  float x2 = x * x;
  return x2 * x2 * x;
}

// Contraction-free float32 ops: the traversal / camera / floor code is specified op by op so
// primary-hit buffers are bit-identical to the CPU oracle (SURVEY.md Appendix C shows both the
// LOD policy and FMA contraction change hit bits).
HD float xadd(float a, float b) { return __fadd_rn(a, b); }
HD float xsub(float a, float b) { return __fsub_rn(a, b); }
HD float xmul(float a, float b) { return __fmul_rn(a, b); }
HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
HD float xsqrt(float a) { return __fsqrt_rn(a); }
HD float xdot(f3 a, f3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
HD f3 xnormalize(f3 a) {
  float inv = xdiv(1.0f, xsqrt(xdot(a, a)));
  return f3{xmul(inv, a.x), xmul(inv, a.y), xmul(inv, a.z)};
}

// Counter-based RNG (specification shared with the oracle; see DESIGN.md "Sampler").
HD uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
HD uint32_t path_key(uint32_t pixel, uint32_t sample, uint32_t seed) {
  return mix32(mix32(pixel ^ (0x9E3779B9U * (sample + 1U))) + seed);
}
HD float rnd(uint32_t key, uint32_t dim) {
  uint32_t h = mix32(key + 0x9E3779B9U * (dim + 1U));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// Material row in device memory: 20 floats (5 x float4), derived constants precomputed on the host
// (vrt_api.cu pack_materials):
//  [0] base rgb, subsurface           [1] metallic, specular, specular_tint, roughness
//  [2] sheen, sheen_tint, clearcoat, cc_alpha (mix(0.1, 0.001, clearcoat_gloss), bsdf.py:113)
//  [3] dw, sw, cw (lobe probabilities, bsdf.py:351-363), cc_norm = (a2-1)/(pi*log(a2)) (bsdf.py:116-117)
//  [4] ax, ay (bsdf.py:92-95), 1/(pi*ax*ay), -
#define MAT_ROW_F4 5
struct Mat {
  f3 base_col;
  float subsurface, metallic, specular, specular_tint, roughness, sheen, sheen_tint, clearcoat;
  float cc_alpha, dw, sw, cw, cc_norm, ax, ay, inv_pi_axay;
};

struct Params {
  // voxel world
  const unsigned long long* bricks;  // (R/4)^3 64-bit bricks, x fastest; bit = (z&3)*16 + (y&3)*4 + (x&3)
  const uint32_t* upper;             // occupancy bits of LOD 3..n_lods-1, concatenated
  const uint32_t* color;             // RGBA8 per voxel (r | g<<8 | b<<16 | material<<24), brick-major
  int R, n_lods, brick_res, upper_words;
  uint32_t upper_off[8];             // word offset of LOD (3+i) inside `upper`
  float voxel_size, voxel_inv_size, voxel_edges, grid_half;
  // floor / light / background
  float floor_height;
  f3 floor_color;
  int floor_material;
  f3 light_dir;
  float light_cos_max;
  f3 light_color;
  float light_weight;
  f3 background;
  int use_sky;
  // camera
  f3 cam_pos;
  float inv_proj[16], inv_view[16];
  float view[16], proj[16];  // moving-camera temporal path only
  int W, H;
  float inv_w, inv_h;  // 1.0f / W, 1.0f / H (IEEE division on the host, same bits as on the device)
  // sky tables, float4 texels, [x][y] with y fastest
  const float4* sky_scatter;
  const float4* sky_trans;
  const uint4* sky_packed;  // format 1: {scatter rgb, trans rgb, -, -} as 8 x binary16 per texel, or nullptr
  int sky_res;
  // materials
  const float4* mats;
  // accumulation
  float4* accum;
  int accum_overwrite;  // 1: first batch after a reset, the kernel writes the texel instead of adding to it
  // work description
  int first_sample, n_samples, stride;
  const float2* jitter;  // per sample of this batch
  uint32_t seed;
  int max_depth;
  int tile_rank, tile_n, tiles_x, n_tiles;
  unsigned int* work_counter;
  unsigned long long* stats;  // 8 counters or nullptr
};
