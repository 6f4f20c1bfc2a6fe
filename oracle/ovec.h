// ORACLE — TEST INFRASTRUCTURE ONLY (not part of the shipped product path).
// CPU restatement of the float32 vector helpers voxel-rt2's Taichi code relies on.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load anything under oracle/.
//
// Parity status: the reference ships no golden vectors and Taichi is not installable here
// (SURVEY.md §8c); the oracle is pinned against the reference's OWN Python source executed
// through oracle/ti_emu (a float32 Taichi emulator): tests/golden/ref_*.npz, checked by
// tests/test_reference_vectors.py. Semantics below follow Taichi's Python definitions
// (taichi.math: mix/clamp/fract/sign/reflect, Vector.dot/norm/normalized) with IEEE float32,
// no FMA contraction.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

static const float kEps = 1e-6f;                                   // math_utils.py:5
static const float kInf = std::numeric_limits<float>::infinity();  // math_utils.py:6
static const float kPi = 3.14159265358979323846f;                  // np.pi cast to f32

struct V2 {
  float x, y;
};
struct V3 {
  float x, y, z;
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct I3 {
  int x, y, z;
  int operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

static inline V3 v3(float a, float b, float c) { return V3{a, b, c}; }
static inline V3 v3(float a) { return V3{a, a, a}; }
static inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
static inline V3 operator+(V3 a, float s) { return V3{a.x + s, a.y + s, a.z + s}; }
static inline V3 operator-(V3 a, float s) { return V3{a.x - s, a.y - s, a.z - s}; }
static inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
static inline V3& operator+=(V3& a, V3 b) {
  a = a + b;
  return a;
}
static inline V3& operator*=(V3& a, V3 b) {
  a = a * b;
  return a;
}
static inline V3& operator*=(V3& a, float s) {
  a = a * s;
  return a;
}

// Taichi Vector.dot: entries multiplied then summed left to right.
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
static inline V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float length(V3 a) { return std::sqrt(dot(a, a)); }
// Taichi Vector.normalized(): invlen = 1/norm; invlen * v.
static inline V3 normalize(V3 a) {
  float inv = 1.0f / length(a);
  return inv * a;
}
static inline V2 normalize(V2 a) {
  float inv = 1.0f / std::sqrt(a.x * a.x + a.y * a.y);
  return V2{inv * a.x, inv * a.y};
}
static inline float fminf_(float a, float b) { return std::fmin(a, b); }
static inline float fmaxf_(float a, float b) { return std::fmax(a, b); }
static inline float clampf(float x, float lo, float hi) { return fmaxf_(lo, fminf_(hi, x)); }
static inline V3 clamp3(V3 v, float lo, float hi) {
  return V3{clampf(v.x, lo, hi), clampf(v.y, lo, hi), clampf(v.z, lo, hi)};
}
static inline float saturate(float x) { return fminf_(fmaxf_(x, 0.0f), 1.0f); }  // math_utils.py:9-11
static inline V3 saturate3(V3 v) { return V3{saturate(v.x), saturate(v.y), saturate(v.z)}; }
static inline float sqr(float x) { return x * x; }  // math_utils.py:13-15
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
static inline float fractf(float x) { return x - std::floor(x); }
static inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline V3 reflect(V3 i, V3 n) { return i - (2.0f * dot(i, n)) * n; }
static inline V3 max3(V3 a, V3 b) { return V3{fmaxf_(a.x, b.x), fmaxf_(a.y, b.y), fmaxf_(a.z, b.z)}; }
static inline V3 max3(V3 a, float b) { return V3{fmaxf_(a.x, b), fmaxf_(a.y, b), fmaxf_(a.z, b)}; }
static inline V3 exp3(V3 a) { return V3{std::exp(a.x), std::exp(a.y), std::exp(a.z)}; }
static inline bool is_vec_zero(V3 v) { return dot(v, v) < 1e-7f; }  // math_utils.py:17-19
static inline float luminance(V3 c) { return dot(V3{0.2125f, 0.7154f, 0.0721f}, c); }  // math_utils.py:151-153
static inline bool isbad(float x) { return std::isnan(x) || std::isinf(x); }

// ---------------------------------------------------------------------------------------------
// Counter-based RNG shared (by specification, not by code) with the CUDA path.
// The reference uses Taichi's ti.random() (per-thread xorshift; unpinned). Ours:
//   key  = mix32(mix32(pixel ^ 0x9E3779B9*sample) + seed)      per path
//   u(d) = mix32(key + 0x9E3779B9 * (d+1)) >> 8 ) * 2^-24      dimension d
// mix32 = "lowbias32" integer finaliser.
static inline uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
static inline uint32_t path_key(uint32_t pixel, uint32_t sample, uint32_t seed) {
  return mix32(mix32(pixel ^ (0x9E3779B9U * (sample + 1U))) + seed);
}
static inline float rnd(uint32_t key, uint32_t dim) {
  uint32_t h = mix32(key + 0x9E3779B9U * (dim + 1U));
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// IEEE binary16 round-to-nearest-even conversion (Taichi ti.f16 casts), used by the
// transmittance LUT (atmos.py:63,473) and octahedral normal packing (math_utils.py:201-215).
static inline uint16_t f32_to_f16_bits(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t man = x & 0x007fffffu;
  int exp = (int)((x >> 23) & 0xff);
  if (exp == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
  int e = exp - 127 + 15;
  if (e >= 31) return (uint16_t)(sign | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    man |= 0x00800000u;
    int shift = 14 - e;
    uint32_t half = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half & 1u))) half++;
    return (uint16_t)(sign | half);
  }
  uint32_t half = ((uint32_t)e << 10) | (man >> 13);
  uint32_t rem = man & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;
  return (uint16_t)(sign | half);
}
static inline float f16_bits_to_f32(uint16_t h) {
  uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ffu;
  uint32_t x;
  if (exp == 0) {
    if (man == 0) {
      x = sign;
    } else {
      int e = -1;
      do {
        e++;
        man <<= 1;
      } while ((man & 0x400u) == 0);
      man &= 0x3ffu;
      x = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    x = sign | 0x7f800000u | (man << 13);
  } else {
    x = sign | ((exp - 15 + 127) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}
static inline float round_f16(float f) { return f16_bits_to_f32(f32_to_f16_bits(f)); }

}  // namespace orc
