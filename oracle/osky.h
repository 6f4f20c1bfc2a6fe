// ORACLE — TEST INFRASTRUCTURE ONLY. See ovec.h header for the rules.
// CPU restatement of voxel-rt2's physical sky + clouds (renderer/atmos.py, whole file):
//   run time   : sample_skybox :94-115, sample_skybox_transmittance :117-131, project_sky :428-440
//   precompute : generate_transmittance_lut :462-498, compute_cloud_ambient :134-138,
//                accumulate_clouds :140-157, compute_skybox :159-189, clouds :195-349,
//                atmospheric_scattering :355-425, densities :500-527
// The table resolution (3840 in the reference, atmos.py:66-67) is a parameter so the oracle can
// finish in seconds. Random numbers: ti.random() is replaced by the counter RNG of ovec.h with
// key = path_key(texel, pass, seed) and a per-thread running counter.
#pragma once
#include <vector>

#include "obsdf.h"
#include "ovec.h"

namespace orc {

struct AtmosConst {
  // atmos.py:38-83 — python-double arithmetic first, then cast to f32 as Taichi does
  V3 rayleigh_coeff{0.00000519673f, 0.0000121427f, 0.0000296453f};
  float mie_coeff = 8.6e-6f;
  V3 ozone_coeff;
  float mie_ext;  // mie_coeff * 1.11
  float scale_height_rayl = 8500.0f, scale_height_mie = 1200.0f;
  float mie_g = 0.75f;
  float planet_r_offset = 0.0f;
  float planet_r = 6371e3f;
  float atmos_height = 110e3f;
  float cloud_height = 2000.0f;     // 1000 + 1e3
  float cloud_thickness = 340.0f;   // 170 * 2
  float cloud_density = 0.27f;
  float cloud_extinc = 0.075f;
  float cloud_scatter = 0.075f;
  V3 cam_pos;
  AtmosConst() {
    double air = 2.5035422e25, ozone_peak = 8e-6;
    double ozone_num = air * 0.012588 * ozone_peak;
    double cs[3] = {4.51103766177301e-21 * 0.0001, 3.2854797958699e-21 * 0.0001, 1.96774621921165e-22 * 0.0001};
    ozone_coeff = V3{(float)(cs[0] * ozone_num), (float)(cs[1] * ozone_num), (float)(cs[2] * ozone_num)};
    mie_ext = (float)(8.6e-6 * 1.11);
    cam_pos = V3{0.0f, (float)(6371e3 + 0.0 + 1e3), 0.0f};
  }
};

struct Sky {
  AtmosConst c;
  int S = 0;                        // table resolution
  std::vector<float> scatter, trans;  // [x][y][3]
  std::vector<uint16_t> trans_lut;    // f16 [256][128][3]
  std::vector<uint8_t> cloud_tex;     // [256][256][3] indexed [x][y] (imread layout: x, flipped y)
  V3 cloud_ambient{0, 0, 0};
  int use_clouds = 0;
  uint32_t seed = 0;
};

struct Ctr {  // sequential counter RNG for one precompute thread
  uint32_t key, n;
  float next() { return rnd(key, n++); }
};

// atmos.py:9-15. NB sqrt of a negative discriminant is NaN and NaN<0 is false, so a miss returns
// (NaN, NaN), not (-1,-1) — kept.
static inline V2 rsi(V3 pos, V3 dir, float r) {
  float b = dot(pos, dir);
  float discr = b * b - dot(pos, pos) + r * r;
  discr = std::sqrt(discr);
  if (discr < 0.0f) return V2{-1.0f, -1.0f};
  return V2{-b + -discr, -b + discr};
}
static inline float rayleigh_phase(float c) { return 3.0f / (16.0f * kPi) * (1.0f + c * c); }  // :18-20
static inline float mie_phase(float c, float g) {                                               // :22-25
  return (1.0f - g * g) / (4.0f * kPi * std::pow(1.0f + g * g - 2.0f * g * c, 1.5f));
}
static inline V3 get_unit_vec(float rx, float ry) {  // :27-31
  rx *= kPi * 2.0f;
  ry = ry * 2.0f - 1.0f;
  float s = std::sqrt(1.0f - ry * ry);
  return normalize(V3{std::sin(rx) * s, std::cos(rx) * s, ry});
}

// atmos.py:428-440 / :442-455
static inline V2 project_sky(V3 d, float fres) {
  V2 p = normalize(V2{d.x, d.z});
  float azimuth = kPi + std::atan2(p.x, -p.y);
  float elevation = kPi * 0.5f - std::acos(d.y);
  float cx = azimuth / (kPi * 2.0f);
  float cy = 0.5f + 0.5f * signf(elevation) * std::sqrt(2.0f / kPi * std::fabs(elevation));
  return V2{cx * (1.0f - fres) + 0.5f * fres, cy * (1.0f - fres) + 0.5f * fres};
}
static inline V3 unproject_sky(V2 uv, float fres) {
  float cx = (uv.x - 0.5f * fres) / (1.0f - 1.0f * fres);
  float cy = (uv.y - 0.5f * fres) / (1.0f - 1.0f * fres);
  cy = cy < 0.5f ? -sqr(1.0f - 2.0f * cy) : sqr(2.0f * cy - 1.0f);
  float azimuth = cx * 2.0f * kPi - kPi;
  float elevation = cy * 0.5f * kPi;
  float ce = std::cos(elevation), se = std::sin(elevation), ca = std::cos(azimuth), sa = std::sin(azimuth);
  return normalize(V3{ce * sa, se, -ce * ca});
}

// Manual bilinear with wrap on both axes (atmos.py:97-113). icoord is clamped to
// the table (pinned: the reference would index out of range on a rounding overshoot).
static inline V3 bilinear(const float* tab, int S, V2 tc) {
  float fx = tc.x * (float)S - 0.5f, fy = tc.y * (float)S - 0.5f;
  int ix = (int)fx, iy = (int)fy;
  float wx = fractf(fx), wy = fractf(fy);
  ix = ix < 0 ? 0 : (ix > S - 1 ? S - 1 : ix);
  iy = iy < 0 ? 0 : (iy > S - 1 ? S - 1 : iy);
  int ix1 = (ix + 1) % S, iy1 = (iy + 1) % S;
  auto at = [&](int x, int y) {
    const float* p = tab + ((size_t)x * S + y) * 3;
    return V3{p[0], p[1], p[2]};
  };
  V3 bl = at(ix, iy), br = at(ix1, iy), tl = at(ix, iy1), tr = at(ix1, iy1);
  return mix3(mix3(bl, br, wx), mix3(tl, tr, wx), wy);
}
// atmos.py:94-115: direction jittered by 0.0015*rand^3 first
static inline void sample_skybox(const float* scat, const float* trans, int S, V3 d, float r0, float r1, float r2, V3& sc,
                                 V3& tr) {
  V3 dj = normalize(d + V3{r0, r1, r2} * 0.0015f);
  V2 tc = project_sky(dj, 1.0f / (float)S);
  sc = bilinear(scat, S, tc);
  tr = bilinear(trans, S, tc);
}
static inline V3 sample_skybox_transmittance(const float* trans, int S, V3 d) {
  return bilinear(trans, S, project_sky(d, 1.0f / (float)S));
}

// ------------------------------------------------------------------------------ precompute
static inline float get_elevation(const AtmosConst& c, V3 p) {  // :525-527
  return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z) - c.planet_r;
}
static inline float get_ozone_density(float h) {  // :500-517
  float h_km = h * 0.001f;
  float rel = h_km - 25.0f;
  rel = rel * rel;
  float d = (1.0f - 0.375f) * std::exp(-rel / 49.0f);
  d += 0.375f * std::exp(-rel / 256.0f);
  d += fmaxf_(0.0f, -0.000015f * std::pow(h_km - 15.0f, 3.0f));
  return d * 4.0f;
}
static inline V3 get_density(const AtmosConst& c, float h) {  // :519-522
  h = fmaxf_(h, 0.0f);
  return V3{std::exp(-h / c.scale_height_rayl), std::exp(-h / c.scale_height_mie), get_ozone_density(h)};
}
static inline V3 extinc_mul(const AtmosConst& c, V3 v) {  // extinc_mat @ v  (:46-48)
  return V3{(c.rayleigh_coeff.x * v.x + c.mie_ext * v.y) + c.ozone_coeff.x * v.z,
            (c.rayleigh_coeff.y * v.x + c.mie_ext * v.y) + c.ozone_coeff.y * v.z,
            (c.rayleigh_coeff.z * v.x + c.mie_ext * v.y) + c.ozone_coeff.z * v.z};
}
static inline V3 read_trans_lut(const Sky& s, float cos_theta, float h) {  // :457-460
  int ux = (int)clampf((cos_theta * 0.5f + 0.5f) * 256.0f, 0.0f, 255.0f);
  int uy = (int)clampf((h / s.c.atmos_height) * 128.0f, 0.0f, 127.0f);
  const uint16_t* p = &s.trans_lut[((size_t)ux * 128 + uy) * 3];
  return V3{f16_bits_to_f32(p[0]), f16_bits_to_f32(p[1]), f16_bits_to_f32(p[2])};
}
static inline V3 get_ray_transmittance(const AtmosConst& c, V3 ray_pos, V3 ray_dir) {  // :475-498
  const int steps = 128;
  const float fsteps = 1.0f / 128.0f;
  float step_delta = rsi(ray_pos, ray_dir, c.planet_r + c.atmos_height).y * fsteps;
  V3 ray_step = ray_dir * step_delta;
  ray_pos = ray_pos + ray_step * (0.5f * (fmaxf_(ray_dir.y, 0.0f) * 0.5f + 0.5f));
  V3 od{0, 0, 0};
  for (int i = 0; i < steps; i++) {
    float e = get_elevation(c, ray_pos);
    V3 dens = get_density(c, e);
    od += dens * step_delta;
    ray_pos += ray_step;
  }
  od = extinc_mul(c, od);
  V3 T = exp3(-od);
  if (rsi(ray_pos, ray_dir, c.planet_r).x > 0.0f) T *= 0.0f;
  return T;
}
static inline void generate_transmittance_lut(Sky& s) {  // :462-473
  s.trans_lut.assign((size_t)256 * 128 * 3, 0);
  for (int x = 0; x < 256; x++)
    for (int y = 0; y < 128; y++) {
      float cos_theta = ((float)x / 256.0f) * 2.0f - 1.0f;
      float h = s.c.atmos_height * (float)y / 128.0f;
      float theta = std::acos(cos_theta);
      float sin_theta = std::sin(theta);
      V3 T = get_ray_transmittance(s.c, V3{0.0f, s.c.planet_r + h, 0.0f}, V3{sin_theta, cos_theta, 0.0f});
      uint16_t* p = &s.trans_lut[((size_t)x * 128 + y) * 3];
      p[0] = f32_to_f16_bits(T.x), p[1] = f32_to_f16_bits(T.y), p[2] = f32_to_f16_bits(T.z);
    }
}

// atmos.py:355-425. depth is a compile-time constant in the reference (ti.template); depth 2
// returns (0, 1) immediately, so the recursive samples issued at depth 1 add exact zeros and
// are skipped here.
static inline void atmospheric_scattering(const Sky& s, V3 ray_origin, V3 ray_dir, V3 sun_dir, V3 sun_col, float cosmax,
                                          int depth, int steps, Ctr& rng, V3& in_scatter_col, V3& transmittance) {
  const AtmosConst& c = s.c;
  float fsteps = 1.0f / (float)steps;
  V2 air = rsi(ray_origin, ray_dir, c.planet_r + c.atmos_height);
  V2 planet = rsi(ray_origin, ray_dir, c.planet_r);
  air.y = planet.x > 0.0f ? fminf_(air.y, planet.x) : air.y;
  float step_delta = (air.y - fmaxf_(air.x, 0.0f)) * fsteps;
  V3 ray_step = ray_dir * step_delta;
  V3 ray_pos = ray_origin + ray_step * 0.5f;
  transmittance = V3{1, 1, 1};
  in_scatter_col = V3{0, 0, 0};
  for (int i = 0; i < steps; i++) {
    float h = get_elevation(c, ray_pos);
    V3 density = get_density(c, h);
    V3 step_od = extinc_mul(c, density * step_delta);
    V3 step_T = saturate3(exp3(-step_od));
    V3 visible = transmittance * saturate3((V3{1, 1, 1} - step_T) / step_od);
    const int DIRECT = 8;
    for (int j = 0; j < DIRECT; j++) {
      float u0 = rng.next(), u1 = rng.next();
      V3 sample_dir = sample_cone_oriented(cosmax, sun_dir, u0, u1);
      float cos_theta = dot(ray_dir, sample_dir);
      float ph_r = rayleigh_phase(cos_theta), ph_m = mie_phase(cos_theta, c.mie_g);
      V3 sun_T = read_trans_lut(s, dot(normalize(ray_pos), sample_dir), h);
      in_scatter_col += c.rayleigh_coeff * sun_col * sun_T * visible * ph_r * density.x * step_delta / (float)DIRECT;
      in_scatter_col += c.mie_coeff * sun_col * sun_T * visible * ph_m * density.y * step_delta / (float)DIRECT;
    }
    if (depth + 1 <= 1) {
      const float ms_energy = 5.3f;
      const int MS = 8;
      for (int j = 0; j < MS; j++) {
        V3 sample_dir = get_unit_vec(((float)j + 0.5f) / (float)MS, fractf((float)j * 1.618033988749f));
        float cos_theta = dot(ray_dir, sample_dir);
        float ph_m = mie_phase(cos_theta, c.mie_g);
        V3 amb, amb_T;
        atmospheric_scattering(s, ray_pos, sample_dir, sun_dir, sun_col, cosmax, depth + 1, 5, rng, amb, amb_T);
        in_scatter_col += ms_energy * c.rayleigh_coeff * amb * visible * density.x * step_delta / (float)MS;
        in_scatter_col += ms_energy * c.mie_coeff * amb * visible * ph_m * density.y * step_delta / (float)MS;
      }
    }
    transmittance *= step_T;
    ray_pos += ray_step;
  }
  if (planet.x > 0.0f) transmittance *= 0.0f;
}

// atmos.py:195-230. The xz offset is applied to the local copy *before* the height is measured.
static inline float sample_cloud_density(const Sky& s, V3 ray_pos) {
  const AtmosConst& c = s.c;
  const float tile_size = 29000.0f;
  ray_pos.x += tile_size * 0.65f;
  ray_pos.z += tile_size * 0.65f;
  float ux = (ray_pos.x - tile_size * std::floor(ray_pos.x / tile_size)) / tile_size;  // mod(x, y) = x - y*floor(x/y)
  float uz = (ray_pos.z - tile_size * std::floor(ray_pos.z / tile_size)) / tile_size;
  int cx = (int)(ux * 256.0f), cy = (int)(uz * 256.0f);
  cx = cx < 0 ? 0 : (cx > 255 ? 255 : cx);  // pinned: rounding can give exactly 256
  cy = cy < 0 ? 0 : (cy > 255 ? 255 : cy);
  float relative_height = length(ray_pos) - c.planet_r - c.planet_r_offset;
  const uint8_t* t = &s.cloud_tex[((size_t)cx * 256 + cy) * 3];
  float tx = (float)t[0] / 255.0f, ty = (float)t[1] / 255.0f, tz = (float)t[2] / 255.0f;
  if (tx < 0.7f) tx = 0.0f;
  if (ty < 0.7f) ty = 0.0f;
  if (tz < 0.7f) tz = 0.0f;
  float cloud = relative_height < c.cloud_height + c.cloud_thickness * 0.65f ? tx : ty;
  float coverage = tz;
  bool in_layer = relative_height > c.cloud_height && relative_height < c.cloud_height + c.cloud_thickness;
  return in_layer ? c.cloud_density * coverage * cloud : 0.0f;
}
// atmos.py:237-266 (the `continue` skips the position advance — kept)
static inline float clouds_shadow_od(const Sky& s, V3 ray_origin, V3 ray_dir, float dither) {
  const AtmosConst& c = s.c;
  const int steps = 8;
  const float exponent = 1.6f;
  float step_delta = 24.0f / (float)steps;
  float od = 0.0f;
  V3 ray_pos = ray_origin;
  V3 ray_step = ray_dir * step_delta;
  for (int i = 0; i < steps; i++) {
    ray_step *= exponent;
    step_delta *= exponent;
    V3 dp = ray_pos + ray_step * dither;
    float rh = length(dp) - c.planet_r - c.planet_r_offset;
    if (rh < c.cloud_height || rh > c.cloud_height + c.cloud_thickness) continue;
    od += sample_cloud_density(s, dp) * step_delta;
    ray_pos += ray_step;
  }
  return od;
}
static inline float cloud_phase(float cos_theta, float an) {  // :268-273
  float peak = mie_phase(cos_theta, 0.92f * an);
  float front = mie_phase(cos_theta, 0.4f * an);
  float back = mie_phase(cos_theta, -0.55f * an);
  return mixf(mixf(front, back, 0.5f), peak, 0.15f);
}
// atmos.py:275-349
static inline void clouds_scattering(const Sky& s, V3 ray_origin, V3 ray_dir, V3 sun_dir, V3 sun_col, float cosmax,
                                     float dither, Ctr& rng, V3& in_scatter, float& transmittance, float& weighted_dist) {
  const AtmosConst& c = s.c;
  const int steps = 32;
  const float fsteps = 1.0f / (float)steps;
  float bottom = rsi(ray_origin, ray_dir, c.planet_r + c.planet_r_offset + c.cloud_height).y;
  float top = rsi(ray_origin, ray_dir, c.planet_r + c.planet_r_offset + c.cloud_height + c.cloud_thickness).y;
  transmittance = 1.0f;
  in_scatter = V3{0, 0, 0};
  float weight_sum = 0.0f;
  weighted_dist = 0.0f;
  V3 start = ray_origin + ray_dir * bottom;
  float step_delta = (top - bottom) * fsteps;
  V3 ray_step = ray_dir * step_delta;
  V3 ray_pos = start + ray_step * dither;
  float distance_traveled = length(start - ray_origin);
  for (int i = 0; i < steps; i++) {
    float density = sample_cloud_density(s, ray_pos);
    if (density <= 0.0f || transmittance <= 1e-4f) {
      ray_pos += ray_step;
      distance_traveled += step_delta;
      weighted_dist += distance_traveled * transmittance;
      weight_sum += transmittance;
      continue;
    }
    float step_od = c.cloud_extinc * density * step_delta;
    float step_T = saturate(std::exp(-step_od));
    float step_weight = (1.0f - step_T) / c.cloud_extinc;
    float visible = transmittance * step_weight;
    const int DIRECT = 8;
    for (int j = 0; j < DIRECT; j++) {
      float u0 = rng.next(), u1 = rng.next();
      V3 sample_dir = sample_cone_oriented(cosmax, sun_dir, u0, u1);
      float cos_theta = dot(ray_dir, sample_dir);
      float sun_ray_od = clouds_shadow_od(s, ray_pos, sample_dir, dither);
      V3 sun_T = read_trans_lut(s, dot(normalize(ray_pos), sample_dir), get_elevation(c, ray_pos));
      float an = 1.0f;
      for (int k = 0; k < 4; k++) {
        float phase = cloud_phase(cos_theta, an);
        in_scatter += visible * an * c.cloud_scatter * phase * std::exp(-sun_ray_od * c.cloud_extinc * an) * sun_T * sun_col /
                      (float)DIRECT;
        an *= 0.5f;
      }
    }
    float ambient_od = clouds_shadow_od(s, ray_pos, V3{0, 1, 0}, dither);
    float an = 1.0f;
    for (int k = 0; k < 4; k++) {
      in_scatter += visible * an * c.cloud_scatter / (4.0f * kPi) * std::exp(-ambient_od * c.cloud_extinc * an) * s.cloud_ambient;
      an *= 0.5f;
    }
    transmittance *= step_T;
    ray_pos += ray_step;
    distance_traveled += step_delta;
    weighted_dist += distance_traveled * transmittance;
    weight_sum += transmittance;
  }
  weighted_dist /= weight_sum;
}

// Full precompute in the order Scene.finish drives it (scene.py:172,243-253; pathtracer.py:314-329):
// LUT -> cloud ambient -> zero tables -> n_cloud_passes x accumulate_clouds -> compute_skybox.
static inline void precompute_sky(Sky& s, int S, V3 sun_dir, V3 sun_col, float cosmax, int n_cloud_passes) {
  s.S = S;
  const float fres = 1.0f / (float)S;
  generate_transmittance_lut(s);
  {
    Ctr rng{path_key(0xFFFFFFFFu, 2000u, s.seed), 0};
    V3 amb, ambT;
    atmospheric_scattering(s, s.c.cam_pos + V3{0.0f, s.c.cloud_height, 0.0f}, V3{0, 1, 0}, sun_dir, sun_col, cosmax, 0, 64, rng,
                           amb, ambT);
    s.cloud_ambient = amb;
  }
  s.scatter.assign((size_t)S * S * 3, 0.0f);
  s.trans.assign((size_t)S * S * 3, 0.0f);
  const float fmax_samples = 1.0f / (float)n_cloud_passes;
  for (int pass = 0; pass < n_cloud_passes; pass++) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int idx = 0; idx < S * S; idx++) {
      int u = idx / S, v = idx % S;
      V2 tc{((float)u + 0.5f) * fres, ((float)v + 0.5f) * fres};
      V3 ray_dir = unproject_sky(tc, fres);
      Ctr rng{path_key((uint32_t)idx, (uint32_t)pass, s.seed), 0};
      float dither = rng.next();
      V3 insc;
      float T, dist;
      clouds_scattering(s, s.c.cam_pos, ray_dir, sun_dir, sun_col, cosmax, dither, rng, insc, T, dist);
      insc *= 1.2f;
      float* ps = &s.scatter[(size_t)idx * 3];
      float* pt = &s.trans[(size_t)idx * 3];
      ps[0] += insc.x * fmax_samples, ps[1] += insc.y * fmax_samples, ps[2] += insc.z * fmax_samples;
      pt[0] += saturate(T) * fmax_samples;
      pt[1] += dist * fmax_samples;
    }
  }
#pragma omp parallel for schedule(dynamic, 4)
  for (int idx = 0; idx < S * S; idx++) {
    int u = idx / S, v = idx % S;
    V2 tc{((float)u + 0.5f) * fres, ((float)v + 0.5f) * fres};
    V3 ray_dir = unproject_sky(tc, fres);
    float* ps = &s.scatter[(size_t)idx * 3];
    float* pt = &s.trans[(size_t)idx * 3];
    V3 cloud_in_scatter{ps[0], ps[1], ps[2]};
    float cloud_T = pt[0], cloud_dist = pt[1];
    Ctr rng{path_key((uint32_t)idx, 1000u, s.seed), 0};
    V3 sc_total, T_total, sc_from, T_from;
    atmospheric_scattering(s, s.c.cam_pos, ray_dir, sun_dir, sun_col, cosmax, 0, 64, rng, sc_total, T_total);
    V3 cloud_pos = s.c.cam_pos + ray_dir * fmaxf_(cloud_dist, 0.0f);
    atmospheric_scattering(s, cloud_pos, ray_dir, sun_dir, sun_col, cosmax, 0, 64, rng, sc_from, T_from);
    V3 T_to_cloud = T_total / T_from;
    V3 in_scattering = sc_total;
    if (s.use_clouds == 1) {
      in_scattering = in_scattering - sc_from * saturate3(T_to_cloud * fmaxf_(1.0f - cloud_T, 0.0f));
      in_scattering += cloud_in_scatter * saturate3(T_to_cloud);
    }
    V3 Tout = T_total * cloud_T;
    ps[0] = in_scattering.x, ps[1] = in_scattering.y, ps[2] = in_scattering.z;
    pt[0] = Tout.x, pt[1] = Tout.y, pt[2] = Tout.z;
  }
}

}  // namespace orc
