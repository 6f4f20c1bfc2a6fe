// ORACLE — TEST INFRASTRUCTURE ONLY. See ovec.h header for the rules.
// CPU restatement of the live part of voxel-rt2's Disney BSDF and its sampling helpers:
//   renderer/bsdf.py:15-458      (translucent code :460-659 is unreachable and omitted)
//   renderer/math_utils.py:21-63 cosine / cone sampling, orthonormal basis
// pdf formulas are kept exactly as written even where they are not the true density of the
// sampler.
#pragma once
#include "ovec.h"

namespace orc {

enum { LOBE_DIFFUSE = 0, LOBE_SPEC_REFL = 1, LOBE_CLEARC = 2, LOBE_ALL = 9 };  // bsdf.py:15-20

struct Mat {  // bsdf.py:26-37 — 14 floats, same order as the CSV columns after the id
  V3 base_col;
  float subsurface, metallic, specular, specular_tint, roughness, anisotropic, sheen, sheen_tint, clearcoat,
      clearcoat_gloss, ior_minus_one;
};

// Random-number source for one path vertex: fixed dimension slots:
//   base+0,1 cone sample; base+2 lobe; base+3,4 direction; base+5,6,7 sky jitter.
struct Rng {
  uint32_t key;
  uint32_t base;
  float get(uint32_t slot) const { return rnd(key, base + slot); }
};

// math_utils.py:32-42
static inline void make_orthonormal_basis(V3 n, V3& x, V3& y) {
  V3 h = std::fabs(n.y) > 0.9f ? V3{1, 0, 0} : V3{0, 1, 0};
  y = normalize(cross(n, h));
  x = cross(n, y);
}

// math_utils.py:21-30 (u0,u1 supplied)
static inline V3 sample_cosine_weighted_hemisphere(V3 n, float u0, float u1) {
  float a = 1.0f - 2.0f * u0;
  float b = std::sqrt(1.0f - a * a);
  a *= 1.0f - 1e-5f;
  b *= 1.0f - 1e-5f;
  float phi = 2.0f * kPi * u1;
  return normalize(V3{n.x + b * std::cos(phi), n.y + b * std::sin(phi), n.z + a});
}

// math_utils.py:44-59
static inline V3 sample_cone(float cos_theta_max, float u0, float u1) {
  float cos_theta = (1.0f - u0) + u0 * cos_theta_max;
  float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
  float phi = 2.0f * kPi * u1;
  return V3{sin_theta * std::cos(phi), sin_theta * std::sin(phi), cos_theta};
}
static inline V3 sample_cone_oriented(float cos_theta_max, V3 n, float u0, float u1) {
  V3 x, y;
  make_orthonormal_basis(n, x, y);
  V3 s = sample_cone(cos_theta_max, u0, u1);
  // mat3(x, y, n).transpose() @ s : rows (x.i, y.i, n.i)
  return V3{(x.x * s.x + y.x * s.y) + n.x * s.z, (x.y * s.x + y.y * s.y) + n.y * s.z, (x.z * s.x + y.z * s.y) + n.z * s.z};
}
// math_utils.py:61-63
static inline float cone_sample_pdf(float cos_theta_max, float cos_theta) {
  return cos_theta >= cos_theta_max ? 1.0f / (2.0f * kPi * (1.0f - cos_theta_max)) : 0.0f;
}

// bsdf.py:39-47
static inline V3 disney_subsurface(const Mat& m, float n_dot_l, float n_dot_v, float l_dot_h, float F_L, float F_V) {
  float Fss90 = l_dot_h * l_dot_h * m.roughness;
  float Fss = mixf(1.0f, Fss90, F_L) * mixf(1.0f, Fss90, F_V);
  float ss = 1.25f * (Fss * (1.0f / (n_dot_l + n_dot_v) - 0.5f) + 0.5f);
  return ((1.0f / kPi) * ss) * m.base_col;
}

// bsdf.py:49-67
static inline V3 disney_diffuse(const Mat& m, float n_dot_l, float n_dot_v, float l_dot_h) {
  float R_R = 2.0f * m.roughness * sqr(l_dot_h);
  float F_L = std::pow(1.0f - n_dot_l, 5.0f);
  float F_V = std::pow(1.0f - n_dot_v, 5.0f);
  V3 f_lambert = m.base_col / kPi;
  V3 f_retro = f_lambert * R_R * (F_L + F_V + F_L * F_V * (R_R - 1.0f));
  V3 f_d = f_lambert * (1.0f - 0.5f * F_L) * (1.0f - 0.5f * F_V) + f_retro;
  float albedo_lum = dot(m.base_col, V3{0.2125f, 0.7154f, 0.0721f});
  V3 sheen_col = albedo_lum > 0.0f ? m.base_col / albedo_lum : V3{1, 1, 1};
  float sheen_schlick = std::pow(1.0f - l_dot_h, 5.0f);
  V3 sheen = m.sheen * mix3(V3{1, 1, 1}, sheen_col, m.sheen_tint) * sheen_schlick;
  V3 ss = disney_subsurface(m, n_dot_l, n_dot_v, l_dot_h, F_L, F_V);
  return mix3(f_d, ss, m.subsurface) + sheen;
}

// bsdf.py:69-75
static inline float GTR2_anisotropic(float n_dot_h, float h_dot_x, float h_dot_y, float ax, float ay) {
  return 1.0f / (kPi * ax * ay * sqr(sqr(h_dot_x / ax) + sqr(h_dot_y / ay) + sqr(n_dot_h)));
}
static inline float smithG_GGX_aniso(float n_dot_v, float v_dot_x, float v_dot_y, float ax, float ay) {
  return 1.0f / (n_dot_v + std::sqrt(sqr(v_dot_x * ax) + sqr(v_dot_y * ay) + sqr(n_dot_v)));
}
// bsdf.py:77-83
static inline V3 disney_fresnel(const Mat& m, float l_dot_h) {
  float albedo_lum = dot(m.base_col, V3{0.2125f, 0.7154f, 0.0721f});
  V3 spec_tint = albedo_lum > 0.0f ? m.base_col / albedo_lum : V3{1, 1, 1};
  V3 spec_col = mix3((m.specular * 0.08f) * mix3(V3{1, 1, 1}, spec_tint, m.specular_tint), m.base_col, m.metallic);
  float F_L = std::pow(1.0f - l_dot_h, 5.0f);
  return mix3(spec_col, V3{1, 1, 1}, F_L);
}
static inline void aniso_alphas(const Mat& m, float& ax, float& ay) {
  float aspect = std::sqrt(1.0f - 0.9f * m.anisotropic);
  ax = fmaxf_(sqr(m.roughness) / aspect, 1e-3f);
  ay = fmaxf_(sqr(m.roughness) * aspect, 1e-3f);
}
// bsdf.py:86-105 (no 1/(4 n.l n.v): the Smith terms already carry the denominators)
static inline V3 disney_specular(const Mat& m, float n_dot_l, float n_dot_v, float l_dot_h, float n_dot_h, float h_dot_x,
                                 float h_dot_y, float l_dot_x, float l_dot_y, float v_dot_x, float v_dot_y) {
  float ax, ay;
  aniso_alphas(m, ax, ay);
  float D = GTR2_anisotropic(n_dot_h, h_dot_x, h_dot_y, ax, ay);
  float G = smithG_GGX_aniso(n_dot_l, l_dot_x, l_dot_y, ax, ay) * smithG_GGX_aniso(n_dot_v, v_dot_x, v_dot_y, ax, ay);
  V3 F = disney_fresnel(m, l_dot_h);
  return (D * G) * F;
}
// bsdf.py:112-135
static inline float GTR1(float n_dot_h, float alpha) {
  float a2 = alpha * alpha;
  float t = 1.0f + (a2 - 1.0f) * n_dot_h * n_dot_h;
  float D = (a2 - 1.0f) / (kPi * std::log(a2) * t);
  if (alpha >= 1.0f) D = 1.0f / kPi;
  return D;
}
static inline float smithG_GGX(float n_dot_v, float alpha) {
  float a2 = alpha * alpha;
  float b = n_dot_v * n_dot_v;
  return 1.0f / (n_dot_v + std::sqrt(a2 + b - a2 * b));
}
static inline float disney_clearcoat(const Mat& m, float n_dot_l, float n_dot_v, float n_dot_h, float l_dot_h) {
  float alpha = mixf(0.1f, 0.001f, m.clearcoat_gloss);
  float D = GTR1(std::fabs(n_dot_h), alpha);
  float F = mixf(0.04f, 1.0f, std::pow(1.0f - l_dot_h, 5.0f));
  float G = smithG_GGX(n_dot_l, 0.25f) * smithG_GGX(n_dot_v, 0.25f);
  return m.clearcoat * D * F * G;
}

// bsdf.py:306-349 disney_evaluate_lobewise_split (lobe_id == LOBE_ALL gives :138-177 evaluate_split)
static inline void disney_evaluate_lobewise_split(const Mat& m, V3 v, V3 n, V3 l, V3 tang, V3 bitang, int lobe_id,
                                                  V3& bsdf_d, V3& bsdf_s, float specular_mult = 1.0f) {
  float n_dot_l = dot(n, l);
  float n_dot_v = dot(n, v);
  bsdf_d = V3{0, 0, 0};
  bsdf_s = V3{0, 0, 0};
  if (n_dot_l > 0.0f && n_dot_v > 0.0f) {
    V3 h = normalize(l + v);
    float l_dot_h = dot(l, h), n_dot_h = dot(n, h);
    float h_dot_x = dot(h, tang), h_dot_y = dot(h, bitang);
    float l_dot_x = dot(l, tang), l_dot_y = dot(l, bitang);
    float v_dot_x = dot(v, tang), v_dot_y = dot(v, bitang);
    if (lobe_id == LOBE_DIFFUSE || lobe_id == LOBE_ALL)
      bsdf_d += disney_diffuse(m, n_dot_l, n_dot_v, l_dot_h) * (1.0f - m.metallic);
    if (lobe_id == LOBE_SPEC_REFL || lobe_id == LOBE_ALL)
      bsdf_s += disney_specular(m, n_dot_l, n_dot_v, l_dot_h, n_dot_h, h_dot_x, h_dot_y, l_dot_x, l_dot_y, v_dot_x,
                                v_dot_y) *
                specular_mult;
    if (lobe_id == LOBE_CLEARC || lobe_id == LOBE_ALL)
      bsdf_s += v3(disney_clearcoat(m, n_dot_l, n_dot_v, n_dot_h, l_dot_h) * specular_mult);
  }
}
static inline void disney_evaluate_split(const Mat& m, V3 v, V3 n, V3 l, V3 tang, V3 bitang, V3& d, V3& s) {
  disney_evaluate_lobewise_split(m, v, n, l, tang, bitang, LOBE_ALL, d, s);
}
static inline V3 disney_evaluate(const Mat& m, V3 v, V3 n, V3 l, V3 tang, V3 bitang) {
  V3 d, s;
  disney_evaluate_split(m, v, n, l, tang, bitang, d, s);
  return d + s;
}
static inline V3 disney_evaluate_lobewise(const Mat& m, V3 v, V3 n, V3 l, V3 tang, V3 bitang, int lobe) {
  V3 d, s;
  disney_evaluate_lobewise_split(m, v, n, l, tang, bitang, lobe, d, s);
  return d + s;
}

// bsdf.py:179-182
static inline float pdf_diffuse(V3 n, V3 l) { return saturate(dot(l, n)) / kPi; }
// bsdf.py:190-199
static inline float pdf_clearcoat(const Mat& m, V3 v, V3 n, V3 l) {
  float alpha = mixf(0.1f, 0.001f, m.clearcoat_gloss);
  V3 h = normalize(v + l);
  float n_dot_h = std::fabs(dot(n, h));
  float v_dot_h = dot(v, h);
  float D = GTR1(n_dot_h, alpha);
  return D * n_dot_h / (4.0f * v_dot_h);
}
// bsdf.py:254-277
static inline float pdf_specular(const Mat& m, V3 v, V3 n, V3 l, V3 tang, V3 bitang) {
  float ax, ay;
  aniso_alphas(m, ax, ay);
  V3 h = normalize(v + l);
  float n_dot_l = std::fabs(dot(n, l));
  float n_dot_v = dot(n, v);
  float l_dot_h = std::fabs(dot(l, h));
  float n_dot_h = dot(n, h);
  float h_dot_x = dot(h, tang), h_dot_y = dot(h, bitang);
  float v_dot_x = dot(v, tang), v_dot_y = dot(v, bitang);
  float D = GTR2_anisotropic(n_dot_h, h_dot_x, h_dot_y, ax, ay);
  float G = smithG_GGX_aniso(n_dot_v, v_dot_x, v_dot_y, ax, ay);
  return G * l_dot_h * D / n_dot_l;
}

// bsdf.py:351-363
static inline void lobe_probabilities(const Mat& m, float& dw, float& sw, float& cw) {
  dw = (1.0f - m.metallic) * clampf(1.0f - m.specular, 0.4f, 0.9f);
  sw = 1.0f - dw;
  cw = m.clearcoat * 0.7f;
  float w_sum = dw + sw + cw;
  dw /= w_sum;
  sw /= w_sum;
  cw /= w_sum;
}
// bsdf.py:365-380
static inline float pdf_disney_lobewise(const Mat& m, V3 v, V3 n, V3 l, V3 tang, V3 bitang, int lobe) {
  float dw, sw, cw;
  lobe_probabilities(m, dw, sw, cw);
  float pdf = 1.0f;
  if (lobe == LOBE_DIFFUSE)
    pdf *= pdf_diffuse(n, l) * dw;
  else if (lobe == LOBE_SPEC_REFL)
    pdf *= pdf_specular(m, v, n, l, tang, bitang) * sw;
  else
    pdf *= pdf_clearcoat(m, v, n, l) * cw;
  if (isbad(pdf)) pdf = 1.0f;
  return pdf;
}
// bsdf.py:382-393
static inline float pdf_disney(const Mat& m, V3 v, V3 n, V3 l, V3 tang, V3 bitang) {
  float dw, sw, cw;
  lobe_probabilities(m, dw, sw, cw);
  float pdf = 0.0f;
  pdf += pdf_diffuse(n, l) * dw;
  pdf += pdf_specular(m, v, n, l, tang, bitang) * sw;
  pdf += pdf_clearcoat(m, v, n, l) * cw;
  return pdf;
}

// bsdf.py:201-224
static inline V3 sample_clearcoat(const Mat& m, V3 v, V3 n, V3 tang, V3 bitang, float ux, float uy, float& pdf) {
  float alpha = mixf(0.1f, 0.001f, m.clearcoat_gloss);
  float a2 = sqr(alpha);
  float cosTheta = std::sqrt(fmaxf_(1e-4f, (1.0f - std::pow(a2, 1.0f - ux)) / (1.0f - a2)));
  float sinTheta = std::sqrt(fmaxf_(1e-4f, 1.0f - cosTheta * cosTheta));
  float phi = 2.0f * kPi * uy;
  V3 mm{sinTheta * std::cos(phi), cosTheta, sinTheta * std::sin(phi)};
  V3 h = mm.x * tang + mm.z * bitang + mm.y * n;
  if (dot(h, v) < 0.0f) h *= -1.0f;
  V3 dir = reflect(-v, h);
  float n_dot_h = std::fabs(dot(n, h));
  float v_dot_h = dot(v, h);
  float D = GTR1(n_dot_h, alpha);
  pdf = D * n_dot_h / (4.0f * v_dot_h);
  return dir;
}

// bsdf.py:226-252 — VNDF sampling in the (tangent, normal, bitangent) frame
static inline V3 GGX_VNDF_aniso(V3 v, V3 n, V3 tang, V3 bitang, float ax, float ay, float ux, float uy) {
  V3 v_t{dot(tang, v), dot(n, v), dot(bitang, v)};
  V3 V = normalize(V3{v_t.x * ax, v_t.y, v_t.z * ay});
  V3 t1 = V.y < 0.9999f ? normalize(cross(V, V3{0, 1, 0})) : V3{1, 0, 0};
  V3 t2 = cross(t1, V);
  float a = 1.0f / (1.0f + V.y);
  float r = std::sqrt(ux);
  float phi = uy < a ? (uy / a) * kPi : kPi + (uy - a) / (1.0f - a) * kPi;
  float p1 = r * std::cos(phi);
  float p2 = r * std::sin(phi) * (uy < a ? 1.0f : V.y);
  V3 mm = p1 * t1 + p2 * t2 + std::sqrt(fmaxf_(0.0f, 1.0f - p1 * p1 - p2 * p2)) * V;
  mm = normalize(V3{ax * mm.x, mm.y, ay * mm.z});
  V3 h = mm.x * tang + mm.z * bitang + mm.y * n;
  if (dot(h, v) < 0.0f) h *= -1.0f;
  return h;
}
// bsdf.py:279-304
static inline V3 sample_specular(const Mat& m, V3 v, V3 n, V3 tang, V3 bitang, float ux, float uy, float& pdf) {
  float ax, ay;
  aniso_alphas(m, ax, ay);
  V3 h = GGX_VNDF_aniso(v, n, tang, bitang, ax, ay, ux, uy);
  V3 dir = reflect(-v, h);
  float n_dot_l = std::fabs(dot(n, dir));
  float n_dot_v = dot(n, v);
  float l_dot_h = std::fabs(dot(dir, h));
  float n_dot_h = dot(n, h);
  float h_dot_x = dot(h, tang), h_dot_y = dot(h, bitang);
  float v_dot_x = dot(v, tang), v_dot_y = dot(v, bitang);
  float D = GTR2_anisotropic(n_dot_h, h_dot_x, h_dot_y, ax, ay);
  float G = smithG_GGX_aniso(n_dot_v, v_dot_x, v_dot_y, ax, ay);
  pdf = G * l_dot_h * D / n_dot_l;
  return dir;
}

// bsdf.py:395-458 sample_disney. u_lobe, ux, uy = the three ti.random() draws (:405 then the sampler's two).
static inline V3 sample_disney(const Mat& m, V3 v, V3 n, V3 tang, V3 bitang, float u_lobe, float ux, float uy, V3& brdf,
                               float& pdf, int& lobe) {
  float dw, sw, cw;
  lobe_probabilities(m, dw, sw, cw);
  V3 dir{1, 1, 1};
  brdf = V3{0, 0, 0};
  pdf = 1.0f;
  if (u_lobe <= dw) {
    dir = sample_cosine_weighted_hemisphere(n, ux, uy);
    pdf = saturate(dot(dir, n)) / kPi;
    lobe = LOBE_DIFFUSE;
  } else if (u_lobe <= dw + sw) {
    dir = sample_specular(m, v, n, tang, bitang, ux, uy, pdf);
    lobe = LOBE_SPEC_REFL;
  } else {
    dir = sample_clearcoat(m, v, n, tang, bitang, ux, uy, pdf);
    lobe = LOBE_CLEARC;
  }
  float n_dot_l = dot(n, dir);
  float n_dot_v = dot(n, v);
  V3 h = normalize(dir + v);
  float l_dot_h = dot(dir, h), n_dot_h = dot(n, h);
  float h_dot_x = dot(h, tang), h_dot_y = dot(h, bitang);
  float l_dot_x = dot(dir, tang), l_dot_y = dot(dir, bitang);
  float v_dot_x = dot(v, tang), v_dot_y = dot(v, bitang);
  if (lobe == LOBE_DIFFUSE) {
    brdf += disney_diffuse(m, n_dot_l, n_dot_v, l_dot_h) * (1.0f - m.metallic);
    pdf *= dw;
  } else if (lobe == LOBE_SPEC_REFL) {
    brdf += disney_specular(m, n_dot_l, n_dot_v, l_dot_h, n_dot_h, h_dot_x, h_dot_y, l_dot_x, l_dot_y, v_dot_x, v_dot_y);
    pdf *= sw;
  } else {
    brdf += v3(disney_clearcoat(m, n_dot_l, n_dot_v, n_dot_h, l_dot_h));
    pdf *= cw;
  }
  if (isbad(pdf)) pdf = 1.0f;
  return dir;
}

}  // namespace orc
