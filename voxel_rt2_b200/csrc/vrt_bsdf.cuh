// Disney BSDF evaluation / pdf / sampling for the fused bounce-shade-NEE kernel.
// Replaces the live part of renderer/bsdf.py (:22-458) and the sampling helpers of
// renderer/math_utils.py:21-63. The estimator must converge to the reference's image, so the
// samplers and the pdf *expressions* are the reference's (including the ones that are not the
// exact density of their sampler); only the arithmetic is reorganised: x^5 by
// multiplication, shared half-vector terms, per-material constants hoisted.
#pragma once
#include "vrt_common.cuh"

#define VRT_COLD static __device__ __noinline__

enum { LOBE_DIFFUSE = 0, LOBE_SPEC_REFL = 1, LOBE_CLEARC = 2 };

HD Mat load_mat(const float4* __restrict__ mats, int id) {
  id = min(max(id, 0), 127);
  const float4* r = mats + id * MAT_ROW_F4;
  float4 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4];
  Mat m;
  m.base_col = f3{a.x, a.y, a.z};
  m.subsurface = a.w;
  m.metallic = b.x, m.specular = b.y, m.specular_tint = b.z, m.roughness = b.w;
  m.sheen = c.x, m.sheen_tint = c.y, m.clearcoat = c.z, m.cc_alpha = c.w;
  m.dw = d.x, m.sw = d.y, m.cw = d.z, m.cc_norm = d.w;
  m.ax = e.x, m.ay = e.y, m.inv_pi_axay = e.z;
  return m;
}

// math_utils.py:32-42
HD void make_orthonormal_basis(f3 n, f3& x, f3& y) {
  f3 h = fabsf(n.y) > 0.9f ? f3{1.0f, 0.0f, 0.0f} : f3{0.0f, 1.0f, 0.0f};
  y = normalize(cross(n, h));
  x = cross(n, y);
}

// math_utils.py:21-30
HD f3 sample_cosine_weighted_hemisphere(f3 n, float u0, float u1) {
  float a = 1.0f - 2.0f * u0;
  float b = fsqrt(1.0f - a * a);
  a *= 1.0f - 1e-5f;
  b *= 1.0f - 1e-5f;
  float s, c;
  __sincosf(2.0f * VRT_PI * u1, &s, &c);
  return normalize(f3{n.x + b * c, n.y + b * s, n.z + a});
}

// math_utils.py:44-59 (basis supplied by the caller: the sun axis is a per-launch constant)
HD f3 sample_cone_oriented(float cos_theta_max, f3 n, f3 bx, f3 by, float u0, float u1) {
  float cos_theta = (1.0f - u0) + u0 * cos_theta_max;
  float sin_theta = fsqrt(1.0f - cos_theta * cos_theta);
  float s, c;
  __sincosf(2.0f * VRT_PI * u1, &s, &c);
  float sx = sin_theta * c, sy = sin_theta * s, sz = cos_theta;
  return f3{(bx.x * sx + by.x * sy) + n.x * sz, (bx.y * sx + by.y * sy) + n.y * sz, (bx.z * sx + by.z * sy) + n.z * sz};
}
// math_utils.py:61-63
HD float cone_sample_pdf(float cos_theta_max, float cos_theta) {
  return cos_theta >= cos_theta_max ? 1.0f / (2.0f * VRT_PI * (1.0f - cos_theta_max)) : 0.0f;
}

// Sheen and subsurface terms of disney_diffuse (bsdf.py:56-67). Kept out of line (VRT_COLD): most
// material rows have neither, and the path kernel is instruction-fetch bound — code that is rarely
// executed must not sit inside its main loop.
VRT_COLD f3 diffuse_sheen_subsurface(f3 base_col, float sheen_w, float sheen_tint, float subsurface, float roughness, f3 f_d, float n_dot_l,
                                     float n_dot_v, float l_dot_h, float F_L, float F_V) {
  f3 sheen = mk3(0.0f);
  if (sheen_w != 0.0f) {
    float albedo_lum = luminance(base_col);
    f3 sheen_col = albedo_lum > 0.0f ? base_col / albedo_lum : mk3(1.0f);
    sheen = sheen_w * mix3(mk3(1.0f), sheen_col, sheen_tint) * pow5(1.0f - l_dot_h);
  }
  if (subsurface != 0.0f) {
    float Fss90 = l_dot_h * l_dot_h * roughness;
    float Fss = mixf(1.0f, Fss90, F_L) * mixf(1.0f, Fss90, F_V);
    float ss = 1.25f * (Fss * (frcp(n_dot_l + n_dot_v) - 0.5f) + 0.5f);
    f3 sub = ((1.0f / VRT_PI) * ss) * base_col;
    f_d = mix3(f_d, sub, subsurface);
  }
  return f_d + sheen;
}

// bsdf.py:39-67 diffuse + retro-reflection + sheen + subsurface
HD f3 disney_diffuse(const Mat& m, float n_dot_l, float n_dot_v, float l_dot_h) {
  float R_R = 2.0f * m.roughness * sqr(l_dot_h);
  float F_L = pow5(1.0f - n_dot_l);
  float F_V = pow5(1.0f - n_dot_v);
  f3 f_lambert = m.base_col * (1.0f / VRT_PI);
  f3 f_retro = f_lambert * (R_R * (F_L + F_V + F_L * F_V * (R_R - 1.0f)));
  f3 f_d = f_lambert * ((1.0f - 0.5f * F_L) * (1.0f - 0.5f * F_V)) + f_retro;
  if (m.sheen != 0.0f || m.subsurface != 0.0f)
    return diffuse_sheen_subsurface(m.base_col, m.sheen, m.sheen_tint, m.subsurface, m.roughness, f_d, n_dot_l, n_dot_v, l_dot_h, F_L, F_V);
  return f_d;
}

// bsdf.py:69-71: 1 / (pi ax ay (hx^2/ax^2 + hy^2/ay^2 + nh^2)^2); 1/(pi ax ay) comes from the table
HD float GTR2_anisotropic(const Mat& m, float n_dot_h, float h_dot_x, float h_dot_y) {
  float q = sqr(fdiv(h_dot_x, m.ax)) + sqr(fdiv(h_dot_y, m.ay)) + sqr(n_dot_h);
  return m.inv_pi_axay * frcp(q * q);
}
HD float smithG_GGX_aniso(float n_dot_v, float v_dot_x, float v_dot_y, float ax, float ay) {  // bsdf.py:73-75
  return frcp(n_dot_v + fsqrt(sqr(v_dot_x * ax) + sqr(v_dot_y * ay) + sqr(n_dot_v)));
}
HD f3 disney_fresnel(const Mat& m, float l_dot_h) {  // bsdf.py:77-83
  f3 tint = mk3(1.0f);
  if (m.specular_tint != 0.0f) {  // mix(1, spec_tint, 0) == 1 otherwise
    float albedo_lum = luminance(m.base_col);
    f3 spec_tint = albedo_lum > 0.0f ? m.base_col / albedo_lum : mk3(1.0f);
    tint = mix3(mk3(1.0f), spec_tint, m.specular_tint);
  }
  f3 spec_col = mix3((m.specular * 0.08f) * tint, m.base_col, m.metallic);
  return mix3(spec_col, mk3(1.0f), pow5(1.0f - l_dot_h));
}
// bsdf.py:112-121 with (a2-1)/(pi log a2) hoisted into the material row (cc_norm; 1/pi if alpha >= 1)
HD float GTR1(const Mat& m, float n_dot_h) {
  float a2 = m.cc_alpha * m.cc_alpha;
  float t = 1.0f + (a2 - 1.0f) * n_dot_h * n_dot_h;
  return m.cc_alpha >= 1.0f ? m.cc_norm : fdiv(m.cc_norm, t);
}
HD float smithG_GGX(float n_dot_v, float alpha) {  // bsdf.py:123-127
  float a2 = alpha * alpha;
  float b = n_dot_v * n_dot_v;
  return frcp(n_dot_v + fsqrt(a2 + b - a2 * b));
}
HD float disney_clearcoat(const Mat& m, float n_dot_l, float n_dot_v, float n_dot_h, float l_dot_h) {  // :129-135
  float D = GTR1(m, fabsf(n_dot_h));
  float F = mixf(0.04f, 1.0f, pow5(1.0f - l_dot_h));
  float G = smithG_GGX(n_dot_l, 0.25f) * smithG_GGX(n_dot_v, 0.25f);
  return m.clearcoat * D * F * G;
}

// Clear-coat code of the path kernel, out of line (VRT_COLD, see diffuse_sheen_subsurface): only
// material rows 21, 22, 32 and 54 of the default set have a clear coat. Arguments by value so the
// caller's material stays in registers.
struct CoatParams {
  float clearcoat, cc_alpha, cc_norm, cw;
};
HD CoatParams coat_of(const Mat& m) { return CoatParams{m.clearcoat, m.cc_alpha, m.cc_norm, m.cw}; }
HD Mat coat_mat(CoatParams c) {
  Mat m{};
  m.clearcoat = c.clearcoat, m.cc_alpha = c.cc_alpha, m.cc_norm = c.cc_norm, m.cw = c.cw;
  return m;
}
VRT_COLD float cold_clearcoat_eval(CoatParams c, float n_dot_l, float n_dot_v, float n_dot_h, float l_dot_h) {
  return disney_clearcoat(coat_mat(c), n_dot_l, n_dot_v, n_dot_h, l_dot_h);
}
VRT_COLD float cold_clearcoat_pdf(CoatParams c, float n_dot_h, float v_dot_h) {  // bsdf.py:190-199
  float ndh = fabsf(n_dot_h);
  return fdiv(GTR1(coat_mat(c), ndh) * ndh, 4.0f * v_dot_h) * c.cw;
}
// bsdf.py:201-224 sample_clearcoat: returns (direction, lobe pdf x lobe probability)
VRT_COLD float4 cold_clearcoat_sample(CoatParams c, f3 v, f3 n, f3 tang, f3 bitang, float ux, float uy) {
  const Mat m = coat_mat(c);
  float a2 = sqr(m.cc_alpha);
  float cosTheta = fsqrt(fmaxf(1e-4f, fdiv(1.0f - __powf(a2, 1.0f - ux), 1.0f - a2)));
  float sinTheta = fsqrt(fmaxf(1e-4f, 1.0f - cosTheta * cosTheta));
  float s, co;
  __sincosf(2.0f * VRT_PI * uy, &s, &co);
  f3 h = (sinTheta * co) * tang + (sinTheta * s) * bitang + cosTheta * n;
  if (dot(h, v) < 0.0f) h *= -1.0f;
  f3 dir = reflect(-v, h);
  float ndh = fabsf(dot(n, h));
  float pdf = fdiv(GTR1(m, ndh) * ndh, 4.0f * dot(v, h)) * m.cw;
  return make_float4(dir.x, dir.y, dir.z, pdf);
}
// brdf of the sampled clear-coat direction, half vector recomputed from (dir, v) as the reference does
VRT_COLD float cold_clearcoat_brdf(CoatParams c, f3 v, f3 n, f3 dir) {
  f3 h = normalize(dir + v);
  return disney_clearcoat(coat_mat(c), dot(n, dir), dot(n, v), dot(n, h), dot(dir, h));
}

// Shared dot products of one (v, n, l) configuration.
struct Geo {
  float n_dot_l, n_dot_v, l_dot_h, n_dot_h, h_dot_x, h_dot_y, l_dot_x, l_dot_y, v_dot_x, v_dot_y, v_dot_h;
};
HD Geo make_geo(f3 v, f3 n, f3 l, f3 tang, f3 bitang) {
  Geo g;
  f3 h = normalize(l + v);
  g.n_dot_l = dot(n, l), g.n_dot_v = dot(n, v);
  g.l_dot_h = dot(l, h), g.n_dot_h = dot(n, h), g.v_dot_h = dot(v, h);
  g.h_dot_x = dot(h, tang), g.h_dot_y = dot(h, bitang);
  g.l_dot_x = dot(l, tang), g.l_dot_y = dot(l, bitang);
  g.v_dot_x = dot(v, tang), g.v_dot_y = dot(v, bitang);
  return g;
}
HD f3 disney_specular(const Mat& m, const Geo& g) {  // bsdf.py:86-105
  float D = GTR2_anisotropic(m, g.n_dot_h, g.h_dot_x, g.h_dot_y);
  float G = smithG_GGX_aniso(g.n_dot_l, g.l_dot_x, g.l_dot_y, m.ax, m.ay) * smithG_GGX_aniso(g.n_dot_v, g.v_dot_x, g.v_dot_y, m.ax, m.ay);
  return (D * G) * disney_fresnel(m, g.l_dot_h);
}

// bsdf.py:138-177 disney_evaluate_split and :382-393 pdf_disney for the same (v,n,l): the NEE
// sample needs both, and they share the half vector and the GGX D term.
HD void eval_and_pdf(const Mat& m, f3 v, f3 n, f3 l, f3 tang, f3 bitang, f3& bsdf_d, f3& bsdf_s, float& pdf) {
  Geo g = make_geo(v, n, l, tang, bitang);
  bsdf_d = mk3(0.0f);
  bsdf_s = mk3(0.0f);
  float D = GTR2_anisotropic(m, g.n_dot_h, g.h_dot_x, g.h_dot_y);
  float Gv = smithG_GGX_aniso(g.n_dot_v, g.v_dot_x, g.v_dot_y, m.ax, m.ay);
  if (g.n_dot_l > 0.0f && g.n_dot_v > 0.0f) {
    bsdf_d = disney_diffuse(m, g.n_dot_l, g.n_dot_v, g.l_dot_h) * (1.0f - m.metallic);
    float Gl = smithG_GGX_aniso(g.n_dot_l, g.l_dot_x, g.l_dot_y, m.ax, m.ay);
    bsdf_s = (D * (Gl * Gv)) * disney_fresnel(m, g.l_dot_h);
    if (m.clearcoat != 0.0f) bsdf_s += mk3(cold_clearcoat_eval(coat_of(m), g.n_dot_l, g.n_dot_v, g.n_dot_h, g.l_dot_h));
  }
  // pdf_disney = sum of lobe pdfs times lobe probabilities
  pdf = (saturate(g.n_dot_l) * (1.0f / VRT_PI)) * m.dw;
  pdf += fdiv(Gv * fabsf(g.l_dot_h) * D, fabsf(g.n_dot_l)) * m.sw;  // bsdf.py:254-277
  if (m.cw != 0.0f) pdf += cold_clearcoat_pdf(coat_of(m), g.n_dot_h, g.v_dot_h);  // bsdf.py:190-199 (cw == 0 adds 0)
}

// bsdf.py:226-252 VNDF sampling in the (tangent, normal, bitangent) frame
HD f3 GGX_VNDF_aniso(f3 v, f3 n, f3 tang, f3 bitang, float ax, float ay, float ux, float uy) {
  f3 v_t{dot(tang, v), dot(n, v), dot(bitang, v)};
  f3 V = normalize(f3{v_t.x * ax, v_t.y, v_t.z * ay});
  f3 t1 = V.y < 0.9999f ? normalize(cross(V, f3{0.0f, 1.0f, 0.0f})) : f3{1.0f, 0.0f, 0.0f};
  f3 t2 = cross(t1, V);
  float a = frcp(1.0f + V.y);
  float r = fsqrt(ux);
  float phi = uy < a ? fdiv(uy, a) * VRT_PI : VRT_PI + fdiv(uy - a, 1.0f - a) * VRT_PI;
  float s, c;
  __sincosf(phi, &s, &c);
  float p1 = r * c;
  float p2 = r * s * (uy < a ? 1.0f : V.y);
  f3 mm = p1 * t1 + p2 * t2 + fsqrt(fmaxf(0.0f, 1.0f - p1 * p1 - p2 * p2)) * V;
  mm = normalize(f3{ax * mm.x, mm.y, ay * mm.z});
  f3 h = mm.x * tang + mm.z * bitang + mm.y * n;
  if (dot(h, v) < 0.0f) h *= -1.0f;
  return h;
}

// The specular-lobe sampler of sample_disney, out of line by default (VRT_COLD_SPEC): on the dense benchmark scene
// 2-3 lanes of a warp take it per shading stage (200 SASS instructions in the middle of the main loop otherwise), and
// with 24 warps per SM the loop is instruction-fetch sensitive: +2.0 % on config 3 (profiles/r03m_ab_coldspec.log).
// Results move in the last ulp (no FMA contraction across the call), inside every radiance tolerance.
struct SpecParams {
  f3 base_col;
  float metallic, specular, specular_tint, ax, ay, inv_pi_axay, sw;
};
struct SpecSample {
  f3 dir, brdf;
  float pdf;
};
HD Mat spec_mat(const SpecParams& p) {
  Mat m{};
  m.base_col = p.base_col, m.metallic = p.metallic, m.specular = p.specular, m.specular_tint = p.specular_tint;
  m.ax = p.ax, m.ay = p.ay, m.inv_pi_axay = p.inv_pi_axay, m.sw = p.sw;
  return m;
}
HD SpecSample spec_sample_body(const Mat& m, f3 v, f3 n, f3 tang, f3 bitang, float ux, float uy) {
  SpecSample r;
  f3 h = GGX_VNDF_aniso(v, n, tang, bitang, m.ax, m.ay, ux, uy);
  r.dir = reflect(-v, h);
  // pdf with the sampled micro-normal (bsdf.py:290-302)
  float D = GTR2_anisotropic(m, dot(n, h), dot(h, tang), dot(h, bitang));
  float Gv = smithG_GGX_aniso(dot(n, v), dot(v, tang), dot(v, bitang), m.ax, m.ay);
  r.pdf = fdiv(Gv * fabsf(dot(r.dir, h)) * D, fabsf(dot(n, r.dir))) * m.sw;
  // brdf with the half vector recomputed from (dir, v) as the reference does (:430-450)
  Geo g = make_geo(v, n, r.dir, tang, bitang);
  r.brdf = disney_specular(m, g);
  return r;
}
#ifndef VRT_COLD_SPEC
#define VRT_COLD_SPEC 1
#endif
#if VRT_COLD_SPEC
VRT_COLD SpecSample cold_spec_sample(SpecParams p, f3 v, f3 n, f3 tang, f3 bitang, float ux, float uy) {
  return spec_sample_body(spec_mat(p), v, n, tang, bitang, ux, uy);
}
#endif

// bsdf.py:395-458 sample_disney: returns direction, brdf of the chosen lobe, pdf (lobe pdf x
// lobe probability; inf/NaN -> 1) and the lobe id.
HD f3 sample_disney(const Mat& m, f3 v, f3 n, f3 tang, f3 bitang, float u_lobe, float ux, float uy, f3& brdf, float& pdf, int& lobe) {
  f3 dir;
  if (u_lobe <= m.dw) {
    dir = sample_cosine_weighted_hemisphere(n, ux, uy);
    float n_dot_l = dot(n, dir);
    pdf = (saturate(n_dot_l) * (1.0f / VRT_PI)) * m.dw;
    lobe = LOBE_DIFFUSE;
    f3 h = normalize(dir + v);
    brdf = disney_diffuse(m, n_dot_l, dot(n, v), dot(dir, h)) * (1.0f - m.metallic);
  } else if (u_lobe <= m.dw + m.sw) {
#if VRT_COLD_SPEC
    const SpecSample r = cold_spec_sample(SpecParams{m.base_col, m.metallic, m.specular, m.specular_tint, m.ax, m.ay, m.inv_pi_axay, m.sw}, v, n, tang,
                                          bitang, ux, uy);
#else
    const SpecSample r = spec_sample_body(m, v, n, tang, bitang, ux, uy);
#endif
    dir = r.dir, pdf = r.pdf, brdf = r.brdf;
    lobe = LOBE_SPEC_REFL;
  } else {
    const float4 r = cold_clearcoat_sample(coat_of(m), v, n, tang, bitang, ux, uy);
    dir = f3{r.x, r.y, r.z};
    pdf = r.w;
    lobe = LOBE_CLEARC;
    brdf = mk3(cold_clearcoat_brdf(coat_of(m), v, n, dir));
  }
  if (isbad(pdf)) pdf = 1.0f;
  return dir;
}

// ---- lobe-wise variants used by the ReSTIR reconnection shift (pathtracer.py:672-812)
enum { LOBE_ALL = 9 };

// bsdf.py:306-349 disney_evaluate_lobewise_split (lobe LOBE_ALL == disney_evaluate_split)
HD void disney_evaluate_lobewise_split(const Mat& m, f3 v, f3 n, f3 l, f3 tang, f3 bitang, int lobe_id, f3& bsdf_d, f3& bsdf_s) {
  Geo g = make_geo(v, n, l, tang, bitang);
  bsdf_d = mk3(0.0f);
  bsdf_s = mk3(0.0f);
  if (g.n_dot_l > 0.0f && g.n_dot_v > 0.0f) {
    if (lobe_id == LOBE_DIFFUSE || lobe_id == LOBE_ALL) bsdf_d = disney_diffuse(m, g.n_dot_l, g.n_dot_v, g.l_dot_h) * (1.0f - m.metallic);
    if (lobe_id == LOBE_SPEC_REFL || lobe_id == LOBE_ALL) bsdf_s = disney_specular(m, g);
    if ((lobe_id == LOBE_CLEARC || lobe_id == LOBE_ALL) && m.clearcoat != 0.0f)
      bsdf_s += mk3(cold_clearcoat_eval(coat_of(m), g.n_dot_l, g.n_dot_v, g.n_dot_h, g.l_dot_h));
  }
}
HD f3 disney_evaluate_lobewise(const Mat& m, f3 v, f3 n, f3 l, f3 tang, f3 bitang, int lobe_id) {
  f3 d, s;
  disney_evaluate_lobewise_split(m, v, n, l, tang, bitang, lobe_id, d, s);
  return d + s;
}
// bsdf.py:365-380 pdf_disney_lobewise (inf/NaN -> 1)
HD float pdf_disney_lobewise(const Mat& m, f3 v, f3 n, f3 l, f3 tang, f3 bitang, int lobe_id) {
  Geo g = make_geo(v, n, l, tang, bitang);
  float pdf;
  if (lobe_id == LOBE_DIFFUSE) {
    pdf = (saturate(g.n_dot_l) * (1.0f / VRT_PI)) * m.dw;
  } else if (lobe_id == LOBE_SPEC_REFL) {
    float D = GTR2_anisotropic(m, g.n_dot_h, g.h_dot_x, g.h_dot_y);
    float Gv = smithG_GGX_aniso(g.n_dot_v, g.v_dot_x, g.v_dot_y, m.ax, m.ay);
    pdf = fdiv(Gv * fabsf(g.l_dot_h) * D, fabsf(g.n_dot_l)) * m.sw;
  } else {
    pdf = cold_clearcoat_pdf(coat_of(m), g.n_dot_h, g.v_dot_h);
  }
  if (isbad(pdf)) pdf = 1.0f;
  return pdf;
}
