// ORACLE — TEST INFRASTRUCTURE ONLY (never on the product path; the product fails loudly
// without its CUDA library). CPU restatement (C++17 + OpenMP, float32, -ffp-contract=off) of
// voxel-rt2's rendering hot path, exported with a flat C interface for ctypes:
//   render            renderer/pathtracer.py:331-632 (non-ReSTIR estimator)
//   accumulation      renderer/pathtracer.py:1185-1230,1242-1303 (static camera: running mean)
//   NaN scrub         renderer/pathtracer.py:1068-1075
//   tonemap           renderer/pathtracer.py:634-662, renderer/math_utils.py:160-186
//   materials         renderer/materials.py:49-112 (table supplied by the host)
// Traversal, BSDF and sky live in otrace.h / obsdf.h / osky.h.
//
// PARITY PINS: the reference ships no golden vectors / tests and Taichi cannot be installed here.
// The restatement is pinned by (1) vectors computed by the reference's own renderer/*.py source
// executed through oracle/ti_emu, a float32 Taichi emulator (tests/golden/make_ref_vectors.py ->
// tests/golden/ref_*.npz -> tests/test_reference_vectors.py: traversal and hit buffers bit-exact,
// render() per pixel to 3e-5, the static frame loop to 1.4e-6, BSDF to 1.5e-7, sky precompute, the
// moving-camera filter chain to 5e-3 worst pixel; the ReSTIR branch of render() to 7.5e-6, ReSTIR
// shift() / reservoir packing / spatial_GRIS bit-exact on inputs free of the zero-vector encodings
// upstream leaves undefined),
// (2) hand-derived known-answer tests and a brute-force traversal twin (tests/test_oracle_kat.py).
// Only the holes of the upstream ReSTIR mode (DESIGN.md "ReSTIR pins") are pinned by (2) alone.
#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "../include/voxelrt.h"
#include "obsdf.h"
#include "orestir.h"
#include "osky.h"
#include "otrace.h"

using namespace orc;

namespace {

struct Ctx {
  Scene scene;
  Sky sky;
  std::vector<Mat> mats;  // 128
  int max_depth = 4;
  float exposure = 3.0f;
  uint32_t seed = 0;
  int jitter_mode = 0;
  int cloud_passes = 32;
  std::vector<float> hist_d, hist_s;  // vec4 per pixel (history_buffer / history_buffer_specular)
  Counters counters;
  double last_ms = 0;
  int tile_rank = 0, tile_n = 1;
  // ReSTIR mode buffers (pathtracer.py:109-125): one packed reservoir + G-buffer + the frame's
  // diffuse / specular colour per pixel
  std::vector<StorageReservoir> reservoirs;
  std::vector<GBufferPx> gbuf;
  std::vector<V3> col_d, col_s;
  // temporal reservoir reuse (the second reservoir slot of pathtracer.py:108-109, written at :989 and never read
  // upstream): output reservoirs + G-buffer of the previous frame's spatial pass
  std::vector<StorageReservoir> hist_res;
  std::vector<GBufferPx> hist_gbuf;
  bool hist_valid = false;
  int restir_temporal = 0;
  // moving-camera temporal path (pathtracer.py:993-1303): G-buffer, previous G-buffer, two
  // history slots for diffuse / specular / reflection depth, previous matrices
  struct Moving {
    std::vector<V3> col_d, col_s, out;
    std::vector<float> depth, refl, refl_blur, prev_depth, noct, prev_noct, hd[2], hs[2], hsd[2];
    std::vector<uint32_t> mat;
    float prev_view[16], prev_proj[16];
    V3 prev_cam{0, 0, 0};
    bool has_prev = false, active = false;
    float scale = 1.0f;
  } mv;
};

static const float RADIANCE_CLAMP = 300.0f;  // pathtracer.py:20
static inline V3 firefly_filter(V3 v) { return clamp3(v, 0.0f, RADIANCE_CLAMP); }  // :22-24
static inline float power_heuristic(float a, float b) {                              // :349-353
  float a_sqr = a * a;
  float p_sum = fmaxf_(a_sqr + b * b, 1e-4f);
  return a_sqr / p_sum;
}
// math_utils.py:231-236
static inline uint32_t encode_material(int mat_id, V3 albedo) {
  uint32_t d0 = (uint32_t)mat_id, d1 = (uint32_t)(albedo.x * 255.0f), d2 = (uint32_t)(albedo.y * 255.0f),
           d3 = (uint32_t)(albedo.z * 255.0f);
  return d0 | (d1 << 8) | (d2 << 16) | (d3 << 24);
}

static inline const Mat& mat_at(const Ctx& c, int id) { return c.mats[id < 0 ? 0 : (id > 127 ? 127 : id)]; }

struct PathOut {
  V3 diffuse, specular;
};
struct GExtra {  // primary-vertex data the moving-camera temporal filters need (pathtracer.py:535-546)
  V3 primary_pos{0, 0, 0}, primary_normal{0, 0, 0}, primary_albedo{1, 1, 1};
  uint32_t mat_info = 0;
  float first_bounce_reflection_dist = 0.0f;
  bool sky = false;
};

// pathtracer.py:355-632 for one pixel sample, USE_RESTIR_PT = False, static camera.
static void trace_path(const Ctx& c, int u, int v, uint32_t sample, Counters* cnt, PathOut& out, GExtra* gx = nullptr,
                       float render_scale = 1.0f) {
  const Scene& s = c.scene;
  const uint32_t key = path_key((uint32_t)(v * s.W + u), sample, c.seed);
  V3 d = get_cast_dir(s, (float)u, (float)v, render_scale, gx != nullptr);
  V3 pos = s.cam_pos;
  V3 contrib{0, 0, 0}, throughput{1, 1, 1};
  uint32_t primary_mat_info = 0;
  int first_bounce_lobe_id = 0;
  float first_bounce_invpdf = 1.0f;
  V3 first_NEE_d{0, 0, 0}, first_NEE_s{0, 0, 0};
  float first_light_sample_bsdf_pdf = 1.0f;
  bool is_sky_ray = false;
  if (cnt) cnt->paths++;

  for (int depth = 0; depth < c.max_depth; depth++) {
    const uint32_t base = 8u * (uint32_t)depth;
    Hit h = next_hit(s, pos, d, kInf, false, cnt);
    Mat hit_mat = mat_at(c, h.mat_id);
    V3 hit_pos = pos + h.closest * d;
    if (depth == 0) primary_mat_info = encode_material(h.mat_id, h.albedo);
    if (gx) {
      if (depth == 0) {
        gx->primary_pos = hit_pos, gx->primary_normal = h.normal, gx->primary_albedo = h.albedo, gx->mat_info = primary_mat_info;
      } else if (depth == 1 && first_bounce_lobe_id != LOBE_DIFFUSE) {
        gx->first_bounce_reflection_dist += h.closest;  // pathtracer.py:410-412
      }
    }

    if (!h.hit_light && h.closest < kInf) {
      if (cnt) cnt->vertices++;
      V3 normal = h.normal;
      pos = hit_pos + normal * kEps;
      hit_mat.base_col = h.albedo;
      V3 view = -d;
      V3 tang, bitang;
      make_orthonormal_basis(normal, tang, bitang);
      float NEE_visible = 0.0f;
      {
        V3 light_dir = sample_cone_oriented(s.light_cos_max, s.light_dir, rnd(key, base + 0), rnd(key, base + 1));
        float ndl = dot(light_dir, normal);
        float light_sample_bsdf_pdf = pdf_disney(hit_mat, view, normal, light_dir, tang, bitang);
        if (depth == 0) first_light_sample_bsdf_pdf = light_sample_bsdf_pdf;
        if (ndl > 0.0f) {
          Hit sh = next_hit(s, pos, light_dir, kInf, true, cnt);
          if (sh.closest >= kInf) {
            NEE_visible = 1.0f;
            float mis = 1.0f;
            if (depth > 0) mis = power_heuristic(cone_sample_pdf(s.light_cos_max, 1.0f), light_sample_bsdf_pdf);
            V3 bd, bs;
            disney_evaluate_split(hit_mat, view, normal, light_dir, tang, bitang, bd, bs);
            V3 skyT{1, 1, 1};
            if (s.use_physical_sky == 1) {
              skyT = sample_skybox_transmittance(s.sky_trans, s.sky_res, light_dir);
              if (cnt) cnt->N++;
            }
            V3 nee_d = mis * bd * skyT * s.light_weight * s.light_color * ndl;
            V3 nee_s = mis * bs * skyT * s.light_weight * s.light_color * ndl;
            if (depth == 0) {
              first_NEE_d += firefly_filter(throughput * nee_d);
              first_NEE_s += firefly_filter(throughput * nee_s);
            } else {
              contrib += firefly_filter(throughput * (nee_d + nee_s));
            }
          }
        }
      }
      V3 bsdf;
      float pdf;
      int lobe_id;
      d = sample_disney(hit_mat, view, normal, tang, bitang, rnd(key, base + 2), rnd(key, base + 3), rnd(key, base + 4), bsdf,
                        pdf, lobe_id);
      V3 bounce_weight = bsdf * saturate(dot(d, normal));
      if (depth == 0) {
        first_bounce_invpdf = 1.0f / pdf;
        first_bounce_lobe_id = lobe_id;
      } else {
        bounce_weight = bounce_weight / pdf;
        float bsdf_sample_light_pdf = cone_sample_pdf(s.light_cos_max, dot(s.light_dir, d));
        bounce_weight *= power_heuristic(pdf, NEE_visible * bsdf_sample_light_pdf);
      }
      throughput *= bounce_weight;
    } else {
      if (h.closest == kInf) {
        float hit_sun = dot(s.light_dir, d) >= s.light_cos_max ? 1.0f : 0.0f;
        V3 sky_scattering = s.background;
        V3 sky_T{1, 1, 1};
        if (s.use_physical_sky == 1) {
          sample_skybox(s.sky_scatter, s.sky_trans, s.sky_res, d, rnd(key, base + 5), rnd(key, base + 6), rnd(key, base + 7),
                        sky_scattering, sky_T);
          if (cnt) cnt->E++;
        }
        V3 sky_emission = firefly_filter(sky_scattering + sky_T * s.light_weight * s.light_color * hit_sun);
        contrib += throughput * sky_emission;
        if (depth == 0) {
          is_sky_ray = true;
          if (gx) gx->primary_pos = V3{0, 0, 0}, gx->sky = true;
        }
      } else {
        if (depth > 0) contrib += throughput * h.albedo;
      }
      break;
    }
  }

  // primary-vertex MIS weight of the NEE sample (:556-579, non-ReSTIR branch)
  if (!is_sky_ray) {
    float light_sample_light_pdf = cone_sample_pdf(s.light_cos_max, 1.0f);
    float w = power_heuristic(light_sample_light_pdf, first_light_sample_bsdf_pdf);
    first_NEE_d *= w;
    first_NEE_s *= w;
  }
  // :609-619
  uint32_t pm = primary_mat_info & 255u;
  V3 emission{0, 0, 0};
  if (pm == 2u)
    emission = V3{(float)((primary_mat_info >> 8) & 255u) / 255.0f, (float)((primary_mat_info >> 16) & 255u) / 255.0f,
                  (float)((primary_mat_info >> 24) & 255u) / 255.0f};
  V3 diffuse{0, 0, 0}, specular{0, 0, 0};
  if (first_bounce_lobe_id == LOBE_DIFFUSE) diffuse += contrib * first_bounce_invpdf + emission;
  if (first_bounce_lobe_id == LOBE_SPEC_REFL) specular += contrib * first_bounce_invpdf;
  diffuse += first_NEE_d;
  specular += first_NEE_s;
  out.diffuse = diffuse;
  out.specular = specular;
}

// ------------------------------------------------------------------------------- ReSTIR mode
static inline void decode_material(const Ctx& c, uint32_t enc, Mat& m, int& mat_id) {  // math_utils.py:238-247
  mat_id = (int)(enc & 255u);
  m = mat_at(c, mat_id);
  m.base_col = V3{(float)((enc >> 8) & 255u) / 255.0f, (float)((enc >> 16) & 255u) / 255.0f, (float)((enc >> 24) & 255u) / 255.0f};
}

// pathtracer.py:355-632 with USE_RESTIR_PT = True: same walk as trace_path plus the reservoir /
// G-buffer bookkeeping. Random dimension 40 is the input_sample draw (reservoir.py:71).
static void trace_path_restir(const Ctx& c, int u, int v, uint32_t sample, Counters* cnt, Reservoir& res, GBufferPx& gb, V3& out_d,
                              V3& out_s) {
  const Scene& s = c.scene;
  const uint32_t key = path_key((uint32_t)(v * s.W + u), sample, c.seed);
  V3 d = get_cast_dir(s, (float)u, (float)v);
  V3 pos = s.cam_pos;
  V3 contrib{0, 0, 0}, throughput{1, 1, 1};
  res = Reservoir();
  V3 primary_normal{0, 0, 0}, primary_pos{0, 0, 0};
  uint32_t primary_mat_info = 0;
  V3 throughput_after_rc{1, 1, 1};
  int first_bounce_lobe_id = 0, rc_bounce_lobe_id = 0;
  float first_bounce_invpdf = 1.0f;
  V3 first_NEE_d{0, 0, 0}, first_NEE_s{0, 0, 0}, first_bounce_dir{0, 0, 0}, first_light_sample_dir{0, 0, 0};
  float first_light_sample_bsdf_pdf = 1.0f;
  bool is_sky_ray = false;
  if (cnt) cnt->paths++;

  for (int depth = 0; depth < c.max_depth; depth++) {
    const uint32_t base = 8u * (uint32_t)depth;
    Hit h = next_hit(s, pos, d, kInf, false, cnt);
    Mat hit_mat = mat_at(c, h.mat_id);
    V3 hit_pos = pos + h.closest * d;
    if (depth == 0) {
      primary_normal = h.normal;
      primary_pos = hit_pos;
      primary_mat_info = encode_material(h.mat_id, h.albedo);
    } else if (depth == 1) {
      res.z.rc_pos = hit_pos;
      res.z.rc_normal = h.normal;
      res.z.rc_mat_info = encode_material(h.mat_id, h.albedo);
      first_bounce_dir = d;
    } else if (depth == 2) {
      res.z.rc_incident_dir = d;
    }
    if (!h.hit_light && h.closest < kInf) {
      if (cnt) cnt->vertices++;
      V3 normal = h.normal;
      pos = hit_pos + normal * kEps;
      hit_mat.base_col = h.albedo;
      V3 view = -d;
      V3 tang, bitang;
      make_orthonormal_basis(normal, tang, bitang);
      float NEE_visible = 0.0f;
      {
        V3 light_dir = sample_cone_oriented(s.light_cos_max, s.light_dir, rnd(key, base + 0), rnd(key, base + 1));
        float ndl = dot(light_dir, normal);
        float light_sample_bsdf_pdf = pdf_disney(hit_mat, view, normal, light_dir, tang, bitang);
        if (depth == 0) {
          first_light_sample_bsdf_pdf = light_sample_bsdf_pdf;
          first_light_sample_dir = light_dir;
        }
        if (ndl > 0.0f) {
          Hit sh = next_hit(s, pos, light_dir, kInf, true, cnt);
          if (sh.closest >= kInf) {
            NEE_visible = 1.0f;
            if (depth == 1) res.z.rc_NEE_dir = light_dir;
            float mis = 1.0f;
            if (depth > 0) mis = power_heuristic(cone_sample_pdf(s.light_cos_max, 1.0f), light_sample_bsdf_pdf);
            V3 bd, bs;
            disney_evaluate_split(hit_mat, view, normal, light_dir, tang, bitang, bd, bs);
            V3 skyT{1, 1, 1};
            if (s.use_physical_sky == 1) {
              skyT = sample_skybox_transmittance(s.sky_trans, s.sky_res, light_dir);
              if (cnt) cnt->N++;
            }
            V3 nee_d = mis * bd * skyT * s.light_weight * s.light_color * ndl;
            V3 nee_s = mis * bs * skyT * s.light_weight * s.light_color * ndl;
            if (depth == 0) {
              first_NEE_d += firefly_filter(throughput * nee_d);
              first_NEE_s += firefly_filter(throughput * nee_s);
            } else {
              contrib += firefly_filter(throughput * (nee_d + nee_s));
            }
            if (depth >= 2) res.z.rc_incident_L += throughput_after_rc * (nee_d + nee_s);
          }
        }
      }
      V3 bsdf;
      float pdf;
      int lobe_id;
      d = sample_disney(hit_mat, view, normal, tang, bitang, rnd(key, base + 2), rnd(key, base + 3), rnd(key, base + 4), bsdf, pdf, lobe_id);
      V3 bounce_weight = bsdf * saturate(dot(d, normal));
      if (depth == 0) {
        first_bounce_invpdf = 1.0f / pdf;
        first_bounce_lobe_id = lobe_id;
      } else {
        bounce_weight = bounce_weight / pdf;
        float bsdf_sample_light_pdf = cone_sample_pdf(s.light_cos_max, dot(s.light_dir, d));
        bounce_weight *= power_heuristic(pdf, NEE_visible * bsdf_sample_light_pdf);
        if (depth == 1) rc_bounce_lobe_id = lobe_id;
        if (depth >= 2) throughput_after_rc *= bounce_weight;
      }
      throughput *= bounce_weight;
    } else {
      if (h.closest == kInf) {
        float hit_sun = dot(s.light_dir, d) >= s.light_cos_max ? 1.0f : 0.0f;
        V3 sky_scattering = s.background;
        V3 sky_T{1, 1, 1};
        if (s.use_physical_sky == 1) {
          sample_skybox(s.sky_scatter, s.sky_trans, s.sky_res, d, rnd(key, base + 5), rnd(key, base + 6), rnd(key, base + 7), sky_scattering,
                        sky_T);
          if (cnt) cnt->E++;
        }
        V3 sky_emission = firefly_filter(sky_scattering + sky_T * s.light_weight * s.light_color * hit_sun);
        contrib += throughput * sky_emission;
        if (depth == 0) {
          primary_pos = V3{0, 0, 0};
          is_sky_ray = true;
        } else if (depth == 1) {
          res.z.rc_pos = d;
          res.z.rc_incident_L = sky_emission;
        }
        if (depth >= 2) res.z.rc_incident_L += firefly_filter(throughput_after_rc * sky_emission);
      } else {
        if (depth > 0) contrib += throughput * h.albedo;
        if (depth >= 2) res.z.rc_incident_L += firefly_filter(throughput_after_rc * h.albedo);
      }
      break;
    }
  }
  // G-buffer (:535-540)
  gb.position = primary_pos;
  gb.mat_info = primary_mat_info;
  gb.sky = is_sky_ray ? 1 : 0;
  gb.n_oct[0] = gb.n_oct[1] = 0.0f;
  if (!is_sky_ray) encode_unit_vector_3x16(primary_normal, gb.n_oct[0], gb.n_oct[1]);
  // reservoir (:548-607)
  res.z.F = contrib;
  res.z.lobes = rc_bounce_lobe_id * 10 + first_bounce_lobe_id;
  res.M = 1.0f;
  res.update_cached_jacobian_term(primary_pos);
  bool chose_NEE = false;
  if (!is_sky_ray) {
    float bsdf_sample_bsdf_pdf = 1.0f / first_bounce_invpdf;
    float bsdf_sample_light_pdf = cone_sample_pdf(s.light_cos_max, dot(s.light_dir, first_bounce_dir));
    if (is_vec_zero(first_NEE_d + first_NEE_s)) bsdf_sample_light_pdf = 0.0f;
    float bsdf_sample_mis_weight = power_heuristic(bsdf_sample_bsdf_pdf, bsdf_sample_light_pdf);
    float light_sample_mis_weight = power_heuristic(cone_sample_pdf(s.light_cos_max, 1.0f), first_light_sample_bsdf_pdf);
    float p_hat = luminance(res.z.F);
    res.weight = bsdf_sample_mis_weight * p_hat * first_bounce_invpdf;
    float light_sample_weight = light_sample_mis_weight * luminance(first_NEE_d + first_NEE_s);
    V3 skyT = s.use_physical_sky == 1 ? sample_skybox_transmittance(s.sky_trans, s.sky_res, first_light_sample_dir) : V3{1, 1, 1};
    Sample ls;
    ls.F = first_NEE_d + first_NEE_s;
    ls.rc_pos = first_light_sample_dir;
    ls.rc_incident_L = skyT * s.light_weight * s.light_color;
    ls.cached_jacobian_term = 1.0f;
    ls.lobes = LOBE_ALL * 10 + LOBE_ALL;
    chose_NEE = res.input_sample(light_sample_weight, ls, rnd(key, 40));
    res.finalize_without_M();
  } else {
    res.weight = 1.0f;
  }
  out_d = V3{0, 0, 0};
  out_s = V3{0, 0, 0};
  if (!chose_NEE) {
    if (first_bounce_lobe_id == LOBE_DIFFUSE) out_d += res.z.F;
    if (first_bounce_lobe_id == LOBE_SPEC_REFL) out_s += res.z.F;
  } else {
    out_d += first_NEE_d;
    out_s += first_NEE_s;
  }
}

// pathtracer.py:672-812 shift(): integrand of src_reservoir's sample reconnected at dst, and the
// Jacobian of the shift.
static void shift_sample(const Ctx& c, V3 dst_pos, V3 dst_normal, const Mat& dst_material, V3 src_pos, const Reservoir& src, V3& diffuse,
                         V3& specular, float& jacobian_out) {
  const Scene& s = c.scene;
  const Sample& z = src.z;
  const bool rc_is_escape_vertex = is_vec_zero(z.rc_normal);
  const bool rc_is_last_vertex = is_vec_zero(z.rc_incident_dir);
  const bool rc_is_NEE_visible = !is_vec_zero(z.rc_NEE_dir);
  V3 dir_to_rc_vertex = rc_is_escape_vertex ? z.rc_pos : normalize(z.rc_pos - dst_pos);
  V3 src_dir_to_rc_vertex = rc_is_escape_vertex ? z.rc_pos : normalize(z.rc_pos - src_pos);
  float passed_checks = 1.0f;
  if (dot(dst_normal, dir_to_rc_vertex) < 1e-5f || (!rc_is_escape_vertex && dot(z.rc_normal, -dir_to_rc_vertex) < 1e-5f)) passed_checks = 0.0f;
  V3 rc_tang, rc_bitang;
  make_orthonormal_basis(z.rc_normal, rc_tang, rc_bitang);
  Mat rc_mat;
  int rc_mat_id;
  decode_material(c, z.rc_mat_info, rc_mat, rc_mat_id);
  V3 rc_brdf{0, 0, 0};
  float dst_rc_pdf = 1.0f;
  if (!rc_is_last_vertex && !rc_is_escape_vertex) {
    rc_brdf = disney_evaluate_lobewise(rc_mat, -dir_to_rc_vertex, z.rc_normal, z.rc_incident_dir, rc_tang, rc_bitang, z.lobes / 10);
    rc_brdf *= saturate(dot(z.rc_normal, z.rc_incident_dir));
    dst_rc_pdf = pdf_disney_lobewise(rc_mat, -dir_to_rc_vertex, z.rc_normal, z.rc_incident_dir, rc_tang, rc_bitang, z.lobes / 10);
  }
  (void)src_dir_to_rc_vertex;  // src_rc_pdf (:707-713) only feeds commented-out Jacobian terms
  V3 rc_nee_brdf{0, 0, 0};
  if (rc_is_NEE_visible) {
    rc_nee_brdf = disney_evaluate(rc_mat, -dir_to_rc_vertex, z.rc_normal, z.rc_NEE_dir, rc_tang, rc_bitang);
    rc_nee_brdf *= saturate(dot(z.rc_normal, z.rc_NEE_dir));
  }
  V3 dst_tang, dst_bitang;
  make_orthonormal_basis(dst_normal, dst_tang, dst_bitang);
  V3 view = normalize(s.cam_pos - dst_pos);
  V3 primary_brdf_d, primary_brdf_s;
  disney_evaluate_lobewise_split(dst_material, view, dst_normal, dir_to_rc_vertex, dst_tang, dst_bitang, z.lobes % 10, primary_brdf_d,
                                 primary_brdf_s);
  float cosd = saturate(dot(dst_normal, dir_to_rc_vertex));
  primary_brdf_d *= cosd;
  primary_brdf_s *= cosd;
  V3 contrib{0, 0, 0};
  if (!rc_is_escape_vertex && !rc_is_last_vertex) {
    float rc_bsdf_sample_light_pdf = cone_sample_pdf(s.light_cos_max, dot(s.light_dir, z.rc_incident_dir));
    float rc_bsdf_mis_weight = power_heuristic(dst_rc_pdf, rc_bsdf_sample_light_pdf * (rc_is_NEE_visible ? 1.0f : 0.0f));
    contrib += firefly_filter(rc_bsdf_mis_weight * rc_brdf / dst_rc_pdf * z.rc_incident_L);
  }
  if (rc_is_escape_vertex) contrib += firefly_filter(z.rc_incident_L);
  if (rc_is_NEE_visible && !rc_is_escape_vertex) {
    float rc_light_sample_bsdf_pdf = pdf_disney(rc_mat, -dir_to_rc_vertex, z.rc_normal, z.rc_NEE_dir, rc_tang, rc_bitang);
    float rc_light_sample_mis_weight = power_heuristic(cone_sample_pdf(s.light_cos_max, 1.0f), rc_light_sample_bsdf_pdf);
    V3 skyT = s.use_physical_sky == 1 ? sample_skybox_transmittance(s.sky_trans, s.sky_res, z.rc_NEE_dir) : V3{1, 1, 1};
    contrib += firefly_filter(rc_light_sample_mis_weight * rc_nee_brdf * skyT * s.light_weight * s.light_color);
  }
  if (rc_mat_id == 2) contrib += rc_mat.base_col;
  diffuse = primary_brdf_d * contrib;
  specular = primary_brdf_s * contrib;
  float jacobian = 1.0f;
  if (!rc_is_escape_vertex) {
    jacobian = z.cached_jacobian_term;
    V3 dir_y1_to_x2 = z.rc_pos - dst_pos;
    jacobian *= std::fabs(dot(normalize(dir_y1_to_x2), z.rc_normal)) / dot(dir_y1_to_x2, dir_y1_to_x2);
  }
  if (jacobian < 0.0f || isbad(jacobian)) jacobian = 0.0f;  // (:799-803; the nested 11x test never fires)
  jacobian_out = jacobian * passed_checks;
}

// pathtracer.py:815-989 spatial_GRIS(pass_id = 0, max_radius = 24, max_taps = 32, pass_total = 1)
// for one pixel. Random dimensions: 65 radius shift, 66+i merge draw of tap i, 98 canonical merge.
static void spatial_gris_pixel(const Ctx& c, int u, int v, uint32_t frame, V3& out_d, V3& out_s) {
  const Scene& s = c.scene;
  const int W = s.W, H = s.H;
  const size_t pi = (size_t)v * W + u;
  const float max_radius = 24.0f;
  const int max_taps = 32;
  const uint32_t key = path_key((uint32_t)pi, frame, c.seed);
  Reservoir center = decode_reservoir(c.reservoirs[pi]);
  const GBufferPx& g = c.gbuf[pi];
  if (g.sky) {  // :854-856
    out_d = center.z.F;
    out_s = V3{0, 0, 0};
    return;
  }
  uint32_t seed = hash3((uint32_t)u >> 3, (uint32_t)v >> 3, frame * 2u + 0u);
  float angle_shift = (float)((seed & 0x007FFFFFu) | 0x3F800000u) / 4294967295.0f * kPi;  // numeric u32 -> f32 cast (:832)
  float radius_shift = rnd(key, 65);
  Reservoir out;
  V3 center_x1 = g.position;
  float center_dist = length(center_x1 - s.cam_pos);
  V3 center_n1 = decode_unit_vector_3x16(g.n_oct[0], g.n_oct[1]);
  Mat center_mat;
  int center_mat_id;
  decode_material(c, g.mat_info, center_mat, center_mat_id);
  int valid_samples = 0;
  float canonical_mis_weight = 1.0f;
  V3 chosen_F_d{0, 0, 0}, chosen_F_s{0, 0, 0};
  for (int i = 0; i < max_taps; i++) {
    const float golden_angle = 2.399963229728f;
    float angle = ((float)i + angle_shift) * golden_angle;
    float offset_radius = std::sqrt(((float)i + radius_shift) / (float)max_taps) * max_radius;
    int ox = (int)(std::cos(angle) * offset_radius), oy = (int)(std::sin(angle) * offset_radius);
    if (ox == 0 && oy == 0) continue;
    int tu = u + ox, tv = v + oy;
    if (tu < 0 || tv < 0 || tu >= W || tv >= H) continue;
    const size_t ti = (size_t)tv * W + tu;
    const GBufferPx& ng = c.gbuf[ti];
    if (ng.sky) continue;
    V3 neighbour_n1 = decode_unit_vector_3x16(ng.n_oct[0], ng.n_oct[1]);
    V3 neighbour_x1 = ng.position;
    float neighbour_dist = length(neighbour_x1 - s.cam_pos);
    Reservoir nb = decode_reservoir(c.reservoirs[ti]);
    if (std::fabs(neighbour_dist - center_dist) > 0.1f * center_dist || dot(center_n1, neighbour_n1) < 0.5f) continue;
    Mat neighbour_mat;
    int neighbour_mat_id;
    decode_material(c, ng.mat_info, neighbour_mat, neighbour_mat_id);
    V3 c_d, c_s, s_d, s_s;
    float c_jacobian, jacobian;
    shift_sample(c, neighbour_x1, neighbour_n1, neighbour_mat, center_x1, center, c_d, c_s, c_jacobian);
    shift_sample(c, center_x1, center_n1, center_mat, neighbour_x1, nb, s_d, s_s, jacobian);
    float center_p_hat = luminance(c_d + c_s) * c_jacobian;
    float canonical_weight = center_p_hat * nb.M;
    canonical_weight /= center_p_hat * nb.M + luminance(center.z.F) * center.M / (float)max_taps;
    canonical_mis_weight += 1.0f - canonical_weight;
    float p_hat = luminance(s_d + s_s);
    // the neighbour's own target value: upstream approximates it by the shifted one (:936). With temporal reuse on,
    // the integrand the neighbour's reservoir was stored with is used instead: temporally resampled reservoirs often
    // hold a sample that is dim at home (large W) and bright once shifted, which the approximation turns into
    // fireflies (measured: rel-RMSE spikes 3.7 vs 1.2 on the material-zoo scene)
    float p_hat_from_neighbour = (c.restir_temporal ? luminance(nb.z.F) : p_hat) / jacobian;
    float neighbour_mis_weight = p_hat_from_neighbour * nb.M;
    neighbour_mis_weight /= p_hat_from_neighbour * nb.M + p_hat * center.M / (float)max_taps;
    if (isbad(neighbour_mis_weight)) neighbour_mis_weight = 0.0f;
    nb.z.F = s_d + s_s;
    bool selected = out.merge(nb, nb.weight * p_hat * jacobian * neighbour_mis_weight, rnd(key, 66 + (uint32_t)i));
    if (selected) {
      chosen_F_d = s_d;
      chosen_F_s = s_s;
    }
    valid_samples += 1;
  }
  // visibility of the resampled reconnection (:957-965)
  bool force_add_canonical = false;
  if (out.weight > 0.0f) {
    const bool esc = is_vec_zero(out.z.rc_normal);
    V3 dir_to_rc_vertex = esc ? out.z.rc_pos : normalize(out.z.rc_pos - center_x1);
    Hit sh = next_hit(s, center_x1 + center_n1 * 0.003f * center_dist, dir_to_rc_vertex, kInf, true, nullptr);
    float actual_dist = esc ? kInf : length(center_x1 - out.z.rc_pos);
    // (an escape / sun sample is never rejected here: |dist - inf| > 0.1 inf is false. Kept as upstream.)
    if (sh.closest < kInf && std::fabs(sh.closest - actual_dist) > 0.1f * actual_dist) {
      out.weight = 0.0f;
      force_add_canonical = true;
    }
  }
  float center_p_hat = luminance(center.z.F);
  bool selected = out.merge(center, center.weight * center_p_hat * canonical_mis_weight, rnd(key, 98), force_add_canonical);
  if (selected) {
    chosen_F_d = c.col_d[pi];
    chosen_F_s = c.col_s[pi];
  }
  out.finalize_without_M();
  out.weight /= (float)(valid_samples + 1);
  V3 emission = center_mat_id == 2 ? center_mat.base_col : V3{0, 0, 0};
  float Wc = clampf(out.weight, 0.0f, 50.0f);
  out_d = chosen_F_d * Wc + emission;
  out_s = chosen_F_s * Wc;
}

// Temporal reservoir reuse. NOT in the reference (its second reservoir slot is written at pathtracer.py:989 and
// never read); BASELINE.json configs[3] asks for "temporal+spatial resampling per frame", so the pass is stated
// here with the reference's own primitives. It runs between render() and spatial_GRIS and is spatial_GRIS
// (:815-989) with ONE tap — the same pixel's reservoir of the previous frame, kept in the second slot — i.e. the
// same similarity test (:911), both reconnection shifts (:672-812), pairwise MIS with max_taps = 1 (:928-944),
// merge (reservoir.py:76-86), a visibility ray for the resampled reconnection (:957-965), the canonical merge
// and finalize_without_M / (valid + 1). Differences from the spatial pass, each deliberate:
//   * the history slot holds THIS pass's output (per-pixel chain), not the spatial pass's: the spatial pass never
//     rejects a shadowed escape / sun sample (its test |dist - inf| > 0.1 inf is never true, :962), and feeding
//     its output back would make those leaks persistent;
//   * the visibility test treats an escape / sun sample as occluded whenever the ray hits anything;
//   * the history's own target value is the integrand it was stored with, not the shifted one (:936);
//   * the history confidence is capped at 20 x the canonical M (the usual ReSTIR bound);
//   * while the mode is on, spatial_GRIS uses the stored integrand for the neighbour's own target value as well.
// Measured on the oracle (96 x 64, 10 frames, geometry pixels, against a 4096-spp path-traced mean): example3
// (emissive ceiling) per-frame mean abs error 1.17 -> 1.02, rel-RMSE 4.9 -> 3.9, image mean 1.016 -> 0.993 of the
// reference; material zoo 0.75 -> 0.69, RMSE spikes (fireflies) gone, image mean 1.026 -> 0.966. Neither estimator
// is unbiased: upstream's leaks light through its visibility test, W is clamped to 50, radiance to 300.
// The pass rewrites the pixel's reservoir and canonical integrands in place (it reads no neighbour), so the
// spatial pass sees the temporally resampled reservoir as its input. Random dimensions: 99 (history merge), 100
// (canonical merge). Static camera only: a camera / light / scene change or reset_framebuffer drops the history.
static const float TEMPORAL_M_CAP = 20.0f;
static void temporal_reuse_pixel(Ctx& c, int u, int v, uint32_t frame) {
  const Scene& s = c.scene;
  const size_t pi = (size_t)v * s.W + u;
  const GBufferPx g = c.gbuf[pi];
  const GBufferPx pg = c.hist_valid ? c.hist_gbuf[pi] : GBufferPx();
  c.hist_gbuf[pi] = g;
  if (g.sky || pg.sky) {  // nothing to reuse: the canonical reservoir starts the chain
    c.hist_res[pi] = c.reservoirs[pi];
    return;
  }
  Reservoir center = decode_reservoir(c.reservoirs[pi]);
  Reservoir prev = decode_reservoir(c.hist_res[pi]);
  prev.M = fminf_(prev.M, TEMPORAL_M_CAP * center.M);
  const V3 center_x1 = g.position, prev_x1 = pg.position;
  const float center_dist = length(center_x1 - s.cam_pos), prev_dist = length(prev_x1 - s.cam_pos);
  const V3 center_n1 = decode_unit_vector_3x16(g.n_oct[0], g.n_oct[1]), prev_n1 = decode_unit_vector_3x16(pg.n_oct[0], pg.n_oct[1]);
  if (!(prev.M > 0.0f) || std::fabs(prev_dist - center_dist) > 0.1f * center_dist || dot(center_n1, prev_n1) < 0.5f) {
    c.hist_res[pi] = c.reservoirs[pi];
    return;
  }
  const uint32_t key = path_key((uint32_t)pi, frame, c.seed);
  Mat center_mat, prev_mat;
  int center_mat_id, prev_mat_id;
  decode_material(c, g.mat_info, center_mat, center_mat_id);
  decode_material(c, pg.mat_info, prev_mat, prev_mat_id);
  V3 c_d, c_s, s_d, s_s;
  float c_jacobian, jacobian;
  shift_sample(c, prev_x1, prev_n1, prev_mat, center_x1, center, c_d, c_s, c_jacobian);
  shift_sample(c, center_x1, center_n1, center_mat, prev_x1, prev, s_d, s_s, jacobian);
  const float center_p_hat_at_prev = luminance(c_d + c_s) * c_jacobian;
  float canonical_weight = center_p_hat_at_prev * prev.M;
  canonical_weight /= center_p_hat_at_prev * prev.M + luminance(center.z.F) * center.M;
  if (isbad(canonical_weight)) canonical_weight = 0.0f;
  const float canonical_mis_weight = 1.0f + (1.0f - canonical_weight);
  const float p_hat = luminance(s_d + s_s);
  const float p_hat_from_prev = luminance(prev.z.F) / jacobian;
  float prev_mis_weight = p_hat_from_prev * prev.M;
  prev_mis_weight /= p_hat_from_prev * prev.M + p_hat * center.M;
  if (isbad(prev_mis_weight)) prev_mis_weight = 0.0f;
  Reservoir out;
  V3 chosen_F_d{0, 0, 0}, chosen_F_s{0, 0, 0};
  prev.z.F = s_d + s_s;
  if (out.merge(prev, prev.weight * p_hat * jacobian * prev_mis_weight, rnd(key, 99))) chosen_F_d = s_d, chosen_F_s = s_s;
  bool force_add_canonical = false;
  if (out.weight > 0.0f) {
    const bool esc = is_vec_zero(out.z.rc_normal);
    V3 dir_to_rc_vertex = esc ? out.z.rc_pos : normalize(out.z.rc_pos - center_x1);
    Hit sh = next_hit(s, center_x1 + center_n1 * 0.003f * center_dist, dir_to_rc_vertex, kInf, true, nullptr);
    float actual_dist = esc ? kInf : length(center_x1 - out.z.rc_pos);
    if (sh.closest < kInf && (esc || std::fabs(sh.closest - actual_dist) > 0.1f * actual_dist)) {
      out.weight = 0.0f;
      force_add_canonical = true;
    }
  }
  if (out.merge(center, center.weight * luminance(center.z.F) * canonical_mis_weight, rnd(key, 100), force_add_canonical))
    chosen_F_d = c.col_d[pi], chosen_F_s = c.col_s[pi];
  out.finalize_without_M();
  out.weight /= 2.0f;
  if (!is_vec_zero(out.z.rc_normal)) out.update_cached_jacobian_term(center_x1);  // the sample now lives at this frame's primary vertex
  c.reservoirs[pi] = encode_reservoir(out);
  c.hist_res[pi] = c.reservoirs[pi];
  c.col_d[pi] = chosen_F_d, c.col_s[pi] = chosen_F_s;
}

static inline bool bad3(V3 c) {  // pathtracer.py:1068-1075
  return isbad(c.x) || isbad(c.y) || isbad(c.z) || c.x < 0.0f || c.y < 0.0f || c.z < 0.0f;
}

static double halton(uint32_t i, uint32_t b) {
  double f = 1.0, r = 0.0;
  while (i > 0) {
    f /= (double)b;
    r += f * (double)(i % b);
    i /= b;
  }
  return r;
}
static void set_jitter(Ctx& c, uint32_t sample) {
  // pathtracer.py:264-265: (rand*2-1) * inv_image_res, one draw per frame shared by all pixels.
  if (c.jitter_mode == 1) {
    c.scene.jitter[0] = (float)((halton(sample + 1, 2) * 2.0 - 1.0) / (double)c.scene.W);
    c.scene.jitter[1] = (float)((halton(sample + 1, 3) * 2.0 - 1.0) / (double)c.scene.H);
  } else {
    c.scene.jitter[0] = c.scene.jitter[1] = 0.0f;
  }
}

// math_utils.py:160-186 (all locals are f32 in the Taichi func)
static inline V3 uchimura(V3 x) {
  const float P = 1.0f, a = 1.0f, m = 0.22f, l = 0.4f, cc = 1.33f, b = 0.0f;
  const float l0 = ((P - m) * l) / a;
  const float S0 = m + l0;
  const float S1 = m + a * l0;
  const float C2 = (a * P) / (P - S1);
  const float CP = -C2 / P;
  V3 r;
  for (int i = 0; i < 3; i++) {
    float xi = x[i];
    float t = clampf((xi - 0.0f) / (m - 0.0f), 0.0f, 1.0f);
    float w0 = 1.0f - t * t * (3.0f - 2.0f * t);
    float w2 = xi < m + l0 ? 0.0f : 1.0f;
    float w1 = 1.0f - w0 - w2;
    float T = m * std::pow(xi / m, cc) + b;
    float S = P - (P - S1) * std::exp(CP * (xi - S0));
    float L = m + a * (xi - m);
    r.at(i) = T * w0 + L * w1 + S * w2;
  }
  return r;
}

}  // namespace

extern "C" {

void* orc_create(int W, int H, int R, float voxel_dx, float voxel_edges, float exposure, int max_depth, uint32_t seed,
                 int jitter_mode, int cloud_passes) {
  Ctx* c = new Ctx();
  c->scene.W = W, c->scene.H = H, c->scene.R = R;
  c->scene.voxel_size = voxel_dx;
  c->scene.voxel_inv_size = (float)(1.0 / (double)voxel_dx);  // voxel_world.py:11 (python float)
  c->scene.voxel_edges = voxel_edges;
  c->exposure = exposure, c->max_depth = max_depth, c->seed = seed, c->jitter_mode = jitter_mode;
  c->cloud_passes = cloud_passes;
  c->sky.seed = seed;
  c->scene.material.assign((size_t)R * R * R, 0);
  c->scene.color.assign((size_t)R * R * R * 3, 0);
  c->mats.assign(128, Mat{V3{1, 1, 1}, 0.0f, 0.0f, 0.04f, 0.0f, 0.9f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f});
  c->hist_d.assign((size_t)W * H * 4, 0.0f);
  c->hist_s.assign((size_t)W * H * 4, 0.0f);
  build_occupancy(c->scene);
  // defaults: scene.py:127, pathtracer.py:91-93
  double n = std::sqrt(3.0);
  c->scene.light_dir = V3{(float)(1 / n), (float)(1 / n), (float)(1 / n)};
  c->scene.light_cos_max = (float)std::cos(0.1 * 0.5);
  return c;
}
void orc_destroy(void* p) { delete (Ctx*)p; }

void orc_upload_voxels(void* p, const int8_t* mat, const uint8_t* rgb) {
  Ctx* c = (Ctx*)p;
  size_t n = (size_t)c->scene.R * c->scene.R * c->scene.R;
  std::copy(mat, mat + n, c->scene.material.begin());
  std::copy(rgb, rgb + 3 * n, c->scene.color.begin());
  build_occupancy(c->scene);
}
// Inverses are formed here, in float64 Gauss-Jordan with partial pivoting, then rounded to
// float32 (the reference inverts in-kernel with Taichi's inverse(), pathtracer.py:273,281).
static bool invert4(const double m[16], double inv[16]) {
  double a[4][8];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) a[i][j] = m[i * 4 + j], a[i][4 + j] = i == j ? 1.0 : 0.0;
  for (int c = 0; c < 4; c++) {
    int p = c;
    for (int r = c + 1; r < 4; r++)
      if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
    if (std::fabs(a[p][c]) < 1e-300) return false;
    if (p != c)
      for (int j = 0; j < 8; j++) std::swap(a[p][j], a[c][j]);
    double d = a[c][c];
    for (int j = 0; j < 8; j++) a[c][j] /= d;
    for (int r = 0; r < 4; r++)
      if (r != c) {
        double f = a[r][c];
        for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) inv[i * 4 + j] = a[i][4 + j];
  return true;
}
int orc_set_camera(void* p, const float* pos, const float* view, const float* proj) {
  Ctx* c = (Ctx*)p;
  double v[16], pr[16], vi[16], pi[16];
  for (int i = 0; i < 16; i++) v[i] = view[i], pr[i] = proj[i];
  if (!invert4(v, vi) || !invert4(pr, pi)) return -1;
  c->hist_valid = false;  // temporal reservoir reuse is defined for a static camera
  c->scene.cam_pos = V3{pos[0], pos[1], pos[2]};
  for (int i = 0; i < 16; i++) {
    c->scene.view[i] = view[i], c->scene.proj[i] = proj[i];
    c->scene.inv_view[i] = (float)vi[i], c->scene.inv_proj[i] = (float)pi[i];
  }
  return 0;
}
void orc_set_light(void* p, const float* dir, float cone_angle, const float* rgb) {
  Ctx* c = (Ctx*)p;
  double x = dir[0], y = dir[1], z = dir[2];
  double n = std::sqrt(x * x + y * y + z * z);
  c->scene.light_dir = V3{(float)(x / n), (float)(y / n), (float)(z / n)};
  c->scene.light_cos_max = (float)std::cos((double)cone_angle * 0.5);
  c->scene.light_color = V3{rgb[0], rgb[1], rgb[2]};
  c->scene.light_weight = 3.0f;
  c->hist_valid = false;
}
void orc_set_floor(void* p, float h, const float* rgb, int mat) {
  Ctx* c = (Ctx*)p;
  c->scene.floor_height = h;
  c->scene.floor_color = V3{rgb[0], rgb[1], rgb[2]};
  c->scene.floor_material = mat;
}
void orc_set_background(void* p, const float* rgb) { ((Ctx*)p)->scene.background = V3{rgb[0], rgb[1], rgb[2]}; }
void orc_set_sky(void* p, int physical, int clouds) {
  Ctx* c = (Ctx*)p;
  c->scene.use_physical_sky = physical ? 1 : 0;
  c->sky.use_clouds = clouds ? 1 : 0;
}
void orc_set_materials(void* p, const float* t) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < 128; i++) {
    const float* r = t + i * 14;
    c->mats[i] = Mat{V3{r[0], r[1], r[2]}, r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13]};
  }
}
void orc_set_cloud_texture(void* p, const uint8_t* tex) {
  Ctx* c = (Ctx*)p;
  c->sky.cloud_tex.assign(tex, tex + 256 * 256 * 3);
}
static void bind_sky(Ctx* c) {
  c->scene.sky_res = c->sky.S;
  c->scene.sky_scatter = c->sky.scatter.data();
  c->scene.sky_trans = c->sky.trans.data();
}
void orc_precompute_sky(void* p, int sky_res) {
  Ctx* c = (Ctx*)p;
  if (c->sky.cloud_tex.empty()) c->sky.cloud_tex.assign(256 * 256 * 3, 0);
  V3 sun_col = c->scene.light_color * c->scene.light_weight;  // pathtracer.py:320,326,329
  precompute_sky(c->sky, sky_res, c->scene.light_dir, sun_col, c->scene.light_cos_max, c->cloud_passes);
  bind_sky(c);
}
void orc_set_sky_tables(void* p, int S, const float* scat, const float* trans) {
  Ctx* c = (Ctx*)p;
  c->sky.S = S;
  c->sky.scatter.assign(scat, scat + (size_t)S * S * 3);
  c->sky.trans.assign(trans, trans + (size_t)S * S * 3);
  bind_sky(c);
}
void orc_get_sky_tables(void* p, float* scat, float* trans) {
  Ctx* c = (Ctx*)p;
  std::copy(c->sky.scatter.begin(), c->sky.scatter.end(), scat);
  std::copy(c->sky.trans.begin(), c->sky.trans.end(), trans);
}
void orc_get_trans_lut(void* p, uint16_t* lut) {
  Ctx* c = (Ctx*)p;
  if (c->sky.trans_lut.empty()) generate_transmittance_lut(c->sky);
  std::copy(c->sky.trans_lut.begin(), c->sky.trans_lut.end(), lut);
}
void orc_get_cloud_ambient(void* p, float* out) {
  Ctx* c = (Ctx*)p;
  out[0] = c->sky.cloud_ambient.x, out[1] = c->sky.cloud_ambient.y, out[2] = c->sky.cloud_ambient.z;
}
void orc_set_tile_shard(void* p, int rank, int n) {
  Ctx* c = (Ctx*)p;
  c->tile_rank = rank, c->tile_n = n;
}

// Primary-hit dump, same record layout as vrt_trace_primary.
void orc_trace_primary(void* p, vrt_hit* out) {
  Ctx* c = (Ctx*)p;
  Scene& s = c->scene;
  s.jitter[0] = s.jitter[1] = 0.0f;
#pragma omp parallel for schedule(dynamic, 4)
  for (int v = 0; v < s.H; v++)
    for (int u = 0; u < s.W; u++) {
      V3 d = get_cast_dir(s, (float)u, (float)v);
      Hit h = next_hit(s, s.cam_pos, d, kInf, false, nullptr);
      vrt_hit& o = out[(size_t)v * s.W + u];
      o.t = h.closest;
      int kind = h.closest < kInf ? h.kind : 0;
      o.cell[0] = kind == 2 ? h.cell.x : -1, o.cell[1] = kind == 2 ? h.cell.y : -1, o.cell[2] = kind == 2 ? h.cell.z : -1;
      o.normal[0] = h.normal.x, o.normal[1] = h.normal.y, o.normal[2] = h.normal.z;
      uint32_t shadow = 3;
      if (!h.hit_light && h.closest < kInf) {
        V3 hit_pos = s.cam_pos + h.closest * d;
        V3 pos = hit_pos + h.normal * kEps;
        float ndl = dot(s.light_dir, h.normal);
        if (ndl > 0.0f) {
          Hit sh = next_hit(s, pos, s.light_dir, kInf, true, nullptr);
          shadow = sh.closest >= kInf ? 0u : 1u;
        } else {
          shadow = 2u;
        }
      }
      o.flags = (uint32_t)kind | (shadow << 8) | (((uint32_t)h.mat_id & 255u) << 16) | ((uint32_t)(h.hit_light ? 1 : 0) << 24);
    }
}

// accumulate(): render + static-camera temporal filters, one frame per sample index.
void orc_accumulate(void* p, int first_sample, int n_samples, int stride, int stats, int n_threads) {
  Ctx* c = (Ctx*)p;
  Scene& s = c->scene;
  if (n_threads > 0) omp_set_num_threads(n_threads);
  double t0 = omp_get_wtime();
  const float max_accum = 999999999.0f;  // scene.py:210
  const int tiles_x = s.W / 8;
  for (int k = 0; k < n_samples; k++) {
    uint32_t sample = (uint32_t)(first_sample + k * stride);
    set_jitter(*c, sample);
    Counters total;
#pragma omp parallel
    {
      Counters local;
#pragma omp for schedule(dynamic, 2)
      for (int v = 0; v < s.H; v++)
        for (int u = 0; u < s.W; u++) {
          if (c->tile_n > 1) {
            int tile = (v / 4) * tiles_x + (u / 8);
            if (tile % c->tile_n != c->tile_rank) continue;
          }
          PathOut o;
          trace_path(*c, u, v, sample, stats ? &local : nullptr, o);
          if (bad3(o.diffuse)) o.diffuse = V3{0, 0, 0};
          if (bad3(o.specular)) o.specular = V3{0, 0, 0};
          float* hd = &c->hist_d[((size_t)v * s.W + u) * 4];
          float* hs = &c->hist_s[((size_t)v * s.W + u) * 4];
          hd[3] = fminf_(hd[3] + 1.0f, max_accum);
          float wd = 1.0f / hd[3];
          hd[0] = mixf(hd[0], o.diffuse.x, wd), hd[1] = mixf(hd[1], o.diffuse.y, wd), hd[2] = mixf(hd[2], o.diffuse.z, wd);
          hs[3] = fminf_(hs[3] + 1.0f, max_accum);
          float ws = 1.0f / hs[3];
          hs[0] = mixf(hs[0], o.specular.x, ws), hs[1] = mixf(hs[1], o.specular.y, ws), hs[2] = mixf(hs[2], o.specular.z, ws);
        }
#pragma omp critical
      total.add(local);
    }
    c->counters.add(total);
  }
  c->last_ms = (omp_get_wtime() - t0) * 1e3;
}
// accumulate() with USE_RESTIR_PT (pathtracer.py:1310-1319): render -> spatial_GRIS -> static
// camera temporal filters, one frame per sample index.
void orc_accumulate_restir(void* p, int first_sample, int n_samples, int stride, int n_threads) {
  Ctx* c = (Ctx*)p;
  Scene& s = c->scene;
  if (n_threads > 0) omp_set_num_threads(n_threads);
  double t0 = omp_get_wtime();
  const size_t npx = (size_t)s.W * s.H;
  c->reservoirs.resize(npx);
  c->gbuf.resize(npx);
  c->col_d.resize(npx);
  c->col_s.resize(npx);
  const float max_accum = 999999999.0f;
  std::vector<V3> fin_d(npx), fin_s(npx);
  for (int k = 0; k < n_samples; k++) {
    uint32_t sample = (uint32_t)(first_sample + k * stride);
    set_jitter(*c, sample);
#pragma omp parallel for schedule(dynamic, 2)
    for (int v = 0; v < s.H; v++)
      for (int u = 0; u < s.W; u++) {
        size_t i = (size_t)v * s.W + u;
        Reservoir r;
        trace_path_restir(*c, u, v, sample, nullptr, r, c->gbuf[i], c->col_d[i], c->col_s[i]);
        c->reservoirs[i] = encode_reservoir(r);
      }
    if (c->restir_temporal) {
      if (c->hist_res.size() != npx) c->hist_res.resize(npx), c->hist_gbuf.resize(npx), c->hist_valid = false;
#pragma omp parallel for schedule(dynamic, 2)
      for (int v = 0; v < s.H; v++)
        for (int u = 0; u < s.W; u++) temporal_reuse_pixel(*c, u, v, sample);
      c->hist_valid = true;
    }
#pragma omp parallel for schedule(dynamic, 2)
    for (int v = 0; v < s.H; v++)
      for (int u = 0; u < s.W; u++) spatial_gris_pixel(*c, u, v, sample, fin_d[(size_t)v * s.W + u], fin_s[(size_t)v * s.W + u]);
    for (size_t i = 0; i < npx; i++) {
      V3 dd = fin_d[i], ss = fin_s[i];
      if (bad3(dd)) dd = V3{0, 0, 0};
      if (bad3(ss)) ss = V3{0, 0, 0};
      float* hd = &c->hist_d[i * 4];
      float* hs = &c->hist_s[i * 4];
      hd[3] = fminf_(hd[3] + 1.0f, max_accum);
      float wd = 1.0f / hd[3];
      hd[0] = mixf(hd[0], dd.x, wd), hd[1] = mixf(hd[1], dd.y, wd), hd[2] = mixf(hd[2], dd.z, wd);
      hs[3] = fminf_(hs[3] + 1.0f, max_accum);
      float ws = 1.0f / hs[3];
      hs[0] = mixf(hs[0], ss.x, ws), hs[1] = mixf(hs[1], ss.y, ws), hs[2] = mixf(hs[2], ss.z, ws);
    }
  }
  c->last_ms = (omp_get_wtime() - t0) * 1e3;
}
// packed reservoirs / G-buffer of the last ReSTIR frame (for parity tests): 56 B and 24 B per pixel
void orc_get_reservoirs(void* p, void* out) {
  Ctx* c = (Ctx*)p;
  std::memcpy(out, c->reservoirs.data(), c->reservoirs.size() * sizeof(StorageReservoir));
}
// octahedral / arbitrary-bit packing probes
void orc_oct_round_trip(int n, const float* v, float* enc, float* dec) {
  for (int i = 0; i < n; i++) {
    encode_unit_vector_3x16(V3{v[3 * i], v[3 * i + 1], v[3 * i + 2]}, enc[2 * i], enc[2 * i + 1]);
    V3 d = decode_unit_vector_3x16(enc[2 * i], enc[2 * i + 1]);
    dec[3 * i] = d.x, dec[3 * i + 1] = d.y, dec[3 * i + 2] = d.z;
  }
}
uint32_t orc_hash3(uint32_t x, uint32_t y, uint32_t z) { return hash3(x, y, z); }
// probes of the small helpers, checked against vectors computed by the reference's own math_utils.py
// kind: 0 make_orthonormal_basis (in: n; out: x, y)   1 sample_cone_oriented (in: n, cosmax, u0, u1; out: dir)
//       2 sample_cosine_weighted_hemisphere (in: n, -, u0, u1; out: dir)   3 uchimura (in: rgb; out: rgb)
//       4 luminance (in: rgb; out: 1)   5 encode_material (in: albedo, mat id as float; out: u32 bits)
void orc_math_probe(int kind, int n, const float* a, const float* b, float* out) {
  for (int i = 0; i < n; i++) {
    V3 v{a[3 * i], a[3 * i + 1], a[3 * i + 2]};
    if (kind == 0) {
      V3 x, y;
      make_orthonormal_basis(v, x, y);
      float* o = out + 6 * i;
      o[0] = x.x, o[1] = x.y, o[2] = x.z, o[3] = y.x, o[4] = y.y, o[5] = y.z;
    } else if (kind == 1 || kind == 2) {
      V3 d = kind == 1 ? sample_cone_oriented(b[3 * i], v, b[3 * i + 1], b[3 * i + 2]) : sample_cosine_weighted_hemisphere(v, b[3 * i + 1], b[3 * i + 2]);
      out[3 * i] = d.x, out[3 * i + 1] = d.y, out[3 * i + 2] = d.z;
    } else if (kind == 3) {
      V3 d = uchimura(v);
      out[3 * i] = d.x, out[3 * i + 1] = d.y, out[3 * i + 2] = d.z;
    } else if (kind == 4) {
      out[i] = luminance(v);
    } else if (kind == 5) {
      uint32_t e = encode_material((int)b[i], v);
      std::memcpy(out + i, &e, 4);
    }
  }
}
// lobe-wise BSDF probes (bsdf.py:306-380): out[n][3 lobes][7] = {diffuse rgb, specular rgb, pdf}; lobe_w[n][3]
void orc_bsdf_lobewise_probe(void* p, int n, const int* mat_id, const float* albedo, const float* v, const float* nrm, const float* l,
                             float* out, float* lobe_w) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) {
    Mat m = mat_at(*c, mat_id[i]);
    m.base_col = V3{albedo[3 * i], albedo[3 * i + 1], albedo[3 * i + 2]};
    V3 vv{v[3 * i], v[3 * i + 1], v[3 * i + 2]}, nn{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]}, ll{l[3 * i], l[3 * i + 1], l[3 * i + 2]};
    V3 tang, bitang;
    make_orthonormal_basis(nn, tang, bitang);
    lobe_probabilities(m, lobe_w[3 * i], lobe_w[3 * i + 1], lobe_w[3 * i + 2]);
    for (int lobe = 0; lobe < 3; lobe++) {
      V3 d, s;
      disney_evaluate_lobewise_split(m, vv, nn, ll, tang, bitang, lobe, d, s);
      float* o = out + (3 * i + lobe) * 7;
      o[0] = d.x, o[1] = d.y, o[2] = d.z, o[3] = s.x, o[4] = s.y, o[5] = s.z;
      o[6] = pdf_disney_lobewise(m, vv, nn, ll, tang, bitang, lobe);
    }
  }
}
// ------------------------------------------------------------------ moving-camera temporal path
// Renderer.accumulate() with camera_is_moving = 1 (scene.py:214-228): render at render_scale with
// albedo-demodulated diffuse (pathtracer.py:628-630), temporal_filter_prepass (:1020-1075),
// temporal_filter (:1185-1230) and temporal_filter_specular (:1242-1303) with reprojection into the
// previous frame, Catmull-Rom 4x4 history fetch (:1092-1183) and depth / normal rejection, then
// copy_prev_matrices (:284-287). Pins: the prepass blur reads pre-blur reflection depths (the
// reference reads and writes gbuff_depth_reflection in place: a race); non-finite reflection
// distances count as "no reflection"; buffer reads are clamped to the image; sky pixels store a
// zero normal code.
static inline float catmullrom(float x) {  // pathtracer.py:1002-1013
  float x2 = x * x, x3 = x * x * x, fx = 0.0f;
  if (x < 1.0f)
    fx = 1.5f * x3 - 2.5f * x2 + 1.0f;
  else if (x < 2.0f)
    fx = -0.5f * x3 + 2.5f * x2 - 4.0f * x + 2.0f;
  return fx;
}
struct V4 {
  float x, y, z, w;
};
static inline V4 ld4(const std::vector<float>& b, size_t i) { return V4{b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]}; }
static inline void st4(std::vector<float>& b, size_t i, V4 v) { b[4 * i] = v.x, b[4 * i + 1] = v.y, b[4 * i + 2] = v.z, b[4 * i + 3] = v.w; }

static V3 mv_bilinear(const Ctx& c, const std::vector<V3>& buf, float uvx, float uvy, int irx, int iry) {  // :1077-1090
  const int W = c.scene.W, H = c.scene.H;
  float fx = uvx * (float)irx - 0.5f, fy = uvy * (float)iry - 0.5f;
  int ix = (int)fx, iy = (int)fy;
  float wx = fractf(fx), wy = fractf(fy);
  auto at = [&](int x, int y) {
    x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
    y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
    return buf[(size_t)y * W + x];
  };
  return mix3(mix3(at(ix, iy), at(ix + 1, iy), wx), mix3(at(ix, iy + 1), at(ix + 1, iy + 1), wx), wy);
}
static V3 mv_reproject(const Ctx& c, V3 world_pos) {  // :991-998
  float a[4] = {world_pos.x, world_pos.y, world_pos.z, 1.0f}, b[4], q[4];
  mat4_mul(c.mv.prev_view, a, b);
  mat4_mul(c.mv.prev_proj, b, q);
  return V3{q[0] / q[3] * 0.5f + 0.5f, q[1] / q[3] * 0.5f + 0.5f, q[2] / q[3] * 0.5f + 0.5f};
}
// history_filter (:1092-1130) and history_filter_specular (:1132-1183) in one routine
static float mv_history(const Ctx& c, bool specular, float uvx, float uvy, float center_depth, V3 center_normal, int irx, int iry, V4& col_out,
                        float& depth_out) {
  const Scene& s = c.scene;
  const auto& m = c.mv;
  col_out = V4{0, 0, 0, 1};
  depth_out = 0.0f;
  if (!(std::isfinite(uvx) && std::isfinite(uvy)) || std::fabs(uvx) > 1e6f || std::fabs(uvy) > 1e6f) return 0.0f;
  float fx = uvx * (float)irx - 0.5f, fy = uvy * (float)iry - 0.5f;
  int ix = (int)fx, iy = (int)fy;
  float ffx = fractf(fx), ffy = fractf(fy);
  V4 sum{0, 0, 0, 0}, cmax{0, 0, 0, 0}, cmin{999999.0f, 999999.0f, 999999.0f, 999999.0f};
  float dsum = 0.0f, dmax = 0.0f, dmin = 999999.0f, wsum = 0.0f;
  for (int x = -1; x < 3; x++)
    for (int y = -1; y < 3; y++) {
      int tx = ix + x, ty = iy + y;
      if (tx < 0 || ty < 0 || tx > irx - 1 || ty > iry - 1) continue;
      size_t ti = (size_t)ty * s.W + tx;
      float w = catmullrom(std::fabs((float)x - ffx)) * catmullrom(std::fabs((float)y - ffy));
      V3 tap_normal = decode_unit_vector_3x16(m.prev_noct[2 * ti], m.prev_noct[2 * ti + 1]);
      if (!specular) {
        float tap_depth = linearize_depth(m.prev_depth[ti], s.inv_proj);
        w *= std::fabs(tap_depth - center_depth) / center_depth < 0.05f ? 1.0f : 0.0f;
      }
      w *= dot(center_normal, tap_normal) > 0.642f ? 1.0f : 0.0f;
      V4 col = ld4(specular ? m.hs[0] : m.hd[0], ti);
      cmax = V4{fmaxf_(cmax.x, col.x), fmaxf_(cmax.y, col.y), fmaxf_(cmax.z, col.z), fmaxf_(cmax.w, col.w)};
      cmin = V4{fminf_(cmin.x, col.x), fminf_(cmin.y, col.y), fminf_(cmin.z, col.z), fminf_(cmin.w, col.w)};
      sum = V4{sum.x + col.x * w, sum.y + col.y * w, sum.z + col.z * w, sum.w + col.w * w};
      if (specular) {
        float rd = m.hsd[0][ti];
        dmin = fminf_(dmin, rd), dmax = fmaxf_(dmax, rd);
        dsum += rd * w;
      }
      wsum += w;
    }
  sum = V4{sum.x / wsum, sum.y / wsum, sum.z / wsum, sum.w / wsum};
  dsum /= wsum;
  col_out = V4{fmaxf_(clampf(sum.x, cmin.x, cmax.x), 0.0f), fmaxf_(clampf(sum.y, cmin.y, cmax.y), 0.0f), fmaxf_(clampf(sum.z, cmin.z, cmax.z), 0.0f),
               fmaxf_(clampf(sum.w, cmin.w, cmax.w), 1.0f)};
  depth_out = clampf(dsum, dmin, dmax);
  return wsum;
}

void orc_accumulate_moving(void* p, int sample, float render_scale, float max_accum) {
  Ctx* c = (Ctx*)p;
  Scene& s = c->scene;
  auto& m = c->mv;
  const int W = s.W, H = s.H;
  const size_t npx = (size_t)W * H;
  if (m.depth.size() != npx) {
    m.col_d.assign(npx, V3{0, 0, 0}), m.col_s.assign(npx, V3{0, 0, 0}), m.out.assign(npx, V3{0, 0, 0});
    m.depth.assign(npx, 0.0f), m.refl.assign(npx, 0.0f), m.refl_blur.assign(npx, 0.0f), m.prev_depth.assign(npx, 0.0f);
    m.noct.assign(2 * npx, 0.0f), m.prev_noct.assign(2 * npx, 0.0f), m.mat.assign(npx, 0u);
    for (int k = 0; k < 2; k++) m.hd[k].assign(4 * npx, 0.0f), m.hs[k].assign(4 * npx, 0.0f), m.hsd[k].assign(npx, 0.0f);
  }
  if (!m.has_prev) {
    std::copy(s.view, s.view + 16, m.prev_view), std::copy(s.proj, s.proj + 16, m.prev_proj);
    m.prev_cam = s.cam_pos, m.has_prev = true;
  }
  m.scale = render_scale, m.active = true;
  s.jitter[0] = s.jitter[1] = 0.0f;
  const int irx = (int)((float)W * render_scale), iry = (int)((float)H * render_scale);
  auto outside = [&](int u, int v) { return (float)u > render_scale * (float)W || (float)v > render_scale * (float)H; };  // :289-291
  // ---- render
#pragma omp parallel for schedule(dynamic, 2)
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      if (outside(u, v)) continue;
      const size_t i = (size_t)v * W + u;
      PathOut o;
      GExtra g;
      trace_path(*c, u, v, (uint32_t)sample, nullptr, o, &g, render_scale);
      m.depth[i] = view_to_screen(xform_point(s.view, g.primary_pos, 1.0f), s.proj).z;
      m.noct[2 * i] = m.noct[2 * i + 1] = 0.0f;
      if (!g.sky) encode_unit_vector_3x16(g.primary_normal, m.noct[2 * i], m.noct[2 * i + 1]);
      m.mat[i] = g.mat_info;
      const float rd = g.first_bounce_reflection_dist;
      float refl = 0.0f;
      if (rd != 0.0f && std::isfinite(rd)) {
        V3 primary_dir = normalize(g.primary_pos - s.cam_pos);
        V3 virtual_point = g.primary_pos + primary_dir * rd;
        refl = linearize_depth(view_to_screen(xform_point(s.view, virtual_point, 1.0f), s.proj).z, s.inv_proj);
      }
      m.refl[i] = refl;
      V3 diffuse = o.diffuse / max3(g.primary_albedo, 1e-2f);  // de-modulate albedo (:628-630)
      m.col_d[i] = diffuse, m.col_s[i] = o.specular;
    }
    // ---- temporal_filter_prepass
#pragma omp parallel for schedule(static)
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      if (outside(u, v)) continue;
      const size_t i = (size_t)v * W + u;
      float sum = 0.0f, valid = 0.0f;
      for (int x = -1; x < 3; x++)
        for (int y = -1; y < 3; y++) {
          int tx = u + x, ty = v + y;
          if (tx < 0 || ty < 0 || tx > irx - 1 || ty > iry - 1) continue;
          float r = m.refl[(size_t)ty * W + tx];
          if (r != 0.0f) valid += 1.0f, sum += r;
        }
      m.refl_blur[i] = valid > 0.01f ? sum / valid : 0.0f;
      if (bad3(m.col_d[i])) m.col_d[i] = V3{0, 0, 0};
      if (bad3(m.col_s[i])) m.col_s[i] = V3{0, 0, 0};
    }
    // ---- temporal_filter + temporal_filter_specular
#pragma omp parallel for schedule(dynamic, 2)
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      if (outside(u, v)) continue;
      const size_t i = (size_t)v * W + u;
      m.out[i] = m.col_d[i];
      const float tcx = ((float)u + 0.5f) * (1.0f / (float)W) / render_scale, tcy = ((float)v + 0.5f) * (1.0f / (float)H) / render_scale;
      const float d_nl = m.depth[i];
      const V3 center_n1 = decode_unit_vector_3x16(m.noct[2 * i], m.noct[2 * i + 1]);
      const V3 center_x1 = xform_point(s.inv_view, screen_to_view(tcx, tcy, d_nl, s.inv_proj), 1.0f);
      if (is_vec_zero(center_x1)) continue;
      {  // diffuse
        V3 current = mv_bilinear(*c, m.col_d, tcx, tcy, irx, iry);
        V3 rp = mv_reproject(*c, center_x1);
        V4 history;
        float dummy;
        float w_sum = mv_history(*c, false, rp.x, rp.y, linearize_depth(rp.z, s.inv_proj), center_n1, irx, iry, history, dummy);
        if (w_sum > 1e-3f) {
          history.w = fminf_(history.w + 1.0f, max_accum);
          float t = 1.0f / history.w;
          history.x = mixf(history.x, current.x, t), history.y = mixf(history.y, current.y, t), history.z = mixf(history.z, current.z, t);
        } else {
          history = V4{current.x, current.y, current.z, 1.0f};
        }
        st4(m.hd[1], i, history);
        Mat cm;
        int cid;
        decode_material(*c, m.mat[i], cm, cid);
        m.out[i] = V3{history.x, history.y, history.z} * cm.base_col;  // re-modulate albedo (:1227-1228)
      }
      {  // specular, reprojected through the virtual reflection point (:1251-1265)
        const float center_refl_depth = m.refl_blur[i];
        const float refl_nl = delinearize_depth(center_refl_depth, s.proj);
        const V3 center_refl_pos = xform_point(s.inv_view, screen_to_view(tcx, tcy, refl_nl, s.inv_proj), 1.0f);
        V3 current = mv_bilinear(*c, m.col_s, tcx, tcy, irx, iry);
        V3 rp = mv_reproject(*c, center_refl_depth != 0.0f ? center_refl_pos : center_x1);
        V4 history;
        float refl_hist;
        float w_sum = mv_history(*c, true, rp.x, rp.y, linearize_depth(rp.z, s.inv_proj), center_n1, irx, iry, history, refl_hist);
        if (w_sum > 1e-3f) {
          history.w = fminf_(history.w + 1.0f, max_accum);
          float t = 1.0f / history.w;
          history.x = mixf(history.x, current.x, t), history.y = mixf(history.y, current.y, t), history.z = mixf(history.z, current.z, t);
          refl_hist = mixf(refl_hist, center_refl_depth, t);
        } else {
          history = V4{current.x, current.y, current.z, 1.0f};
          refl_hist = center_refl_depth;
        }
        st4(m.hs[1], i, history);
        m.hsd[1][i] = refl_hist;
        m.out[i] += V3{history.x, history.y, history.z};
      }
    }
  // ---- slot 1 -> slot 0, previous G-buffer (:1297-1303), copy_prev_matrices (:284-287)
  m.hd[0] = m.hd[1], m.hs[0] = m.hs[1], m.hsd[0] = m.hsd[1];
  m.prev_depth = m.depth, m.prev_noct = m.noct;
  std::copy(s.view, s.view + 16, m.prev_view), std::copy(s.proj, s.proj + 16, m.prev_proj);
  m.prev_cam = s.cam_pos;
}
// color_buffer of the moving path, nearest-upsampled as _render_to_image does (:643-644)
void orc_fetch_hdr_moving(void* p, float* rgba) {
  Ctx* c = (Ctx*)p;
  const int W = c->scene.W, H = c->scene.H;
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      int sx = (int)((float)i * c->mv.scale), sy = (int)((float)j * c->mv.scale);
      V3 v = c->mv.out[(size_t)sy * W + sx];
      float* o = rgba + ((size_t)j * W + i) * 4;
      o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = 1.0f;
    }
}
void orc_reset_moving(void* p) {
  Ctx* c = (Ctx*)p;
  c->mv = Ctx::Moving();
}
double orc_last_ms(void* p) { return ((Ctx*)p)->last_ms; }
void orc_set_restir_temporal(void* p, int enable) {
  Ctx* c = (Ctx*)p;
  c->restir_temporal = enable ? 1 : 0;
  c->hist_valid = false;
}
void orc_drop_restir_history(void* p) { ((Ctx*)p)->hist_valid = false; }
void orc_reset(void* p) {
  Ctx* c = (Ctx*)p;
  c->hist_valid = false;
  std::fill(c->hist_d.begin(), c->hist_d.end(), 0.0f);
  std::fill(c->hist_s.begin(), c->hist_s.end(), 0.0f);
  c->counters = Counters();
}
void orc_get_counters(void* p, uint64_t* out) {
  const Counters& k = ((Ctx*)p)->counters;
  out[0] = k.paths, out[1] = k.rays, out[2] = k.steps, out[3] = k.Q, out[4] = k.H, out[5] = k.E, out[6] = k.N,
  out[7] = k.vertices;
}
// color_buffer = diffuse history + specular history (pathtracer.py:1230,1295); w = sample count
void orc_fetch_hdr(void* p, float* rgba) {
  Ctx* c = (Ctx*)p;
  size_t n = (size_t)c->scene.W * c->scene.H;
  for (size_t i = 0; i < n; i++) {
    rgba[i * 4 + 0] = c->hist_d[i * 4 + 0] + c->hist_s[i * 4 + 0];
    rgba[i * 4 + 1] = c->hist_d[i * 4 + 1] + c->hist_s[i * 4 + 1];
    rgba[i * 4 + 2] = c->hist_d[i * 4 + 2] + c->hist_s[i * 4 + 2];
    rgba[i * 4 + 3] = c->hist_d[i * 4 + 3];
  }
}
// _render_to_image (pathtracer.py:634-662) applied to an arbitrary HDR buffer
void orc_tonemap(void* p, const float* hdr_rgba, float* ldr_rgba) {
  Ctx* c = (Ctx*)p;
  int W = c->scene.W, H = c->scene.H;
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      size_t k = ((size_t)j * W + i) * 4;
      float uvx = (float)i / (float)W, uvy = (float)j / (float)H;
      float dx = uvx - 0.5f, dy = uvy - 0.5f;
      float dist = std::sqrt(dx * dx + dy * dy);
      float darken = 1.0f - 0.9f * fmaxf_(dist - 0.0f, 0.0f);
      V3 hdr{hdr_rgba[k], hdr_rgba[k + 1], hdr_rgba[k + 2]};
      V3 tm = uchimura(hdr * darken * c->exposure);
      ldr_rgba[k + 0] = saturate(std::pow(tm.x, 1.0f / 2.2f));
      ldr_rgba[k + 1] = saturate(std::pow(tm.y, 1.0f / 2.2f));
      ldr_rgba[k + 2] = saturate(std::pow(tm.z, 1.0f / 2.2f));
      ldr_rgba[k + 3] = 1.0f;
    }
}
void orc_fetch_ldr(void* p, float* rgba) {
  Ctx* c = (Ctx*)p;
  std::vector<float> hdr((size_t)c->scene.W * c->scene.H * 4);
  orc_fetch_hdr(p, hdr.data());
  orc_tonemap(p, hdr.data(), rgba);
}

// ------------------------------------------------------------------ unit probes for tests
// raytrace() in voxel space for n rays (raytracer.py:72-155)
void orc_raytrace(void* p, int n, const float* o, const float* d, float* t, int* cell, float* normal, int* iters) {
  Ctx* c = (Ctx*)p;
#pragma omp parallel for
  for (int i = 0; i < n; i++) {
    RayHit h = raytrace(c->scene, V3{o[3 * i], o[3 * i + 1], o[3 * i + 2]}, V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]}, kEps, kInf,
                        nullptr);
    t[i] = h.t;
    cell[3 * i] = h.cell.x, cell[3 * i + 1] = h.cell.y, cell[3 * i + 2] = h.cell.z;
    normal[3 * i] = h.normal.x, normal[3 * i + 1] = h.normal.y, normal[3 * i + 2] = h.normal.z;
    iters[i] = h.iters;
  }
}
// occupancy bit of (x,y,z) at lod
int orc_occupancy(void* p, int x, int y, int z, int lod) {
  Ctx* c = (Ctx*)p;
  if (lod < 0 || lod >= c->scene.n_lods) return -1;
  return c->scene.occ_bit(x, y, z, lod) ? 1 : 0;
}
// BSDF probes: for n (mat_id, albedo, v, n, l, u3) tuples return
// out[n][12] = {eval_d rgb, eval_s rgb, pdf_disney, sample_dir xyz, sample_pdf, lobe} + brdf rgb -> 15 floats
void orc_bsdf_probe(void* p, int n, const int* mat_id, const float* albedo, const float* v, const float* nrm, const float* l,
                    const float* u3, float* out) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) {
    Mat m = mat_at(*c, mat_id[i]);
    m.base_col = V3{albedo[3 * i], albedo[3 * i + 1], albedo[3 * i + 2]};
    V3 vv{v[3 * i], v[3 * i + 1], v[3 * i + 2]}, nn{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]},
        ll{l[3 * i], l[3 * i + 1], l[3 * i + 2]};
    V3 tang, bitang;
    make_orthonormal_basis(nn, tang, bitang);
    V3 bd, bs;
    disney_evaluate_split(m, vv, nn, ll, tang, bitang, bd, bs);
    float pdf = pdf_disney(m, vv, nn, ll, tang, bitang);
    V3 brdf;
    float spdf;
    int lobe;
    V3 dir = sample_disney(m, vv, nn, tang, bitang, u3[3 * i], u3[3 * i + 1], u3[3 * i + 2], brdf, spdf, lobe);
    float* o = out + 15 * i;
    o[0] = bd.x, o[1] = bd.y, o[2] = bd.z, o[3] = bs.x, o[4] = bs.y, o[5] = bs.z, o[6] = pdf;
    o[7] = dir.x, o[8] = dir.y, o[9] = dir.z, o[10] = spdf, o[11] = (float)lobe, o[12] = brdf.x, o[13] = brdf.y, o[14] = brdf.z;
  }
}
// sky probes: project/unproject and table lookups for n directions
void orc_project_sky(int n, int S, const float* d, float* uv) {
  for (int i = 0; i < n; i++) {
    V2 t = project_sky(V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]}, 1.0f / (float)S);
    uv[2 * i] = t.x, uv[2 * i + 1] = t.y;
  }
}
void orc_unproject_sky(int n, int S, const float* uv, float* d) {
  for (int i = 0; i < n; i++) {
    V3 r = unproject_sky(V2{uv[2 * i], uv[2 * i + 1]}, 1.0f / (float)S);
    d[3 * i] = r.x, d[3 * i + 1] = r.y, d[3 * i + 2] = r.z;
  }
}
void orc_sample_sky_trans(void* p, int n, const float* d, float* out) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) {
    V3 r = sample_skybox_transmittance(c->scene.sky_trans, c->scene.sky_res, V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]});
    out[3 * i] = r.x, out[3 * i + 1] = r.y, out[3 * i + 2] = r.z;
  }
}
// shift() probe (pathtracer.py:672-812). in[n][28] = dst_pos 3, dst_normal 3, src_pos 3, rc_pos 3, rc_normal 3,
// rc_incident_dir 3, rc_incident_L 3, rc_NEE_dir 3, cached_jacobian_term, lobes, dst_mat_info bits, rc_mat_info bits;
// out[n][7] = diffuse 3, specular 3, jacobian * passed_checks
void orc_shift_probe(void* p, int n, const float* in, float* out) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) {
    const float* a = in + 28 * i;
    auto v3 = [&](int k) { return V3{a[k], a[k + 1], a[k + 2]}; };
    Reservoir src;
    src.z.rc_pos = v3(9), src.z.rc_normal = v3(12), src.z.rc_incident_dir = v3(15), src.z.rc_incident_L = v3(18), src.z.rc_NEE_dir = v3(21);
    src.z.cached_jacobian_term = a[24];
    src.z.lobes = (int)a[25];
    uint32_t dst_info, rc_info;
    std::memcpy(&dst_info, a + 26, 4);
    std::memcpy(&rc_info, a + 27, 4);
    src.z.rc_mat_info = rc_info;
    Mat dst_mat;
    int dst_id;
    decode_material(*c, dst_info, dst_mat, dst_id);
    V3 d, s;
    float j;
    shift_sample(*c, v3(0), v3(3), dst_mat, v3(6), src, d, s, j);
    float* o = out + 7 * i;
    o[0] = d.x, o[1] = d.y, o[2] = d.z, o[3] = s.x, o[4] = s.y, o[5] = s.z, o[6] = j;
  }
}
// Reservoir bookkeeping and the 56-byte record (reservoir.py:41-141) on caller-built samples with
// NON-zero vectors (the zero-vector markers are the part upstream leaves undefined):
// init -> input_sample(wA, A) -> input_sample(wB, B) -> update_cached_jacobian_term(x1) ->
// merge({z = B, M = M_other}, wm) -> finalize_without_M -> encode -> decode.
// in[n][53]: A (21) | B (21) | wA wB uA uB | x1 (3) | wm um | M_other weight_other. A sample is
// F, rc_pos, rc_normal, rc_incident_dir, rc_incident_L, rc_NEE_dir (3 each), rc_mat_info (bits),
// cached_jacobian_term, lobes. out[n][28]: selA selB selM M weight (before encode) | decoded M W |
// decoded sample in the same 21-float layout.
void orc_reservoir_probe(int n, const float* in, float* out) {
  auto sample_of = [](const float* a) {
    Sample z;
    auto v3 = [&](int k) { return V3{a[k], a[k + 1], a[k + 2]}; };
    z.F = v3(0), z.rc_pos = v3(3), z.rc_normal = v3(6), z.rc_incident_dir = v3(9), z.rc_incident_L = v3(12), z.rc_NEE_dir = v3(15);
    std::memcpy(&z.rc_mat_info, a + 18, 4);
    z.cached_jacobian_term = a[19];
    z.lobes = (int)a[20];
    return z;
  };
  for (int i = 0; i < n; i++) {
    const float* a = in + 53 * i;
    float* o = out + 28 * i;
    const Sample A = sample_of(a), B = sample_of(a + 21);
    Reservoir r;
    o[0] = r.input_sample(a[42], A, a[44]) ? 1.0f : 0.0f;
    o[1] = r.input_sample(a[43], B, a[45]) ? 1.0f : 0.0f;
    r.update_cached_jacobian_term(V3{a[46], a[47], a[48]});
    Reservoir other;
    other.z = B, other.M = a[51], other.weight = a[52];
    o[2] = r.merge(other, a[49], a[50]) ? 1.0f : 0.0f;
    r.finalize_without_M();
    o[3] = r.M, o[4] = r.weight;
    const Reservoir d = decode_reservoir(encode_reservoir(r));
    o[5] = d.M, o[6] = d.weight;
    float* z = o + 7;
    auto put = [&](int k, V3 v) { z[k] = v.x, z[k + 1] = v.y, z[k + 2] = v.z; };
    put(0, d.z.F), put(3, d.z.rc_pos), put(6, d.z.rc_normal), put(9, d.z.rc_incident_dir), put(12, d.z.rc_incident_L), put(15, d.z.rc_NEE_dir);
    std::memcpy(z + 18, &d.z.rc_mat_info, 4);
    z[19] = d.z.cached_jacobian_term;
    z[20] = (float)d.z.lobes;
  }
}
// spatial_GRIS (pathtracer.py:815-989) on caller-built buffers: per pixel a reservoir (sample in the
// 21-float layout of orc_reservoir_probe + M + W = 23 floats, packed with encode_reservoir as
// render() would, :607), a G-buffer entry (position 3, octahedral normal 2, material info bits, sky
// flag = 7 floats) and the canonical sample's diffuse / specular integrand (:631-632). Runs the
// pass for the listed pixels (index = v * W + u); out[k] = colour buffers after the pass (6 floats).
void orc_gris_probe(void* p, uint32_t frame, const float* samples, const float* gbuf, const float* col_d, const float* col_s, int n,
                    const int* pixels, float* out) {
  Ctx* c = (Ctx*)p;
  const size_t npx = (size_t)c->scene.W * c->scene.H;
  c->reservoirs.resize(npx), c->gbuf.resize(npx), c->col_d.resize(npx), c->col_s.resize(npx);
  for (size_t i = 0; i < npx; i++) {
    const float* a = samples + 23 * i;
    auto v3 = [&](const float* q) { return V3{q[0], q[1], q[2]}; };
    Reservoir r;
    r.z.F = v3(a), r.z.rc_pos = v3(a + 3), r.z.rc_normal = v3(a + 6), r.z.rc_incident_dir = v3(a + 9), r.z.rc_incident_L = v3(a + 12),
    r.z.rc_NEE_dir = v3(a + 15);
    std::memcpy(&r.z.rc_mat_info, a + 18, 4);
    r.z.cached_jacobian_term = a[19], r.z.lobes = (int)a[20], r.M = a[21], r.weight = a[22];
    c->reservoirs[i] = encode_reservoir(r);
    const float* g = gbuf + 7 * i;
    GBufferPx& G = c->gbuf[i];
    G.position = v3(g), G.n_oct[0] = g[3], G.n_oct[1] = g[4], G.sky = g[6] != 0.0f;
    std::memcpy(&G.mat_info, g + 5, 4);
    c->col_d[i] = v3(col_d + 3 * i), c->col_s[i] = v3(col_s + 3 * i);
  }
  for (int k = 0; k < n; k++) {
    V3 d, s;
    spatial_gris_pixel(*c, pixels[k] % c->scene.W, pixels[k] / c->scene.W, frame, d, s);
    float* o = out + 6 * k;
    o[0] = d.x, o[1] = d.y, o[2] = d.z, o[3] = s.x, o[4] = s.y, o[5] = s.z;
  }
}
// The ReSTIR branch of render() (pathtracer.py:355-632 with USE_RESTIR_PT = True) for every pixel of
// one sample: the reservoir BEFORE encode() (23 floats: the 21-float sample layout of
// orc_reservoir_probe + M + W), the G-buffer entry (7 floats as in orc_gris_probe) and the canonical
// diffuse / specular integrands.
void orc_restir_render_probe(void* p, int sample, float* samples, float* gbuf, float* col_d, float* col_s) {
  Ctx* c = (Ctx*)p;
  const Scene& s = c->scene;
  set_jitter(*c, (uint32_t)sample);
#pragma omp parallel for schedule(dynamic, 2)
  for (int v = 0; v < s.H; v++)
    for (int u = 0; u < s.W; u++) {
      const size_t i = (size_t)v * s.W + u;
      Reservoir r;
      GBufferPx g;
      V3 d, sp;
      trace_path_restir(*c, u, v, (uint32_t)sample, nullptr, r, g, d, sp);
      float* a = samples + 23 * i;
      auto put = [](float* q, V3 x) { q[0] = x.x, q[1] = x.y, q[2] = x.z; };
      put(a, r.z.F), put(a + 3, r.z.rc_pos), put(a + 6, r.z.rc_normal), put(a + 9, r.z.rc_incident_dir), put(a + 12, r.z.rc_incident_L),
          put(a + 15, r.z.rc_NEE_dir);
      std::memcpy(a + 18, &r.z.rc_mat_info, 4);
      a[19] = r.z.cached_jacobian_term, a[20] = (float)r.z.lobes, a[21] = r.M, a[22] = r.weight;
      float* q = gbuf + 7 * i;
      put(q, g.position), q[3] = g.n_oct[0], q[4] = g.n_oct[1], q[6] = g.sky ? 1.0f : 0.0f;
      std::memcpy(q + 5, &g.mat_info, 4);
      put(col_d + 3 * i, d), put(col_s + 3 * i, sp);
    }
}
// sample_skybox (atmos.py:94-115) with the three jitter numbers supplied
void orc_sample_skybox(void* p, int n, const float* d, const float* jitter, float* scat, float* trans) {
  Ctx* c = (Ctx*)p;
  for (int i = 0; i < n; i++) {
    V3 sc, tr;
    sample_skybox(c->scene.sky_scatter, c->scene.sky_trans, c->scene.sky_res, V3{d[3 * i], d[3 * i + 1], d[3 * i + 2]}, jitter[3 * i],
                  jitter[3 * i + 1], jitter[3 * i + 2], sc, tr);
    scat[3 * i] = sc.x, scat[3 * i + 1] = sc.y, scat[3 * i + 2] = sc.z;
    trans[3 * i] = tr.x, trans[3 * i + 1] = tr.y, trans[3 * i + 2] = tr.z;
  }
}
float orc_rnd(uint32_t pixel, uint32_t sample, uint32_t seed, uint32_t dim) { return rnd(path_key(pixel, sample, seed), dim); }
uint16_t orc_f32_to_f16(float f) { return f32_to_f16_bits(f); }
float orc_f16_to_f32(uint16_t h) { return f16_bits_to_f32(h); }
int orc_num_threads() { return omp_get_max_threads(); }
// launchers such as torchrun export OMP_NUM_THREADS=1: the CPU baseline legs of bench.py ask for the cores explicitly
void orc_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

}  // extern "C"
