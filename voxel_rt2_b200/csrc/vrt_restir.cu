// Spatial reservoir resampling (GRIS) for the ReSTIR-PT mode. Replaces, in
// renderer/pathtracer.py: shift() :672-812 and spatial_GRIS :815-989 (pass 0 of 1, radius 24,
// 32 taps), followed by the static-camera accumulation of :1185-1303 (running mean == sum).
// One thread per pixel in 8x4 tiles; per tap it reads the neighbour's 56-byte reservoir and
// 24-byte G-buffer record (the pass is the bandwidth-heavy kernel of the mode: 32 x 80 B per
// pixel, mostly L2 hits because 48x48-pixel footprints of neighbouring pixels overlap).
// Pinned holes of the upstream code: see DESIGN.md "ReSTIR pins" and oracle/orestir.h.
#include "vrt_bsdf.cuh"
#include "vrt_internal.h"
#include "vrt_restir.cuh"
#include "vrt_sky.cuh"
#include "vrt_trace.cuh"

namespace {

#define RADIANCE_CLAMP 300.0f
HD f3 firefly_filter(f3 v) { return clamp3(v, 0.0f, RADIANCE_CLAMP); }
HD float power_heuristic(float a, float b) {
  float a_sqr = a * a;
  return __fdividef(a_sqr, fmaxf(a_sqr + b * b, 1e-4f));
}
HD bool bad3(f3 c) { return isbad(c.x) || isbad(c.y) || isbad(c.z) || c.x < 0.0f || c.y < 0.0f || c.z < 0.0f; }
HD uint32_t hash3(uint32_t x, uint32_t y, uint32_t z) {  // math_utils.py:217-229
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

struct GrisCtx {
  const float4* mats;
  const float* unorm8;
  f3 cam_pos, light_dir, sun_rad;
  float light_cos_max, light_pdf_axis;
};

HD Mat decode_material(const GrisCtx& G, uint32_t enc, int& mat_id) {  // math_utils.py:238-247
  mat_id = (int)(enc & 255u);
  Mat m = load_mat(G.mats, mat_id);
  m.base_col = f3{G.unorm8[(enc >> 8) & 255u], G.unorm8[(enc >> 16) & 255u], G.unorm8[(enc >> 24) & 255u]};
  return m;
}

// Per-sample and per-destination terms of shift() that do not depend on the other end of the
// reconnection: hoisted so the 32-tap loop computes them once for the centre pixel.
struct RcPre {   // reconnection vertex of a sample (pathtracer.py:676-689)
  Mat rc_mat;
  int rc_mat_id;
  f3 tang, bitang;
  const float4* sky_T;  // this reservoir's entry of RestirBuffers::rc_skyT (loaded only if the shift survives the early-out)
  bool esc, last, nee;
};
struct DstPre {  // primary vertex the sample is shifted to (pathtracer.py:731-733)
  f3 tang, bitang, view;
};
HD f3 rc_sky_T(const float4* sky_trans, int sky_res, const RSample& z) {  // shift(), pathtracer.py:780
  return sky_fetch(sky_trans, sky_tap(sky_res, project_sky(z.rc_NEE_dir, 1.0f / (float)sky_res)));
}
HD RcPre prep_rc(const GrisCtx& G, const RSample& z, const float4* sky_T) {
  RcPre r;
  r.esc = is_vec_zero(z.rc_normal), r.last = is_vec_zero(z.rc_incident_dir), r.nee = !is_vec_zero(z.rc_NEE_dir);
  make_orthonormal_basis(z.rc_normal, r.tang, r.bitang);
  r.rc_mat = decode_material(G, z.rc_mat_info, r.rc_mat_id);
  r.sky_T = sky_T;
  return r;
}
HD DstPre prep_dst(const GrisCtx& G, f3 dst_pos, f3 dst_normal) {
  DstPre d;
  make_orthonormal_basis(dst_normal, d.tang, d.bitang);
  d.view = normalize(G.cam_pos - dst_pos);
  return d;
}

// pathtracer.py:672-812. The geometric terms (N.L tests, Jacobian) are evaluated first: when
// their product is zero every use of the shifted integrand is multiplied by it (p_hat * jacobian
// in the merge weight, center_p_hat in the canonical weight), so the BSDF work is skipped.
HD void shift_sample(const GrisCtx& G, f3 dst_pos, f3 dst_normal, const Mat& dst_material, const DstPre& D, const RReservoir& src, const RcPre& R,
                     f3& diffuse, f3& specular, float& jacobian_out) {
  const RSample& z = src.z;
  const f3 dir = R.esc ? z.rc_pos : normalize(z.rc_pos - dst_pos);
  float passed_checks = 1.0f;
  if (dot(dst_normal, dir) < 1e-5f || (!R.esc && dot(z.rc_normal, -dir) < 1e-5f)) passed_checks = 0.0f;
  float jacobian = 1.0f;
  if (!R.esc) {
    const f3 y = z.rc_pos - dst_pos;
    jacobian = z.cached_jacobian_term * fdiv(fabsf(dot(normalize(y), z.rc_normal)), dot(y, y));
  }
  if (jacobian < 0.0f || isbad(jacobian)) jacobian = 0.0f;
  jacobian_out = jacobian * passed_checks;
  diffuse = mk3(0.0f), specular = mk3(0.0f);
  if (jacobian_out == 0.0f) return;
  f3 contrib = mk3(0.0f);
  if (!R.last && !R.esc) {
    f3 rc_brdf = disney_evaluate_lobewise(R.rc_mat, -dir, z.rc_normal, z.rc_incident_dir, R.tang, R.bitang, z.lobes / 10);
    rc_brdf *= saturate(dot(z.rc_normal, z.rc_incident_dir));
    const float dst_rc_pdf = pdf_disney_lobewise(R.rc_mat, -dir, z.rc_normal, z.rc_incident_dir, R.tang, R.bitang, z.lobes / 10);
    const float lp = cone_sample_pdf(G.light_cos_max, dot(G.light_dir, z.rc_incident_dir));
    const float w = power_heuristic(dst_rc_pdf, lp * (R.nee ? 1.0f : 0.0f));
    contrib += firefly_filter((w * rc_brdf) * frcp(dst_rc_pdf) * z.rc_incident_L);
  }
  if (R.esc) contrib += firefly_filter(z.rc_incident_L);
  if (R.nee && !R.esc) {
    f3 bd, bs;
    float lpdf;
    eval_and_pdf(R.rc_mat, -dir, z.rc_normal, z.rc_NEE_dir, R.tang, R.bitang, bd, bs, lpdf);
    const f3 rc_nee_brdf = (bd + bs) * saturate(dot(z.rc_normal, z.rc_NEE_dir));
    const float w = power_heuristic(G.light_pdf_axis, lpdf);
    const float4 t4 = __ldg(R.sky_T);  // looked up once per reservoir by k_rc_sky (1,1,1 without the physical sky)
    const f3 sky_T{t4.x, t4.y, t4.z};
    contrib += firefly_filter((w * rc_nee_brdf) * sky_T * G.sun_rad);
  }
  if (R.rc_mat_id == 2) contrib += R.rc_mat.base_col;
  f3 pd, ps;
  disney_evaluate_lobewise_split(dst_material, D.view, dst_normal, dir, D.tang, D.bitang, z.lobes % 10, pd, ps);
  const float cosd = saturate(dot(dst_normal, dir));
  diffuse = (pd * cosd) * contrib;
  specular = (ps * cosd) * contrib;
}

HD void load_reservoir(const uint2* __restrict__ base, size_t pidx, const float* unorm8, RReservoir& r) {
  uint32_t w[14];
  const uint2* src = base + pidx * 7;
#pragma unroll
  for (int i = 0; i < 7; i++) {
    uint2 t = __ldg(src + i);
    w[2 * i] = t.x, w[2 * i + 1] = t.y;
  }
  decode_reservoir(w, unorm8, r);
}

// Sun transmittance at every reservoir's reconnection vertex (the sample_skybox_transmittance call
// of shift(), pathtracer.py:780). shift() runs up to 64 times per pixel in spatial_GRIS but the
// value depends on the reservoir alone, so it is looked up once per pixel here, between the path
// kernel and k_gris, which then reads 16 bytes per surviving tap instead of projecting the
// direction and fetching four texels of the 236 MB table.
__global__ void __launch_bounds__(128) k_rc_sky(const __grid_constant__ Params P, RestirBuffers RB, size_t first_px, size_t n_px) {
  __shared__ float s_unorm[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_unorm[i] = xdiv((float)i, 255.0f);
  __syncthreads();
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_px) return;
  const size_t pidx = first_px + k;  // the rows this context renders (the whole frame unless row-sharded)
  RReservoir r;
  load_reservoir(RB.reservoirs, pidx, s_unorm, r);
  f3 T = mk3(1.0f);
  if (P.use_sky && !is_vec_zero(r.z.rc_NEE_dir) && !is_vec_zero(r.z.rc_normal)) T = rc_sky_T(P.sky_trans, P.sky_res, r.z);
  RB.rc_skyT[pidx] = make_float4(T.x, T.y, T.z, 0.0f);
}

#ifndef VRT_GRIS_THREADS
#define VRT_GRIS_THREADS 512  // ONE CTA of 16 warps per SM instead of four of 4: 5.19 -> 3.80 ms at 1080p (profiles/r02g_ab_gris.log)
#endif
#ifndef VRT_GRIS_MIN_BLOCKS
#define VRT_GRIS_MIN_BLOCKS (512 / VRT_GRIS_THREADS)  // 128 registers, 200 B of spills: measured 3.6 % faster than fewer resident warps
#endif
// The kernel is instruction-fetch bound (ncu r02f: 4.5 no-instruction stalls per issued instruction, issue slots 35 %):
// each tap-loop body is ~17 KB of BSDF code and the resident warps drift apart inside it. Launching the 16 resident warps
// of an SM as ONE CTA (they start together and stay loosely in step) took 27 % off the kernel. VRT_GRIS_SYNC = k
// additionally puts a CTA barrier after every k-th tap of both loops; measured slower for every k (the barrier waits
// for the slowest warp of 16), so it is off.
#ifndef VRT_GRIS_SYNC
#define VRT_GRIS_SYNC 0
#endif
__global__ void __launch_bounds__(VRT_GRIS_THREADS, VRT_GRIS_MIN_BLOCKS) k_gris(const __grid_constant__ Params P, RestirBuffers RB, uint32_t frame, int upper_in_smem, int fixed_words) {
  extern __shared__ uint32_t smem[];
  // same staging layout as the render kernels: materials, UNORM8 table, upper pyramid
  float4* s_mats = reinterpret_cast<float4*>(smem);
  for (int i = threadIdx.x; i < 128 * MAT_ROW_F4; i += blockDim.x) s_mats[i] = P.mats[i];
  float* s_unorm = reinterpret_cast<float*>(smem + 128 * MAT_ROW_F4 * 4);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_unorm[i] = xdiv((float)i, 255.0f);
  const uint32_t* upper = P.upper;
  if (upper_in_smem) {
    uint32_t* s_upper = smem + fixed_words;
    for (int i = threadIdx.x; i < P.upper_words; i += blockDim.x) s_upper[i] = P.upper[i];
    upper = s_upper;
  }
  __syncthreads();

  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
#if VRT_GRIS_SYNC
  const bool in_range = warp < P.n_tiles;  // every thread of the CTA has to reach the barriers
  const int tile = P.tile_rank + P.tile_n * (in_range ? warp : 0);
#else
  if (warp >= P.n_tiles) return;
  const int tile = P.tile_rank + P.tile_n * warp;
#endif
  const int u = (tile % P.tiles_x) * 8 + (lane & 7), v = (tile / P.tiles_x) * 4 + (lane >> 3);
  const int W = P.W, H = P.H;
  const size_t pidx = (size_t)v * W + u;

  GrisCtx G;
  G.mats = s_mats, G.unorm8 = s_unorm;
  G.cam_pos = P.cam_pos, G.light_dir = P.light_dir, G.sun_rad = P.light_weight * P.light_color;
  G.light_cos_max = P.light_cos_max, G.light_pdf_axis = cone_sample_pdf(P.light_cos_max, 1.0f);

  const float max_radius = 24.0f;
  const int max_taps = 32;
  const uint32_t key = path_key((uint32_t)pidx, frame, P.seed);
  RReservoir center;
  load_reservoir(RB.reservoirs, pidx, s_unorm, center);
  const float4 gp = RB.gpos[pidx];
  // a pixel whose primary ray escaped passes its sample through (pathtracer.py:854-856) and takes no part in the tap loops
#if VRT_GRIS_SYNC
  const bool live = in_range && gp.w == 0.0f;
#else
  const bool live = gp.w == 0.0f;
#endif
  f3 out_d = center.z.F, out_s = mk3(0.0f);
  const uint2 ga = RB.gattr[pidx];
  const uint32_t seed = hash3((uint32_t)u >> 3, (uint32_t)v >> 3, frame * 2u);
  const float angle_shift = (float)((seed & 0x007FFFFFu) | 0x3F800000u) / 4294967295.0f * VRT_PI;
  const float radius_shift = rnd(key, 65);
  RReservoir out;
  rinit(out);
  const f3 center_x1{gp.x, gp.y, gp.z};
  const float center_dist = length(center_x1 - P.cam_pos);
  const f3 center_n1 = decode_unit_vector_3x16(h16val(ga.x), h16val(ga.x >> 16));
  int center_mat_id;
  const Mat center_mat = decode_material(G, ga.y, center_mat_id);
  int valid_samples = 0;
  float canonical_mis_weight = 1.0f;
  f3 chosen_F_d = mk3(0.0f), chosen_F_s = mk3(0.0f);
  const float center_F_lum = luminance(center.z.F);
  const RcPre center_rc = prep_rc(G, center.z, RB.rc_skyT + pidx);
  const DstPre center_dst = prep_dst(G, center_x1, center_n1);
  // The tap loop is split in two so that each loop body inlines ONE copy of shift() (~17 KB of
  // BSDF code): with both shifts in one body the loop was 40 KB, beyond the SM's instruction
  // cache, and the kernel was instruction-fetch bound (ncu: 3.0 no-instruction stalls per issue).
  // Pass A shifts the centre sample to every accepted neighbour (needs the neighbour's G-buffer
  // record only) and keeps p_hat(centre -> tap) and the tap's pixel index in local memory;
  // pass B shifts each neighbour's sample to the centre and merges, in the reference's tap order.
  // Each body sits in a do { } while (0) so that `continue` leaves the tap, not the iteration: with
  // VRT_GRIS_SYNC every thread of the CTA reaches the barrier that closes the iteration.
#if VRT_GRIS_SYNC
#define GRIS_TAP_SYNC(i)                                \
  do {                                                  \
    if (((i) + 1) % VRT_GRIS_SYNC == 0) __syncthreads(); \
  } while (0)
#else
#define GRIS_TAP_SYNC(i) \
  do {                   \
  } while (0)
#endif
  float tap_center_p_hat[32];
  int tap_index[32];
#pragma unroll 1
  for (int i = 0; i < max_taps; i++) {
    do {
      tap_index[i] = -1;
      if (!live) continue;
      const float golden_angle = 2.399963229728f;
      const float angle = ((float)i + angle_shift) * golden_angle;
      const float offset_radius = sqrtf(((float)i + radius_shift) / (float)max_taps) * max_radius;
      float sn, cs;
      sincosf(angle, &sn, &cs);
      const int ox = (int)(cs * offset_radius), oy = (int)(sn * offset_radius);
      if (ox == 0 && oy == 0) continue;
      const int tu = u + ox, tv = v + oy;
      if (tu < 0 || tv < 0 || tu >= W || tv >= H) continue;
      const size_t ti = (size_t)tv * W + tu;
      const float4 ngp = __ldg(RB.gpos + ti);
      const uint2 nga = __ldg(RB.gattr + ti);
      if (ngp.w != 0.0f) continue;
      const f3 neighbour_n1 = decode_unit_vector_3x16(h16val(nga.x), h16val(nga.x >> 16));
      const f3 neighbour_x1{ngp.x, ngp.y, ngp.z};
      const float neighbour_dist = length(neighbour_x1 - P.cam_pos);
      if (fabsf(neighbour_dist - center_dist) > 0.1f * center_dist || dot(center_n1, neighbour_n1) < 0.5f) continue;
      int neighbour_mat_id;
      const Mat neighbour_mat = decode_material(G, nga.y, neighbour_mat_id);
      f3 c_d, c_s;
      float c_jacobian;
      shift_sample(G, neighbour_x1, neighbour_n1, neighbour_mat, prep_dst(G, neighbour_x1, neighbour_n1), center, center_rc, c_d, c_s, c_jacobian);
      tap_center_p_hat[i] = luminance(c_d + c_s) * c_jacobian;
      tap_index[i] = (int)ti;
    } while (0);
    GRIS_TAP_SYNC(i);
  }
#pragma unroll 1
  for (int i = 0; i < max_taps; i++) {
    do {
      const int ti = tap_index[i];
      if (ti < 0) continue;
      RReservoir nb;
      load_reservoir(RB.reservoirs, (size_t)ti, s_unorm, nb);
      f3 s_d, s_s;
      float jacobian;
      shift_sample(G, center_x1, center_n1, center_mat, center_dst, nb, prep_rc(G, nb.z, RB.rc_skyT + tap_index[i]), s_d, s_s, jacobian);
      const float center_p_hat = tap_center_p_hat[i];
      float canonical_weight = center_p_hat * nb.M;
      canonical_weight = canonical_weight / (center_p_hat * nb.M + center_F_lum * center.M / (float)max_taps);
      canonical_mis_weight += 1.0f - canonical_weight;
      const float p_hat = luminance(s_d + s_s);
      // the neighbour's own target value: upstream approximates it by the shifted one (pathtracer.py:936); with temporal
      // reuse on, the integrand the neighbour's reservoir was stored with is used (see k_temporal)
      const float p_hat_from_neighbour = (RB.temporal ? luminance(nb.z.F) : p_hat) / jacobian;
      float neighbour_mis_weight = p_hat_from_neighbour * nb.M;
      neighbour_mis_weight = neighbour_mis_weight / (p_hat_from_neighbour * nb.M + p_hat * center.M / (float)max_taps);
      if (isbad(neighbour_mis_weight)) neighbour_mis_weight = 0.0f;
      // merge (reservoir.py:76-86)
      const float in_w = nb.weight * p_hat * jacobian * neighbour_mis_weight;
      out.M += nb.M;
      if (in_w > 0.0f) {
        out.weight += in_w;
        if (rnd(key, 66u + (uint32_t)i) * out.weight <= in_w) {
          out.z = nb.z;
          out.z.F = s_d + s_s;
          chosen_F_d = s_d;
          chosen_F_s = s_s;
        }
      }
      valid_samples += 1;
    } while (0);
    GRIS_TAP_SYNC(i);
  }
#undef GRIS_TAP_SYNC
#if VRT_GRIS_SYNC
  if (!in_range) return;
#endif
  if (live) {
    // visibility of the resampled reconnection (pathtracer.py:957-965)
    bool force_add_canonical = false;
    if (out.weight > 0.0f) {
      const bool esc = is_vec_zero(out.z.rc_normal);
      const f3 dir = esc ? out.z.rc_pos : normalize(out.z.rc_pos - center_x1);
      const f3 org = center_x1 + center_n1 * (0.003f * center_dist);
      Hit sh = next_hit<false>(P, upper, s_unorm, org, dir, true, nullptr, nullptr);
      const float actual_dist = esc ? VRT_INF : length(center_x1 - out.z.rc_pos);
      if (sh.closest < VRT_INF && fabsf(sh.closest - actual_dist) > 0.1f * actual_dist) {
        out.weight = 0.0f;
        force_add_canonical = true;
      }
    }
    {
      const float in_w = center.weight * center_F_lum * canonical_mis_weight;
      out.M += center.M;
      if (in_w > 0.0f) {
        out.weight += in_w;
        if (rnd(key, 98) * out.weight <= in_w || force_add_canonical) {
          out.z = center.z;
          const float4 cd = RB.col_d[pidx], cs4 = RB.col_s[pidx];
          chosen_F_d = f3{cd.x, cd.y, cd.z};
          chosen_F_s = f3{cs4.x, cs4.y, cs4.z};
        }
      }
    }
    const float p_hat = luminance(out.z.F);  // finalize_without_M, then / (valid + 1)
    float Wt = p_hat < 1e-6f ? 0.0f : out.weight / p_hat;
    Wt = Wt / (float)(valid_samples + 1);
    const f3 emission = center_mat_id == 2 ? center_mat.base_col : mk3(0.0f);
    const float Wc = clampf(Wt, 0.0f, 50.0f);
    out_d = chosen_F_d * Wc + emission;
    out_s = chosen_F_s * Wc;
  }
  if (bad3(out_d)) out_d = mk3(0.0f);
  if (bad3(out_s)) out_s = mk3(0.0f);
  float4 a = P.accum[pidx];
  a.x += out_d.x + out_s.x, a.y += out_d.y + out_s.y, a.z += out_d.z + out_s.z, a.w += 1.0f;
  P.accum[pidx] = a;
}

// Temporal reservoir reuse (vrt_set_restir_temporal). NOT in the reference — its second reservoir slot is written at
// pathtracer.py:989 and never read — but BASELINE.json configs[3] asks for "temporal+spatial resampling per frame",
// so the pass is built from the reference's own primitives: it runs between the path kernel and k_gris and is
// spatial_GRIS (:815-989) with ONE tap, the same pixel's reservoir of the previous frame: the same similarity test
// (:911), both reconnection shifts (:672-812), pairwise MIS with max_taps = 1 (:928-944), merge (reservoir.py:76-86),
// a visibility ray for the resampled reconnection (:957-965), the canonical merge, finalize_without_M / 2.
// Deliberate differences (oracle/oracle.cpp temporal_reuse_pixel states the same algorithm and the measurements):
// the history slot holds this pass's output (a per-pixel chain; the spatial pass leaks shadowed sun samples and must
// not be fed back), an escape / sun sample is occluded whenever its ray hits anything, the history's own target
// value is the integrand it was stored with, and its confidence is capped at 20 x the canonical M.
// One thread per pixel, no neighbour reads: the pixel's reservoir, canonical integrands and sun transmittance are
// rewritten in place, the history slot and the previous G-buffer record are updated for the next frame.
#define VRT_TEMPORAL_M_CAP 20.0f
#ifndef VRT_TEMPORAL_THREADS
#define VRT_TEMPORAL_THREADS 128
#endif
__global__ void __launch_bounds__(VRT_TEMPORAL_THREADS, 512 / VRT_TEMPORAL_THREADS) k_temporal(const __grid_constant__ Params P, RestirBuffers RB, uint32_t frame, int hist_valid, int upper_in_smem,
                                                     int fixed_words) {
  extern __shared__ uint32_t smem[];
  float4* s_mats = reinterpret_cast<float4*>(smem);
  for (int i = threadIdx.x; i < 128 * MAT_ROW_F4; i += blockDim.x) s_mats[i] = P.mats[i];
  float* s_unorm = reinterpret_cast<float*>(smem + 128 * MAT_ROW_F4 * 4);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_unorm[i] = xdiv((float)i, 255.0f);
  const uint32_t* upper = P.upper;
  if (upper_in_smem) {
    uint32_t* s_upper = smem + fixed_words;
    for (int i = threadIdx.x; i < P.upper_words; i += blockDim.x) s_upper[i] = P.upper[i];
    upper = s_upper;
  }
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.n_tiles) return;
  const int tile = P.tile_rank + P.tile_n * warp;
  const int u = (tile % P.tiles_x) * 8 + (lane & 7), v = (tile / P.tiles_x) * 4 + (lane >> 3);
  const size_t pidx = (size_t)v * P.W + u;

  const float4 gp = RB.gpos[pidx];
  const uint2 ga = RB.gattr[pidx];
  float4 pgp = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
  uint2 pga = make_uint2(0u, 0u);
  if (hist_valid) pgp = RB.hist_gpos[pidx], pga = RB.hist_gattr[pidx];
  RB.hist_gpos[pidx] = gp;
  RB.hist_gattr[pidx] = ga;
  bool reuse = gp.w == 0.0f && pgp.w == 0.0f;
  RReservoir center, prev;
  f3 center_x1 = mk3(0.0f), prev_x1 = mk3(0.0f), center_n1 = mk3(0.0f), prev_n1 = mk3(0.0f);
  float center_dist = 0.0f;
  if (reuse) {
    load_reservoir(RB.reservoirs, pidx, s_unorm, center);
    load_reservoir(RB.hist_res, pidx, s_unorm, prev);
    prev.M = fminf(prev.M, VRT_TEMPORAL_M_CAP * center.M);
    center_x1 = f3{gp.x, gp.y, gp.z}, prev_x1 = f3{pgp.x, pgp.y, pgp.z};
    center_dist = length(center_x1 - P.cam_pos);
    const float prev_dist = length(prev_x1 - P.cam_pos);
    center_n1 = decode_unit_vector_3x16(h16val(ga.x), h16val(ga.x >> 16));
    prev_n1 = decode_unit_vector_3x16(h16val(pga.x), h16val(pga.x >> 16));
    reuse = prev.M > 0.0f && !(fabsf(prev_dist - center_dist) > 0.1f * center_dist) && !(dot(center_n1, prev_n1) < 0.5f);
  }
  if (!reuse) {  // nothing to reuse: the canonical reservoir starts the chain
#pragma unroll
    for (int i = 0; i < 7; i++) RB.hist_res[pidx * 7 + i] = RB.reservoirs[pidx * 7 + i];
    RB.hist_skyT[pidx] = RB.rc_skyT[pidx];
    return;
  }
  GrisCtx G;
  G.mats = s_mats, G.unorm8 = s_unorm;
  G.cam_pos = P.cam_pos, G.light_dir = P.light_dir, G.sun_rad = P.light_weight * P.light_color;
  G.light_cos_max = P.light_cos_max, G.light_pdf_axis = cone_sample_pdf(P.light_cos_max, 1.0f);
  const uint32_t key = path_key((uint32_t)pidx, frame, P.seed);
  int center_mat_id, prev_mat_id;
  const Mat center_mat = decode_material(G, ga.y, center_mat_id);
  const Mat prev_mat = decode_material(G, pga.y, prev_mat_id);
  f3 c_d, c_s, s_d, s_s;
  float c_jacobian, jacobian;
  shift_sample(G, prev_x1, prev_n1, prev_mat, prep_dst(G, prev_x1, prev_n1), center, prep_rc(G, center.z, RB.rc_skyT + pidx), c_d, c_s, c_jacobian);
  shift_sample(G, center_x1, center_n1, center_mat, prep_dst(G, center_x1, center_n1), prev, prep_rc(G, prev.z, RB.hist_skyT + pidx), s_d, s_s, jacobian);
  const float center_F_lum = luminance(center.z.F);
  const float center_p_hat_at_prev = luminance(c_d + c_s) * c_jacobian;
  float canonical_weight = center_p_hat_at_prev * prev.M;
  canonical_weight = canonical_weight / (center_p_hat_at_prev * prev.M + center_F_lum * center.M);
  if (isbad(canonical_weight)) canonical_weight = 0.0f;
  const float canonical_mis_weight = 1.0f + (1.0f - canonical_weight);
  const float p_hat = luminance(s_d + s_s);
  const float p_hat_from_prev = luminance(prev.z.F) / jacobian;
  float prev_mis_weight = p_hat_from_prev * prev.M;
  prev_mis_weight = prev_mis_weight / (p_hat_from_prev * prev.M + p_hat * center.M);
  if (isbad(prev_mis_weight)) prev_mis_weight = 0.0f;
  RReservoir out;
  rinit(out);
  f3 chosen_F_d = mk3(0.0f), chosen_F_s = mk3(0.0f);
  bool from_history = false;
  {
    const float in_w = prev.weight * p_hat * jacobian * prev_mis_weight;
    out.M += prev.M;
    if (in_w > 0.0f) {
      out.weight += in_w;
      if (rnd(key, 99) * out.weight <= in_w) {
        out.z = prev.z;
        out.z.F = s_d + s_s;
        chosen_F_d = s_d, chosen_F_s = s_s;
        from_history = true;
      }
    }
  }
  bool force_add_canonical = false;
  if (out.weight > 0.0f) {
    const bool esc = is_vec_zero(out.z.rc_normal);
    const f3 dir = esc ? out.z.rc_pos : normalize(out.z.rc_pos - center_x1);
    const f3 org = center_x1 + center_n1 * (0.003f * center_dist);
    Hit sh = next_hit<false>(P, upper, s_unorm, org, dir, true, nullptr, nullptr);
    const float actual_dist = esc ? VRT_INF : length(center_x1 - out.z.rc_pos);
    if (sh.closest < VRT_INF && (esc || fabsf(sh.closest - actual_dist) > 0.1f * actual_dist)) {
      out.weight = 0.0f;
      force_add_canonical = true;
    }
  }
  {
    const float in_w = center.weight * center_F_lum * canonical_mis_weight;
    out.M += center.M;
    if (in_w > 0.0f) {
      out.weight += in_w;
      if (rnd(key, 100) * out.weight <= in_w || force_add_canonical) {
        out.z = center.z;
        const float4 cd = RB.col_d[pidx], cs4 = RB.col_s[pidx];
        chosen_F_d = f3{cd.x, cd.y, cd.z};
        chosen_F_s = f3{cs4.x, cs4.y, cs4.z};
        from_history = false;
      }
    }
  }
  const float ph = luminance(out.z.F);  // finalize_without_M, then / (valid + 1)
  out.weight = (ph < 1e-6f ? 0.0f : out.weight / ph) / 2.0f;
  if (!is_vec_zero(out.z.rc_normal)) out.z.cached_jacobian_term = jacobian_term(out.z.rc_pos, out.z.rc_normal, center_x1);
  uint32_t w[14];
  encode_reservoir(out, w);
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const uint2 t = make_uint2(w[2 * i], w[2 * i + 1]);
    RB.reservoirs[pidx * 7 + i] = t;
    RB.hist_res[pidx * 7 + i] = t;
  }
  RB.col_d[pidx] = make_float4(chosen_F_d.x, chosen_F_d.y, chosen_F_d.z, 0.0f);
  RB.col_s[pidx] = make_float4(chosen_F_s.x, chosen_F_s.y, chosen_F_s.z, 0.0f);
  const float4 T = from_history ? RB.hist_skyT[pidx] : RB.rc_skyT[pidx];
  RB.rc_skyT[pidx] = T;
  RB.hist_skyT[pidx] = T;
}

}  // namespace

// The tiles of P form whole tile rows when tile_n == 1 (the whole frame, or the contiguous range of a row shard): the pixel
// range k_rc_sky covers.
static void rc_sky_range(const Params& P, size_t* first_px, size_t* n_px) {
  *first_px = (size_t)(P.tile_rank / P.tiles_x) * 4 * P.W;
  *n_px = (size_t)(P.n_tiles / P.tiles_x) * 4 * P.W;
}
cudaError_t vrt_launch_rc_sky(const Params& P, const RestirBuffers& RB, cudaStream_t st) {
  size_t f, n;
  rc_sky_range(P, &f, &n);
  if (n) k_rc_sky<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, RB, f, n);
  return cudaGetLastError();
}

cudaError_t vrt_launch_gris(const Params& P, const RestirBuffers& RB, uint32_t frame, cudaStream_t st) {
  int uis;
  size_t sm = vrt_render_smem_bytes(P, &uis);
  const int fixed_words = 128 * MAT_ROW_F4 * 4 + 256;
  int blocks = (P.n_tiles * 32 + VRT_GRIS_THREADS - 1) / VRT_GRIS_THREADS;
  if (blocks > 0) k_gris<<<blocks, VRT_GRIS_THREADS, sm, st>>>(P, RB, frame, uis, fixed_words);  // k_rc_sky has run (vrt_launch_rc_sky / vrt_launch_temporal)
  return cudaGetLastError();
}

cudaError_t vrt_launch_temporal(const Params& P, const RestirBuffers& RB, uint32_t frame, int hist_valid, cudaStream_t st) {
  int uis;
  size_t sm = vrt_render_smem_bytes(P, &uis);
  const int fixed_words = 128 * MAT_ROW_F4 * 4 + 256;
  int blocks = (P.n_tiles * 32 + VRT_TEMPORAL_THREADS - 1) / VRT_TEMPORAL_THREADS;
  if (cudaError_t e = vrt_launch_rc_sky(P, RB, st)) return e;
  if (blocks > 0) k_temporal<<<blocks, VRT_TEMPORAL_THREADS, sm, st>>>(P, RB, frame, hist_valid, uis, fixed_words);
  return cudaGetLastError();
}
