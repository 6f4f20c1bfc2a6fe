// Render kernels of libvoxelrt (sm_100a):
//   k_primary  primary-hit dump: next_hit for the camera ray + sun shadow ray on the cone axis
//              (renderer/pathtracer.py:218-244, :435-450)
//   k_path     fused bounce / shade / NEE path kernel: persistent warps, per-lane path
//              regeneration from a warp-local tile queue, segment and shadow rays of different
//              lanes traced by one shared traversal loop (renderer/pathtracer.py:355-632,
//              non-ReSTIR estimator, static camera; accumulation :1185-1303 reduced to a sum).
//              Shape of one outer iteration (every stage is issued once per iteration whatever
//              the number of lanes that need it, so the design goal is few iterations per path
//              and a main loop that fits the SM's instruction cache):
//                (0) retire finished paths   (1) refill idle lanes, start paths
//                (2) trace pass 0: every lane's current ray; classify (escape / emissive / surface)
//                    trace pass 1: the shadow rays spawned by pass 0, if >= 14 lanes have one
//                (3b) ONE sky-table site for escaped segments and visible sun samples
//                (4) shade: NEE term + BSDF sample (rare lobes in out-of-line functions)
//   k_resolve  mean + vignette + exposure + Uchimura + gamma (renderer/pathtracer.py:634-662,
//              renderer/math_utils.py:160-186), float4 in / float4 out
#include "vrt_bsdf.cuh"
#include "vrt_internal.h"
#include "vrt_kshared.cuh"
#include "vrt_restir.cuh"
#include "vrt_sky.cuh"
#include "vrt_trace.cuh"

// ------------------------------------------------------------------------------- k_primary
__global__ void __launch_bounds__(128) k_primary(const __grid_constant__ Params P, vrt_hit* __restrict__ out, int upper_in_smem) {
  extern __shared__ uint32_t smem[];
  const uint32_t* upper = stage_shared(P, smem, upper_in_smem);
  const float* unorm8 = reinterpret_cast<const float*>(smem + SMEM_MAT_WORDS);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.n_tiles) return;
  const int tile = P.tile_rank + P.tile_n * warp;
  const int u = (tile % P.tiles_x) * 8 + (lane & 7), v = (tile / P.tiles_x) * 4 + (lane >> 3);
  f3 d = get_cast_dir(P, (float)u, (float)v, 0.0f, 0.0f);
  Hit h = next_hit<false>(P, upper, unorm8, P.cam_pos, d, false, nullptr, nullptr);
  uint32_t shadow = 3u;
  const int kind = h.closest < VRT_INF ? h.kind : 0;
  if (!h.hit_light && h.closest < VRT_INF) {
    f3 n{h.nx, h.ny, h.nz};
    f3 pos{xadd(xadd(P.cam_pos.x, xmul(h.closest, d.x)), xmul(n.x, VRT_EPS)), xadd(xadd(P.cam_pos.y, xmul(h.closest, d.y)), xmul(n.y, VRT_EPS)),
           xadd(xadd(P.cam_pos.z, xmul(h.closest, d.z)), xmul(n.z, VRT_EPS))};
    float ndl = xdot(P.light_dir, n);
    if (ndl > 0.0f) {
      Hit sh = next_hit<false>(P, upper, unorm8, pos, P.light_dir, true, nullptr, nullptr);
      shadow = sh.closest >= VRT_INF ? 0u : 1u;
    } else {
      shadow = 2u;
    }
  }
  vrt_hit o;
  o.t = h.closest;
  o.cell[0] = kind == 2 ? h.cx : -1, o.cell[1] = kind == 2 ? h.cy : -1, o.cell[2] = kind == 2 ? h.cz : -1;
  o.normal[0] = h.nx, o.normal[1] = h.ny, o.normal[2] = h.nz;
  o.flags = (uint32_t)kind | (shadow << 8) | (((uint32_t)h.mat_id & 255u) << 16) | ((uint32_t)(h.hit_light ? 1 : 0) << 24);
  float4* dst = reinterpret_cast<float4*>(out + (size_t)v * P.W + u);
  const float4* src = reinterpret_cast<const float4*>(&o);
  dst[0] = src[0];
  dst[1] = src[1];
}

// ---------------------------------------------------------------------------------- k_path
enum { ST_SEGMENT = 0, ST_SHADOW = 1 };
enum { PIX_IDLE = -1, PIX_DONE = -2 };

#ifndef VRT_PASS1_MIN_LANES
#define VRT_PASS1_MIN_LANES 14 // lanes with a fresh shadow ray that justify a second trace pass in the same iteration
#endif
#ifndef VRT_PATH_THREADS
#define VRT_PATH_THREADS 768   // threads per CTA of the static-camera path kernel: ONE CTA per SM (the staged tables are loaded once per
                               // SM, co-scheduled warps share instruction-cache lines). Measured, config 3 / example6 / city in G paths/s
                               // (profiles/r02h_ab_path.log, r02r_ab_threads.log): 5 x 128 threads 3.36 / 1.86 / 3.48; 1 x 640 (96 regs)
                               // 3.40 / 1.90 / 3.64; 1 x 768 (24 warps, 80 regs, 48 B of spills) 3.44 / 1.92 / 3.67; 1 x 896 (72 regs)
                               // 3.30 / 1.94 / 3.66; 1 x 1024 (64 regs) 3.14 / 1.94 / 3.65
#endif
#ifndef VRT_PATH_MIN_BLOCKS
#define VRT_PATH_MIN_BLOCKS 1
#endif
#ifndef VRT_RESTIR_THREADS
#define VRT_RESTIR_THREADS 512  // ... of the ReSTIR variant: ONE CTA of 16 warps per SM at 123 registers, no spills. Measured at 1080p
                                // (profiles/r04c_ab_restir_threads.log), path + reservoir pass of example6 / example3 in ms: 3 x 128 threads
                                // (145 regs, 12 warps) 1.861 / 2.119; 1 x 384 1.840 / 2.091; 1 x 512 1.711 / 1.908; 1 x 640 (96 regs, 178 B
                                // of spills) 1.712 / 1.888; 768 threads would spill 422 B
#endif
#ifndef VRT_MOVING_THREADS
#define VRT_MOVING_THREADS 128  // ... of the moving-camera variant (12 resident warps per SM)
#endif

// RESTIR = true is the USE_RESTIR_PT variant of render (pathtracer.py:15): besides the pixel
// colour it records the reconnection data of the path (first vertex after the primary hit) into a
// packed reservoir, merges the primary-vertex NEE sample into it (reservoir.py:64-74) and writes
// the G-buffer the spatial pass needs. It runs one sample per launch and never retires
// zero-throughput paths early (their reconnection data is still used by the shift).
// MODE 2 is render with camera_is_moving = 1 (pathtracer.py:146,628-630): the frame is rendered at
// render_scale into the lower-left corner, the diffuse part is divided by the primary albedo, and
// the G-buffer the reprojecting temporal filters need (NDC depth, octahedral normal, material,
// virtual reflection depth; pathtracer.py:535-546) is written next to the two colour buffers.
template <bool STATS, int MODE, bool SKY16 = false>
__global__ void __launch_bounds__(MODE == 1 ? VRT_RESTIR_THREADS : (MODE == 2 ? VRT_MOVING_THREADS : VRT_PATH_THREADS),
                                  MODE == 1 ? 1 : (MODE == 2 ? 384 / VRT_MOVING_THREADS : VRT_PATH_MIN_BLOCKS)) k_path(const __grid_constant__ Params P, int upper_in_smem, RestirBuffers RB, MovingOut MO) {
  constexpr bool RESTIR = MODE == 1;
  constexpr bool MOVING = MODE == 2;
  extern __shared__ uint32_t smem[];
  const uint32_t* upper = stage_shared(P, smem, upper_in_smem);
  const float4* s_mats = reinterpret_cast<const float4*>(smem);
  const float* unorm8 = reinterpret_cast<const float*>(smem + SMEM_MAT_WORDS);
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;

  // per-launch constants
  f3 sun_bx, sun_by;
  make_orthonormal_basis(P.light_dir, sun_bx, sun_by);
  const float light_pdf_axis = cone_sample_pdf(P.light_cos_max, 1.0f);
  const f3 sun_rad = P.light_weight * P.light_color;
  const float sky_fres = 1.0f / (float)P.sky_res;

  // warp-local work queue: one tile (32 pixels) per global atomic
  int chunk_base = 0, chunk_rem = 0;

  // lane state
  int pix = PIX_IDLE;
  int s_i = 0, depth = 0, state = ST_SEGMENT, f_lobe = 0;
  uint32_t key = 0, pm_info = 0;
  f3 pos = mk3(0.0f), d = mk3(0.0f), thr = mk3(1.0f), contrib = mk3(0.0f), acc = mk3(0.0f);
  f3 fnee_d = mk3(0.0f), fnee_s = mk3(0.0f);
  float f_invpdf = 1.0f;
  // surface stash while the shadow ray is in flight
  f3 s_n = mk3(0.0f), s_alb = mk3(0.0f), s_view = mk3(0.0f);
  int s_mat = 0;
  TraceCounters tc{0, 0, 0};
  uint32_t c_hits = 0, c_escapes = 0, c_nee = 0, c_vertices = 0, c_paths = 0;
  // ReSTIR bookkeeping (dead code when !RESTIR)
  RSample rz;
  f3 thr_after_rc = mk3(1.0f), primary_pos = mk3(0.0f);
  float f_lpdf = 1.0f, f_bounce_light_pdf = 0.0f;
  uint32_t primary_noct = 0u;
  int rc_lobe = 0;
  bool sky_ray = false;
  // moving-camera bookkeeping (dead code unless MOVING)
  f3 primary_albedo = mk3(1.0f);
  float refl_dist = 0.0f;

  bool finished = false, restart = false;

  for (;;) {
    // ---- (0) retire finished paths at one converged site: pixel sample value
    // (pathtracer.py:609-619), NaN scrub (:1068-1075), then either the next sample of this
    // pixel or the single read-modify-write of its accumulation texel.
    if (finished) {
      finished = false;
      f3 emission = mk3(0.0f);
      if ((pm_info & 255u) == 2u)
        emission = f3{unorm8[(pm_info >> 8) & 255u], unorm8[(pm_info >> 16) & 255u], unorm8[(pm_info >> 24) & 255u]};
      f3 diffuse, specular;
      const int u = pix & 0xffff, v = pix >> 16;
      const size_t pidx = (size_t)v * P.W + u;
      if (MOVING) {
        diffuse = fnee_d, specular = fnee_s;
        if (f_lobe == LOBE_DIFFUSE) diffuse += contrib * f_invpdf + emission;
        if (f_lobe == LOBE_SPEC_REFL) specular += contrib * f_invpdf;
        diffuse = diffuse / f3{fmaxf(primary_albedo.x, 1e-2f), fmaxf(primary_albedo.y, 1e-2f), fmaxf(primary_albedo.z, 1e-2f)};
        MO.col_d[pidx] = make_float4(diffuse.x, diffuse.y, diffuse.z, 0.0f);
        MO.col_s[pidx] = make_float4(specular.x, specular.y, specular.z, 0.0f);
        MO.depth[pidx] = view_to_screen_z(P, primary_pos);
        MO.attr[pidx] = make_uint2(primary_noct, pm_info);
        float refl = 0.0f;
        if (refl_dist != 0.0f && !isbad(refl_dist)) {
          const f3 virtual_point = primary_pos + normalize(primary_pos - P.cam_pos) * refl_dist;
          refl = linearize_depth(P, view_to_screen_z(P, virtual_point));
        }
        MO.refl[pidx] = refl;
      } else if (!RESTIR) {
        diffuse = fnee_d, specular = fnee_s;
        if (f_lobe == LOBE_DIFFUSE) diffuse += contrib * f_invpdf + emission;
        if (f_lobe == LOBE_SPEC_REFL) specular += contrib * f_invpdf;
        if (bad3(diffuse)) diffuse = mk3(0.0f);
        if (bad3(specular)) specular = mk3(0.0f);
        acc += diffuse + specular;
      } else {
        // pathtracer.py:548-607: reservoir of the BSDF-sampled path, RIS-merged with the NEE sample
        RReservoir res;
        res.z = rz;
        res.z.F = contrib;
        res.z.lobes = rc_lobe * 10 + f_lobe;
        res.M = 1.0f;
        res.z.cached_jacobian_term = jacobian_term(res.z.rc_pos, res.z.rc_normal, primary_pos);
        bool chose_NEE = false;
        if (!sky_ray) {
          float bsdf_light_pdf = f_bounce_light_pdf;
          if (is_vec_zero(fnee_d + fnee_s)) bsdf_light_pdf = 0.0f;
          const float bsdf_mis = power_heuristic(frcp(f_invpdf), bsdf_light_pdf);
          const float light_mis = power_heuristic(light_pdf_axis, f_lpdf);
          res.weight = bsdf_mis * luminance(res.z.F) * f_invpdf;
          const float in_w = light_mis * luminance(fnee_d + fnee_s);
          res.M += 1.0f;  // input_sample (reservoir.py:64-74)
          if (in_w > 0.0f) {
            res.weight += in_w;
            if (rnd(key, 40) * res.weight <= in_w) {
              chose_NEE = true;
              const f3 first_light_dir = sample_cone_oriented(P.light_cos_max, P.light_dir, sun_bx, sun_by, rnd(key, 0), rnd(key, 1));
              f3 sky_T = mk3(1.0f);
              if (P.use_sky) sky_T = sky_fetch(P.sky_trans, sky_tap(P.sky_res, project_sky(first_light_dir, sky_fres)));
              rinit(res);
              res.M = 2.0f;
              res.weight = bsdf_mis * luminance(contrib) * f_invpdf + in_w;
              res.z.F = fnee_d + fnee_s;
              res.z.rc_pos = first_light_dir;
              res.z.rc_incident_L = sky_T * sun_rad;
              res.z.lobes = LOBE_ALL * 10 + LOBE_ALL;
            }
          }
          const float p_hat = luminance(res.z.F);  // finalize_without_M
          res.weight = p_hat < 1e-6f ? 0.0f : fdiv(res.weight, p_hat);
        } else {
          res.weight = 1.0f;
        }
        uint32_t w[14];
        encode_reservoir(res, w);
        uint2* dst = RB.reservoirs + pidx * 7;
#pragma unroll
        for (int i = 0; i < 7; i++) dst[i] = make_uint2(w[2 * i], w[2 * i + 1]);
        RB.gpos[pidx] = make_float4(primary_pos.x, primary_pos.y, primary_pos.z, sky_ray ? 1.0f : 0.0f);
        RB.gattr[pidx] = make_uint2(primary_noct, pm_info);
        diffuse = mk3(0.0f), specular = mk3(0.0f);
        if (!chose_NEE) {
          if (f_lobe == LOBE_DIFFUSE) diffuse = res.z.F;
          if (f_lobe == LOBE_SPEC_REFL) specular = res.z.F;
        } else {
          diffuse = fnee_d, specular = fnee_s;
        }
        RB.col_d[pidx] = make_float4(diffuse.x, diffuse.y, diffuse.z, 0.0f);
        RB.col_s[pidx] = make_float4(specular.x, specular.y, specular.z, 0.0f);
      }
      s_i++;
      if (s_i < P.n_samples && MODE == 0) {
        restart = true;
      } else {
        if (MODE == 0) {
          // first batch after reset_framebuffer: the texel is written, not read-modified (no 33 MB memset, no read)
          float4* dst = P.accum + pidx;
          float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (!P.accum_overwrite) a = *dst;
          a.x += acc.x, a.y += acc.y, a.z += acc.z, a.w += (float)P.n_samples;
          *dst = a;
        }
        pix = PIX_IDLE;
      }
    }
    // ---- (1) refill idle lanes from the warp's tile queue
    for (;;) {
      const unsigned need = __ballot_sync(FULL, pix == PIX_IDLE);
      if (need == 0u) break;
      if (chunk_rem == 0) {
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(P.work_counter, 1u);
        c = __shfl_sync(FULL, c, 0);
        if (c >= (unsigned)P.n_tiles) {
          if (pix == PIX_IDLE) pix = PIX_DONE;
          break;
        }
        chunk_base = (int)c * 32;
        chunk_rem = 32;
      }
      const int rank = __popc(need & lt_mask);
      if (pix == PIX_IDLE && rank < chunk_rem) {
        // pix = u | v << 16 of the pixel this lane now owns (one integer division per pixel)
        const int item = chunk_base + (32 - chunk_rem) + rank;
        const int tile = P.tile_rank + P.tile_n * (item >> 5);
        const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
        pix = (tx * 8 + (item & 7)) | ((ty * 4 + ((item >> 3) & 3)) << 16);
        s_i = 0;
        acc = mk3(0.0f);
        restart = true;
      }
      chunk_rem -= min(__popc(need), chunk_rem);
    }
    // ---- (1b) start the next path (new pixel or next sample of the same pixel), one site
    if (restart) {
      restart = false;
      const int u = pix & 0xffff, v = pix >> 16;
      const uint32_t sample = (uint32_t)(P.first_sample + s_i * P.stride);
      key = path_key((uint32_t)(v * P.W + u), sample, P.seed);
      if (MOVING) {
        // is_outside_render_area (pathtracer.py:289-291): such pixels are not rendered this frame
        if ((float)u > MO.scale * (float)P.W || (float)v > MO.scale * (float)P.H) pix = PIX_IDLE;
        d = get_cast_dir_scaled(P, (float)u, (float)v, MO.scale);
      } else {
        const float2 j = P.jitter[s_i];
        // IEEE camera ray in every mode: an SFU-arithmetic variant was 2.7 % faster but flipped 1-2 edge pixels per
        // 10^4 against the oracle (profiles/r01e_ifetch_experiments.md); per-pixel parity is worth more
        d = get_cast_dir(P, (float)u, (float)v, j.x, j.y);
      }
      pos = P.cam_pos;
      thr = mk3(1.0f), contrib = mk3(0.0f), fnee_d = mk3(0.0f), fnee_s = mk3(0.0f);
      f_invpdf = 1.0f, f_lobe = 0, pm_info = 0, depth = 0, state = ST_SEGMENT;
      if (MOVING) primary_albedo = mk3(1.0f), refl_dist = 0.0f, primary_pos = mk3(0.0f), primary_noct = 0u, sky_ray = false;
      if (RESTIR) {
        rz.F = rz.rc_pos = rz.rc_normal = rz.rc_incident_dir = rz.rc_incident_L = rz.rc_NEE_dir = mk3(0.0f);
        rz.rc_mat_info = 0u, rz.cached_jacobian_term = 1.0f, rz.lobes = 0;
        thr_after_rc = mk3(1.0f), primary_pos = mk3(0.0f), f_lpdf = 1.0f, f_bounce_light_pdf = 0.0f, primary_noct = 0u, rc_lobe = 0;
        sky_ray = false;
      }
      if (STATS) c_paths++;
    }
    if (__all_sync(FULL, pix == PIX_DONE)) break;
    const bool active = pix >= 0;

    // ---- (2)+(3) up to two trace passes per iteration through ONE copy of the traversal code: pass 0
    // traces every live lane's current ray, pass 1 the sun shadow rays that pass 0 spawned. The
    // stages around the passes (retire, refill, restart, sky site, shade) cost the same whether
    // few or many lanes need them, so a path vertex should take one outer iteration, not two.
    bool do_shade = false, escaped = false, sky_need = false;
    float visible = 0.0f;
    f3 light_dir = d, sky_dir = d;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      const bool tracing = active && (pass == 0 || state == ST_SHADOW);
      // pass 1 only pays when enough lanes have a fresh shadow ray: below the threshold the rays wait
      // (state stays ST_SHADOW) and are traced by pass 0 of the next iteration together with the segments
      if (pass == 1 && __popc(__ballot_sync(FULL, tracing)) < VRT_PASS1_MIN_LANES) break;
      Hit h;
      h.closest = VRT_INF, h.hit_light = 0, h.mat_id = 0, h.kind = 0, h.nx = h.ny = h.nz = 0.0f, h.albedo = mk3(1.0f);
      if (tracing) h = next_hit<STATS>(P, upper, unorm8, pos, d, state == ST_SHADOW, &tc, &c_hits);

      // ---- (3) classify
      if (tracing) {
        if (state == ST_SEGMENT) {
          const uint32_t base = 8u * (uint32_t)depth;
          if (MOVING) {  // pathtracer.py:402-412
            if (depth == 0) {
              primary_pos = pos + h.closest * d;
              primary_albedo = h.albedo;
              pm_info = encode_material(h.mat_id, h.albedo);
              float ex = 0.0f, ey = 0.0f;
              if (h.closest < VRT_INF) encode_unit_vector_3x16(f3{h.nx, h.ny, h.nz}, ex, ey);
              primary_noct = h16bits(ex) | (h16bits(ey) << 16);
            } else if (depth == 1 && f_lobe != LOBE_DIFFUSE) {
              refl_dist += h.closest;
            }
          }
          if (RESTIR) {  // pathtracer.py:402-417
            const f3 hit_pos = pos + h.closest * d;
            if (depth == 0) {
              primary_pos = hit_pos;
              pm_info = encode_material(h.mat_id, h.albedo);
              float ex = 0.0f, ey = 0.0f;
              if (h.closest < VRT_INF) encode_unit_vector_3x16(f3{h.nx, h.ny, h.nz}, ex, ey);
              primary_noct = h16bits(ex) | (h16bits(ey) << 16);
            } else if (depth == 1) {
              rz.rc_pos = hit_pos;
              rz.rc_normal = f3{h.nx, h.ny, h.nz};
              rz.rc_mat_info = encode_material(h.mat_id, h.albedo);
              f_bounce_light_pdf = cone_sample_pdf(P.light_cos_max, dot(P.light_dir, d));
            } else if (depth == 2) {
              rz.rc_incident_dir = d;
            }
          }
          if (h.closest == VRT_INF) {
            // escaped: background or sky tables + sun disk (pathtracer.py:499-511); the table
            // lookup happens at the merged sky site (3b) below
            escaped = true;
            if (P.use_sky) {
              sky_dir = normalize(d + f3{rnd(key, base + 5), rnd(key, base + 6), rnd(key, base + 7)} * 0.0015f);
              sky_need = true;
              if (STATS) c_escapes++;
            }
            finished = true;
          } else if (h.hit_light) {
            // emissive voxel / floor terminates the path (pathtracer.py:519-525)
            if (depth > 0) contrib += thr * h.albedo;
            if (depth == 0) pm_info = encode_material(h.mat_id, h.albedo);
            if (RESTIR && depth >= 2) rz.rc_incident_L += firefly_filter(thr_after_rc * h.albedo);
            finished = true;
          } else {
            if (STATS) c_vertices++;
            s_n = f3{h.nx, h.ny, h.nz};
            s_alb = h.albedo;
            s_mat = h.mat_id;
            s_view = -d;
            pos = (pos + h.closest * d) + s_n * VRT_EPS;
            light_dir = sample_cone_oriented(P.light_cos_max, P.light_dir, sun_bx, sun_by, rnd(key, base + 0), rnd(key, base + 1));
            if (dot(light_dir, s_n) > 0.0f) {
              d = light_dir;
              state = ST_SHADOW;
            } else {
              do_shade = true;
            }
          }
        } else {
          visible = h.closest >= VRT_INF ? 1.0f : 0.0f;
          state = ST_SEGMENT;
          do_shade = true;
          if (visible != 0.0f && P.use_sky) {
            sky_need = true;
            sky_dir = d;  // the sun sample the shadow ray was traced along
            if (STATS) c_nee++;
          }
        }
      }
    }  // trace passes

    // ---- (3b) merged sky site: one projection + bilinear footprint per lane for both users of
    // the tables, escaped segments (scattering + transmittance, atmos.py:94-115) and visible sun
    // samples (transmittance only, atmos.py:117-131)
    f3 sky_T = mk3(1.0f), sky_scattering = P.background;
    if (sky_need) {
      const SkyTap t = sky_tap(P.sky_res, project_sky(sky_dir, sky_fres));
      if (SKY16) {
        f3 sc;
        sky_fetch_packed(P.sky_packed, t, sc, sky_T);
        if (escaped) sky_scattering = sc;
      } else {
        sky_T = sky_fetch(P.sky_trans, t);
        if (escaped) sky_scattering = sky_fetch(P.sky_scatter, t);
      }
    }
    if (escaped) {
      const float hit_sun = dot(P.light_dir, d) >= P.light_cos_max ? 1.0f : 0.0f;
      f3 sky_emission = firefly_filter(sky_scattering + sky_T * sun_rad * hit_sun);
      contrib += thr * sky_emission;
      if (MOVING && depth == 0) primary_pos = mk3(0.0f), sky_ray = true;
      if (RESTIR) {  // pathtracer.py:509-517
        if (depth == 0) {
          primary_pos = mk3(0.0f);
          sky_ray = true;
        } else if (depth == 1) {
          rz.rc_pos = d;
          rz.rc_incident_L = sky_emission;
        } else {
          rz.rc_incident_L += firefly_filter(thr_after_rc * sky_emission);
        }
      }
    }

    // ---- (4) shade: NEE contribution (if the sun is visible) + BSDF sample
    if (do_shade) {
      const uint32_t base = 8u * (uint32_t)depth;
      Mat m = load_mat(s_mats, s_mat);
      m.base_col = s_alb;
      f3 tang, bitang;
      make_orthonormal_basis(s_n, tang, bitang);
      if (visible != 0.0f) {
        f3 bd, bs;
        float lpdf;
        eval_and_pdf(m, s_view, s_n, light_dir, tang, bitang, bd, bs, lpdf);
        const float mis = power_heuristic(light_pdf_axis, lpdf);
        const float ndl = dot(light_dir, s_n);
        const f3 lr = sky_T * sun_rad * ndl;
        if (depth == 0) {
          // primary vertex: unweighted at accumulation time, MIS weight applied once at the end
          // (pathtracer.py:470-472, :561-568); thr == 1 here. The ReSTIR variant keeps the
          // unweighted value: the MIS weight goes into the RIS weight instead (:565-568).
          fnee_d = firefly_filter(thr * (bd * lr)) * (RESTIR ? 1.0f : mis);
          fnee_s = firefly_filter(thr * (bs * lr)) * (RESTIR ? 1.0f : mis);
          if (RESTIR) f_lpdf = lpdf;
        } else {
          contrib += firefly_filter(thr * ((mis * (bd + bs)) * lr));
          if (RESTIR) {
            if (depth == 1) rz.rc_NEE_dir = light_dir;
            if (depth >= 2) rz.rc_incident_L += thr_after_rc * ((mis * (bd + bs)) * lr);
          }
        }
      }
      f3 brdf;
      float pdf;
      int lobe;
      const f3 nd = sample_disney(m, s_view, s_n, tang, bitang, rnd(key, base + 2), rnd(key, base + 3), rnd(key, base + 4), brdf, pdf, lobe);
      f3 bounce_weight = brdf * saturate(dot(nd, s_n));
      if (depth == 0) {
        f_invpdf = frcp(pdf);
        f_lobe = lobe;
      } else {
        bounce_weight = bounce_weight * frcp(pdf);
        const float bsdf_sample_light_pdf = cone_sample_pdf(P.light_cos_max, dot(P.light_dir, nd));
        bounce_weight *= power_heuristic(pdf, visible * bsdf_sample_light_pdf);
        if (RESTIR) {
          if (depth == 1) rc_lobe = lobe;
          if (depth >= 2) thr_after_rc *= bounce_weight;
        }
      }
      thr *= bounce_weight;
      d = nd;
      depth++;
      // The reference keeps tracing zero-throughput paths; they add exact zeros, so stop here.
      const bool dead = MODE == 0 && thr.x == 0.0f && thr.y == 0.0f && thr.z == 0.0f;
      if (depth >= P.max_depth || dead) finished = true;
    }
  }

  if (STATS) {
    unsigned long long vals[8] = {c_paths, tc.rays, tc.steps, tc.queries, c_hits, c_escapes, c_nee, c_vertices};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      unsigned long long x = vals[i];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
      if (lane == 0 && x) atomicAdd(P.stats + i, x);
    }
  }
}

// ------------------------------------------------------------------------------- k_resolve
HD float uchimura1(float x) {  // math_utils.py:160-186
  const float Pm = 1.0f, a = 1.0f, m = 0.22f, l = 0.4f, c = 1.33f, b = 0.0f;
  const float l0 = ((Pm - m) * l) / a;
  const float S0 = m + l0;
  const float S1 = m + a * l0;
  const float C2 = (a * Pm) / (Pm - S1);
  const float CP = -C2 / Pm;
  float t = clampf(x / m, 0.0f, 1.0f);
  float w0 = 1.0f - t * t * (3.0f - 2.0f * t);
  float w2 = x < m + l0 ? 0.0f : 1.0f;
  float w1 = 1.0f - w0 - w2;
  float T = m * powf(x / m, c) + b;
  float S = Pm - (Pm - S1) * expf(CP * (x - S0));
  float L = m + a * (x - m);
  return T * w0 + L * w1 + S * w2;
}

__global__ void __launch_bounds__(256) k_resolve(const float4* __restrict__ accum, float4* __restrict__ hdr, float4* __restrict__ ldr, int W,
                                                 int H, float exposure) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * H) return;
  const int i = idx % W, j = idx / W;
  float4 a = accum[idx];
  const float inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
  f3 c{a.x * inv, a.y * inv, a.z * inv};
  if (hdr) hdr[idx] = make_float4(c.x, c.y, c.z, a.w);
  if (ldr) {
    float ux = (float)i / (float)W - 0.5f, uy = (float)j / (float)H - 0.5f;
    float darken = 1.0f - 0.9f * fmaxf(sqrtf(ux * ux + uy * uy), 0.0f);
    float s = darken * exposure;
    ldr[idx] = make_float4(saturate(powf(uchimura1(c.x * s), 1.0f / 2.2f)), saturate(powf(uchimura1(c.y * s), 1.0f / 2.2f)),
                           saturate(powf(uchimura1(c.z * s), 1.0f / 2.2f)), 1.0f);
  }
}

// Fused multi-GPU merge + tonemap: rank-local sums plus up to 7 peers' partial sums read through
// NVLink peer mappings (ld.global on cudaIpc-mapped pointers), then the same tonemap as k_resolve.
// 16-byte loads, one pixel per thread: each peer contributes one coalesced 512-byte request per warp.
struct PeerPtrs {
  const float4* p[8];
  int n;
};
__global__ void __launch_bounds__(256) k_resolve_merged(const float4* __restrict__ accum, PeerPtrs peers, float4* __restrict__ ldr, int W, int H,
                                                        float exposure) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * H) return;
  float4 a = accum[idx];
#pragma unroll 1
  for (int k = 0; k < peers.n; k++) {
    const float4 b = peers.p[k][idx];
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
  }
  const int i = idx % W, j = idx / W;
  const float inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
  const float ux = (float)i / (float)W - 0.5f, uy = (float)j / (float)H - 0.5f;
  const float s = (1.0f - 0.9f * fmaxf(sqrtf(ux * ux + uy * uy), 0.0f)) * exposure;
  ldr[idx] = make_float4(saturate(powf(uchimura1(a.x * inv * s), 1.0f / 2.2f)), saturate(powf(uchimura1(a.y * inv * s), 1.0f / 2.2f)),
                         saturate(powf(uchimura1(a.z * inv * s), 1.0f / 2.2f)), 1.0f);
}

cudaError_t vrt_launch_resolve_merged(const float4* accum, const float4* const* peers, int n_peers, float4* ldr, int W, int H, float exposure,
                                      cudaStream_t st) {
  PeerPtrs pp;
  pp.n = n_peers;
  for (int k = 0; k < 8; k++) pp.p[k] = k < n_peers ? peers[k] : nullptr;
  const int n = W * H;
  k_resolve_merged<<<(n + 255) / 256, 256, 0, st>>>(accum, pp, ldr, W, H, exposure);
  return cudaGetLastError();
}

// Per-sample TAA jitter of a batch, computed on the device so that vrt_accumulate needs neither a host
// buffer nor a host-to-device copy (the call is fully asynchronous). pathtracer.py:264-265 draws one
// (rand*2-1) * inv_image_res pair per frame; ours is the Halton(2,3) point of the sample index. Same double
// arithmetic, op for op, as the host / oracle formulation (halton() in oracle.cpp), no contraction.
__device__ double halton_dev(uint32_t i, uint32_t b) {
  double f = 1.0, r = 0.0;
  while (i > 0) {
    f = __ddiv_rn(f, (double)b);
    r = __dadd_rn(r, __dmul_rn(f, (double)(i % b)));
    i /= b;
  }
  return r;
}
__global__ void k_jitter(float2* __restrict__ out, int first_sample, int stride, int n, int W, int H, int mode, unsigned int* work_counter) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0 && work_counter) *work_counter = 0u;  // the tile queue of the path kernel launched next on this stream
  if (k >= n) return;
  const uint32_t s = (uint32_t)(first_sample + k * stride);
  float2 j = make_float2(0.0f, 0.0f);
  if (mode == 1) {
    j.x = __double2float_rn(__ddiv_rn(__dsub_rn(__dmul_rn(halton_dev(s + 1, 2), 2.0), 1.0), (double)W));
    j.y = __double2float_rn(__ddiv_rn(__dsub_rn(__dmul_rn(halton_dev(s + 1, 3), 2.0), 1.0), (double)H));
  }
  out[k] = j;
}
cudaError_t vrt_launch_jitter(float2* out, int first_sample, int stride, int n, int W, int H, int mode, unsigned int* work_counter, cudaStream_t st) {
  k_jitter<<<(n + 127) / 128, 128, 0, st>>>(out, first_sample, stride, n, W, H, mode, work_counter);
  return cudaGetLastError();
}

// Fused multi-GPU reduce-scatter + tonemap (SURVEY.md §8e "fused variant", spread over the ranks): every rank
// owns a contiguous slice of the pixels, sums the partial accumulation buffers of ALL ranks for that slice —
// its own from HBM, the others through NVLink peer mappings (ld.global on cudaIpc-mapped pointers) — and
// writes (a) the merged float4 sums to its own `sum_out` slice (optional) and (b) the tonemapped pixels
// (pathtracer.py:634-662) to `ldr_out`, which may itself be a peer mapping of the displaying rank's image
// buffer (st.global over NVLink). Per rank and frame that is (N-1)/N x 16 B/pixel of NVLink reads instead of
// the 2 (N-1)/N x 16 B/pixel an all-reduce moves, and no rank handles more than 1/N of the frame.
__global__ void __launch_bounds__(256) k_merge_slice(const float4* accum, PeerPtrs peers, float4* sum_out,
                                                     float4* __restrict__ ldr_out, int first, int count, int W, int H, float exposure) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int idx = first + k;
  float4 a = accum[idx];
#pragma unroll 1
  for (int p = 0; p < peers.n; p++) {
    const float4 b = peers.p[p][idx];
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
  }
  if (sum_out) sum_out[idx] = a;
  if (ldr_out) {
    const int i = idx % W, j = idx / W;
    const float inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
    const float ux = (float)i / (float)W - 0.5f, uy = (float)j / (float)H - 0.5f;
    const float s = (1.0f - 0.9f * fmaxf(sqrtf(ux * ux + uy * uy), 0.0f)) * exposure;
    ldr_out[idx] = make_float4(saturate(powf(uchimura1(a.x * inv * s), 1.0f / 2.2f)), saturate(powf(uchimura1(a.y * inv * s), 1.0f / 2.2f)),
                               saturate(powf(uchimura1(a.z * inv * s), 1.0f / 2.2f)), 1.0f);
  }
}
cudaError_t vrt_launch_merge_slice(const float4* accum, const float4* const* peers, int n_peers, float4* sum_out, float4* ldr_out, int first,
                                   int count, int W, int H, float exposure, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  PeerPtrs pp;
  pp.n = n_peers;
  for (int k = 0; k < 8; k++) pp.p[k] = k < n_peers ? peers[k] : nullptr;
  k_merge_slice<<<(count + 255) / 256, 256, 0, st>>>(accum, pp, sum_out, ldr_out, first, count, W, H, exposure);
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------- launchers
static size_t smem_bytes(const Params& P, int* upper_in_smem) {
  size_t need = (size_t)SMEM_FIXED_WORDS * 4 + (size_t)P.upper_words * 4;
  *upper_in_smem = need <= 48 * 1024 ? 1 : 0;
  return *upper_in_smem ? need : (size_t)SMEM_FIXED_WORDS * 4;
}

cudaError_t vrt_launch_primary(const Params& P, vrt_hit* out, cudaStream_t st) {
  int uis;
  size_t sm = smem_bytes(P, &uis);
  int warps = P.n_tiles;
  int blocks = (warps * 32 + 127) / 128;
  k_primary<<<blocks, 128, sm, st>>>(P, out, uis);
  return cudaGetLastError();
}

template <bool STATS, int MODE, bool SKY16 = false>
static cudaError_t launch_path_t(const Params& P, int sm_count, cudaStream_t st, int* blocks_out, const RestirBuffers& RB, const MovingOut& MO) {
  int uis;
  size_t sm = smem_bytes(P, &uis);
  int per_sm = 0;
  constexpr int threads = MODE == 1 ? VRT_RESTIR_THREADS : (MODE == 2 ? VRT_MOVING_THREADS : VRT_PATH_THREADS);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_path<STATS, MODE, SKY16>, threads, sm);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  int blocks = sm_count * per_sm;  // persistent: one wave, a multiple of the SM count
  int max_useful = (P.n_tiles + threads / 32 - 1) / (threads / 32);
  if (blocks > max_useful) blocks = max_useful > 0 ? max_useful : 1;
  if (blocks_out) *blocks_out = blocks;
  k_path<STATS, MODE, SKY16><<<blocks, threads, sm, st>>>(P, uis, RB, MO);
  return cudaGetLastError();
}

cudaError_t vrt_launch_path(const Params& P, bool stats, int sm_count, cudaStream_t st, int* blocks_out) {
  RestirBuffers none{nullptr, nullptr, nullptr, nullptr, nullptr};
  MovingOut mo{nullptr, nullptr, nullptr, nullptr, nullptr, 1.0f};
  if (stats) return launch_path_t<true, 0>(P, sm_count, st, blocks_out, none, mo);  // the counting build reads the float tables
  if (P.sky_packed && P.use_sky) return launch_path_t<false, 0, true>(P, sm_count, st, blocks_out, none, mo);
  return launch_path_t<false, 0>(P, sm_count, st, blocks_out, none, mo);
}

cudaError_t vrt_launch_path_restir(const Params& P, const RestirBuffers& RB, int sm_count, cudaStream_t st) {
  MovingOut mo{nullptr, nullptr, nullptr, nullptr, nullptr, 1.0f};
  return launch_path_t<false, 1>(P, sm_count, st, nullptr, RB, mo);
}

cudaError_t vrt_launch_path_moving(const Params& P, const MovingOut& MO, int sm_count, cudaStream_t st) {
  RestirBuffers none{nullptr, nullptr, nullptr, nullptr, nullptr};
  return launch_path_t<false, 2>(P, sm_count, st, nullptr, none, MO);
}

// float tables -> one packed binary16 table (format 1 of vrt_set_sky_format)
__global__ void __launch_bounds__(256) k_pack_sky(const float4* __restrict__ scatter, const float4* __restrict__ trans, uint4* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 s = scatter[i], t = trans[i];
  const __half2 a = __floats2half2_rn(s.x, s.y), b = __floats2half2_rn(s.z, t.x), c = __floats2half2_rn(t.y, t.z);
  uint4 o;
  o.x = *reinterpret_cast<const uint32_t*>(&a), o.y = *reinterpret_cast<const uint32_t*>(&b), o.z = *reinterpret_cast<const uint32_t*>(&c), o.w = 0u;
  out[i] = o;
}
cudaError_t vrt_launch_pack_sky(const float4* scatter, const float4* trans, uint4* out, size_t n, cudaStream_t st) {
  k_pack_sky<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scatter, trans, out, n);
  return cudaGetLastError();
}

size_t vrt_render_smem_bytes(const Params& P, int* upper_in_smem) { return smem_bytes(P, upper_in_smem); }

cudaError_t vrt_launch_resolve(const float4* accum, float4* hdr, float4* ldr, int W, int H, float exposure, cudaStream_t st) {
  int n = W * H;
  k_resolve<<<(n + 255) / 256, 256, 0, st>>>(accum, hdr, ldr, W, H, exposure);
  return cudaGetLastError();
}
