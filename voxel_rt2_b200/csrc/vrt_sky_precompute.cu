// Physical sky + cloud precompute (one-off, at vrt_prepare). Replaces, in renderer/atmos.py:
//   generate_transmittance_lut :462-498   -> k_trans_lut
//   compute_cloud_ambient      :134-138   -> k_cloud_ambient
//   accumulate_clouds          :140-157   -> k_clouds (all passes of a texel in one thread)
//   compute_skybox             :159-189   -> k_skybox
// with clouds_scattering/clouds_shadow_od/sample_cloud_density/cloud_phase :195-349,
// atmospheric_scattering :355-425 and the density helpers :500-527.
// Output tables are float4 texels [x][y]; resolution is a parameter (reference: 3840).
// Random numbers: counter RNG keyed by (texel, pass, seed) — same specification as the oracle.
#include "vrt_bsdf.cuh"
#include "vrt_internal.h"
#include "vrt_sky.cuh"

namespace {

struct Atm {  // atmos.py:38-83
  f3 rayleigh_coeff;
  float mie_coeff;
  f3 ozone_coeff;
  float mie_ext;
  float scale_height_rayl, scale_height_mie, mie_g;
  float planet_r, planet_r_offset, atmos_height;
  float cloud_height, cloud_thickness, cloud_density, cloud_extinc, cloud_scatter;
  f3 cam_pos;
};

__constant__ Atm c_atm;

struct Ctr {
  uint32_t key, n;
  HD float next() { return rnd(key, n++); }
};

// atmos.py:9-15 — contraction-free: pos.pos ~ 4e13 cancels against r*r, one ulp there is 4e6.
HD f2 rsi(f3 pos, f3 dir, float r) {
  float b = xdot(pos, dir);
  float discr = xadd(xsub(xmul(b, b), xdot(pos, pos)), xmul(r, r));
  discr = xsqrt(discr);
  if (discr < 0.0f) return f2{-1.0f, -1.0f};
  return f2{xadd(-b, -discr), xadd(-b, discr)};
}
HD float rayleigh_phase(float c) { return 3.0f / (16.0f * VRT_PI) * (1.0f + c * c); }
// atmos.py:22-25; x^1.5 as x*sqrt(x): powf was ~40 % of the skybox kernel's instructions
HD float mie_phase(float c, float g) {
  const float t = 1.0f + g * g - 2.0f * g * c;
  return (1.0f - g * g) * frcp(4.0f * VRT_PI * (t * fsqrt(t)));
}
HD f3 get_unit_vec(float rx, float ry) {
  rx *= VRT_PI * 2.0f;
  ry = ry * 2.0f - 1.0f;
  float s = sqrtf(1.0f - ry * ry);
  float sn, cs;
  sincosf(rx, &sn, &cs);
  return normalize(f3{sn * s, cs * s, ry});
}
HD float xlength(f3 p) { return xsqrt(xadd(xadd(xmul(p.x, p.x), xmul(p.y, p.y)), xmul(p.z, p.z))); }
HD float get_elevation(f3 p) { return xsub(xlength(p), c_atm.planet_r); }
HD float get_ozone_density(float h) {
  float h_km = h * 0.001f;
  float rel = h_km - 25.0f;
  rel = rel * rel;
  float d = (1.0f - 0.375f) * expf(-rel / 49.0f);
  d += 0.375f * expf(-rel / 256.0f);
  float q = h_km - 15.0f;
  d += fmaxf(0.0f, -0.000015f * (q * q * q));
  return d * 4.0f;
}
HD f3 get_density(float h) {
  h = fmaxf(h, 0.0f);
  return f3{expf(-h / c_atm.scale_height_rayl), expf(-h / c_atm.scale_height_mie), get_ozone_density(h)};
}
HD f3 extinc_mul(f3 v) {
  return f3{(c_atm.rayleigh_coeff.x * v.x + c_atm.mie_ext * v.y) + c_atm.ozone_coeff.x * v.z,
            (c_atm.rayleigh_coeff.y * v.x + c_atm.mie_ext * v.y) + c_atm.ozone_coeff.y * v.z,
            (c_atm.rayleigh_coeff.z * v.x + c_atm.mie_ext * v.y) + c_atm.ozone_coeff.z * v.z};
}
HD f3 read_trans_lut(const __half* __restrict__ lut, float cos_theta, float h) {
  int ux = (int)clampf((cos_theta * 0.5f + 0.5f) * 256.0f, 0.0f, 255.0f);
  int uy = (int)clampf((h / c_atm.atmos_height) * 128.0f, 0.0f, 127.0f);
  const __half* p = lut + (ux * 128 + uy) * 3;
  return f3{__half2float(p[0]), __half2float(p[1]), __half2float(p[2])};
}
HD f3 sun_basis_sample(float cosmax, f3 sun_dir, f3 bx, f3 by, Ctr& rng) {
  float u0 = rng.next(), u1 = rng.next();
  return sample_cone_oriented(cosmax, sun_dir, bx, by, u0, u1);
}

// atmos.py:475-498
HD f3 get_ray_transmittance(f3 ray_pos, f3 ray_dir) {
  const float fsteps = 1.0f / 128.0f;
  float step_delta = rsi(ray_pos, ray_dir, c_atm.planet_r + c_atm.atmos_height).y * fsteps;
  f3 ray_step = ray_dir * step_delta;
  ray_pos = ray_pos + ray_step * (0.5f * (fmaxf(ray_dir.y, 0.0f) * 0.5f + 0.5f));
  f3 od = mk3(0.0f);
  for (int i = 0; i < 128; i++) {
    f3 dens = get_density(get_elevation(ray_pos));
    od += dens * step_delta;
    ray_pos += ray_step;
  }
  od = extinc_mul(od);
  f3 T = exp3(-od);
  if (rsi(ray_pos, ray_dir, c_atm.planet_r).x > 0.0f) T *= 0.0f;
  return T;
}

__global__ void __launch_bounds__(128) k_trans_lut(__half* __restrict__ lut) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 256 * 128) return;
  const int x = i / 128, y = i % 128;
  float cos_theta = ((float)x / 256.0f) * 2.0f - 1.0f;
  float h = c_atm.atmos_height * (float)y / 128.0f;
  float theta = acosf(cos_theta);
  float sin_theta = sinf(theta);
  f3 T = get_ray_transmittance(f3{0.0f, c_atm.planet_r + h, 0.0f}, f3{sin_theta, cos_theta, 0.0f});
  lut[i * 3 + 0] = __float2half_rn(T.x);
  lut[i * 3 + 1] = __float2half_rn(T.y);
  lut[i * 3 + 2] = __float2half_rn(T.z);
}

// atmos.py:355-425; DEPTH is the reference's ti.template() recursion depth (depth 2 returns
// (0,1) at once, so depth-1 marches skip the multiple-scattering samples: they add zeros).
template <int DEPTH>
__device__ void atmospheric_scattering(const SkyBuild& B, f3 sbx, f3 sby, f3 ray_origin, f3 ray_dir, int steps, Ctr& rng, f3& in_scatter_col,
                                       f3& transmittance) {
  const float fsteps = 1.0f / (float)steps;
  f2 air = rsi(ray_origin, ray_dir, c_atm.planet_r + c_atm.atmos_height);
  f2 planet = rsi(ray_origin, ray_dir, c_atm.planet_r);
  air.y = planet.x > 0.0f ? fminf(air.y, planet.x) : air.y;
  const float step_delta = (air.y - fmaxf(air.x, 0.0f)) * fsteps;
  const f3 ray_step = ray_dir * step_delta;
  f3 ray_pos = ray_origin + ray_step * 0.5f;
  transmittance = mk3(1.0f);
  in_scatter_col = mk3(0.0f);
  for (int i = 0; i < steps; i++) {
    const float h = get_elevation(ray_pos);
    const f3 density = get_density(h);
    const f3 step_od = extinc_mul(density * step_delta);
    const f3 step_T = saturate3(exp3(-step_od));
    const f3 visible = transmittance * saturate3((mk3(1.0f) - step_T) / step_od);
    const f3 npos = normalize(ray_pos);
    // per-step factors of the two in-scatter terms hoisted out of the 8 sun samples (the
    // reference multiplies left to right inside the loop; same product, different rounding order)
    const f3 a_r = c_atm.rayleigh_coeff * B.sun_col * visible * (density.x * step_delta * 0.125f);
    const f3 a_m = c_atm.mie_coeff * B.sun_col * visible * (density.y * step_delta * 0.125f);
    for (int j = 0; j < 8; j++) {
      f3 sample_dir = sun_basis_sample(B.cosmax, B.sun_dir, sbx, sby, rng);
      float cos_theta = dot(ray_dir, sample_dir);
      float ph_r = rayleigh_phase(cos_theta), ph_m = mie_phase(cos_theta, c_atm.mie_g);
      f3 sun_T = read_trans_lut(B.trans_lut, dot(npos, sample_dir), h);
      in_scatter_col += (a_r * ph_r + a_m * ph_m) * sun_T;
    }
    if (DEPTH == 0) {
      const float ms_energy = 5.3f;
      for (int j = 0; j < 8; j++) {
        f3 sample_dir = get_unit_vec(((float)j + 0.5f) / 8.0f, fractf((float)j * 1.618033988749f));
        float cos_theta = dot(ray_dir, sample_dir);
        float ph_m = mie_phase(cos_theta, c_atm.mie_g);
        f3 amb, amb_T;
        atmospheric_scattering<1>(B, sbx, sby, ray_pos, sample_dir, 5, rng, amb, amb_T);
        in_scatter_col += ms_energy * c_atm.rayleigh_coeff * amb * visible * density.x * step_delta / 8.0f;
        in_scatter_col += ms_energy * c_atm.mie_coeff * amb * visible * ph_m * density.y * step_delta / 8.0f;
      }
    }
    transmittance *= step_T;
    ray_pos += ray_step;
  }
  if (planet.x > 0.0f) transmittance *= 0.0f;
}

__global__ void k_cloud_ambient(SkyBuild B) {
  f3 sbx, sby;
  make_orthonormal_basis(B.sun_dir, sbx, sby);
  Ctr rng{path_key(0xFFFFFFFFu, 2000u, B.seed), 0};
  f3 amb, ambT;
  atmospheric_scattering<0>(B, sbx, sby, c_atm.cam_pos + f3{0.0f, c_atm.cloud_height, 0.0f}, f3{0.0f, 1.0f, 0.0f}, 64, rng, amb, ambT);
  B.cloud_ambient[0] = amb.x, B.cloud_ambient[1] = amb.y, B.cloud_ambient[2] = amb.z;
}

// atmos.py:195-230
HD float sample_cloud_density(const uint8_t* __restrict__ tex, f3 ray_pos) {
  const float tile_size = 29000.0f;
  ray_pos.x = xadd(ray_pos.x, xmul(tile_size, 0.65f));
  ray_pos.z = xadd(ray_pos.z, xmul(tile_size, 0.65f));
  float ux = xdiv(xsub(ray_pos.x, xmul(tile_size, floorf(xdiv(ray_pos.x, tile_size)))), tile_size);
  float uz = xdiv(xsub(ray_pos.z, xmul(tile_size, floorf(xdiv(ray_pos.z, tile_size)))), tile_size);
  int cx = min(max((int)xmul(ux, 256.0f), 0), 255), cy = min(max((int)xmul(uz, 256.0f), 0), 255);
  float relative_height = xsub(xsub(xlength(ray_pos), c_atm.planet_r), c_atm.planet_r_offset);
  const uint8_t* t = tex + (cx * 256 + cy) * 3;
  float tx = (float)t[0] / 255.0f, ty = (float)t[1] / 255.0f, tz = (float)t[2] / 255.0f;
  if (tx < 0.7f) tx = 0.0f;
  if (ty < 0.7f) ty = 0.0f;
  if (tz < 0.7f) tz = 0.0f;
  float cloud = relative_height < c_atm.cloud_height + c_atm.cloud_thickness * 0.65f ? tx : ty;
  bool in_layer = relative_height > c_atm.cloud_height && relative_height < c_atm.cloud_height + c_atm.cloud_thickness;
  return in_layer ? c_atm.cloud_density * tz * cloud : 0.0f;
}
// atmos.py:237-266
HD float clouds_shadow_od(const uint8_t* __restrict__ tex, f3 ray_origin, f3 ray_dir, float dither) {
  const float exponent = 1.6f;
  float step_delta = 24.0f / 8.0f;
  float od = 0.0f;
  f3 ray_pos = ray_origin;
  f3 ray_step{xmul(ray_dir.x, step_delta), xmul(ray_dir.y, step_delta), xmul(ray_dir.z, step_delta)};
  for (int i = 0; i < 8; i++) {
    ray_step = f3{xmul(ray_step.x, exponent), xmul(ray_step.y, exponent), xmul(ray_step.z, exponent)};
    step_delta = xmul(step_delta, exponent);
    f3 dp{xadd(ray_pos.x, xmul(ray_step.x, dither)), xadd(ray_pos.y, xmul(ray_step.y, dither)), xadd(ray_pos.z, xmul(ray_step.z, dither))};
    float rh = xsub(xsub(xlength(dp), c_atm.planet_r), c_atm.planet_r_offset);
    if (rh < c_atm.cloud_height || rh > c_atm.cloud_height + c_atm.cloud_thickness) continue;
    od += sample_cloud_density(tex, dp) * step_delta;
    ray_pos = f3{xadd(ray_pos.x, ray_step.x), xadd(ray_pos.y, ray_step.y), xadd(ray_pos.z, ray_step.z)};
  }
  return od;
}
HD float cloud_phase(float cos_theta, float an) {
  float peak = mie_phase(cos_theta, 0.92f * an);
  float front = mie_phase(cos_theta, 0.4f * an);
  float back = mie_phase(cos_theta, -0.55f * an);
  return mixf(mixf(front, back, 0.5f), peak, 0.15f);
}
// atmos.py:275-349. Positions are advanced with contraction-free ops: at |p| ~ 6.4e6 one ulp is
// 0.5 m and the layer tests (340 m thick) would otherwise flip between oracle and device.
__device__ void clouds_scattering(const SkyBuild& B, f3 sbx, f3 sby, f3 ray_origin, f3 ray_dir, float dither, Ctr& rng, f3& in_scatter,
                                  float& transmittance, float& weighted_dist) {
  const float fsteps = 1.0f / 32.0f;
  float bottom = rsi(ray_origin, ray_dir, c_atm.planet_r + c_atm.planet_r_offset + c_atm.cloud_height).y;
  float top = rsi(ray_origin, ray_dir, c_atm.planet_r + c_atm.planet_r_offset + c_atm.cloud_height + c_atm.cloud_thickness).y;
  transmittance = 1.0f;
  in_scatter = mk3(0.0f);
  float weight_sum = 0.0f;
  weighted_dist = 0.0f;
  f3 start{xadd(ray_origin.x, xmul(ray_dir.x, bottom)), xadd(ray_origin.y, xmul(ray_dir.y, bottom)), xadd(ray_origin.z, xmul(ray_dir.z, bottom))};
  const float step_delta = xmul(xsub(top, bottom), fsteps);
  const f3 ray_step{xmul(ray_dir.x, step_delta), xmul(ray_dir.y, step_delta), xmul(ray_dir.z, step_delta)};
  f3 ray_pos{xadd(start.x, xmul(ray_step.x, dither)), xadd(start.y, xmul(ray_step.y, dither)), xadd(start.z, xmul(ray_step.z, dither))};
  float distance_traveled = xlength(f3{xsub(start.x, ray_origin.x), xsub(start.y, ray_origin.y), xsub(start.z, ray_origin.z)});
  for (int i = 0; i < 32; i++) {
    float density = sample_cloud_density(B.cloud_tex, ray_pos);
    if (density <= 0.0f || transmittance <= 1e-4f) {
      ray_pos = f3{xadd(ray_pos.x, ray_step.x), xadd(ray_pos.y, ray_step.y), xadd(ray_pos.z, ray_step.z)};
      distance_traveled += step_delta;
      weighted_dist += distance_traveled * transmittance;
      weight_sum += transmittance;
      continue;
    }
    float step_od = c_atm.cloud_extinc * density * step_delta;
    float step_T = saturate(expf(-step_od));
    float step_weight = (1.0f - step_T) / c_atm.cloud_extinc;
    float visible = transmittance * step_weight;
    const f3 npos = normalize(ray_pos);
    const float elev = get_elevation(ray_pos);
    for (int j = 0; j < 8; j++) {
      f3 sample_dir = sun_basis_sample(B.cosmax, B.sun_dir, sbx, sby, rng);
      float cos_theta = dot(ray_dir, sample_dir);
      float sun_ray_od = clouds_shadow_od(B.cloud_tex, ray_pos, sample_dir, dither);
      f3 sun_T = read_trans_lut(B.trans_lut, dot(npos, sample_dir), elev);
      float an = 1.0f;
      for (int k = 0; k < 4; k++) {
        float phase = cloud_phase(cos_theta, an);
        in_scatter += (visible * an * c_atm.cloud_scatter * phase * expf(-sun_ray_od * c_atm.cloud_extinc * an)) * sun_T * B.sun_col / 8.0f;
        an *= 0.5f;
      }
    }
    float ambient_od = clouds_shadow_od(B.cloud_tex, ray_pos, f3{0.0f, 1.0f, 0.0f}, dither);
    f3 amb{B.cloud_ambient[0], B.cloud_ambient[1], B.cloud_ambient[2]};
    float an = 1.0f;
    for (int k = 0; k < 4; k++) {
      in_scatter += (visible * an * c_atm.cloud_scatter / (4.0f * VRT_PI) * expf(-ambient_od * c_atm.cloud_extinc * an)) * amb;
      an *= 0.5f;
    }
    transmittance *= step_T;
    ray_pos = f3{xadd(ray_pos.x, ray_step.x), xadd(ray_pos.y, ray_step.y), xadd(ray_pos.z, ray_step.z)};
    distance_traveled += step_delta;
    weighted_dist += distance_traveled * transmittance;
    weight_sum += transmittance;
  }
  weighted_dist /= weight_sum;
}

// accumulate_clouds, all passes of one texel: scatter.xyz += 1.2*in_scatter/n, trans.x += sat(T)/n,
// trans.y += mean distance / n  (atmos.py:140-157)
__global__ void __launch_bounds__(128) k_clouds(SkyBuild B) {
  const int idx = B.first_texel + blockIdx.x * blockDim.x + threadIdx.x;  // this rank's slice of the table (vrt_set_sky_shard)
  if (idx >= B.first_texel + B.n_texels) return;
  const float fres = 1.0f / (float)B.S;
  const int u = idx / B.S, v = idx % B.S;
  f3 sbx, sby;
  make_orthonormal_basis(B.sun_dir, sbx, sby);
  f3 ray_dir = unproject_sky(f2{((float)u + 0.5f) * fres, ((float)v + 0.5f) * fres}, fres);
  const float fmax_samples = 1.0f / (float)B.cloud_passes;
  f3 sc = mk3(0.0f);
  float tx = 0.0f, ty = 0.0f;
  for (int pass = 0; pass < B.cloud_passes; pass++) {
    Ctr rng{path_key((uint32_t)idx, (uint32_t)pass, B.seed), 0};
    float dither = rng.next();
    f3 insc;
    float T, dist;
    clouds_scattering(B, sbx, sby, c_atm.cam_pos, ray_dir, dither, rng, insc, T, dist);
    insc *= 1.2f;
    sc += insc * fmax_samples;
    tx += saturate(T) * fmax_samples;
    ty += dist * fmax_samples;
  }
  B.scatter[idx] = make_float4(sc.x, sc.y, sc.z, 0.0f);
  B.trans[idx] = make_float4(tx, ty, 0.0f, 0.0f);
}

// compute_skybox (atmos.py:159-189)
__global__ void __launch_bounds__(128) k_skybox(SkyBuild B) {
  const int idx = B.first_texel + blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B.first_texel + B.n_texels) return;
  const float fres = 1.0f / (float)B.S;
  const int u = idx / B.S, v = idx % B.S;
  f3 sbx, sby;
  make_orthonormal_basis(B.sun_dir, sbx, sby);
  f3 ray_dir = unproject_sky(f2{((float)u + 0.5f) * fres, ((float)v + 0.5f) * fres}, fres);
  float4 cs = B.scatter[idx], ct = B.trans[idx];
  f3 cloud_in_scatter{cs.x, cs.y, cs.z};
  float cloud_T = ct.x, cloud_dist = ct.y;
  Ctr rng{path_key((uint32_t)idx, 1000u, B.seed), 0};
  f3 sc_total, T_total, sc_from, T_from;
  atmospheric_scattering<0>(B, sbx, sby, c_atm.cam_pos, ray_dir, 64, rng, sc_total, T_total);
  f3 cloud_pos = c_atm.cam_pos + ray_dir * fmaxf(cloud_dist, 0.0f);
  atmospheric_scattering<0>(B, sbx, sby, cloud_pos, ray_dir, 64, rng, sc_from, T_from);
  f3 T_to_cloud = T_total / T_from;
  f3 in_scattering = sc_total;
  if (B.use_clouds == 1) {
    in_scattering = in_scattering - sc_from * saturate3(T_to_cloud * fmaxf(1.0f - cloud_T, 0.0f));
    in_scattering += cloud_in_scatter * saturate3(T_to_cloud);
  }
  f3 Tout = T_total * cloud_T;
  B.scatter[idx] = make_float4(in_scattering.x, in_scattering.y, in_scattering.z, 0.0f);
  B.trans[idx] = make_float4(Tout.x, Tout.y, Tout.z, 0.0f);
}

}  // namespace

cudaError_t vrt_launch_sky_precompute(const SkyBuild& B, cudaStream_t st) {
  Atm a;
  a.rayleigh_coeff = f3{0.00000519673f, 0.0000121427f, 0.0000296453f};
  a.mie_coeff = 8.6e-6f;
  {
    double air = 2.5035422e25, ozone_peak = 8e-6;
    double ozone_num = air * 0.012588 * ozone_peak;
    a.ozone_coeff = f3{(float)(4.51103766177301e-21 * 0.0001 * ozone_num), (float)(3.2854797958699e-21 * 0.0001 * ozone_num),
                       (float)(1.96774621921165e-22 * 0.0001 * ozone_num)};
  }
  a.mie_ext = (float)(8.6e-6 * 1.11);
  a.scale_height_rayl = 8500.0f, a.scale_height_mie = 1200.0f, a.mie_g = 0.75f;
  a.planet_r = 6371e3f, a.planet_r_offset = 0.0f, a.atmos_height = 110e3f;
  a.cloud_height = 2000.0f, a.cloud_thickness = 340.0f, a.cloud_density = 0.27f, a.cloud_extinc = 0.075f, a.cloud_scatter = 0.075f;
  a.cam_pos = f3{0.0f, (float)(6371e3 + 0.0 + 1e3), 0.0f};
  cudaError_t e = cudaMemcpyToSymbolAsync(c_atm, &a, sizeof(Atm), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  k_trans_lut<<<(256 * 128 + 127) / 128, 128, 0, st>>>(B.trans_lut);
  k_cloud_ambient<<<1, 1, 0, st>>>(B);
  const int n = B.n_texels;
  if (n <= 0) return cudaGetLastError();
  k_clouds<<<(n + 127) / 128, 128, 0, st>>>(B);
  k_skybox<<<(n + 127) / 128, 128, 0, st>>>(B);
  return cudaGetLastError();
}
