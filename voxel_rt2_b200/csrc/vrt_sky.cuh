// Run-time sky lookups: renderer/atmos.py:94-131 (sample_skybox, sample_skybox_transmittance)
// and :428-455 (project_sky / unproject_sky). Tables are float4 texels [x][y] (y fastest), so a
// bilinear footprint is two 32-byte pairs per table: (x, y..y+1) and (x+1, y..y+1).
#pragma once
#include <cuda_fp16.h>

#include "vrt_common.cuh"

// Polynomial atan2 / acos (max abs error 1.7e-7 / 3.3e-7 rad = 1e-4 / 2e-4 sky texels at 3840^2;
// coefficients: Chebyshev fits evaluated in float32, see tests/test_host.py::test_fast_trig_error).
HD float fast_atan2(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float a = __fdividef(mn, mx);
  const float z = a * a;
  float r = -0.004668773151934147f;
  r = r * z + 0.02416618913412094f;
  r = r * z - 0.0593671016395092f;
  r = r * z + 0.09906096756458282f;
  r = r * z - 0.14016585052013397f;
  r = r * z + 0.19969235360622406f;
  r = r * z - 0.33331960439682007f;
  r = r * z + 0.9999998807907104f;
  r *= a;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.0f) r = VRT_PI - r;
  return y < 0.0f ? -r : r;
}
HD float fast_acos(float x) {
  const float a = fabsf(x);
  float r = 0.002251368248835206f;
  r = r * a - 0.011012386530637741f;
  r = r * a + 0.02674933150410652f;
  r = r * a - 0.048724401742219925f;
  r = r * a + 0.08873733133077621f;
  r = r * a - 0.21458369493484497f;
  r = r * a + 1.5707961320877075f;
  r *= fsqrt(fmaxf(1.0f - a, 0.0f));
  return x < 0.0f ? VRT_PI - r : r;
}

HD f2 project_sky(f3 d, float fres) {
  // atan2 is scale invariant: the normalisation of d.xz (atmos.py:430) is not needed
  float azimuth = VRT_PI + fast_atan2(d.x, -d.z);
  float elevation = VRT_PI * 0.5f - fast_acos(d.y);
  float cx = azimuth * (1.0f / (VRT_PI * 2.0f));
  float cy = 0.5f + 0.5f * signf(elevation) * fsqrt(2.0f / VRT_PI * fabsf(elevation));
  return f2{cx * (1.0f - fres) + 0.5f * fres, cy * (1.0f - fres) + 0.5f * fres};
}

HD f3 unproject_sky(f2 uv, float fres) {
  float cx = (uv.x - 0.5f * fres) / (1.0f - 1.0f * fres);
  float cy = (uv.y - 0.5f * fres) / (1.0f - 1.0f * fres);
  cy = cy < 0.5f ? -sqr(1.0f - 2.0f * cy) : sqr(2.0f * cy - 1.0f);
  float azimuth = cx * 2.0f * VRT_PI - VRT_PI;
  float elevation = cy * 0.5f * VRT_PI;
  float se, ce, sa, ca;
  sincosf(elevation, &se, &ce);
  sincosf(azimuth, &sa, &ca);
  return normalize(f3{ce * sa, se, -ce * ca});
}

struct SkyTap {
  int i00, i10, i01, i11;
  float wx, wy;
};
HD SkyTap sky_tap(int S, f2 tc) {
  float fx = tc.x * (float)S - 0.5f, fy = tc.y * (float)S - 0.5f;
  int ix = (int)fx, iy = (int)fy;
  SkyTap t;
  t.wx = fractf(fx), t.wy = fractf(fy);
  ix = min(max(ix, 0), S - 1);
  iy = min(max(iy, 0), S - 1);
  int ix1 = ix + 1 == S ? 0 : ix + 1, iy1 = iy + 1 == S ? 0 : iy + 1;
  t.i00 = ix * S + iy, t.i10 = ix1 * S + iy, t.i01 = ix * S + iy1, t.i11 = ix1 * S + iy1;
  return t;
}
HD float4 sky_ld(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
HD f3 sky_fetch(const float4* __restrict__ tab, const SkyTap& t) {
  // the tables are 2 x 236 MB walked at random: their texels must not evict the occupancy bricks from L1
  // (no-allocate loads: +1 % dense, +2 % example6)
  float4 bl = sky_ld(tab + t.i00), br = sky_ld(tab + t.i10), tl = sky_ld(tab + t.i01), tr = sky_ld(tab + t.i11);
  f3 a = mix3(f3{bl.x, bl.y, bl.z}, f3{br.x, br.y, br.z}, t.wx);
  f3 b = mix3(f3{tl.x, tl.y, tl.z}, f3{tr.x, tr.y, tr.z}, t.wx);
  return mix3(a, b, t.wy);
}

// Format-1 table: one 16-byte texel holds scattering and transmittance as binary16, so an escaped
// segment needs 4 loads instead of 8 and a visible sun sample reads the same 4 texels.
HD void sky_unpack(uint4 t, f3& sc, f3& tr) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y)),
               c = __half22float2(*reinterpret_cast<const __half2*>(&t.z));
  sc = f3{a.x, a.y, b.x};
  tr = f3{b.y, c.x, c.y};
}
HD uint4 sky_ld_packed(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
HD void sky_fetch_packed(const uint4* __restrict__ tab, const SkyTap& t, f3& sc, f3& tr) {
  const uint4 bl = sky_ld_packed(tab + t.i00), br = sky_ld_packed(tab + t.i10), tl = sky_ld_packed(tab + t.i01), trr = sky_ld_packed(tab + t.i11);
  f3 s0, t0, s1, t1, s2, t2, s3, t3;
  sky_unpack(bl, s0, t0), sky_unpack(br, s1, t1), sky_unpack(tl, s2, t2), sky_unpack(trr, s3, t3);
  sc = mix3(mix3(s0, s1, t.wx), mix3(s2, s3, t.wx), t.wy);
  tr = mix3(mix3(t0, t1, t.wx), mix3(t2, t3, t.wx), t.wy);
}
