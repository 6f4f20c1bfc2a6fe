// Run-time sky lookups: renderer/atmos.py:94-131 (sample_skybox, sample_skybox_transmittance)
// and :428-455 (project_sky / unproject_sky). Tables are float4 texels [x][y] (y fastest), so a
// bilinear footprint is two 32-byte pairs per table: (x, y..y+1) and (x+1, y..y+1).
#pragma once
#include "vrt_common.cuh"

HD f2 project_sky(f3 d, float fres) {
  float il = 1.0f / sqrtf(d.x * d.x + d.z * d.z);
  float px = il * d.x, py = il * d.z;
  float azimuth = VRT_PI + atan2f(px, -py);
  float elevation = VRT_PI * 0.5f - acosf(d.y);
  float cx = azimuth / (VRT_PI * 2.0f);
  float cy = 0.5f + 0.5f * signf(elevation) * sqrtf(2.0f / VRT_PI * fabsf(elevation));
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
HD f3 sky_fetch(const float4* __restrict__ tab, const SkyTap& t) {
  float4 bl = __ldg(tab + t.i00), br = __ldg(tab + t.i10), tl = __ldg(tab + t.i01), tr = __ldg(tab + t.i11);
  f3 a = mix3(f3{bl.x, bl.y, bl.z}, f3{br.x, br.y, br.z}, t.wx);
  f3 b = mix3(f3{tl.x, tl.y, tl.z}, f3{tr.x, tr.y, tr.z}, t.wx);
  return mix3(a, b, t.wy);
}
