// Moving-camera temporal path. Replaces, in renderer/pathtracer.py: temporal_filter_prepass
// :1020-1075, bilinear_sample :1077-1090, history_filter(_specular) :1092-1183, reproject :991-998,
// temporal_filter :1185-1230, temporal_filter_specular :1242-1303 (the slot-1 -> slot-0 and
// previous-G-buffer copies of :1297-1303 are pointer swaps on the host) and the nearest-neighbour
// up-sampling of _render_to_image :643-644. One thread per pixel of the render area; history taps
// read the previous frame's slot, so there is no intra-kernel hazard. Pins: see DESIGN.md
// ("moving-camera pins") and oracle.cpp.
#include "vrt_internal.h"
#include "vrt_restir.cuh"
#include "vrt_trace.cuh"

namespace {

HD bool bad3(f3 c) { return isbad(c.x) || isbad(c.y) || isbad(c.z) || c.x < 0.0f || c.y < 0.0f || c.z < 0.0f; }
HD bool outside_area(const Params& P, float scale, int u, int v) { return (float)u > scale * (float)P.W || (float)v > scale * (float)P.H; }

__global__ void __launch_bounds__(256) k_mv_prepass(const __grid_constant__ Params P, MovingFrame F) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.W * P.H) return;
  const int u = idx % P.W, v = idx / P.W;
  if (outside_area(P, F.scale, u, v)) return;
  const int irx = (int)((float)P.W * F.scale), iry = (int)((float)P.H * F.scale);
  float sum = 0.0f, valid = 0.0f;
  for (int x = -1; x < 3; x++)
    for (int y = -1; y < 3; y++) {
      const int tx = u + x, ty = v + y;
      if (tx < 0 || ty < 0 || tx > irx - 1 || ty > iry - 1) continue;
      const float r = F.refl[(size_t)ty * P.W + tx];
      if (r != 0.0f) valid += 1.0f, sum += r;
    }
  F.refl_blur[idx] = valid > 0.01f ? sum / valid : 0.0f;
  float4 d = F.col_d[idx], s = F.col_s[idx];
  if (bad3(f3{d.x, d.y, d.z})) F.col_d[idx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (bad3(f3{s.x, s.y, s.z})) F.col_s[idx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

HD float catmullrom(float x) {
  const float x2 = x * x, x3 = x * x * x;
  float fx = 0.0f;
  if (x < 1.0f)
    fx = 1.5f * x3 - 2.5f * x2 + 1.0f;
  else if (x < 2.0f)
    fx = -0.5f * x3 + 2.5f * x2 - 4.0f * x + 2.0f;
  return fx;
}
HD f3 ld3(const float4* b, int W, int H, int x, int y) {
  x = min(max(x, 0), W - 1), y = min(max(y, 0), H - 1);
  const float4 t = b[(size_t)y * W + x];
  return f3{t.x, t.y, t.z};
}
HD f3 mv_bilinear(const float4* buf, int W, int H, float uvx, float uvy, int irx, int iry) {
  const float fx = uvx * (float)irx - 0.5f, fy = uvy * (float)iry - 0.5f;
  const int ix = (int)fx, iy = (int)fy;
  const float wx = fractf(fx), wy = fractf(fy);
  return mix3(mix3(ld3(buf, W, H, ix, iy), ld3(buf, W, H, ix + 1, iy), wx), mix3(ld3(buf, W, H, ix, iy + 1), ld3(buf, W, H, ix + 1, iy + 1), wx), wy);
}
HD f3 mv_reproject(const MovingFrame& F, f3 p) {
  float a[4], q[4];
  mat4_mul(F.prev_view, p.x, p.y, p.z, 1.0f, a);
  mat4_mul(F.prev_proj, a[0], a[1], a[2], a[3], q);
  return f3{q[0] / q[3] * 0.5f + 0.5f, q[1] / q[3] * 0.5f + 0.5f, q[2] / q[3] * 0.5f + 0.5f};
}
template <bool SPECULAR>
HD float mv_history(const Params& P, const MovingFrame& F, float uvx, float uvy, float center_depth, f3 center_normal, int irx, int iry, float4& col_out,
                    float& depth_out) {
  col_out = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
  depth_out = 0.0f;
  if (isbad(uvx) || isbad(uvy) || fabsf(uvx) > 1e6f || fabsf(uvy) > 1e6f) return 0.0f;
  const float fx = uvx * (float)irx - 0.5f, fy = uvy * (float)iry - 0.5f;
  const int ix = (int)fx, iy = (int)fy;
  const float ffx = fractf(fx), ffy = fractf(fy);
  float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f), cmax = sum, cmin = make_float4(999999.0f, 999999.0f, 999999.0f, 999999.0f);
  float dsum = 0.0f, dmax = 0.0f, dmin = 999999.0f, wsum = 0.0f;
  const float4* hist = SPECULAR ? F.hs_prev : F.hd_prev;
  for (int x = -1; x < 3; x++)
    for (int y = -1; y < 3; y++) {
      const int tx = ix + x, ty = iy + y;
      if (tx < 0 || ty < 0 || tx > irx - 1 || ty > iry - 1) continue;
      const size_t ti = (size_t)ty * P.W + tx;
      float w = catmullrom(fabsf((float)x - ffx)) * catmullrom(fabsf((float)y - ffy));
      const uint32_t no = F.attr_prev[ti].x;
      const f3 tap_normal = decode_unit_vector_3x16(h16val(no), h16val(no >> 16));
      if (!SPECULAR) {
        const float tap_depth = linearize_depth(P, F.depth_prev[ti]);
        w *= fabsf(tap_depth - center_depth) / center_depth < 0.05f ? 1.0f : 0.0f;
      }
      w *= dot(center_normal, tap_normal) > 0.642f ? 1.0f : 0.0f;
      const float4 col = hist[ti];
      cmax = make_float4(fmaxf(cmax.x, col.x), fmaxf(cmax.y, col.y), fmaxf(cmax.z, col.z), fmaxf(cmax.w, col.w));
      cmin = make_float4(fminf(cmin.x, col.x), fminf(cmin.y, col.y), fminf(cmin.z, col.z), fminf(cmin.w, col.w));
      sum = make_float4(sum.x + col.x * w, sum.y + col.y * w, sum.z + col.z * w, sum.w + col.w * w);
      if (SPECULAR) {
        const float rd = F.hsd_prev[ti];
        dmin = fminf(dmin, rd), dmax = fmaxf(dmax, rd);
        dsum += rd * w;
      }
      wsum += w;
    }
  sum = make_float4(sum.x / wsum, sum.y / wsum, sum.z / wsum, sum.w / wsum);
  dsum /= wsum;
  col_out = make_float4(fmaxf(clampf(sum.x, cmin.x, cmax.x), 0.0f), fmaxf(clampf(sum.y, cmin.y, cmax.y), 0.0f), fmaxf(clampf(sum.z, cmin.z, cmax.z), 0.0f),
                        fmaxf(clampf(sum.w, cmin.w, cmax.w), 1.0f));
  depth_out = clampf(dsum, dmin, dmax);
  return wsum;
}

__global__ void __launch_bounds__(128) k_mv_filter(const __grid_constant__ Params P, MovingFrame F, float max_accum) {
  __shared__ float s_unorm[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_unorm[i] = xdiv((float)i, 255.0f);
  __syncthreads();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P.W * P.H) return;
  const int u = idx % P.W, v = idx / P.W;
  if (outside_area(P, F.scale, u, v)) return;
  const int irx = (int)((float)P.W * F.scale), iry = (int)((float)P.H * F.scale);
  const float4 cd = F.col_d[idx];
  f3 out{cd.x, cd.y, cd.z};
  const float tcx = ((float)u + 0.5f) * P.inv_w / F.scale, tcy = ((float)v + 0.5f) * P.inv_h / F.scale;
  const float d_nl = F.depth[idx];
  const uint2 at = F.attr[idx];
  const f3 center_n1 = decode_unit_vector_3x16(h16val(at.x), h16val(at.x >> 16));
  const f3 center_x1 = view_to_world(P, screen_to_view(P, tcx, tcy, d_nl));
  if (!is_vec_zero(center_x1)) {
    {  // diffuse (pathtracer.py:1185-1230)
      const f3 current = mv_bilinear(F.col_d, P.W, P.H, tcx, tcy, irx, iry);
      const f3 rp = mv_reproject(F, center_x1);
      float4 history;
      float dummy;
      const float w_sum = mv_history<false>(P, F, rp.x, rp.y, linearize_depth(P, rp.z), center_n1, irx, iry, history, dummy);
      if (w_sum > 1e-3f) {
        history.w = fminf(history.w + 1.0f, max_accum);
        const float t = 1.0f / history.w;
        history.x = mixf(history.x, current.x, t), history.y = mixf(history.y, current.y, t), history.z = mixf(history.z, current.z, t);
      } else {
        history = make_float4(current.x, current.y, current.z, 1.0f);
      }
      F.hd[idx] = history;
      const f3 albedo{s_unorm[(at.y >> 8) & 255u], s_unorm[(at.y >> 16) & 255u], s_unorm[(at.y >> 24) & 255u]};
      out = f3{history.x, history.y, history.z} * albedo;
    }
    {  // specular through the virtual reflection point (pathtracer.py:1242-1295)
      const float center_refl_depth = F.refl_blur[idx];
      const f3 center_refl_pos = view_to_world(P, screen_to_view(P, tcx, tcy, delinearize_depth(P, center_refl_depth)));
      const f3 current = mv_bilinear(F.col_s, P.W, P.H, tcx, tcy, irx, iry);
      const f3 rp = mv_reproject(F, center_refl_depth != 0.0f ? center_refl_pos : center_x1);
      float4 history;
      float refl_hist;
      const float w_sum = mv_history<true>(P, F, rp.x, rp.y, linearize_depth(P, rp.z), center_n1, irx, iry, history, refl_hist);
      if (w_sum > 1e-3f) {
        history.w = fminf(history.w + 1.0f, max_accum);
        const float t = 1.0f / history.w;
        history.x = mixf(history.x, current.x, t), history.y = mixf(history.y, current.y, t), history.z = mixf(history.z, current.z, t);
        refl_hist = mixf(refl_hist, center_refl_depth, t);
      } else {
        history = make_float4(current.x, current.y, current.z, 1.0f);
        refl_hist = center_refl_depth;
      }
      F.hs[idx] = history;
      F.hsd[idx] = refl_hist;
      out += f3{history.x, history.y, history.z};
    }
  }
  F.out[idx] = make_float4(out.x, out.y, out.z, 1.0f);
}

// _render_to_image's nearest-neighbour fetch at render_scale (pathtracer.py:643-644): expands the
// half-resolution colour buffer to a full-resolution "accumulation" image with w = 1.
__global__ void __launch_bounds__(256) k_mv_upsample(const float4* __restrict__ out, float4* __restrict__ full, int W, int H, float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * H) return;
  const int i = idx % W, j = idx / W;
  const int sx = (int)((float)i * scale), sy = (int)((float)j * scale);
  float4 c = out[(size_t)sy * W + sx];
  c.w = 1.0f;
  full[idx] = c;
}

}  // namespace

cudaError_t vrt_launch_moving_filters(const Params& P, const MovingFrame& F, float max_accum, cudaStream_t st) {
  const int n = P.W * P.H;
  k_mv_prepass<<<(n + 255) / 256, 256, 0, st>>>(P, F);
  k_mv_filter<<<(n + 127) / 128, 128, 0, st>>>(P, F, max_accum);
  return cudaGetLastError();
}
cudaError_t vrt_launch_moving_upsample(const float4* out, float4* full, int W, int H, float scale, cudaStream_t st) {
  const int n = W * H;
  k_mv_upsample<<<(n + 255) / 256, 256, 0, st>>>(out, full, W, H, scale);
  return cudaGetLastError();
}
