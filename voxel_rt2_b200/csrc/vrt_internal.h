// Internal launcher interface between the translation units of libvoxelrt.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/voxelrt.h"
#include "vrt_common.cuh"
#include "vrt_restir.cuh"

// vrt_render.cu
cudaError_t vrt_launch_primary(const Params& P, vrt_hit* out, cudaStream_t st);
cudaError_t vrt_launch_path(const Params& P, bool stats, int sm_count, cudaStream_t st, int* blocks_out);
cudaError_t vrt_launch_jitter(float2* out, int first_sample, int stride, int n, int W, int H, int mode, unsigned int* work_counter, cudaStream_t st);
cudaError_t vrt_launch_merge_slice(const float4* accum, const float4* const* peers, int n_peers, float4* sum_out, float4* ldr_out, int first,
                                   int count, int W, int H, float exposure, cudaStream_t st);
cudaError_t vrt_launch_path_restir(const Params& P, const RestirBuffers& RB, int sm_count, cudaStream_t st);
cudaError_t vrt_launch_path_moving(const Params& P, const MovingOut& MO, int sm_count, cudaStream_t st);
// vrt_temporal.cu — moving-camera temporal filters (pathtracer.py:993-1303)
cudaError_t vrt_launch_moving_filters(const Params& P, const MovingFrame& F, float max_accum, cudaStream_t st);
cudaError_t vrt_launch_moving_upsample(const float4* out, float4* full, int W, int H, float scale, cudaStream_t st);
size_t vrt_render_smem_bytes(const Params& P, int* upper_in_smem);
cudaError_t vrt_launch_pack_sky(const float4* scatter, const float4* trans, uint4* out, size_t n, cudaStream_t st);
// vrt_restir.cu — spatial_GRIS (pathtracer.py:815-989); adds the frame's colour into P.accum
cudaError_t vrt_launch_gris(const Params& P, const RestirBuffers& RB, uint32_t frame, cudaStream_t st);
// sun transmittance at every reservoir's reconnection vertex, once per frame before the resampling passes (vrt_launch_temporal
// includes it; without temporal reuse the frame loop calls it before vrt_launch_gris)
cudaError_t vrt_launch_rc_sky(const Params& P, const RestirBuffers& RB, cudaStream_t st);
// temporal reservoir reuse between the path kernel and k_gris (k_rc_sky + k_temporal)
cudaError_t vrt_launch_temporal(const Params& P, const RestirBuffers& RB, uint32_t frame, int hist_valid, cudaStream_t st);
cudaError_t vrt_launch_resolve_merged(const float4* accum, const float4* const* peers, int n_peers, float4* ldr, int W, int H, float exposure,
                                      cudaStream_t st);
cudaError_t vrt_launch_resolve(const float4* accum, float4* hdr, float4* ldr, int W, int H, float exposure, cudaStream_t st);

// vrt_build.cu — voxel arrays ([x][y][z], z fastest) -> bricks, colour SoA, upper pyramid
cudaError_t vrt_launch_build(const int8_t* d_mat, const uint8_t* d_rgb, int R, unsigned long long* bricks, uint32_t* color,
                             uint32_t* upper, const uint32_t* upper_off_host, int n_lods, int upper_words, cudaStream_t st);

// vrt_sky_precompute.cu
struct SkyBuild {
  int S;
  float4* scatter;  // [S][S]
  float4* trans;
  __half* trans_lut;          // [256][128][3]
  const uint8_t* cloud_tex;   // [256][256][3]
  float* cloud_ambient;       // 3 floats (device)
  f3 sun_dir, sun_col;
  float cosmax;
  int use_clouds;
  int cloud_passes;
  uint32_t seed;
  int first_texel, n_texels;  // the slice of the S x S tables this call computes (texel = x * S + y)
};
cudaError_t vrt_launch_sky_precompute(const SkyBuild& B, cudaStream_t st);
