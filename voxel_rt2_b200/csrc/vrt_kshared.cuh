// Pieces shared by the path kernels (vrt_render.cu: k_path, vrt_wave.cu: k_wave): radiance clamp, MIS
// heuristic, material packing and the per-CTA shared-memory staging of the small tables.
#pragma once
#include "vrt_bsdf.cuh"
#include "vrt_common.cuh"

#define RADIANCE_CLAMP 300.0f  // pathtracer.py:20
HD f3 firefly_filter(f3 v) { return clamp3(v, 0.0f, RADIANCE_CLAMP); }
HD float power_heuristic(float a, float b) {  // pathtracer.py:349-353
  float a_sqr = a * a;
  return __fdividef(a_sqr, fmaxf(a_sqr + b * b, 1e-4f));
}
HD uint32_t encode_material(int mat_id, f3 albedo) {  // math_utils.py:231-236
  return (uint32_t)mat_id | ((uint32_t)(albedo.x * 255.0f) << 8) | ((uint32_t)(albedo.y * 255.0f) << 16) |
         ((uint32_t)(albedo.z * 255.0f) << 24);
}
HD bool bad3(f3 c) { return isbad(c.x) || isbad(c.y) || isbad(c.z) || c.x < 0.0f || c.y < 0.0f || c.z < 0.0f; }

#define SMEM_MAT_WORDS (128 * MAT_ROW_F4 * 4)  // 128 material rows x 20 floats
#define SMEM_UNORM_WORDS 256                   // k / 255.0f for the RGBA8 colour decode
#define SMEM_FIXED_WORDS (SMEM_MAT_WORDS + SMEM_UNORM_WORDS)

// Stage the material table, the UNORM8 decode table and (when it fits) the upper occupancy
// pyramid in shared memory.
HD const uint32_t* stage_shared(const Params& P, uint32_t* smem, int upper_in_smem) {
  float4* s_mats = reinterpret_cast<float4*>(smem);
  for (int i = threadIdx.x; i < 128 * MAT_ROW_F4; i += blockDim.x) s_mats[i] = P.mats[i];
  float* s_unorm = reinterpret_cast<float*>(smem + SMEM_MAT_WORDS);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_unorm[i] = xdiv((float)i, 255.0f);
  const uint32_t* upper = P.upper;
  if (upper_in_smem) {
    uint32_t* s_upper = smem + SMEM_FIXED_WORDS;
    for (int i = threadIdx.x; i < P.upper_words; i += blockDim.x) s_upper[i] = P.upper[i];
    upper = s_upper;
  }
  __syncthreads();
  return upper;
}

