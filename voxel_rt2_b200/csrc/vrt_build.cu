// Voxel upload -> device layout. Replaces VoxelWorld._make_texture (renderer/voxel_world.py:69-87)
// and VoxelOctreeRaytracer._update_lods (renderer/raytracer.py:46-70).
//
// Input: material int8 [x][y][z] and colour uint8 [x][y][z][3] (z fastest), the NumPy layout of
// the host Scene. Output:
//   bricks : one 64-bit word per 4^3 block, bit = (z&3)*16 + (y&3)*4 + (x&3), set iff material > 0
//   color  : RGBA8 per voxel, brick-major (brick*64 + bit): r | g<<8 | b<<16 | max(material,0)<<24
//   upper  : bit arrays of LOD 3..n_lods-1 (bit = OR of the 2^3 children)
#include "vrt_internal.h"

__global__ void __launch_bounds__(256) k_build_bricks(const int8_t* __restrict__ mat, const uint8_t* __restrict__ rgb, int R,
                                                      unsigned long long* __restrict__ bricks, uint32_t* __restrict__ color) {
  // one thread per voxel, thread index follows the input order so the reads coalesce
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)R * R * R;
  if (i >= n) return;
  const int z = (int)(i % R), y = (int)((i / R) % R), x = (int)(i / ((size_t)R * R));
  const int m = mat[i];
  const uint32_t c = (uint32_t)rgb[3 * i] | ((uint32_t)rgb[3 * i + 1] << 8) | ((uint32_t)rgb[3 * i + 2] << 16) |
                     ((uint32_t)(m > 0 ? m : 0) << 24);
  const int br = R >> 2;
  const size_t b = ((size_t)(z >> 2) * br + (y >> 2)) * br + (x >> 2);
  const int bit = (z & 3) * 16 + (y & 3) * 4 + (x & 3);
  color[b * 64 + bit] = c;
  if (m > 0) atomicOr(bricks + b, 1ull << bit);
}

// LOD 3 from bricks (each LOD-3 cell = 2^3 bricks)
__global__ void __launch_bounds__(256) k_build_lod3(const unsigned long long* __restrict__ bricks, int R, uint32_t* __restrict__ upper) {
  const int r = R >> 3, br = R >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r * r * r) return;
  const int x = i % r, y = (i / r) % r, z = i / (r * r);
  unsigned long long any = 0ull;
#pragma unroll
  for (int c = 0; c < 8; c++)
    any |= bricks[((size_t)(2 * z + (c >> 2)) * br + (2 * y + ((c >> 1) & 1))) * br + (2 * x + (c & 1))];
  if (any) atomicOr(upper + (i >> 5), 1u << (i & 31));
}

// LOD l (>= 4) from LOD l-1
__global__ void __launch_bounds__(256) k_build_lod(const uint32_t* __restrict__ child, int rc, uint32_t* __restrict__ parent) {
  const int r = rc >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r * r * r) return;
  const int x = i % r, y = (i / r) % r, z = i / (r * r);
  bool any = false;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    uint32_t idx = (uint32_t)(((2 * z + (c >> 2)) * rc + (2 * y + ((c >> 1) & 1))) * rc + (2 * x + (c & 1)));
    any |= (child[idx >> 5] >> (idx & 31)) & 1u;
  }
  if (any) atomicOr(parent + (i >> 5), 1u << (i & 31));
}

cudaError_t vrt_launch_build(const int8_t* d_mat, const uint8_t* d_rgb, int R, unsigned long long* bricks, uint32_t* color,
                             uint32_t* upper, const uint32_t* upper_off, int n_lods, int upper_words, cudaStream_t st) {
  const size_t n = (size_t)R * R * R;
  const size_t nb = n / 64;
  cudaError_t e = cudaMemsetAsync(bricks, 0, nb * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  if (upper_words > 0) {
    e = cudaMemsetAsync(upper, 0, (size_t)upper_words * 4, st);
    if (e != cudaSuccess) return e;
  }
  k_build_bricks<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_mat, d_rgb, R, bricks, color);
  if (n_lods > 3) {
    int r = R >> 3;
    k_build_lod3<<<(r * r * r + 255) / 256, 256, 0, st>>>(bricks, R, upper + upper_off[0]);
    for (int l = 4; l < n_lods; l++) {
      int rc = R >> (l - 1);
      int rp = rc >> 1;
      k_build_lod<<<(rp * rp * rp + 255) / 256, 256, 0, st>>>(upper + upper_off[l - 4], rc, upper + upper_off[l - 3]);
    }
  }
  return cudaGetLastError();
}
