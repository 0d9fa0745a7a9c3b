// reduce.cuh — deterministic column reduction of [rows][n] partials (split-K weight gradients).
#pragma once
#include "common.cuh"

namespace mcgm {

static constexpr int kRedLanes = 8;

template <typename V>
__device__ __forceinline__ V red_zero();
template <> __device__ __forceinline__ float4 red_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float red_zero<float>() { return 0.f; }
__device__ __forceinline__ void red_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void red_add(float& a, const float& b) { a += b; }

template <typename V, int LANES = kRedLanes>
__device__ __forceinline__ void reduce_columns(const V* __restrict__ src, int rows, size_t row_stride, size_t n,
                                               size_t i0, V* __restrict__ dst, V (*sm)[32]) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t i = i0 + lane;
  V a = red_zero<V>();
  if (i < n) {
    int s = w;
    for (; s + 3 * LANES < rows; s += 4 * LANES) {
      V b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = src[(size_t)(s + u * LANES) * row_stride + i];
#pragma unroll
      for (int u = 0; u < 4; ++u) red_add(a, b[u]);
    }
    for (; s < rows; s += LANES) red_add(a, src[(size_t)s * row_stride + i]);
  }
  sm[w][lane] = a;
  __syncthreads();
  if (w == 0 && i < n) {
    V t = sm[0][lane];
#pragma unroll
    for (int k = 1; k < LANES; ++k) red_add(t, sm[k][lane]);
    dst[i] = t;
  }
  __syncthreads();
}


}  // namespace mcgm
