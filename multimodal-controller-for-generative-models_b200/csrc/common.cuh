// common.cuh — shared helpers for libmcgm.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mcgm.h"

namespace mcgm {

// ---- error plumbing (thread-local message, negative return codes) ----------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);

#define MCGM_CHECK_ARG(cond, ...)                          \
  do {                                                     \
    if (!(cond)) return ::mcgm::fail(MCGM_E_ARG, __VA_ARGS__); \
  } while (0)

#define MCGM_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return ::mcgm::fail(MCGM_E_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                          cudaGetErrorString(_e), __FILE__, __LINE__);               \
  } while (0)

#define MCGM_LAUNCH_CHECK()                                                          \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      return ::mcgm::fail(MCGM_E_CUDA, "kernel launch failed: %s (%s:%d)",           \
                          cudaGetErrorString(_e), __FILE__, __LINE__);               \
  } while (0)

int sm_count();

// ---- 16-byte vector access ---------------------------------------------------
template <typename T> struct Vec;   // VEC elements of T in one 16-byte access

template <> struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// raw 16-byte load (keeps 4 registers until the values are needed) + conversion
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <typename T> __device__ __forceinline__ void unpack16(const uint4& r, float* v);
template <> __device__ __forceinline__ void unpack16<float>(const uint4& r, float* v) {
  v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& r, float* v) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// pixel-index arithmetic: shifts when the extent is a power of two, else 32-bit division
struct PixDiv {
  uint32_t d;
  int shift;   // >= 0: power of two
  __host__ __device__ static PixDiv make(uint32_t d) {
    PixDiv r; r.d = d; r.shift = -1;
    for (int s = 0; s < 31; ++s) if ((1u << s) == d) r.shift = s;
    return r;
  }
  __device__ __forceinline__ uint32_t div(uint32_t x) const { return shift >= 0 ? (x >> shift) : (x / d); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// bits of channels [c0, c0+n) (n <= 32, c0 % n == 0, n power of two) of mask row `row`
__device__ __forceinline__ uint32_t mask_bits(const uint32_t* __restrict__ bits, int words_per_row,
                                              int row, int c0, int n) {
  uint32_t w = __ldg(bits + (size_t)row * words_per_row + (c0 >> 5));
  w >>= (c0 & 31);
  return n >= 32 ? w : (w & ((1u << n) - 1u));
}

}  // namespace mcgm
