// split.cu — operand splitting for the fp32 PARITY pipeline.
//
// The tensor cores have no fp32 multiply; the parity pipeline (cfg['compute_dtype'] = 'fp32')
// recovers fp32-accurate convolutions from exact low-precision pieces:
//   forward / dgrad : x = xb + xs with xb, xs TF32-representable;   x*w ~ xb*wb + xb*ws + xs*wb
//                     (three kind::tf32 launches chained through the fp32 residual input, "3xTF32")
//   wgrad           : x = x1 + x2 + x3 with bf16 pieces;  the six products of total order <= 4
//                     on the bf16 MN-major wgrad kernel.
// Every piece is exactly representable in its format, so the tensor-core conversion is exact and the
// only error left is the dropped cross terms (~2^-21 relative).
#include "common.cuh"

namespace mcgm {

__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ big, float* __restrict__ small,
                                  long n4) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float in[4] = {v.x, v.y, v.z, v.w};
    float b[4], s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      b[k] = __uint_as_float(__float_as_uint(in[k]) & 0xffffe000u);          // 10-bit mantissa, truncated
      s[k] = __uint_as_float(__float_as_uint(in[k] - b[k]) & 0xffffe000u);   // exact difference, truncated again
    }
    reinterpret_cast<float4*>(big)[i] = make_float4(b[0], b[1], b[2], b[3]);
    reinterpret_cast<float4*>(small)[i] = make_float4(s[0], s[1], s[2], s[3]);
  }
}

__global__ void split_bf16x3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ p1,
                                    __nv_bfloat16* __restrict__ p2, __nv_bfloat16* __restrict__ p3, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 a = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(a);
    const __nv_bfloat16 b = __float2bfloat16(r1);
    const float r2 = r1 - __bfloat162float(b);
    p1[i] = a; p2[i] = b; p3[i] = __float2bfloat16(r2);
  }
}

}  // namespace mcgm

using namespace mcgm;

extern "C" int mcgm_split_tf32(const float* x, float* big, float* small, long n, void* stream) {
  MCGM_CHECK_ARG(x && big && small && n > 0 && n % 4 == 0, "split_tf32: need n > 0 and n %% 4 == 0");
  long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_tf32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, big, small, n / 4);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_split_bf16x3(const float* x, void* p1, void* p2, void* p3, long n, void* stream) {
  MCGM_CHECK_ARG(x && p1 && p2 && p3 && n > 0, "split_bf16x3: bad arguments");
  long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_bf16x3_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)p1, (__nv_bfloat16*)p2,
                                                                    (__nv_bfloat16*)p3, n);
  MCGM_LAUNCH_CHECK();
  return 0;
}
