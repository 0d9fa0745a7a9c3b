// tmap.cuh — TMA descriptor construction shared by the tcgen05 kernels (defined in conv_umma.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace mcgm {

// Tensor map of a dense tensor: dims innermost-first, dense strides, zero OOB fill; cached by (pointer, geometry).
//   estride : TMA traversal stride of dims 1 and 2 of a 4-D NHWC map (1, or 2 for stride-2 gathers / phase stores)
//   f32     : element type fp32 (else bf16)
//   swizzle : 2 = 128-byte swizzle (UMMA operand tiles, staged output tiles), 1 = 64-byte swizzle (32-element bf16
//             k-blocks), 0 = none (plain row-major boxes)
int make_tmap(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box, int estride = 1,
              int f32 = 0, int swizzle = 2);

}  // namespace mcgm
