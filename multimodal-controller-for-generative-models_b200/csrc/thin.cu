// thin.cu — convolutions whose one side has <= 4 channels: the RGB side of
// Conv2d(3,128,3,1,1) / Conv2d(3,128,1,1,0) (src/models/mcgan.py:77,87) and of
// Conv2d(256,3,3,1,1) (src/models/mcgan.py:59), forward, data gradient and
// kernel gradient.  They are HBM-bound (<= 27 MACs per byte of the wide
// tensor), so they run on CUDA cores with 128-bit NHWC accesses; the wide
// tensor is read/written exactly once per call.
//
//   thin : fp32 NHWC [N,H,W,Ct]   (Ct <= 4)      wide : bf16 NHWC [N,H,W,Cw]
//   kern : fp32 [T = R*S][Ct][Cw],  off(t) = (r - pad, s - pad)
#include "common.cuh"

namespace mcgm {

static constexpr int kMaxCt = 4;

// wide[p,c] = sum_{t,j} thin[p+off(t), j] * kern[t][j][c]  (+bias) -> ReLU -> mask -> (aux>0)
__global__ void __launch_bounds__(256)
wide_from_thin_kernel(const float* __restrict__ thin, const float* __restrict__ kern,
                      __nv_bfloat16* __restrict__ wide, int N, int H, int W, int Ct, int Cw, int R, int S,
                      const float* __restrict__ bias, const int32_t* __restrict__ labels,
                      const uint32_t* __restrict__ bits, int words, const __nv_bfloat16* __restrict__ aux,
                      int flags) {
  extern __shared__ float sk[];  // [T*Ct][Cw]
  const int T = R * S, pad = (R - 1) / 2;
  for (int i = threadIdx.x; i < T * Ct * Cw; i += blockDim.x) sk[i] = kern[i];
  __syncthreads();
  const int CG = Cw / 8;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  const long P = (long)N * H * W;
  float b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (flags & MCGM_EPI_BIAS) ? bias[cg * 8 + i] : 0.0f;
  for (long p = (long)blockIdx.x * PL + pl; p < P; p += (long)gridDim.x * PL) {
    const int w = (int)(p % W), h = (int)((p / W) % H), n = (int)(p / ((long)W * H));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = b[i];
    for (int r = 0; r < R; ++r) {
      const int hh = h + r - pad;
      if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < S; ++s) {
        const int ww = w + s - pad;
        if (ww < 0 || ww >= W) continue;
        const float* tp = thin + (((long)n * H + hh) * W + ww) * Ct;
        const float* kp = sk + (size_t)((r * S + s) * Ct) * Cw + cg * 8;
        for (int j = 0; j < Ct; ++j) {
          const float tv = __ldg(tp + j);
          const float4 k0 = *reinterpret_cast<const float4*>(kp + (size_t)j * Cw);
          const float4 k1 = *reinterpret_cast<const float4*>(kp + (size_t)j * Cw + 4);
          acc[0] = fmaf(tv, k0.x, acc[0]); acc[1] = fmaf(tv, k0.y, acc[1]);
          acc[2] = fmaf(tv, k0.z, acc[2]); acc[3] = fmaf(tv, k0.w, acc[3]);
          acc[4] = fmaf(tv, k1.x, acc[4]); acc[5] = fmaf(tv, k1.y, acc[5]);
          acc[6] = fmaf(tv, k1.z, acc[6]); acc[7] = fmaf(tv, k1.w, acc[7]);
        }
      }
    }
    Vec<__nv_bfloat16> o;
    uint32_t m = 0xffu;
    if (flags & MCGM_EPI_MASK) m = mask_bits(bits, words, __ldg(labels + n), cg * 8, 8);
    Vec<__nv_bfloat16> a;
    if (flags & MCGM_EPI_POSMASK) a.load(aux + p * Cw + cg * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[i];
      if (flags & MCGM_EPI_RELU) v = fmaxf(v, 0.0f);
      if (!((m >> i) & 1u)) v = 0.0f;
      if ((flags & MCGM_EPI_POSMASK) && !(a.v[i] > 0.0f)) v = 0.0f;
      o.v[i] = v;
    }
    o.store(wide + p * Cw + cg * 8);
  }
}

// thin[p,j] = sum_{t,c} wide[p+off(t), c] * kern[t][j][c] (+bias) -> tanh
// LP = min(32, Cw/8) lanes cooperate on one pixel; each lane walks channel groups lane, lane+LP, ...
__global__ void __launch_bounds__(256)
thin_from_wide_kernel(const __nv_bfloat16* __restrict__ wide, const float* __restrict__ kern,
                      float* __restrict__ thin, int N, int H, int W, int Ct, int Cw, int R, int S,
                      const float* __restrict__ bias, int flags) {
  extern __shared__ float sk[];  // [T*Ct][Cw]
  const int T = R * S, pad = (R - 1) / 2;
  for (int i = threadIdx.x; i < T * Ct * Cw; i += blockDim.x) sk[i] = kern[i];
  __syncthreads();
  const int CG = Cw / 8;
  const int LP = CG < 32 ? CG : 32;            // lanes per pixel (power of two)
  const int lane = threadIdx.x & 31;
  const int sub = lane % LP;                    // lane within the pixel group
  const int ppw = 32 / LP;                      // pixels per warp
  const long P = (long)N * H * W;
  const long warp_global = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long p0 = warp_global * ppw; p0 < P; p0 += nwarps * ppw) {
    const long p = p0 + lane / LP;
    const bool active = p < P;
    const long pc = active ? p : 0;
    const int w = (int)(pc % W), h = (int)((pc / W) % H), n = (int)(pc / ((long)W * H));
    float acc[kMaxCt] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      for (int r = 0; r < R; ++r) {
        const int hh = h + r - pad;
        if (hh < 0 || hh >= H) continue;
        for (int s = 0; s < S; ++s) {
          const int ww = w + s - pad;
          if (ww < 0 || ww >= W) continue;
          const __nv_bfloat16* wp = wide + (((long)n * H + hh) * W + ww) * Cw;
          for (int cg = sub; cg < CG; cg += LP) {
            Vec<__nv_bfloat16> v;
            v.load(wp + cg * 8);
            const float* kp = sk + (size_t)((r * S + s) * Ct) * Cw + cg * 8;
#pragma unroll
            for (int j = 0; j < kMaxCt; ++j) {
              if (j < Ct) {
                const float4 k0 = *reinterpret_cast<const float4*>(kp + (size_t)j * Cw);
                const float4 k1 = *reinterpret_cast<const float4*>(kp + (size_t)j * Cw + 4);
                float a = acc[j];
                a = fmaf(v.v[0], k0.x, a); a = fmaf(v.v[1], k0.y, a); a = fmaf(v.v[2], k0.z, a); a = fmaf(v.v[3], k0.w, a);
                a = fmaf(v.v[4], k1.x, a); a = fmaf(v.v[5], k1.y, a); a = fmaf(v.v[6], k1.z, a); a = fmaf(v.v[7], k1.w, a);
                acc[j] = a;
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxCt; ++j)
      for (int o = LP >> 1; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    if (active && sub == 0) {
      for (int j = 0; j < Ct; ++j) {
        float v = acc[j] + ((flags & MCGM_EPI_BIAS) ? __ldg(bias + j) : 0.0f);
        if (flags & MCGM_EPI_TANH) v = tanhf(v);
        thin[p * Ct + j] = v;
      }
    }
  }
}

// dkern[t][j][c] = sum_p thin[p+off(t), j] * wide[p, c]
// thread = (channel group cg, tap row r, pixel lane pl); per-block partials, fixed-order finish.
template <int S_>
__global__ void __launch_bounds__(256)
kern_grad_partial_kernel(const float* __restrict__ thin, const __nv_bfloat16* __restrict__ wide,
                         int N, int H, int W, int Ct, int Cw, int R, int PL, float* __restrict__ partial) {
  const int pad = (R - 1) / 2;
  const int CG = Cw / 8;
  const int cg = threadIdx.x % CG;
  const int r = (threadIdx.x / CG) % R;
  const int pl = threadIdx.x / (CG * R);
  const bool worker = pl < PL;
  const long P = (long)N * H * W;
  float acc[S_][kMaxCt][8];
#pragma unroll
  for (int s = 0; s < S_; ++s)
#pragma unroll
    for (int j = 0; j < kMaxCt; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[s][j][i] = 0.0f;
  if (worker) {
    for (long p = (long)blockIdx.x * PL + pl; p < P; p += (long)gridDim.x * PL) {
      const int w = (int)(p % W), h = (int)((p / W) % H), n = (int)(p / ((long)W * H));
      const int hh = h + r - pad;
      if (hh < 0 || hh >= H) continue;
      Vec<__nv_bfloat16> v;
      v.load(wide + p * Cw + cg * 8);
#pragma unroll
      for (int s = 0; s < S_; ++s) {
        const int ww = w + s - pad;
        if (ww < 0 || ww >= W) continue;
        const float* tp = thin + (((long)n * H + hh) * W + ww) * Ct;
#pragma unroll
        for (int j = 0; j < kMaxCt; ++j) {
          if (j < Ct) {
            const float tv = __ldg(tp + j);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[s][j][i] = fmaf(tv, v.v[i], acc[s][j][i]);
          }
        }
      }
    }
  }
  // reduce over pixel lanes through shared memory, one (s,j) slab at a time
  extern __shared__ float sm[];  // [PL][R][Cw]
  float* out = partial + (size_t)blockIdx.x * (R * S_ * Ct) * Cw;
  for (int s = 0; s < S_; ++s) {
    for (int j = 0; j < Ct; ++j) {
      __syncthreads();
      if (worker) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[((size_t)pl * R + r) * Cw + cg * 8 + i] = acc[s][j][i];
      }
      __syncthreads();
      for (int e = threadIdx.x; e < R * Cw; e += blockDim.x) {
        const int rr = e / Cw, c = e % Cw;
        float a = 0.f;
        for (int l = 0; l < PL; ++l) a += sm[((size_t)l * R + rr) * Cw + c];
        out[(size_t)((rr * S_ + s) * Ct + j) * Cw + c] = a;
      }
    }
  }
}

__global__ void kern_grad_finalize_kernel(const float* __restrict__ partial, int nblocks, int total,
                                          float* __restrict__ dkern) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double a = 0.0;
  for (int b = 0; b < nblocks; ++b) a += (double)partial[(size_t)b * total + i];
  dkern[i] = (float)a;
}

// ---------------------------------------------------------------------------
// im2col / col2im of the THIN tensor.  With them every thin convolution becomes a
// 1x1 GEMM on the tensor cores that touches the wide tensor exactly once:
//   U[p, t*Ct + j] = thin[p + sgn*off(t), j]          (bf16, KU = 64 columns, zero padded)
//   wide_from_thin : wide = U x Wmat^T            (conv_fwd 1x1, Cin = 64)
//   kernel grads   : dWmat = dY^T x U             (conv_wgrad 1x1)
//   thin_from_wide : T = wide x Wmat^T (fp32 [P, CT]);  thin[p, j] = sum_t T[p + off(t), t*Ct + j]
// ---------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(256)
thin_im2col_kernel(const float* __restrict__ thin, OT* __restrict__ U, int N, int H, int W, int Ho, int Wo,
                   int Ct, int R, int S, int stride, int pad, int sgn, int KU, PixDiv dw, PixDiv dhw) {
  // column -> (row offset, column offset, channel) tables: no division in the hot loop
  __shared__ signed char tdr[128], tds[128], tj[128];
  const int T = R * S;
  for (int col = threadIdx.x; col < KU; col += blockDim.x) {
    if (col < T * Ct) {
      const int t = col / Ct, r = t / S;
      tdr[col] = (signed char)(sgn * (r - pad));
      tds[col] = (signed char)(sgn * (t - r * S - pad));
      tj[col] = (signed char)(col - t * Ct);
    } else {
      tj[col] = -1;
    }
  }
  __syncthreads();
  constexpr int V = Vec<OT>::N;
  const uint32_t groups = KU / V, used = (T * Ct + V - 1) / V;
  const uint32_t HWo = (uint32_t)Ho * Wo;
  const uint32_t total = (uint32_t)N * HWo * groups;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const uint32_t p = e / groups, gidx = e - p * groups;
    Vec<OT> o;
#pragma unroll
    for (int i = 0; i < V; ++i) o.v[i] = 0.0f;
    if (gidx < used) {
      const uint32_t n = dhw.div(p), hw = p - n * HWo, h = dw.div(hw), w = hw - h * Wo;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int col = gidx * V + i, j = tj[col];
        const int hh = (int)h * stride + tdr[col], ww = (int)w * stride + tds[col];
        if (j >= 0 && hh >= 0 && hh < H && ww >= 0 && ww < W)
          o.v[i] = __ldg(thin + ((size_t)(n * H + hh) * W + ww) * Ct + j);
      }
    }
    o.store(U + (size_t)p * KU + gidx * V);
  }
}

__global__ void thin_col2im_kernel(const float* __restrict__ Tm, int CT, float* __restrict__ thin, int N, int H,
                                   int W, int Ct, int R, int S, const float* __restrict__ bias, int flags) {
  const int pad = (R - 1) / 2;
  const long total = (long)N * H * W;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long)gridDim.x * blockDim.x) {
    const int w = (int)(p % W), h = (int)((p / W) % H), n = (int)(p / ((long)W * H));
    float acc[kMaxCt];
#pragma unroll
    for (int j = 0; j < kMaxCt; ++j) acc[j] = (j < Ct && (flags & MCGM_EPI_BIAS)) ? __ldg(bias + j) : 0.0f;
    for (int r = 0; r < R; ++r) {
      const int hh = h + r - pad;
      if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < S; ++s) {
        const int ww = w + s - pad;
        if (ww < 0 || ww >= W) continue;
        const float* tp = Tm + (((long)n * H + hh) * W + ww) * CT + (r * S + s) * Ct;
#pragma unroll
        for (int j = 0; j < kMaxCt; ++j)
          if (j < Ct) acc[j] += __ldg(tp + j);
      }
    }
    for (int j = 0; j < Ct; ++j) thin[p * Ct + j] = (flags & MCGM_EPI_TANH) ? tanhf(acc[j]) : acc[j];
  }
}

static constexpr int kKernGradBlocks = 148 * 2;

}  // namespace mcgm

using namespace mcgm;

static int thin_check(int Ct, int Cw, int R, int S, const char* who) {
  if (Ct < 1 || Ct > kMaxCt) return fail(MCGM_E_UNSUPPORTED, "%s: Ct=%d must be in [1,4]", who, Ct);
  if (Cw % 8 || Cw < 8) return fail(MCGM_E_UNSUPPORTED, "%s: Cw=%d must be a multiple of 8", who, Cw);
  const int CG = Cw / 8;
  if ((CG & (CG - 1)) || CG > 256) return fail(MCGM_E_UNSUPPORTED, "%s: Cw/8=%d must be a power of two <= 256", who, CG);
  if (!((R == 1 && S == 1) || (R == 3 && S == 3))) return fail(MCGM_E_UNSUPPORTED, "%s: only 1x1 and 3x3 kernels", who);
  return 0;
}

extern "C" int mcgm_thin_wide_from_thin(const float* thin, const float* kern, void* wide, int N, int H, int W,
                                        int Ct, int Cw, int R, int S, const float* bias, const int32_t* labels,
                                        const uint32_t* bits, int M, const void* aux, int flags, void* stream) {
  MCGM_CHECK_ARG(thin && kern && wide && N > 0 && H > 0 && W > 0, "wide_from_thin: bad arguments");
  if (int e = thin_check(Ct, Cw, R, S, "wide_from_thin")) return e;
  if (flags & MCGM_EPI_BIAS) MCGM_CHECK_ARG(bias, "wide_from_thin: BIAS without bias");
  if (flags & MCGM_EPI_MASK) MCGM_CHECK_ARG(labels && bits && M > 0, "wide_from_thin: MASK needs labels/bits");
  if (flags & MCGM_EPI_POSMASK) MCGM_CHECK_ARG(aux, "wide_from_thin: POSMASK needs aux");
  const size_t smem = (size_t)R * S * Ct * Cw * sizeof(float);
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(wide_from_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  if (smem > 160 * 1024) return fail(MCGM_E_UNSUPPORTED, "wide_from_thin: kernel table too large");
  const int PL = 256 / (Cw / 8);
  const long P = (long)N * H * W;
  long blocks = (P + PL - 1) / PL;
  if (blocks > 148 * 6) blocks = 148 * 6;
  wide_from_thin_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>(
      thin, kern, (__nv_bfloat16*)wide, N, H, W, Ct, Cw, R, S, bias, labels, bits, (Cw + 31) / 32,
      (const __nv_bfloat16*)aux, flags);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_thin_thin_from_wide(const void* wide, const float* kern, float* thin, int N, int H, int W,
                                        int Ct, int Cw, int R, int S, const float* bias, int flags, void* stream) {
  MCGM_CHECK_ARG(wide && kern && thin && N > 0 && H > 0 && W > 0, "thin_from_wide: bad arguments");
  if (int e = thin_check(Ct, Cw, R, S, "thin_from_wide")) return e;
  if (flags & MCGM_EPI_BIAS) MCGM_CHECK_ARG(bias, "thin_from_wide: BIAS without bias");
  const size_t smem = (size_t)R * S * Ct * Cw * sizeof(float);
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(thin_from_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  if (smem > 160 * 1024) return fail(MCGM_E_UNSUPPORTED, "thin_from_wide: kernel table too large");
  const int CG = Cw / 8, LP = CG < 32 ? CG : 32, ppw = 32 / LP;
  const long P = (long)N * H * W;
  long blocks = (P + (long)8 * ppw - 1) / ((long)8 * ppw);
  if (blocks > 148 * 6) blocks = 148 * 6;
  thin_from_wide_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)wide, kern, thin, N, H,
                                                                         W, Ct, Cw, R, S, bias, flags);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mcgm_thin_kern_grad_workspace_bytes(int Ct, int Cw, int R, int S) {
  return (size_t)kKernGradBlocks * R * S * Ct * Cw * sizeof(float);
}

extern "C" int mcgm_thin_kern_grad(const float* thin, const void* wide, float* dkern, int N, int H, int W, int Ct,
                                   int Cw, int R, int S, void* workspace, void* stream) {
  MCGM_CHECK_ARG(thin && wide && dkern && workspace && N > 0 && H > 0 && W > 0, "kern_grad: bad arguments");
  if (int e = thin_check(Ct, Cw, R, S, "kern_grad")) return e;
  const int CG = Cw / 8;
  if (CG * R > 256) return fail(MCGM_E_UNSUPPORTED, "kern_grad: Cw=%d too wide for R=%d", Cw, R);
  const int PL = 256 / (CG * R);
  const long P = (long)N * H * W;
  long blocks = (P + PL - 1) / PL;
  if (blocks > kKernGradBlocks) blocks = kKernGradBlocks;
  const size_t smem = (size_t)PL * R * Cw * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  if (S == 3)
    kern_grad_partial_kernel<3><<<(int)blocks, 256, smem, st>>>(thin, (const __nv_bfloat16*)wide, N, H, W, Ct, Cw, R, PL, partial);
  else
    kern_grad_partial_kernel<1><<<(int)blocks, 256, smem, st>>>(thin, (const __nv_bfloat16*)wide, N, H, W, Ct, Cw, R, PL, partial);
  MCGM_LAUNCH_CHECK();
  const int total = R * S * Ct * Cw;
  kern_grad_finalize_kernel<<<(total + 255) / 256, 256, 0, st>>>(partial, (int)blocks, total, dkern);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_thin_im2col(const float* thin, void* U_bf16, int out_dtype, int N, int H, int W, int Ct, int R,
                                int S, int stride, int pad, int flip, int KU, void* stream) {
  MCGM_CHECK_ARG(thin && U_bf16 && N > 0 && H > 0 && W > 0, "thin_im2col: bad arguments");
  MCGM_CHECK_ARG(out_dtype == MCGM_BF16 || out_dtype == MCGM_F32, "thin_im2col: out_dtype %d unsupported", out_dtype);
  MCGM_CHECK_ARG(Ct >= 1 && Ct <= kMaxCt && KU % 8 == 0 && KU <= 128 && R * S * Ct <= KU, "thin_im2col: Ct=%d R=%d S=%d do not fit KU=%d", Ct, R, S, KU);
  MCGM_CHECK_ARG(stride >= 1 && stride <= 2 && pad >= 0 && (!flip || stride == 1), "thin_im2col: stride %d / pad %d / flip unsupported", stride, pad);
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  MCGM_CHECK_ARG(Ho > 0 && Wo > 0, "thin_im2col: empty output");
  const long total = (long)N * Ho * Wo * (KU / 8);
  MCGM_CHECK_ARG(total < (1L << 32), "thin_im2col: tensor too large");
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (out_dtype == MCGM_BF16)
    thin_im2col_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        thin, (__nv_bfloat16*)U_bf16, N, H, W, Ho, Wo, Ct, R, S, stride, pad, flip ? -1 : 1, KU,
        PixDiv::make((uint32_t)Wo), PixDiv::make((uint32_t)(Ho * Wo)));
  else
    thin_im2col_kernel<float><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        thin, (float*)U_bf16, N, H, W, Ho, Wo, Ct, R, S, stride, pad, flip ? -1 : 1, KU,
        PixDiv::make((uint32_t)Wo), PixDiv::make((uint32_t)(Ho * Wo)));
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_thin_col2im(const float* T, int CT, float* thin, int N, int H, int W, int Ct, int R, int S,
                                const float* bias, int flags, void* stream) {
  MCGM_CHECK_ARG(T && thin && N > 0 && H > 0 && W > 0, "thin_col2im: bad arguments");
  MCGM_CHECK_ARG(Ct >= 1 && Ct <= kMaxCt && R * S * Ct <= CT, "thin_col2im: Ct=%d R=%d S=%d do not fit CT=%d", Ct, R, S, CT);
  if (flags & MCGM_EPI_BIAS) MCGM_CHECK_ARG(bias, "thin_col2im: BIAS without bias");
  const long total = (long)N * H * W;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  thin_col2im_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(T, CT, thin, N, H, W, Ct, R, S, bias, flags);
  MCGM_LAUNCH_CHECK();
  return 0;
}
