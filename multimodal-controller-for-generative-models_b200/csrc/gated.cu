// gated.cu — the MC-gated activation of the gated PixelCNN:
//
//   MCGatedActivation.forward    src/models/mcpixelcnn.py:9-20
//     x, y = h.chunk(2, dim=1);  out = mc(relu(bn(x)) * sigmoid(y))
//
// h is the 2C-channel NHWC output of a conv ([N,H,W,2C] bf16): x = channels [0,C), y = channels [C,2C) of the
// SAME pixel row, so one kernel reads each row once and writes C channels.  Three HBM-bound kernels replace the
// chunk / BatchNorm2d / ReLU / sigmoid / mul / MultimodalController chain (7 eager launches forward, ~12 backward):
//   stats  : train-mode batch statistics of the x half (row stride 2C) — bn_partial with a leading dimension,
//            then the shared finalize of gate.cu's layout (mean / rstd / running stats, unbiased var, momentum)
//   fwd    : out = mask * relu(a*x + b) * sigmoid(y)                      reads 2C, writes C per pixel
//   bwd    : g  = mask * dout * sigmoid(y) * 1[z > 0]   (gradient at the BatchNorm output z = a*x + b)
//            dy = mask * dout * relu(z) * s * (1 - s)
//            dbeta = sum g, dgamma = sum g*xhat, dx = a*(g - dbeta/P - xhat*dgamma/P)   (eval: dx = a*g)
//            as a reduce pass (reads dout, h) and an apply pass (reads dout, h; writes dh = [dx | dy]).
// Algorithmic bytes per pixel and channel pair (bf16): stats 2, fwd 4 + 2, bwd 2*(2 + 4) + 4.
// Thread = one 16-byte channel group of the x half (and the matching group of the y half), walking pixels.
#include "common.cuh"

namespace mcgm {

static constexpr int kSgThreads = 256;
static constexpr int kSgMaxBlocks = 148 * 4;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }

// per-block partial sums of channels [0, C) of a tensor with row stride ld (elements): [grid][2][C]
__global__ void __launch_bounds__(kSgThreads)
sg_stats_partial_kernel(const __nv_bfloat16* __restrict__ h, long P, int C, int ld, float* __restrict__ partial) {
  constexpr int V = 8;
  const int CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = q[i] = 0.0f;
  if (pl < PL) {
    for (long p = (long)blockIdx.x * PL + pl; p < P; p += (long)gridDim.x * PL) {
      float v[V];
      unpack16<__nv_bfloat16>(ldg16(h + p * ld + cg * V), v);
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] += v[i]; q[i] = fmaf(v[i], v[i], q[i]); }
    }
  }
  extern __shared__ float sm[];          // [2][PL][C]
  float* ss = sm;
  float* sq = sm + (size_t)PL * C;
  if (pl < PL) {
#pragma unroll
    for (int i = 0; i < V; ++i) { ss[pl * C + cg * V + i] = s[i]; sq[pl * C + cg * V + i] = q[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < PL; ++l) { a += ss[l * C + c]; b += sq[l * C + c]; }
    partial[((size_t)blockIdx.x * 2 + 0) * C + c] = a;
    partial[((size_t)blockIdx.x * 2 + 1) * C + c] = b;
  }
}

// fixed-order fp64 finalize of [nblocks][2][C] partials -> (S, Q) per channel; one thread per channel
__global__ void sg_stats_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, double count,
                                         float* __restrict__ mean, float* __restrict__ rstd,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         float momentum, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double S = 0.0, Q = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    S += (double)partial[((size_t)b * 2 + 0) * C + c];
    Q += (double)partial[((size_t)b * 2 + 1) * C + c];
  }
  const double m = S / count;
  double var = Q / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
  }
}

__global__ void __launch_bounds__(kSgThreads)
sg_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out, long P, int HW, int C,
              const int32_t* __restrict__ labels, const uint32_t* __restrict__ bits, int words,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              const float* __restrict__ beta) {
  constexpr int V = 8;
  const int CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  if (pl >= PL) return;
  float a[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    a[i] = gamma[c] * rstd[c];
    b[i] = fmaf(-mean[c], a[i], beta[c]);
  }
  for (long p = (long)blockIdx.x * PL + pl; p < P; p += (long)gridDim.x * PL) {
    const uint4 xr = ldg16(h + p * 2 * C + cg * V), yr = ldg16(h + p * 2 * C + C + cg * V);
    const uint32_t m = mask_bits(bits, words, __ldg(labels + (int)(p / HW)), cg * V, V);
    float xv[V], yv[V];
    unpack16<__nv_bfloat16>(xr, xv);
    unpack16<__nv_bfloat16>(yr, yv);
    Vec<__nv_bfloat16> o;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float z = fmaxf(fmaf(a[i], xv[i], b[i]), 0.0f);
      o.v[i] = ((m >> i) & 1u) ? z * sigmoidf_(yv[i]) : 0.0f;
    }
    o.store(out + p * C + cg * V);
  }
}

// pass 1 of the backward: per-block partials of sum g and sum g*xhat
__global__ void __launch_bounds__(kSgThreads)
sg_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ h, long P, int HW, int C,
                     const int32_t* __restrict__ labels, const uint32_t* __restrict__ bits, int words,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ partial) {
  constexpr int V = 8;
  const int CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  float a[V], b[V], mu[V], rs[V], s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    mu[i] = mean[c]; rs[i] = rstd[c];
    a[i] = gamma[c] * rs[i];
    b[i] = fmaf(-mu[i], a[i], beta[c]);
    s1[i] = s2[i] = 0.0f;
  }
  if (pl < PL) {
    for (long p = (long)blockIdx.x * PL + pl; p < P; p += (long)gridDim.x * PL) {
      const uint4 xr = ldg16(h + p * 2 * C + cg * V), yr = ldg16(h + p * 2 * C + C + cg * V);
      const uint4 gr = ldg16(dout + p * C + cg * V);
      const uint32_t m = mask_bits(bits, words, __ldg(labels + (int)(p / HW)), cg * V, V);
      float xv[V], yv[V], gv[V];
      unpack16<__nv_bfloat16>(xr, xv);
      unpack16<__nv_bfloat16>(yr, yv);
      unpack16<__nv_bfloat16>(gr, gv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float z = fmaf(a[i], xv[i], b[i]);
        const float gi = (((m >> i) & 1u) && z > 0.0f) ? gv[i] * sigmoidf_(yv[i]) : 0.0f;
        s1[i] += gi;
        s2[i] = fmaf(gi, (xv[i] - mu[i]) * rs[i], s2[i]);
      }
    }
  }
  extern __shared__ float sm[];
  float* ss = sm;
  float* sq = sm + (size_t)PL * C;
  if (pl < PL) {
#pragma unroll
    for (int i = 0; i < V; ++i) { ss[pl * C + cg * V + i] = s1[i]; sq[pl * C + cg * V + i] = s2[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float u = 0.f, w = 0.f;
    for (int l = 0; l < PL; ++l) { u += ss[l * C + c]; w += sq[l * C + c]; }
    partial[((size_t)blockIdx.x * 2 + 0) * C + c] = u;
    partial[((size_t)blockIdx.x * 2 + 1) * C + c] = w;
  }
}

__global__ void sg_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double S = 0.0, Q = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    S += (double)partial[((size_t)b * 2 + 0) * C + c];
    Q += (double)partial[((size_t)b * 2 + 1) * C + c];
  }
  dbeta[c] = (float)S;
  dgamma[c] = (float)Q;
}

// pass 2: dh = [dx | dy]
__global__ void __launch_bounds__(kSgThreads)
sg_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ h,
                    __nv_bfloat16* __restrict__ dh, long P, int HW, int C,
                    const int32_t* __restrict__ labels, const uint32_t* __restrict__ bits, int words,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                    int batch_stats) {
  constexpr int V = 8;
  const int CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  if (pl >= PL) return;
  const float invP = 1.0f / (float)P;
  float a[V], b[V], c0[V], c1[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    const float mu = mean[c], rs = rstd[c];
    a[i] = gamma[c] * rs;
    b[i] = fmaf(-mu, a[i], beta[c]);
    if (batch_stats) {                       // dx = a*(g - dbeta/P - xhat*dgamma/P) = a*g + c0 + c1*x
      c1[i] = -a[i] * rs * (dgamma[c] * invP);
      c0[i] = -a[i] * (dbeta[c] * invP) - c1[i] * mu;
    } else {
      c0[i] = c1[i] = 0.0f;
    }
  }
  for (long p = (long)blockIdx.x * PL + pl; p < P; p += (long)gridDim.x * PL) {
    const uint4 xr = ldg16(h + p * 2 * C + cg * V), yr = ldg16(h + p * 2 * C + C + cg * V);
    const uint4 gr = ldg16(dout + p * C + cg * V);
    const uint32_t m = mask_bits(bits, words, __ldg(labels + (int)(p / HW)), cg * V, V);
    float xv[V], yv[V], gv[V];
    unpack16<__nv_bfloat16>(xr, xv);
    unpack16<__nv_bfloat16>(yr, yv);
    unpack16<__nv_bfloat16>(gr, gv);
    Vec<__nv_bfloat16> ox, oy;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float z = fmaf(a[i], xv[i], b[i]);
      const float s = sigmoidf_(yv[i]);
      const float gm = ((m >> i) & 1u) ? gv[i] : 0.0f;
      const float gi = z > 0.0f ? gm * s : 0.0f;
      ox.v[i] = fmaf(a[i], gi, fmaf(c1[i], xv[i], c0[i]));
      oy.v[i] = gm * fmaxf(z, 0.0f) * s * (1.0f - s);
    }
    ox.store(dh + p * 2 * C + cg * V);
    oy.store(dh + p * 2 * C + C + cg * V);
  }
}

static int sg_check(int C, int* cg_out) {
  if (C <= 0 || C % 8) return fail(MCGM_E_ARG, "sgate: C=%d must be a positive multiple of 8", C);
  const int CG = C / 8;
  if (CG > kSgThreads) return fail(MCGM_E_UNSUPPORTED, "sgate: C=%d too wide (<= %d)", C, 8 * kSgThreads);
  *cg_out = CG;
  return 0;
}

static int sg_grid(long P, int PL, int cap) {
  long need = (P + PL - 1) / PL;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace mcgm

using namespace mcgm;

extern "C" size_t mcgm_sgate_workspace_bytes(int C) {
  return (size_t)kSgMaxBlocks * 2 * (size_t)(C > 0 ? C : 0) * sizeof(float);
}

// Train-mode statistics of channels [0, C) of h [P, 2C] bf16 -> mean, rstd (+ running stats update).
extern "C" int mcgm_sgate_stats(const void* h, long P, int C, float* mean, float* rstd, float* running_mean,
                                float* running_var, float momentum, float eps, void* workspace, void* stream) {
  MCGM_CHECK_ARG(h && mean && rstd && workspace && P > 0, "sgate_stats: bad arguments");
  MCGM_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "sgate_stats: running stats must come in pairs");
  int CG;
  if (int e = sg_check(C, &CG)) return e;
  const int PL = kSgThreads / CG;
  const int grid = sg_grid(P, PL * 4, kSgMaxBlocks);
  cudaStream_t st = (cudaStream_t)stream;
  sg_stats_partial_kernel<<<grid, kSgThreads, (size_t)2 * PL * C * sizeof(float), st>>>((const __nv_bfloat16*)h, P, C, 2 * C,
                                                                                     (float*)workspace);
  MCGM_LAUNCH_CHECK();
  sg_stats_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>((const float*)workspace, grid, C, (double)P, mean, rstd,
                                                         running_mean, running_var, momentum, eps);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_sgate_fwd(const void* h, void* out, int N, int HW, int C, const int32_t* labels, const uint32_t* bits,
                              int M, const float* mean, const float* rstd, const float* gamma, const float* beta,
                              void* stream) {
  MCGM_CHECK_ARG(h && out && labels && bits && M > 0 && mean && rstd && gamma && beta && N >= 0 && HW > 0, "sgate_fwd: bad arguments");
  if (N == 0) return 0;
  int CG;
  if (int e = sg_check(C, &CG)) return e;
  const int PL = kSgThreads / CG;
  const long P = (long)N * HW;
  sg_fwd_kernel<<<sg_grid(P, PL * 4, 148 * 8), kSgThreads, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)h, (__nv_bfloat16*)out, P, HW, C, labels, bits, (C + 31) / 32, mean, rstd, gamma, beta);
  MCGM_LAUNCH_CHECK();
  return 0;
}

// batch_stats = 1: (mean, rstd) are this batch's statistics (training) — dgamma / dbeta are produced and the
// BatchNorm backward couples the pixels; 0: running statistics (eval) — dx = a * g.
extern "C" int mcgm_sgate_bwd(const void* dout, const void* h, void* dh, int N, int HW, int C, const int32_t* labels,
                              const uint32_t* bits, int M, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, float* dgamma, float* dbeta, int batch_stats, void* workspace,
                              void* stream) {
  MCGM_CHECK_ARG(dout && h && dh && labels && bits && M > 0 && mean && rstd && gamma && beta && dgamma && dbeta && workspace &&
                     N >= 0 && HW > 0, "sgate_bwd: bad arguments");
  if (N == 0) return 0;
  int CG;
  if (int e = sg_check(C, &CG)) return e;
  const int PL = kSgThreads / CG, words = (C + 31) / 32;
  const long P = (long)N * HW;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = sg_grid(P, PL * 4, kSgMaxBlocks);
  sg_bwd_reduce_kernel<<<grid, kSgThreads, (size_t)2 * PL * C * sizeof(float), st>>>(
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)h, P, HW, C, labels, bits, words, mean, rstd, gamma, beta, (float*)workspace);
  MCGM_LAUNCH_CHECK();
  sg_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>((const float*)workspace, grid, C, dgamma, dbeta);
  MCGM_LAUNCH_CHECK();
  sg_bwd_apply_kernel<<<sg_grid(P, PL * 4, 148 * 8), kSgThreads, 0, st>>>(
      (const __nv_bfloat16*)dout, (const __nv_bfloat16*)h, (__nv_bfloat16*)dh, P, HW, C, labels, bits, words, mean, rstd, gamma,
      beta, dgamma, dbeta, batch_stats);
  MCGM_LAUNCH_CHECK();
  return 0;
}
