// sn.cu — spectral normalisation of the discriminator's weights, batched over layers.
//
// Restates the forward pre-hook of torch's legacy spectral_norm that the reference applies
// to every Conv/Linear of D (src/models/utils.py:17-21 -> torch/nn/utils/spectral_norm.py
// compute_weight): per layer and per training forward
//     v <- normalize(W^T u), u <- normalize(W v)      (eps 1e-12, in place on the buffers)
//     sigma = u . (W v),  W_bar = W / sigma,  u and v treated as constants in backward
// In eager PyTorch that is ~12 tiny kernels per layer (x 11 layers x 11 forwards per
// iteration); here ONE launch handles all layers (one CTA each), writes the updated u, v
// in place, and also snapshots [1/sigma | u | v] per layer into a `state` buffer, because
// the two D forwards of one D update use different u, v and each backward needs its own.
//
//   backward:  dW = (dW_bar - <dW_bar, W_bar> u v^T) / sigma
#include "common.cuh"

namespace mcgm {

static constexpr int kSnMaxLayers = 16;
static constexpr int kSnThreads = 1024;

struct SnBatch {
  const float* W[kSnMaxLayers];
  float* u[kSnMaxLayers];
  float* v[kSnMaxLayers];
  int K[kSnMaxLayers];
  int CRS[kSnMaxLayers];
  long off[kSnMaxLayers];
};

__device__ __forceinline__ float block_sum(float x, float* red) {
  // red: >= 33 floats of shared scratch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  x = warp_sum(x);
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__global__ void __launch_bounds__(kSnThreads)
sn_power_iter_kernel(const SnBatch b, float* __restrict__ state, int do_iter, float eps) {
  extern __shared__ float sm[];          // [CRS] v | [K] u | [K] W v | [40] reduction scratch
  const int L = blockIdx.x;
  const float* __restrict__ W = b.W[L];
  const int K = b.K[L], CRS = b.CRS[L];
  float* sv = sm;
  float* su = sm + CRS;
  float* swv = su + K;
  float* red = swv + K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;

  for (int k = tid; k < K; k += blockDim.x) su[k] = b.u[L][k];
  for (int j = tid; j < CRS; j += blockDim.x) sv[j] = b.v[L][j];
  __syncthreads();
  if (do_iter) {
    // v = normalize(W^T u): thread j walks a column (coalesced across threads for every row k)
    float nrm = 0.0f;
    for (int j = tid; j < CRS; j += blockDim.x) {
      float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;      // 4 independent chains x unroll 2: loads in flight
      int k = 0;
#pragma unroll 2
      for (; k + 3 < K; k += 4) {
        a0 = fmaf(__ldg(W + (size_t)k * CRS + j), su[k], a0);
        a1 = fmaf(__ldg(W + (size_t)(k + 1) * CRS + j), su[k + 1], a1);
        a2 = fmaf(__ldg(W + (size_t)(k + 2) * CRS + j), su[k + 2], a2);
        a3 = fmaf(__ldg(W + (size_t)(k + 3) * CRS + j), su[k + 3], a3);
      }
      for (; k < K; ++k) a0 = fmaf(__ldg(W + (size_t)k * CRS + j), su[k], a0);
      const float a = (a0 + a1) + (a2 + a3);
      sv[j] = a;
      nrm = fmaf(a, a, nrm);
    }
    nrm = block_sum(nrm, red);
    const float inv = 1.0f / fmaxf(sqrtf(nrm), eps);
    for (int j = tid; j < CRS; j += blockDim.x) sv[j] *= inv;
    __syncthreads();
  }
  // W v: one warp per row
  for (int k = warp; k < K; k += nw) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    int j = lane;
#pragma unroll 2
    for (; j + 96 < CRS; j += 128) {
      a0 = fmaf(__ldg(W + (size_t)k * CRS + j), sv[j], a0);
      a1 = fmaf(__ldg(W + (size_t)k * CRS + j + 32), sv[j + 32], a1);
      a2 = fmaf(__ldg(W + (size_t)k * CRS + j + 64), sv[j + 64], a2);
      a3 = fmaf(__ldg(W + (size_t)k * CRS + j + 96), sv[j + 96], a3);
    }
    for (; j < CRS; j += 32) a0 = fmaf(__ldg(W + (size_t)k * CRS + j), sv[j], a0);
    const float a = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) swv[k] = a;
  }
  __syncthreads();
  if (do_iter) {
    // u = normalize(W v)
    float nrm = 0.0f;
    for (int k = tid; k < K; k += blockDim.x) nrm = fmaf(swv[k], swv[k], nrm);
    nrm = block_sum(nrm, red);
    const float inv = 1.0f / fmaxf(sqrtf(nrm), eps);
    for (int k = tid; k < K; k += blockDim.x) su[k] = swv[k] * inv;
    __syncthreads();
  }
  float dot = 0.0f;                       // sigma = u . (W v)
  for (int k = tid; k < K; k += blockDim.x) dot = fmaf(su[k], swv[k], dot);
  dot = block_sum(dot, red);
  float* st = state + b.off[L];
  if (tid == 0) st[0] = 1.0f / dot;
  for (int k = tid; k < K; k += blockDim.x) {
    st[1 + k] = su[k];
    if (do_iter) b.u[L][k] = su[k];
  }
  for (int j = tid; j < CRS; j += blockDim.x) {
    st[1 + K + j] = sv[j];
    if (do_iter) b.v[L][j] = sv[j];
  }
}

// dW[k,c,r,s] = (dWbar[k,(r,s,c)] - s * u[k] * v[(c,r,s)]) * inv_sigma,  s = <dWbar, W> * inv_sigma
// dWbar is [K][row_stride] with the first R*S*C entries of a row in (r,s,c) order (KRSC), or in
// (c,r,s) order when krsc == 0.  Two launches of kSnBwdBlocks CTAs: per-block partial dot products,
// then every block re-sums the partials in a fixed order (deterministic) and applies its slice.
static constexpr int kSnBwdBlocks = 64;
static constexpr int kSnBwdThreads = 256;

__device__ __forceinline__ long sn_src_index(long o, int CRS, int C, int RS, long row_stride, int krsc) {
  const int k = (int)(o / CRS), rem = (int)(o - (long)k * CRS);   // o indexes OIHW
  if (!krsc) return (long)k * row_stride + rem;
  const int c = rem / RS, rs = rem - c * RS;
  return (long)k * row_stride + (long)rs * C + c;
}

__global__ void __launch_bounds__(kSnBwdThreads)
sn_bwd_dot_kernel(const float* __restrict__ dwbar, long row_stride, int krsc, const float* __restrict__ W,
                  int K, int C, int RS, float* __restrict__ partial) {
  __shared__ float red[40];
  const int CRS = C * RS;
  const long total = (long)K * CRS;
  float acc = 0.0f;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x)
    acc = fmaf(dwbar[sn_src_index(o, CRS, C, RS, row_stride, krsc)], W[o], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kSnBwdThreads)
sn_bwd_apply_kernel(const float* __restrict__ dwbar, long row_stride, int krsc, const float* __restrict__ st,
                    const float* __restrict__ partial, int nparts, int K, int C, int RS, float* __restrict__ dW) {
  const int CRS = C * RS;
  const long total = (long)K * CRS;
  const float inv_sigma = st[0];
  const float* u = st + 1;
  const float* v = st + 1 + K;
  float s = 0.0f;
  for (int i = 0; i < nparts; ++i) s += partial[i];
  s *= inv_sigma;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int k = (int)(o / CRS), rem = (int)(o - (long)k * CRS);
    dW[o] = (dwbar[sn_src_index(o, CRS, C, RS, row_stride, krsc)] - s * u[k] * v[rem]) * inv_sigma;
  }
}

// ---- batched form: every spectral-normed layer of one backward pass in TWO launches ---------------
// A D update runs D twice (real / fake batch) with different (sigma, u, v) snapshots: 22 items, which as
// per-layer launches were 44 kernels of ~150 k elements each (launch-bound: 6 % of a D update).  Items are
// grouped by the parameter they belong to; a group's items are applied by the same threads, one after the
// other, and ADDED to the parameter's gradient (grad += dW: the AccumulateGrad pass is fused away too).
static constexpr int kSnMaxItems = 32;
static constexpr int kSnMaxGroupItems = 4;
static constexpr int kSnBatchBlocks = 32;     // dot-product blocks per item / apply blocks per group

struct SnBwdBatch {
  const float* dwbar[kSnMaxItems];
  const float* W[kSnMaxItems];
  const float* st[kSnMaxItems];
  long row_stride[kSnMaxItems];
  int krsc[kSnMaxItems], K[kSnMaxItems], C[kSnMaxItems], RS[kSnMaxItems];
  float* grad[kSnMaxItems];                   // per group
  int group_items[kSnMaxItems][kSnMaxGroupItems];
  int group_n[kSnMaxItems];
};

__global__ void __launch_bounds__(kSnBwdThreads)
sn_bwd_dot_batched_kernel(const SnBwdBatch b, float* __restrict__ partial) {
  __shared__ float red[40];
  const int it = blockIdx.y;
  const int C = b.C[it], RS = b.RS[it], CRS = C * RS;
  const long total = (long)b.K[it] * CRS;
  const float* __restrict__ dw = b.dwbar[it];
  const float* __restrict__ W = b.W[it];
  float acc = 0.0f;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x)
    acc = fmaf(dw[sn_src_index(o, CRS, C, RS, b.row_stride[it], b.krsc[it])], W[o], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[it * kSnBatchBlocks + blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kSnBwdThreads)
sn_bwd_apply_batched_kernel(const SnBwdBatch b, const float* __restrict__ partial) {
  const int g = blockIdx.y, n = b.group_n[g];
  const int i0 = b.group_items[g][0];
  const int K = b.K[i0], C = b.C[i0], RS = b.RS[i0], CRS = C * RS;
  const long total = (long)K * CRS;
  float s[kSnMaxGroupItems], inv[kSnMaxGroupItems];
#pragma unroll
  for (int q = 0; q < kSnMaxGroupItems; ++q) {
    s[q] = 0.0f; inv[q] = 0.0f;
    if (q < n) {
      const int it = b.group_items[g][q];
      float a = 0.0f;
      for (int i = 0; i < kSnBatchBlocks; ++i) a += partial[it * kSnBatchBlocks + i];
      inv[q] = b.st[it][0];
      s[q] = a * inv[q];
    }
  }
  float* __restrict__ grad = b.grad[g];
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int k = (int)(o / CRS), rem = (int)(o - (long)k * CRS);
    float acc = 0.0f;
#pragma unroll
    for (int q = 0; q < kSnMaxGroupItems; ++q) {
      if (q < n) {
        const int it = b.group_items[g][q];
        const float* st = b.st[it];
        acc += (b.dwbar[it][sn_src_index(o, CRS, C, RS, b.row_stride[it], b.krsc[it])] - s[q] * st[1 + k] * st[1 + K + rem]) * inv[q];
      }
    }
    grad[o] += acc;
  }
}

// bf16 [K][row_stride] <- W[k,c,r,s] * inv_sigma in (r,s,c) order, zero padded (im2col GEMM weights)
__global__ void pack_weight_padded_kernel(const float* __restrict__ w, int K, int C, int R, int S,
                                          const float* __restrict__ inv_sigma, __nv_bfloat16* __restrict__ out,
                                          int row_stride) {
  const long total = (long)K * row_stride;
  const float sc = inv_sigma ? __ldg(inv_sigma) : 1.0f;
  const int RSC = R * S * C;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i / row_stride), col = (int)(i - (long)k * row_stride);
    float v = 0.0f;
    if (col < RSC) {
      const int c = col % C, rs = col / C;
      v = w[((long)k * C + c) * (R * S) + rs] * sc;
    }
    out[i] = __float2bfloat16(v);
  }
}

// ---- batched weight packing of all spectral-normed conv layers of one forward: W/sigma -> bf16 operand layouts.
// mode 0: KRSC (forward) + CRSK with flipped taps (dgrad); mode 1: [K][row_stride] rows in (r,s,c) order, zero
// padded (thin layers).  One launch instead of one per layer (11 per D forward, 11 forwards per iteration).
struct PackBatch {
  const float* W[kSnMaxLayers];
  const float* inv_sigma[kSnMaxLayers];
  __nv_bfloat16* out0[kSnMaxLayers];      // KRSC / padded rows
  __nv_bfloat16* out1[kSnMaxLayers];      // flipped CRSK (mode 0)
  int K[kSnMaxLayers], C[kSnMaxLayers], R[kSnMaxLayers], S[kSnMaxLayers], mode[kSnMaxLayers], row_stride[kSnMaxLayers];
};

__global__ void __launch_bounds__(256)
pack_weight_batched_kernel(const PackBatch b) {
  const int L = blockIdx.y;
  const int K = b.K[L], C = b.C[L], R = b.R[L], S = b.S[L];
  const float* __restrict__ w = b.W[L];
  const float sc = b.inv_sigma[L] ? __ldg(b.inv_sigma[L]) : 1.0f;
  if (b.mode[L] == 0) {
    const long total = (long)K * C * R * S;
    __nv_bfloat16* krsc = b.out0[L];
    __nv_bfloat16* flip = b.out1[L];
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      long t = i;                                        // i indexes the KRSC destination: coalesced bf16 writes
      const int c = (int)(t % C); t /= C;
      const int s_ = (int)(t % S); t /= S;
      const int r = (int)(t % R);
      const int k = (int)(t / R);
      const __nv_bfloat16 v = __float2bfloat16(w[(((long)k * C + c) * R + r) * S + s_] * sc);
      krsc[i] = v;
      flip[(((long)c * R + (R - 1 - r)) * S + (S - 1 - s_)) * K + k] = v;
    }
  } else {
    const int rs = b.row_stride[L], RSC = R * S * C;
    const long total = (long)K * rs;
    __nv_bfloat16* out = b.out0[L];
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
      const int k = (int)(i / rs), col = (int)(i - (long)k * rs);
      float v = 0.0f;
      if (col < RSC) {
        const int c = col % C, t = col / C;
        v = w[((long)k * C + c) * (R * S) + t] * sc;
      }
      out[i] = __float2bfloat16(v);
    }
  }
}

}  // namespace mcgm

using namespace mcgm;

extern "C" int mcgm_sn_power_iter(int n_layers, const float* const* W_host, const int* K_host, const int* CRS_host,
                                  float* const* u_host, float* const* v_host, float* state,
                                  const long* state_off_host, int do_iter, float eps, void* stream) {
  MCGM_CHECK_ARG(n_layers > 0 && n_layers <= kSnMaxLayers, "sn_power_iter: 1..%d layers per call (got %d)", kSnMaxLayers, n_layers);
  MCGM_CHECK_ARG(W_host && K_host && CRS_host && u_host && v_host && state && state_off_host, "sn_power_iter: bad arguments");
  SnBatch b{};
  int smem_floats = 0;
  for (int i = 0; i < n_layers; ++i) {
    MCGM_CHECK_ARG(W_host[i] && u_host[i] && v_host[i] && K_host[i] > 0 && CRS_host[i] > 0, "sn_power_iter: layer %d invalid", i);
    b.W[i] = W_host[i]; b.u[i] = u_host[i]; b.v[i] = v_host[i];
    b.K[i] = K_host[i]; b.CRS[i] = CRS_host[i]; b.off[i] = state_off_host[i];
    const int need = 2 * K_host[i] + CRS_host[i] + 64;
    if (need > smem_floats) smem_floats = need;
  }
  const size_t smem = (size_t)smem_floats * sizeof(float);
  if (smem > 200 * 1024) return fail(MCGM_E_UNSUPPORTED, "sn_power_iter: layer too large for shared memory");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    MCGM_CUDA(cudaFuncSetAttribute(sn_power_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  sn_power_iter_kernel<<<n_layers, kSnThreads, smem, (cudaStream_t)stream>>>(b, state, do_iter, eps);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mcgm_sn_weight_bwd_workspace_bytes(void) { return kSnBwdBlocks * sizeof(float); }

extern "C" int mcgm_sn_weight_bwd(const float* dwbar, long row_stride, int krsc, const float* W_oihw,
                                  const float* sn_state, int K, int C, int R, int S, float* dW_oihw,
                                  void* workspace, void* stream) {
  MCGM_CHECK_ARG(dwbar && W_oihw && sn_state && dW_oihw && K > 0 && C > 0 && R > 0 && S > 0 && row_stride >= (long)C * R * S,
                 "sn_weight_bwd: bad arguments");
  MCGM_CHECK_ARG(workspace, "sn_weight_bwd: needs a workspace of mcgm_sn_weight_bwd_workspace_bytes()");
  const long total = (long)K * C * R * S;
  int blocks = (int)((total + kSnBwdThreads - 1) / kSnBwdThreads);
  if (blocks > kSnBwdBlocks) blocks = kSnBwdBlocks;
  float* partial = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  sn_bwd_dot_kernel<<<blocks, kSnBwdThreads, 0, st>>>(dwbar, row_stride, krsc, W_oihw, K, C, R * S, partial);
  MCGM_LAUNCH_CHECK();
  sn_bwd_apply_kernel<<<blocks, kSnBwdThreads, 0, st>>>(dwbar, row_stride, krsc, sn_state, partial, blocks, K, C, R * S, dW_oihw);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_pack_weight_batched(int n, const float* const* W_host, const float* const* inv_sigma_host,
                                        const int* K_host, const int* C_host, const int* R_host, const int* S_host,
                                        const int* mode_host, const int* row_stride_host, void* const* out0_host,
                                        void* const* out1_host, void* stream) {
  MCGM_CHECK_ARG(n > 0 && n <= kSnMaxLayers, "pack_weight_batched: 1..%d layers per call (got %d)", kSnMaxLayers, n);
  MCGM_CHECK_ARG(W_host && inv_sigma_host && K_host && C_host && R_host && S_host && mode_host && row_stride_host && out0_host && out1_host,
                 "pack_weight_batched: bad arguments");
  PackBatch b{};
  for (int i = 0; i < n; ++i) {
    MCGM_CHECK_ARG(W_host[i] && out0_host[i] && K_host[i] > 0 && C_host[i] > 0 && R_host[i] > 0 && S_host[i] > 0, "pack_weight_batched: layer %d invalid", i);
    MCGM_CHECK_ARG(mode_host[i] == 0 ? out1_host[i] != nullptr : row_stride_host[i] >= C_host[i] * R_host[i] * S_host[i],
                   "pack_weight_batched: layer %d: mode 0 needs the flipped output, mode 1 a row stride >= C*R*S", i);
    b.W[i] = W_host[i]; b.inv_sigma[i] = inv_sigma_host[i]; b.K[i] = K_host[i]; b.C[i] = C_host[i]; b.R[i] = R_host[i]; b.S[i] = S_host[i];
    b.mode[i] = mode_host[i]; b.row_stride[i] = row_stride_host[i];
    b.out0[i] = (__nv_bfloat16*)out0_host[i]; b.out1[i] = (__nv_bfloat16*)out1_host[i];
  }
  pack_weight_batched_kernel<<<dim3(64, n), 256, 0, (cudaStream_t)stream>>>(b);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mcgm_sn_weight_bwd_batched_workspace_bytes(void) { return (size_t)kSnMaxItems * kSnBatchBlocks * sizeof(float); }

extern "C" int mcgm_sn_weight_bwd_batched(int n, const float* const* dwbar_host, const long* row_stride_host,
                                          const int* krsc_host, const float* const* W_host,
                                          const float* const* state_host, const int* K_host, const int* C_host,
                                          const int* RS_host, float* const* grad_host, void* workspace, void* stream) {
  MCGM_CHECK_ARG(n > 0 && n <= kSnMaxItems, "sn_weight_bwd_batched: 1..%d items per call (got %d)", kSnMaxItems, n);
  MCGM_CHECK_ARG(dwbar_host && row_stride_host && krsc_host && W_host && state_host && K_host && C_host && RS_host && grad_host && workspace,
                 "sn_weight_bwd_batched: bad arguments");
  SnBwdBatch b{};
  int groups = 0;
  for (int i = 0; i < n; ++i) {
    MCGM_CHECK_ARG(dwbar_host[i] && W_host[i] && state_host[i] && grad_host[i] && K_host[i] > 0 && C_host[i] > 0 && RS_host[i] > 0 &&
                       row_stride_host[i] >= (long)C_host[i] * RS_host[i], "sn_weight_bwd_batched: item %d invalid", i);
    b.dwbar[i] = dwbar_host[i]; b.W[i] = W_host[i]; b.st[i] = state_host[i]; b.row_stride[i] = row_stride_host[i];
    b.krsc[i] = krsc_host[i]; b.K[i] = K_host[i]; b.C[i] = C_host[i]; b.RS[i] = RS_host[i];
    int g = 0;
    while (g < groups && b.grad[g] != grad_host[i]) ++g;
    if (g == groups) { b.grad[g] = grad_host[i]; b.group_n[g] = 0; ++groups; }
    else MCGM_CHECK_ARG(b.K[b.group_items[g][0]] == K_host[i] && b.C[b.group_items[g][0]] == C_host[i] && b.RS[b.group_items[g][0]] == RS_host[i],
                        "sn_weight_bwd_batched: items of one gradient differ in shape");
    MCGM_CHECK_ARG(b.group_n[g] < kSnMaxGroupItems, "sn_weight_bwd_batched: more than %d items for one gradient", kSnMaxGroupItems);
    b.group_items[g][b.group_n[g]++] = i;
  }
  cudaStream_t st = (cudaStream_t)stream;
  sn_bwd_dot_batched_kernel<<<dim3(kSnBatchBlocks, n), kSnBwdThreads, 0, st>>>(b, (float*)workspace);
  MCGM_LAUNCH_CHECK();
  sn_bwd_apply_batched_kernel<<<dim3(kSnBatchBlocks, groups), kSnBwdThreads, 0, st>>>(b, (const float*)workspace);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_pack_weight_padded(const float* w_oihw, int K, int C, int R, int S, const float* inv_sigma,
                                       void* out_bf16, int row_stride, void* stream) {
  MCGM_CHECK_ARG(w_oihw && out_bf16 && K > 0 && C > 0 && R > 0 && S > 0 && row_stride >= C * R * S, "pack_weight_padded: bad arguments");
  const long total = (long)K * row_stride;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  pack_weight_padded_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, K, C, R, S, inv_sigma,
                                                                          (__nv_bfloat16*)out_bf16, row_stride);
  MCGM_LAUNCH_CHECK();
  return 0;
}
