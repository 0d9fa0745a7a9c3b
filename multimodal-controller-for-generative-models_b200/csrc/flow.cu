// flow.cu — the few-channel (C <= 64) arithmetic of one Glow flow step on fp32 NHWC tensors
// (SURVEY.md §8 a11-a13), forward and backward:
//
//   front  = ActNorm(C) -> InvConv2dLU          src/models/mcglow.py:24-55, 76-116
//            a = scale*(x + loc);  y = W a   (W = calc_weight(), a CxC matrix built on the host side of the
//            C ABI from the LU parameters — O(C^3) parameter work stays with the caller); the kernel also emits
//            the coupling network's input  xa = bf16(y[:, :C/2]) zero-padded to `xa_ch` channels, the
//            operand layout of the tcgen05 conv, so no separate chunk / pad / cast pass exists.
//   tail   = AffineCoupling minus its network   src/models/mcglow.py:119-130, 153-164
//            h = net_out * exp(3*zscale)  (ZeroConv2d's scale);  s = sigmoid(h[:C/2] + 2)
//            out = [y[:C/2], (y[C/2:] + h[C/2:]) * s];  logdet[n] = sum log s     (additive: out_b = in_b + h)
//
// Every kernel is a single pass over tensors that are 10-40x smaller than the 512-channel activations of
// the coupling network: HBM-bound, per-pixel arithmetic C^2 FMAs at most.  The reference runs ~25 eager
// launches per flow step for this work; here it is 2 forward + 4 backward kernels.
// Reductions (dW, dscale, dloc, dzscale, logdet) are two-level with a fixed order: deterministic.
#include "common.cuh"

namespace mcgm {

static constexpr int kFlowThreads = 256;
static constexpr int kFlowMaxC = 64;

// ---- front, forward -----------------------------------------------------------------------------
// thread = (pixel, quad of output channels); W, loc, scale in shared memory
__global__ void __launch_bounds__(kFlowThreads)
flow_front_fwd_kernel(const float* __restrict__ x, const float* __restrict__ loc, const float* __restrict__ scale,
                      const float* __restrict__ W, float* __restrict__ y, __nv_bfloat16* __restrict__ xa,
                      long P, int C, int xa_ch, int PB) {
  __shared__ float sW[kFlowMaxC * kFlowMaxC];
  __shared__ float sl[kFlowMaxC], ss[kFlowMaxC];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sW[i] = W[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sl[i] = loc[i]; ss[i] = scale[i]; }
  __syncthreads();
  const int G = C >> 2;
  const int og = threadIdx.x % G, pl = threadIdx.x / G;
  if (pl >= PB) return;
  const long p = (long)blockIdx.x * PB + pl;
  if (p >= P) return;
  const float4* xp = reinterpret_cast<const float4*>(x + p * C);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < C; i += 4) {
    const float4 xv = __ldg(xp + (i >> 2));
    const float a[4] = {ss[i] * (xv.x + sl[i]), ss[i + 1] * (xv.y + sl[i + 1]), ss[i + 2] * (xv.z + sl[i + 2]),
                        ss[i + 3] * (xv.w + sl[i + 3])};
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const float4 w = *reinterpret_cast<const float4*>(sW + (og * 4 + o) * C + i);
      acc[o] = fmaf(w.x, a[0], acc[o]); acc[o] = fmaf(w.y, a[1], acc[o]);
      acc[o] = fmaf(w.z, a[2], acc[o]); acc[o] = fmaf(w.w, a[3], acc[o]);
    }
  }
  *reinterpret_cast<float4*>(y + p * C + og * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if (xa != nullptr) {
    const int half = C >> 1;
    __nv_bfloat16* xr = xa + p * xa_ch;
    __nv_bfloat16 q[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) q[o] = __float2bfloat16(og * 4 + o < half ? acc[o] : 0.0f);
    *reinterpret_cast<uint2*>(xr + og * 4) = *reinterpret_cast<const uint2*>(q);
    for (int g = G + og; g < (xa_ch >> 2); g += G) *reinterpret_cast<uint2*>(xr + g * 4) = make_uint2(0u, 0u);
  }
}

// ---- front, backward: dx = scale * W^T (dy + [dxa, 0]) ------------------------------------------
__global__ void __launch_bounds__(kFlowThreads)
flow_front_bwd_dx_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ dxa,
                         const float* __restrict__ scale, const float* __restrict__ W, float* __restrict__ dx,
                         long P, int C, int xa_ch, int PB) {
  __shared__ float sWt[kFlowMaxC * kFlowMaxC];            // transposed: sWt[i*C + o] = W[o*C + i]
  __shared__ float ss[kFlowMaxC];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) sWt[(i % C) * C + i / C] = W[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) ss[i] = scale[i];
  __syncthreads();
  const int G = C >> 2;
  const int ig = threadIdx.x % G, pl = threadIdx.x / G;
  if (pl >= PB) return;
  const long p = (long)blockIdx.x * PB + pl;
  if (p >= P) return;
  const int half = C >> 1;
  const float4* gp = reinterpret_cast<const float4*>(dy + p * C);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int o = 0; o < C; o += 4) {
    const float4 gv = __ldg(gp + (o >> 2));
    float g[4] = {gv.x, gv.y, gv.z, gv.w};
    if (dxa != nullptr && o < half) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (o + k < half) g[k] += __bfloat162float(dxa[p * xa_ch + o + k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(sWt + (ig * 4 + k) * C + o);
      acc[k] = fmaf(w.x, g[0], acc[k]); acc[k] = fmaf(w.y, g[1], acc[k]);
      acc[k] = fmaf(w.z, g[2], acc[k]); acc[k] = fmaf(w.w, g[3], acc[k]);
    }
  }
  *reinterpret_cast<float4*>(dx + p * C + ig * 4) =
      make_float4(acc[0] * ss[ig * 4], acc[1] * ss[ig * 4 + 1], acc[2] * ss[ig * 4 + 2], acc[3] * ss[ig * 4 + 3]);
}

// ---- front, backward: G[o][i] = sum_p g[p,o] * (x[p,i] + loc[i]),  G[o][C] = sum_p g[p,o] --------
// (g = dy + [dxa, 0]).  From G the caller gets dW = G[:, :C] * scale, dscale = sum_o W*G, dloc = scale * W^T G[:, C].
static constexpr int kFlowTile = 32;                      // pixels per shared-memory tile
static constexpr int kFlowAccPerThread = (kFlowMaxC * (kFlowMaxC + 1) + kFlowThreads - 1) / kFlowThreads;

__global__ void __launch_bounds__(kFlowThreads)
flow_front_bwd_reduce_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ dxa,
                             const float* __restrict__ x, const float* __restrict__ loc,
                             float* __restrict__ partial, long P, int C, int xa_ch) {
  __shared__ float sA[kFlowTile][kFlowMaxC];              // g
  __shared__ float sB[kFlowTile][kFlowMaxC + 1];          // x + loc | 1
  const int half = C >> 1, C1 = C + 1, nout = C * C1;
  float acc[kFlowAccPerThread];
#pragma unroll
  for (int k = 0; k < kFlowAccPerThread; ++k) acc[k] = 0.0f;
  const long tiles = (P + kFlowTile - 1) / kFlowTile;
  for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long p0 = t * kFlowTile;
    for (int e = threadIdx.x; e < kFlowTile * C; e += blockDim.x) {
      const int pp = e / C, c = e % C;
      const long p = p0 + pp;
      float g = 0.0f, b = 0.0f;
      if (p < P) {
        g = dy[p * C + c];
        if (dxa != nullptr && c < half) g += __bfloat162float(dxa[p * xa_ch + c]);
        b = x[p * C + c] + loc[c];
      }
      sA[pp][c] = g;
      sB[pp][c] = b;
      if (c == 0) sB[pp][C] = p < P ? 1.0f : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFlowAccPerThread; ++k) {
      const int idx = threadIdx.x + k * kFlowThreads;
      if (idx < nout) {
        const int o = idx / C1, i = idx % C1;
        float a = acc[k];
#pragma unroll 8
        for (int pp = 0; pp < kFlowTile; ++pp) a = fmaf(sA[pp][o], sB[pp][i], a);
        acc[k] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kFlowAccPerThread; ++k) {
    const int idx = threadIdx.x + k * kFlowThreads;
    if (idx < nout) partial[(size_t)blockIdx.x * nout + idx] = acc[k];
  }
}

// out[idx] = sum_b partial[b][idx] in block order
__global__ void flow_sum_partials_kernel(const float* __restrict__ partial, int nblocks, int nout, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nout) return;
  float a = 0.0f;
  int b = 0;
  for (; b + 4 <= nblocks; b += 4) {
    const float v0 = partial[(size_t)b * nout + idx], v1 = partial[(size_t)(b + 1) * nout + idx];
    const float v2 = partial[(size_t)(b + 2) * nout + idx], v3 = partial[(size_t)(b + 3) * nout + idx];
    a += v0; a += v1; a += v2; a += v3;
  }
  for (; b < nblocks; ++b) a += partial[(size_t)b * nout + idx];
  out[idx] = a;
}

// ---- tail: one block per sample; thread = (channel of the b-half, pixel lane) ----------------------
__device__ __forceinline__ float block_sum_fixed(float v, float* red) {
  // deterministic block reduction: warp shuffles, then the warp sums in order
  v = warp_sum(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = 0.0f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int k = 0; k < nw; ++k) t += red[k];
  return t;
}

__global__ void __launch_bounds__(kFlowThreads)
flow_tail_fwd_kernel(const float* __restrict__ y, const float* __restrict__ net, const float* __restrict__ zscale,
                     float* __restrict__ out, float* __restrict__ logdet, int HW, int C, int net_ch, int affine) {
  __shared__ float red[kFlowThreads / 32];
  const int n = blockIdx.x, half = C >> 1;
  const int c = threadIdx.x % half, pl = threadIdx.x / half, PL = blockDim.x / half;
  float ld = 0.0f;
  if (pl < PL) {
    const float e_s = affine ? expf(3.0f * zscale[c]) : 0.0f;
    const float e_t = expf(3.0f * zscale[affine ? half + c : c]);
    for (int hw = pl; hw < HW; hw += PL) {
      const long p = (long)n * HW + hw;
      const float in_a = y[p * C + c], in_b = y[p * C + half + c];
      float ob;
      if (affine) {
        const float ls = net[p * net_ch + c] * e_s + 2.0f;
        const float t = net[p * net_ch + half + c] * e_t;
        const float s = 1.0f / (1.0f + expf(-ls));
        ob = (in_b + t) * s;
        ld += logf(s);
      } else {
        ob = in_b + net[p * net_ch + c] * e_t;
      }
      out[p * C + c] = in_a;
      out[p * C + half + c] = ob;
    }
  }
  const float tot = block_sum_fixed(ld, red);
  if (threadIdx.x == 0 && logdet != nullptr) logdet[n] = tot;
}

// The sampling direction (AffineCoupling.reverse, src/models/mcglow.py:166-175): given the coupling's OUTPUT
// `outp` and the network evaluated on its first half, in_b = out_b / sigmoid(log_s + 2) - t  (additive:
// in_b = out_b - net_out); the first half passes through.
__global__ void __launch_bounds__(kFlowThreads)
flow_tail_rev_kernel(const float* __restrict__ outp, const float* __restrict__ net, const float* __restrict__ zscale,
                     float* __restrict__ y, long P, int C, int net_ch, int affine) {
  const int half = C >> 1;
  const long total = P * half;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long p = e / half;
    const int c = (int)(e - p * half);
    const float out_a = outp[p * C + c], out_b = outp[p * C + half + c];
    float ib;
    if (affine) {
      const float ls = net[p * net_ch + c] * expf(3.0f * zscale[c]) + 2.0f;
      const float t = net[p * net_ch + half + c] * expf(3.0f * zscale[half + c]);
      const float s = 1.0f / (1.0f + expf(-ls));
      ib = out_b / s - t;
    } else {
      ib = out_b - net[p * net_ch + c] * expf(3.0f * zscale[c]);
    }
    y[p * C + c] = out_a;
    y[p * C + half + c] = ib;
  }
}

// dy (grad of the front's y), dnet (grad of the network output, all net_ch channels written), and the
// per-sample partial of dzscale [N][nz] (nz = C affine, C/2 additive): dzscale[c] = 3 * sum dh*h.
__global__ void __launch_bounds__(kFlowThreads)
flow_tail_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ dlogdet, const float* __restrict__ y,
                     const float* __restrict__ net, const float* __restrict__ zscale, float* __restrict__ dy,
                     float* __restrict__ dnet, float* __restrict__ dz_partial, int HW, int C, int net_ch, int affine) {
  __shared__ float sz[2][kFlowThreads];
  const int n = blockIdx.x, half = C >> 1;
  const int c = threadIdx.x % half, pl = threadIdx.x / half, PL = blockDim.x / half;
  const int used = PL * half;
  float dzs = 0.0f, dzt = 0.0f;
  if (threadIdx.x < used) {
    const float e_s = affine ? expf(3.0f * zscale[c]) : 0.0f;
    const float e_t = expf(3.0f * zscale[affine ? half + c : c]);
    const float dld = (affine && dlogdet != nullptr) ? dlogdet[n] : 0.0f;
    for (int hw = pl; hw < HW; hw += PL) {
      const long p = (long)n * HW + hw;
      const float in_b = y[p * C + half + c];
      const float dob = dout[p * C + half + c];
      dy[p * C + c] = dout[p * C + c];
      if (affine) {
        const float h_s = net[p * net_ch + c] * e_s, h_t = net[p * net_ch + half + c] * e_t;
        const float s = 1.0f / (1.0f + expf(-(h_s + 2.0f)));
        const float d_s = dob * (in_b + h_t) + dld / s;
        const float dh_s = d_s * s * (1.0f - s), dh_t = dob * s;
        dy[p * C + half + c] = dob * s;
        dnet[p * net_ch + c] = dh_s * e_s;
        dnet[p * net_ch + half + c] = dh_t * e_t;
        dzs = fmaf(dh_s, h_s, dzs);
        dzt = fmaf(dh_t, h_t, dzt);
      } else {
        const float h_t = net[p * net_ch + c] * e_t;
        dy[p * C + half + c] = dob;
        dnet[p * net_ch + c] = dob * e_t;
        dzt = fmaf(dob, h_t, dzt);
      }
    }
  }
  // channels of the network output the coupling never reads: zero gradient
  const int live = affine ? C : half;
  for (long e = threadIdx.x; e < (long)HW * (net_ch - live); e += blockDim.x) {
    const long p = (long)n * HW + e / (net_ch - live);
    dnet[p * net_ch + live + e % (net_ch - live)] = 0.0f;
  }
  sz[0][threadIdx.x] = dzs;
  sz[1][threadIdx.x] = dzt;
  __syncthreads();
  if (threadIdx.x < half) {
    float a = 0.0f, b = 0.0f;
    for (int k = 0; k < PL; ++k) { a += sz[0][k * half + threadIdx.x]; b += sz[1][k * half + threadIdx.x]; }
    const int nz = affine ? C : half;
    if (affine) {
      dz_partial[(size_t)n * nz + threadIdx.x] = 3.0f * a;
      dz_partial[(size_t)n * nz + half + threadIdx.x] = 3.0f * b;
    } else {
      dz_partial[(size_t)n * nz + threadIdx.x] = 3.0f * b;
    }
  }
}

static int flow_check(long P, int C) {
  if (P < 0 || C < 4 || C > kFlowMaxC || (C & 3)) return fail(MCGM_E_UNSUPPORTED, "flow: C=%d must be a multiple of 4 in [4, %d]", C, kFlowMaxC);
  return 0;
}

static constexpr int kFlowReduceBlocks = 148;

}  // namespace mcgm

using namespace mcgm;

extern "C" int mcgm_flow_front_fwd(const float* x, const float* loc, const float* scale, const float* W, float* y,
                                   void* xa_bf16, long P, int C, int xa_ch, void* stream) {
  MCGM_CHECK_ARG(x && loc && scale && W && y, "flow_front_fwd: bad arguments");
  if (int e = flow_check(P, C)) return e;
  if (xa_bf16) MCGM_CHECK_ARG(xa_ch % 4 == 0 && xa_ch >= C, "flow_front_fwd: xa_ch=%d must be a multiple of 4 and >= C", xa_ch);
  if (P == 0) return 0;
  const int G = C / 4, PB = kFlowThreads / G;
  const long blocks = (P + PB - 1) / PB;
  flow_front_fwd_kernel<<<(unsigned)blocks, kFlowThreads, 0, (cudaStream_t)stream>>>(x, loc, scale, W, y, (__nv_bfloat16*)xa_bf16,
                                                                                  P, C, xa_ch, PB);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mcgm_flow_front_bwd_workspace_bytes(int C) {
  return (size_t)kFlowReduceBlocks * C * (C + 1) * sizeof(float);
}

extern "C" int mcgm_flow_front_bwd(const float* dy, const void* dxa_bf16, const float* x, const float* loc,
                                   const float* scale, const float* W, float* dx, float* G_out, long P, int C,
                                   int xa_ch, void* workspace, void* stream) {
  MCGM_CHECK_ARG(dy && x && loc && scale && W && dx && G_out && workspace, "flow_front_bwd: bad arguments");
  if (int e = flow_check(P, C)) return e;
  if (P == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int G = C / 4, PB = kFlowThreads / G;
  const long blocks = (P + PB - 1) / PB;
  flow_front_bwd_dx_kernel<<<(unsigned)blocks, kFlowThreads, 0, st>>>(dy, (const __nv_bfloat16*)dxa_bf16, scale, W, dx, P, C, xa_ch, PB);
  MCGM_LAUNCH_CHECK();
  const long tiles = (P + kFlowTile - 1) / kFlowTile;
  const int rb = (int)(tiles < kFlowReduceBlocks ? tiles : kFlowReduceBlocks);
  const int nout = C * (C + 1);
  flow_front_bwd_reduce_kernel<<<rb, kFlowThreads, 0, st>>>(dy, (const __nv_bfloat16*)dxa_bf16, x, loc, (float*)workspace, P, C, xa_ch);
  MCGM_LAUNCH_CHECK();
  flow_sum_partials_kernel<<<(nout + 255) / 256, 256, 0, st>>>((const float*)workspace, rb, nout, G_out);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_flow_tail_fwd(const float* y, const float* net, const float* zscale, float* out, float* logdet,
                                  int N, int HW, int C, int net_ch, int affine, void* stream) {
  MCGM_CHECK_ARG(y && net && zscale && out && N >= 0 && HW > 0, "flow_tail_fwd: bad arguments");
  if (int e = flow_check((long)N * HW, C)) return e;
  MCGM_CHECK_ARG(net_ch >= (affine ? C : C / 2), "flow_tail_fwd: network output has %d channels, need %d", net_ch, affine ? C : C / 2);
  MCGM_CHECK_ARG(!affine || logdet, "flow_tail_fwd: affine coupling needs a logdet output");
  if (N == 0) return 0;
  flow_tail_fwd_kernel<<<N, kFlowThreads, 0, (cudaStream_t)stream>>>(y, net, zscale, out, logdet, HW, C, net_ch, affine);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_flow_tail_rev(const float* out, const float* net, const float* zscale, float* y, int N, int HW, int C,
                                  int net_ch, int affine, void* stream) {
  MCGM_CHECK_ARG(out && net && zscale && y && N >= 0 && HW > 0, "flow_tail_rev: bad arguments");
  if (int e = flow_check((long)N * HW, C)) return e;
  MCGM_CHECK_ARG(net_ch >= (affine ? C : C / 2), "flow_tail_rev: network output has %d channels, need %d", net_ch, affine ? C : C / 2);
  if (N == 0) return 0;
  const long total = (long)N * HW * (C / 2);
  long blocks = (total + kFlowThreads - 1) / kFlowThreads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  flow_tail_rev_kernel<<<(int)blocks, kFlowThreads, 0, (cudaStream_t)stream>>>(out, net, zscale, y, (long)N * HW, C, net_ch, affine);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_flow_tail_bwd(const float* dout, const float* dlogdet, const float* y, const float* net,
                                  const float* zscale, float* dy, float* dnet, float* dz_partial, int N, int HW, int C,
                                  int net_ch, int affine, void* stream) {
  MCGM_CHECK_ARG(dout && y && net && zscale && dy && dnet && dz_partial && N >= 0 && HW > 0, "flow_tail_bwd: bad arguments");
  if (int e = flow_check((long)N * HW, C)) return e;
  MCGM_CHECK_ARG(net_ch >= (affine ? C : C / 2), "flow_tail_bwd: network output has %d channels, need %d", net_ch, affine ? C : C / 2);
  if (N == 0) return 0;
  flow_tail_bwd_kernel<<<N, kFlowThreads, 0, (cudaStream_t)stream>>>(dout, dlogdet, y, net, zscale, dy, dnet, dz_partial, HW, C,
                                                                     net_ch, affine);
  MCGM_LAUNCH_CHECK();
  return 0;
}
