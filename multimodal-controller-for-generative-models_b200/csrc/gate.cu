// gate.cu — HBM-bound kernels of the MC block: codebook packing, the
// MultimodalController mask (src/modules/modules.py:71-76), train-mode BatchNorm
// statistics, and the fused BN-apply → ReLU → nearest-2x-upsample → mask gate
// (src/models/mcgan.py:15-18,20-22,56-58 and :104-105) forward and backward.
//
// Layout: NHWC; every thread owns one 16-byte channel group (8 bf16 / 4 fp32)
// and walks pixels, so a warp reads whole contiguous pixel rows (128-bit,
// coalesced) and per-channel parameters live in registers.
#include "common.cuh"

namespace mcgm {

static constexpr int kGateThreads = 256;
static constexpr int kMaxPartialBlocks = 148 * 4;

// ---------------------------------------------------------------------------
// codebook packing
// ---------------------------------------------------------------------------
__global__ void pack_codebook_kernel(const float* __restrict__ cb, int M, int C, int words,
                                     uint32_t* __restrict__ bits) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per output word
  if (idx >= M * words) return;
  int row = idx / words, w = idx % words;
  uint32_t v = 0;
  for (int i = 0; i < 32; ++i) {
    int c = w * 32 + i;
    if (c < C && cb[(size_t)row * C + c] != 0.0f) v |= (1u << i);
  }
  bits[idx] = v;
}

// ---------------------------------------------------------------------------
// MultimodalController.forward : y = x * code
// ---------------------------------------------------------------------------
template <typename T>
__global__ void mc_forward_nhwc_kernel(const T* __restrict__ x, T* __restrict__ y, long rows /*N*P*/,
                                       int P, int C, const int32_t* __restrict__ labels,
                                       const uint32_t* __restrict__ bits, int words,
                                       const float* __restrict__ code) {
  constexpr int V = Vec<T>::N;
  const int CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  for (long p = (long)blockIdx.x * PL + pl; p < rows; p += (long)gridDim.x * PL) {
    const int n = (int)(p / P);
    Vec<T> v;
    v.load(x + p * C + cg * V);
    if (labels) {
      uint32_t m = mask_bits(bits, words, __ldg(labels + n), cg * V, V);
#pragma unroll
      for (int i = 0; i < V; ++i) v.v[i] = ((m >> i) & 1u) ? v.v[i] : 0.0f;
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v.v[i] *= __ldg(code + (size_t)n * C + cg * V + i);
    }
    v.store(y + p * C + cg * V);
  }
}

// NCHW-contiguous x: plane (n,c) of HW elements is scaled by code(n,c).
template <typename T>
__global__ void mc_forward_nchw_kernel(const T* __restrict__ x, T* __restrict__ y, long total, int HW,
                                       int C, const int32_t* __restrict__ labels,
                                       const uint32_t* __restrict__ bits, int words,
                                       const float* __restrict__ code) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    long plane = e / HW;
    int n = (int)(plane / C), c = (int)(plane % C);
    float s;
    if (labels) {
      uint32_t w = __ldg(bits + (size_t)__ldg(labels + n) * words + (c >> 5));
      s = ((w >> (c & 31)) & 1u) ? 1.0f : 0.0f;
    } else {
      s = __ldg(code + (size_t)n * C + c);
    }
    if constexpr (sizeof(T) == 4) {
      y[e] = x[e] * s;
    } else {
      y[e] = __float2bfloat16(__bfloat162float(x[e]) * s);
    }
  }
}

// ---------------------------------------------------------------------------
// BatchNorm statistics: per-block partial sums, then one warp per channel
// reduces the partials in a fixed order in double (deterministic).
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGateThreads)
bn_partial_kernel(const T* __restrict__ x, long P, int C, float* __restrict__ partial) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 4;
  const int CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = q[i] = 0.0f;
  const long stride = (long)gridDim.x * PL;
  long p = (long)blockIdx.x * PL + pl;
  for (; p + (U - 1) * stride < P; p += U * stride) {
    Vec<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u].load(x + (p + u * stride) * C + cg * V);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] += v[u].v[i]; q[i] = fmaf(v[u].v[i], v[u].v[i], q[i]); }
  }
  for (; p < P; p += stride) {
    Vec<T> v;
    v.load(x + p * C + cg * V);
#pragma unroll
    for (int i = 0; i < V; ++i) { s[i] += v.v[i]; q[i] = fmaf(v.v[i], v.v[i], q[i]); }
  }
  // block reduction over the PL pixel lanes
  extern __shared__ float sm[];  // [2][PL][C]
  float* ss = sm;
  float* sq = sm + (size_t)PL * C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    ss[pl * C + cg * V + i] = s[i];
    sq[pl * C + cg * V + i] = q[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < PL; ++l) { a += ss[l * C + c]; b += sq[l * C + c]; }
    partial[((size_t)blockIdx.x * 2 + 0) * C + c] = a;
    partial[((size_t)blockIdx.x * 2 + 1) * C + c] = b;
  }
}

// Per-channel sums of the [nblocks][2][C] block partials, in fp64 and in a fixed order (deterministic).
// Block = 32 channels x 32 partial lanes: every warp reads 128 contiguous bytes, four independent loads per
// lane are in flight, and the 32 lane sums of a channel are combined through shared memory in lane order.
// (One warp per channel with lane-strided rows was a chain of ~20 dependent L2 round trips: 10 us.)
__device__ __forceinline__ void channel_partial_sums(const float* __restrict__ partial, int nblocks, int C,
                                                     double& S, double& Q) {
  __shared__ double sS[32][33], sQ[32][33];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  S = 0.0; Q = 0.0;
  if (c < C) {
    int b = bl;
    for (; b + 96 < nblocks; b += 128) {
      float s[4], q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[u] = partial[((size_t)(b + 32 * u) * 2 + 0) * C + c];
        q[u] = partial[((size_t)(b + 32 * u) * 2 + 1) * C + c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { S += (double)s[u]; Q += (double)q[u]; }
    }
    for (; b < nblocks; b += 32) {
      S += (double)partial[((size_t)b * 2 + 0) * C + c];
      Q += (double)partial[((size_t)b * 2 + 1) * C + c];
    }
  }
  sS[bl][cl] = S; sQ[bl][cl] = Q;
  __syncthreads();
  if (bl == 0) {
    S = 0.0; Q = 0.0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) { S += sS[k][cl]; Q += sQ[k][cl]; }
  }
}

__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, double count,
                   float* __restrict__ mean, float* __restrict__ rstd,
                   float* __restrict__ running_mean, float* __restrict__ running_var,
                   float momentum, float eps) {
  double S, Q;
  channel_partial_sums(partial, nblocks, C, S, Q);
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if ((threadIdx.x >> 5) == 0 && c < C) {
    double m = S / count;
    double var = Q / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * m);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
    }
  }
}

// ---------------------------------------------------------------------------
// fused gate kernels.
//
// Work decomposition: a TILE is TP = PL*U consecutive pixels (PL pixel lanes of the
// block x U pixels per thread); a warp always touches whole contiguous pixel rows.
// All U 16-byte loads of a thread are issued back to back as raw uint4 (4 registers
// each) before any arithmetic, so one block keeps PL*U*C*elt bytes in flight.  When
// H*W is a multiple of TP the sample index — hence the codebook row — is uniform per
// tile and looked up once; index arithmetic is 32-bit with shifts for power-of-two
// extents.
// ---------------------------------------------------------------------------
struct GateGeom {
  int N, H, W, C;
  uint32_t P;        // N*H*W
  PixDiv hw, w;      // dividers by H*W and by W
  int uniform_n;     // H*W % TP == 0
};

template <typename T, bool UP2>
__global__ void __launch_bounds__(kGateThreads, 2)
gate_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, T* __restrict__ y_raw, const GateGeom g,
                const int32_t* __restrict__ labels, const uint32_t* __restrict__ bits, int words,
                const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, int flags) {
  constexpr int V = Vec<T>::N;
  constexpr int U = UP2 ? 4 : 8;
  const int C = g.C, CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  const uint32_t TP = PL * U, HW = (uint32_t)g.H * g.W;
  const bool relu = flags & MCGM_GATE_RELU, use_mask = flags & MCGM_GATE_MASK;
  float a[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if (flags & MCGM_GATE_BN) {
      const int c = cg * V + i;
      a[i] = gamma[c] * rstd[c];
      b[i] = fmaf(-mean[c], a[i], beta[c]);
    } else {
      a[i] = 1.0f; b[i] = 0.0f;
    }
  }
  const uint32_t tiles = (g.P + TP - 1) / TP;
  for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint32_t base = tile * TP + pl;
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = base + u * PL;
      if (p < g.P) raw[u] = ldg16(x + (size_t)p * C + cg * V);
    }
    uint32_t m_tile = 0xffffffffu;
    if (use_mask && g.uniform_n) m_tile = mask_bits(bits, words, __ldg(labels + g.hw.div(tile * TP)), cg * V, V);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = base + u * PL;
      if (p >= g.P) continue;
      const uint32_t n = g.hw.div(p);
      uint32_t m = m_tile;
      if (use_mask && !g.uniform_n) m = mask_bits(bits, words, __ldg(labels + n), cg * V, V);
      float v[V];
      unpack16<T>(raw[u], v);
      if (y_raw) {
        Vec<T> r;
#pragma unroll
        for (int i = 0; i < V; ++i) r.v[i] = ((m >> i) & 1u) ? v[i] : 0.0f;
        r.store(y_raw + (size_t)p * C + cg * V);
      }
      Vec<T> o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float z = fmaf(a[i], v[i], b[i]);
        if (relu) z = fmaxf(z, 0.0f);
        o.v[i] = ((m >> i) & 1u) ? z : 0.0f;
      }
      if (UP2) {
        const uint32_t hw = p - n * HW, h = g.w.div(hw), w = hw - h * g.W;
        T* o0 = y + (((size_t)n * 2 * g.H + 2 * h) * 2 * g.W + 2 * w) * C + cg * V;
        o.store(o0);
        o.store(o0 + C);
        o.store(o0 + (size_t)2 * g.W * C);
        o.store(o0 + (size_t)2 * g.W * C + C);
      } else {
        o.store(y + (size_t)p * C + cg * V);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// fused gate backward
//   g  = (UP2 ? sum_{2x2} dy : dy) * 1[z>0] * mask          z = a*x+b
//   BN:  dbeta = sum g ; dgamma = sum g*xhat ; dx = a*(g - dbeta/P - xhat*dgamma/P)
//   else dx = g
//   dx += mask * d_raw   (gradient of the raw-gated second output)
// ---------------------------------------------------------------------------
template <typename T, bool UP2>
struct DyRaw {
  uint4 r[UP2 ? 4 : 1];
  __device__ __forceinline__ void load(const T* __restrict__ dy, const GateGeom& g, uint32_t p, uint32_t n, int coff) {
    if (UP2) {
      const uint32_t HW = (uint32_t)g.H * g.W;
      const uint32_t hw = p - n * HW, h = g.w.div(hw), w = hw - h * g.W;
      const T* b0 = dy + (((size_t)n * 2 * g.H + 2 * h) * 2 * g.W + 2 * w) * g.C + coff;
      r[0] = ldg16(b0);
      r[UP2 ? 1 : 0] = ldg16(b0 + g.C);
      r[UP2 ? 2 : 0] = ldg16(b0 + (size_t)2 * g.W * g.C);
      r[UP2 ? 3 : 0] = ldg16(b0 + (size_t)2 * g.W * g.C + g.C);
    } else {
      r[0] = ldg16(dy + (size_t)p * g.C + coff);
    }
  }
  __device__ __forceinline__ void get(float* v) const {
    constexpr int V = Vec<T>::N;
    unpack16<T>(r[0], v);
    if (UP2) {
      float t1[V], t2[V], t3[V];
      unpack16<T>(r[UP2 ? 1 : 0], t1);
      unpack16<T>(r[UP2 ? 2 : 0], t2);
      unpack16<T>(r[UP2 ? 3 : 0], t3);
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = (v[i] + t1[i]) + (t2[i] + t3[i]);
    }
  }
};

template <typename T, bool UP2>
__global__ void __launch_bounds__(kGateThreads, 2)
gate_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const GateGeom g,
                       const int32_t* __restrict__ labels, const uint32_t* __restrict__ bits, int words,
                       const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta, int flags,
                       float* __restrict__ partial) {
  constexpr int V = Vec<T>::N;
  constexpr int U = UP2 ? 2 : 4;
  const int C = g.C, CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  const uint32_t TP = PL * U;
  const bool relu = flags & MCGM_GATE_RELU, use_mask = flags & MCGM_GATE_MASK;
  float a[V], b[V], mu[V], rs[V], s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    mu[i] = mean[c]; rs[i] = rstd[c];
    a[i] = gamma[c] * rs[i];
    b[i] = fmaf(-mu[i], a[i], beta[c]);
    s1[i] = s2[i] = 0.0f;
  }
  const uint32_t tiles = (g.P + TP - 1) / TP;
  for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint32_t base = tile * TP + pl;
    uint4 xr[U];
    DyRaw<T, UP2> dr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = base + u * PL;
      if (p < g.P) {
        xr[u] = ldg16(x + (size_t)p * C + cg * V);
        dr[u].load(dy, g, p, g.hw.div(p), cg * V);
      }
    }
    uint32_t m_tile = 0xffffffffu;
    if (use_mask && g.uniform_n) m_tile = mask_bits(bits, words, __ldg(labels + g.hw.div(tile * TP)), cg * V, V);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = base + u * PL;
      if (p >= g.P) continue;
      uint32_t m = m_tile;
      if (use_mask && !g.uniform_n) m = mask_bits(bits, words, __ldg(labels + g.hw.div(p)), cg * V, V);
      float xv[V], gv[V];
      unpack16<T>(xr[u], xv);
      dr[u].get(gv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float z = fmaf(a[i], xv[i], b[i]);
        float gi = ((m >> i) & 1u) ? gv[i] : 0.0f;
        if (relu && !(z > 0.0f)) gi = 0.0f;
        s1[i] += gi;
        s2[i] = fmaf(gi, (xv[i] - mu[i]) * rs[i], s2[i]);
      }
    }
  }
  extern __shared__ float sm[];  // [2][PL][C]
  float* ss = sm;
  float* sq = sm + (size_t)PL * C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    ss[pl * C + cg * V + i] = s1[i];
    sq[pl * C + cg * V + i] = s2[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float u = 0.f, w = 0.f;
    for (int l = 0; l < PL; ++l) { u += ss[l * C + c]; w += sq[l * C + c]; }
    partial[((size_t)blockIdx.x * 2 + 0) * C + c] = u;
    partial[((size_t)blockIdx.x * 2 + 1) * C + c] = w;
  }
}

__global__ void __launch_bounds__(1024)
gate_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  double S, Q;
  channel_partial_sums(partial, nblocks, C, S, Q);
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if ((threadIdx.x >> 5) == 0 && c < C) { dbeta[c] = (float)S; dgamma[c] = (float)Q; }
}

template <typename T, bool UP2>
__global__ void __launch_bounds__(kGateThreads, 2)
gate_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ d_raw, const T* __restrict__ x,
                      T* __restrict__ dx, const GateGeom g,
                      const int32_t* __restrict__ labels, const uint32_t* __restrict__ bits, int words,
                      const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ dgamma, const float* __restrict__ dbeta, int flags) {
  constexpr int V = Vec<T>::N;
  constexpr int U = UP2 ? 2 : 4;
  const int C = g.C, CG = C / V;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG, PL = blockDim.x / CG;
  const uint32_t TP = PL * U;
  const bool bn = flags & MCGM_GATE_BN, relu = flags & MCGM_GATE_RELU, use_mask = flags & MCGM_GATE_MASK;
  const float invP = 1.0f / (float)g.P;
  // dx = a*(g - dbeta/P - xhat*dgamma/P) = a*g + c0 + c1*x   (xhat = (x-mu)*rs), z = a*x + b
  float a[V], b[V], c0[V], c1[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    if (bn) {
      const float mu = mean[c], rs = rstd[c];
      a[i] = gamma[c] * rs;
      b[i] = fmaf(-mu, a[i], beta[c]);
      if (flags & MCGM_GATE_AFFINE) {            // ActNorm: (mean, rstd) are parameters, not batch statistics
        c0[i] = c1[i] = 0.0f;
      } else {
        c1[i] = -a[i] * rs * (dgamma[c] * invP);
        c0[i] = -a[i] * (dbeta[c] * invP) - c1[i] * mu;
      }
    } else {
      a[i] = 1.f; b[i] = 0.f; c0[i] = c1[i] = 0.f;
    }
  }
  const uint32_t tiles = (g.P + TP - 1) / TP;
  for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const uint32_t base = tile * TP + pl;
    uint4 xr[U], rr[U];
    DyRaw<T, UP2> dr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = base + u * PL;
      if (p < g.P) {
        xr[u] = ldg16(x + (size_t)p * C + cg * V);
        if (dy) dr[u].load(dy, g, p, g.hw.div(p), cg * V);
        if (d_raw) rr[u] = ldg16(d_raw + (size_t)p * C + cg * V);
      }
    }
    uint32_t m_tile = 0xffffffffu;
    if (use_mask && g.uniform_n) m_tile = mask_bits(bits, words, __ldg(labels + g.hw.div(tile * TP)), cg * V, V);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t p = base + u * PL;
      if (p >= g.P) continue;
      uint32_t m = m_tile;
      if (use_mask && !g.uniform_n) m = mask_bits(bits, words, __ldg(labels + g.hw.div(p)), cg * V, V);
      float xv[V], gv[V], rv[V];
      unpack16<T>(xr[u], xv);
      if (dy) {
        dr[u].get(gv);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) gv[i] = 0.0f;
      }
      if (d_raw) unpack16<T>(rr[u], rv);
      Vec<T> o;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float z = fmaf(a[i], xv[i], b[i]);
        const bool on = (m >> i) & 1u;
        float gi = on ? gv[i] : 0.0f;
        if (relu && !(z > 0.0f)) gi = 0.0f;
        float d = bn ? fmaf(a[i], gi, fmaf(c1[i], xv[i], c0[i])) : gi;
        if (d_raw && on) d += rv[i];
        o.v[i] = d;
      }
      o.store(dx + (size_t)p * C + cg * V);
    }
  }
}

// ---------------------------------------------------------------------------
// 2x2 pooling / un-pooling
// ---------------------------------------------------------------------------
template <typename T>
__global__ void pool2_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C,
                             float scale) {  // H,W are OUTPUT dims
  constexpr int V = Vec<T>::N;
  const int CG = C / V;
  const long total = (long)N * H * W * CG;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(e % CG);
    long p = e / CG;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    const T* base = x + (((long)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + cg * V;
    Vec<T> t0, t1, t2, t3, o;
    t0.load(base); t1.load(base + C); t2.load(base + (long)2 * W * C); t3.load(base + (long)2 * W * C + C);
#pragma unroll
    for (int i = 0; i < V; ++i) o.v[i] = scale * ((t0.v[i] + t1.v[i]) + (t2.v[i] + t3.v[i]));
    o.store(y + (e / CG) * C + cg * V);
  }
}

template <typename T>
__global__ void unpool2_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C,
                               float scale) {  // H,W are dims of dy (the small side)
  constexpr int V = Vec<T>::N;
  const int CG = C / V;
  const long total = (long)N * H * W * CG;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int cg = (int)(e % CG);
    long p = e / CG;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    Vec<T> v;
    v.load(dy + (e / CG) * C + cg * V);
#pragma unroll
    for (int i = 0; i < V; ++i) v.v[i] *= scale;
    T* base = dx + (((long)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + cg * V;
    v.store(base); v.store(base + C); v.store(base + (long)2 * W * C); v.store(base + (long)2 * W * C + C);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int check_cg(int dtype, int C, int* cg_out) {
  const int V = dtype == MCGM_BF16 ? 8 : 4;
  if (dtype != MCGM_BF16 && dtype != MCGM_F32) return fail(MCGM_E_ARG, "dtype %d unsupported", dtype);
  if (C <= 0 || C % V != 0) return fail(MCGM_E_ARG, "C=%d must be a positive multiple of %d", C, V);
  const int CG = C / V;
  if (CG > kGateThreads || kGateThreads % CG != 0)
    return fail(MCGM_E_UNSUPPORTED, "C=%d: C/%d must divide %d", C, V, kGateThreads);
  *cg_out = CG;
  return 0;
}

static int gate_grid(long P, int PL, int max_blocks) {
  long need = (P + PL - 1) / PL;
  if (need < 1) need = 1;
  return (int)(need < max_blocks ? need : max_blocks);
}

}  // namespace mcgm

using namespace mcgm;

extern "C" int mcgm_pack_codebook(const float* codebook, int M, int C, uint32_t* bits, void* stream) {
  MCGM_CHECK_ARG(codebook && bits && M > 0 && C > 0, "pack_codebook: bad arguments");
  const int words = (C + 31) / 32;
  const int total = M * words;
  pack_codebook_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(codebook, M, C, words, bits);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_mc_forward(const void* x, void* y, int dtype, int nchw, int N, int P, int C,
                               const int32_t* labels, const uint32_t* bits, int M, const float* code,
                               void* stream) {
  MCGM_CHECK_ARG(x && y && N >= 0 && P > 0 && C > 0, "mc_forward: bad arguments");
  MCGM_CHECK_ARG((labels && bits && M > 0) || code, "mc_forward: need (labels,bits) or code");
  if (N == 0) return 0;
  const int words = (C + 31) / 32;
  cudaStream_t st = (cudaStream_t)stream;
  const long total = (long)N * P * C;
  if (nchw && P > 1) {
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (dtype == MCGM_F32)
      mc_forward_nchw_kernel<float><<<(int)blocks, 256, 0, st>>>((const float*)x, (float*)y, total, P, C, labels, bits, words, code);
    else if (dtype == MCGM_BF16)
      mc_forward_nchw_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, total, P, C, labels, bits, words, code);
    else
      return fail(MCGM_E_ARG, "dtype %d unsupported", dtype);
    MCGM_LAUNCH_CHECK();
    return 0;
  }
  int CG;
  if (int e = check_cg(dtype, C, &CG)) return e;
  const int PL = kGateThreads / CG;
  const long rows = (long)N * P;
  const int grid = gate_grid(rows, PL, 148 * 8);
  if (dtype == MCGM_F32)
    mc_forward_nhwc_kernel<float><<<grid, kGateThreads, 0, st>>>((const float*)x, (float*)y, rows, P, C, labels, bits, words, code);
  else
    mc_forward_nhwc_kernel<__nv_bfloat16><<<grid, kGateThreads, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, rows, P, C, labels, bits, words, code);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mcgm_bn_stats_workspace_bytes(int C) {
  return (size_t)kMaxPartialBlocks * 2 * (size_t)(C > 0 ? C : 0) * sizeof(float);
}

extern "C" int mcgm_bn_stats(const void* x, int dtype, long P_total, int C, float* mean, float* rstd,
                             float* running_mean, float* running_var, float momentum, float eps,
                             void* workspace, void* stream) {
  MCGM_CHECK_ARG(x && mean && rstd && workspace && P_total > 0, "bn_stats: bad arguments");
  MCGM_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_stats: running stats must come in pairs");
  int CG;
  if (int e = check_cg(dtype, C, &CG)) return e;
  const int PL = kGateThreads / CG;
  const int grid = gate_grid((P_total + 3) / 4, PL, kMaxPartialBlocks);
  const size_t smem = (size_t)2 * PL * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  if (dtype == MCGM_F32)
    bn_partial_kernel<float><<<grid, kGateThreads, smem, st>>>((const float*)x, P_total, C, partial);
  else
    bn_partial_kernel<__nv_bfloat16><<<grid, kGateThreads, smem, st>>>((const __nv_bfloat16*)x, P_total, C, partial);
  MCGM_LAUNCH_CHECK();
  bn_finalize_kernel<<<(C + 31) / 32, 1024, 0, st>>>(partial, grid, C, (double)P_total, mean, rstd,
                                                           running_mean, running_var, momentum, eps);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_bn_finalize(const float* partial, int nblocks, int C, long count, float* mean, float* rstd,
                                float* running_mean, float* running_var, float momentum, float eps, void* stream) {
  MCGM_CHECK_ARG(partial && mean && rstd && nblocks > 0 && C > 0 && count > 0, "bn_finalize: bad arguments");
  MCGM_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats must come in pairs");
  bn_finalize_kernel<<<(C + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(partial, nblocks, C, (double)count, mean, rstd,
                                                                    running_mean, running_var, momentum, eps);
  MCGM_LAUNCH_CHECK();
  return 0;
}

static GateGeom make_geom(int N, int H, int W, int C, int TP) {
  GateGeom g;
  g.N = N; g.H = H; g.W = W; g.C = C;
  g.P = (uint32_t)((long)N * H * W);
  g.hw = PixDiv::make((uint32_t)(H * W));
  g.w = PixDiv::make((uint32_t)W);
  g.uniform_n = ((H * W) % TP) == 0;
  return g;
}

template <typename T>
static int gate_fwd_t(const void* x, void* y, void* y_raw, int N, int H, int W, int C, int CG,
                      const int32_t* labels, const uint32_t* bits, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, int flags, cudaStream_t st) {
  const int PL = kGateThreads / CG, words = (C + 31) / 32;
  const bool up = flags & MCGM_GATE_UP2;
  const int TP = PL * (up ? 4 : 8);
  const GateGeom g = make_geom(N, H, W, C, TP);
  const int grid = gate_grid(g.P, TP, 148 * 8);
  if (up)
    gate_fwd_kernel<T, true><<<grid, kGateThreads, 0, st>>>((const T*)x, (T*)y, (T*)y_raw, g, labels, bits, words, mean, rstd, gamma, beta, flags);
  else
    gate_fwd_kernel<T, false><<<grid, kGateThreads, 0, st>>>((const T*)x, (T*)y, (T*)y_raw, g, labels, bits, words, mean, rstd, gamma, beta, flags);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_gate_fwd(const void* x, void* y, void* y_raw, int dtype, int N, int H, int W, int C,
                             const int32_t* labels, const uint32_t* bits, int M, const float* mean,
                             const float* rstd, const float* gamma, const float* beta, int flags,
                             void* stream) {
  MCGM_CHECK_ARG(x && y && N >= 0 && H > 0 && W > 0, "gate_fwd: bad arguments");
  MCGM_CHECK_ARG((long)N * H * W < (1L << 31), "gate_fwd: more than 2^31 pixels");
  if (flags & MCGM_GATE_MASK) MCGM_CHECK_ARG(labels && bits && M > 0, "gate_fwd: MASK needs labels and bits");
  if (flags & MCGM_GATE_BN) MCGM_CHECK_ARG(mean && rstd && gamma && beta, "gate_fwd: BN needs mean/rstd/gamma/beta");
  if (N == 0) return 0;
  int CG;
  if (int e = check_cg(dtype, C, &CG)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MCGM_F32)
    return gate_fwd_t<float>(x, y, y_raw, N, H, W, C, CG, labels, bits, mean, rstd, gamma, beta, flags, st);
  return gate_fwd_t<__nv_bfloat16>(x, y, y_raw, N, H, W, C, CG, labels, bits, mean, rstd, gamma, beta, flags, st);
}

extern "C" size_t mcgm_gate_bwd_workspace_bytes(int C) {
  return (size_t)kMaxPartialBlocks * 2 * (size_t)(C > 0 ? C : 0) * sizeof(float);
}

template <typename T>
static int gate_bwd_t(const void* dy, const void* d_raw, const void* x, void* dx, int N, int H, int W, int C,
                      int CG, const int32_t* labels, const uint32_t* bits, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, float* dgamma, float* dbeta,
                      int flags, void* workspace, cudaStream_t st) {
  const int PL = kGateThreads / CG, words = (C + 31) / 32;
  const bool up = flags & MCGM_GATE_UP2;
  const int TP = PL * (up ? 2 : 4);
  const GateGeom g = make_geom(N, H, W, C, TP);
  if (flags & MCGM_GATE_BN) {
    const int grid = gate_grid(g.P, TP, kMaxPartialBlocks);
    const size_t smem = (size_t)2 * PL * C * sizeof(float);
    float* partial = (float*)workspace;
    if (up)
      gate_bwd_reduce_kernel<T, true><<<grid, kGateThreads, smem, st>>>((const T*)dy, (const T*)x, g, labels, bits, words, mean, rstd, gamma, beta, flags, partial);
    else
      gate_bwd_reduce_kernel<T, false><<<grid, kGateThreads, smem, st>>>((const T*)dy, (const T*)x, g, labels, bits, words, mean, rstd, gamma, beta, flags, partial);
    MCGM_LAUNCH_CHECK();
    gate_bwd_finalize_kernel<<<(C + 31) / 32, 1024, 0, st>>>(partial, grid, C, dgamma, dbeta);
    MCGM_LAUNCH_CHECK();
  }
  const int grid = gate_grid(g.P, TP, 148 * 8);
  if (up)
    gate_bwd_apply_kernel<T, true><<<grid, kGateThreads, 0, st>>>((const T*)dy, (const T*)d_raw, (const T*)x, (T*)dx, g, labels, bits, words, mean, rstd, gamma, beta, dgamma, dbeta, flags);
  else
    gate_bwd_apply_kernel<T, false><<<grid, kGateThreads, 0, st>>>((const T*)dy, (const T*)d_raw, (const T*)x, (T*)dx, g, labels, bits, words, mean, rstd, gamma, beta, dgamma, dbeta, flags);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_gate_bwd(const void* dy, const void* d_raw, const void* x, void* dx, int dtype, int N,
                             int H, int W, int C, const int32_t* labels, const uint32_t* bits, int M,
                             const float* mean, const float* rstd, const float* gamma, const float* beta,
                             float* dgamma, float* dbeta, int flags, void* workspace, void* stream) {
  MCGM_CHECK_ARG((dy || d_raw) && x && dx && N >= 0 && H > 0 && W > 0, "gate_bwd: bad arguments");
  MCGM_CHECK_ARG((long)N * H * W < (1L << 31), "gate_bwd: more than 2^31 pixels");
  if (flags & MCGM_GATE_MASK) MCGM_CHECK_ARG(labels && bits && M > 0, "gate_bwd: MASK needs labels and bits");
  if (flags & MCGM_GATE_BN)
    MCGM_CHECK_ARG(dy && mean && rstd && gamma && beta && dgamma && dbeta && workspace,
                   "gate_bwd: BN needs dy, mean/rstd/gamma/beta, dgamma/dbeta and a workspace");
  if (N == 0) return 0;
  int CG;
  if (int e = check_cg(dtype, C, &CG)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MCGM_F32)
    return gate_bwd_t<float>(dy, d_raw, x, dx, N, H, W, C, CG, labels, bits, mean, rstd, gamma, beta, dgamma, dbeta, flags, workspace, st);
  return gate_bwd_t<__nv_bfloat16>(dy, d_raw, x, dx, N, H, W, C, CG, labels, bits, mean, rstd, gamma, beta, dgamma, dbeta, flags, workspace, st);
}

extern "C" int mcgm_pool2(const void* x, void* y, int dtype, int N, int H, int W, int C, float scale,
                          void* stream) {
  MCGM_CHECK_ARG(x && y && N >= 0 && H > 0 && W > 0, "pool2: bad arguments (H,W are output dims)");
  if (N == 0) return 0;
  int CG;
  if (int e = check_cg(dtype, C, &CG)) return e;
  const long total = (long)N * H * W * CG;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MCGM_F32)
    pool2_kernel<float><<<(int)blocks, 256, 0, st>>>((const float*)x, (float*)y, N, H, W, C, scale);
  else
    pool2_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C, scale);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_unpool2(const void* dy, void* dx, int dtype, int N, int H, int W, int C, float scale,
                            void* stream) {
  MCGM_CHECK_ARG(dy && dx && N >= 0 && H > 0 && W > 0, "unpool2: bad arguments (H,W are dims of dy)");
  if (N == 0) return 0;
  int CG;
  if (int e = check_cg(dtype, C, &CG)) return e;
  const long total = (long)N * H * W * CG;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MCGM_F32)
    unpool2_kernel<float><<<(int)blocks, 256, 0, st>>>((const float*)dy, (float*)dx, N, H, W, C, scale);
  else
    unpool2_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, N, H, W, C, scale);
  MCGM_LAUNCH_CHECK();
  return 0;
}
