// thin_mma.cu — the RGB-side convolutions of the discriminator without an im2col buffer:
//   Conv2d(3,128,3,1,1) -> ReLU -> MC   and   Conv2d(3,128,1,1,0)      (src/models/mcgan.py:77-79, 87)
// forward and weight gradient.
//
// These two layers hold 0.5 % of the step's flops but every pass over them moves the 128-channel
// tensor (134 MB at batch 512): they are HBM-bound, K = R*S*Ct <= 32.  Feeding them to the tcgen05
// implicit-GEMM kernel meant materialising a 64-column bf16 im2col of the image (67 MB written,
// then read again by the GEMM and once more by the weight gradient) — 3.7 ms of a 30 ms iteration.
// Here the patch matrix lives only in shared memory: a block builds U[128 px][32] from the fp32 NHWC
// image (6 MB, L2 resident), multiplies with warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate:
// at K <= 32 the tensor work is ~1 % of the kernel's time, so the legacy warp MMA costs nothing; the
// tcgen05 path's 128 x N x 64 tile granularity is what wasted bandwidth here) and writes the wide tensor
// exactly once with full-line stores.  The weight gradient reads dY exactly once; its extra patch
// column of ones makes the bias gradient fall out of the same MMA.
#include <mma.h>

#include "common.cuh"
#include "reduce.cuh"

namespace mcgm {

static constexpr int kThinThreads = 256;
static constexpr int kThinK = 32;                 // patch columns (k = (r*S + s)*Ct + j), zero padded
static constexpr int kThinKP = 40;                // smem row pitch in halves (80 B: conflict-free ldmatrix)
static constexpr int kThinTileP = 128;            // pixels per forward tile
static constexpr int kThinWP = 64;                // pixels per weight-gradient tile
static constexpr int kWgStages = 3;               // cp.async ring depth of the weight-gradient kernel's dY tiles

struct ThinGeom {
  int N, H, W, Ct, R, S, pad, Cout;
  PixDiv dw, dhw;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A tile is `np` consecutive pixels of ONE image that form whole image rows (np % W == 0, H*W % np == 0 —
// checked on the host).  Its (np/W + 2 pad) x (W + 2 pad) x Ct fp32 halo window is staged in shared memory once
// (coalesced, zero outside the image); every patch value is then one shared-memory read.  ALL index arithmetic
// is tile-invariant per thread and is done once per block (ncu: with per-tile divisions and bounds tests the
// kernels were issue-bound on IMAD / ISETP, tensor and memory pipes idle).
static constexpr int kThinWinFloats = 1024;              // staged window (+ a 0.0 and a 1.0 slot behind it)
static constexpr int kThinWinIters = kThinWinFloats / kThinThreads;

struct PatchPlan {
  int src_rel[kThinWinIters];      // image offset of window element (thread, it) relative to the window's first row start
  int win_row[kThinWinIters];      // its window row, or -1: outside the image in w / beyond the window -> stays zero
  int idx16[16];                   // window indices of this thread's 16 patch columns (zero / one slots included)
  int pp;                          // patch row (pixel of the tile) this thread builds, -1: none
};

__device__ __forceinline__ void make_plan(const ThinGeom& g, int np, int ones_col, PatchPlan& pl) {
  const int WP = g.W + 2 * g.pad, rows = np / g.W + 2 * g.pad, rowf = WP * g.Ct, total = rows * rowf;
#pragma unroll
  for (int it = 0; it < kThinWinIters; ++it) {
    const int e = threadIdx.x + it * kThinThreads;
    pl.win_row[it] = -1;
    pl.src_rel[it] = 0;
    if (e < total) {
      const int r = e / rowf, c = e - r * rowf, ww = c / g.Ct - g.pad, j = c - (c / g.Ct) * g.Ct;
      if (ww >= 0 && ww < g.W) {
        pl.win_row[it] = r;
        pl.src_rel[it] = (r * g.W + ww) * g.Ct + j;
      }
    }
  }
  const int KT = g.R * g.S * g.Ct;
  const int e = threadIdx.x;
  pl.pp = e < np * 2 ? (e >> 1) : -1;
  const int pp = e >> 1, k0 = (e & 1) * 16, r = pp / g.W, w = pp - r * g.W;
  const int pbase = (r * WP + w) * g.Ct;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int k = k0 + i;
    if (k < KT) {
      const int t = k / g.Ct, tr = t / g.S, ts = t - tr * g.S, j = k - t * g.Ct;
      pl.idx16[i] = pbase + (tr * WP + ts) * g.Ct + j;
    } else {
      pl.idx16[i] = kThinWinFloats + (k == ones_col ? 1 : 0);
    }
  }
}

// window of the tile starting at pixel p0 -> registers (issued one tile ahead: the L2 latency of these loads was
// the top stall of both kernels), then registers -> shared memory at the start of the tile
__device__ __forceinline__ void fetch_window(const float* __restrict__ img, const ThinGeom& g, long p0, long P, int np,
                                             const PatchPlan& pl, float (&v)[kThinWinIters]) {
  const int HW = g.H * g.W;
  const uint32_t n = g.dhw.div((uint32_t)p0), h0 = g.dw.div((uint32_t)p0 - n * HW);
  const float* win0 = img + ((long)n * g.H + (long)h0 - g.pad) * g.W * g.Ct;      // first window row (may lie outside)
  const int r_lo = g.pad - (int)h0, r_hi = g.H + g.pad - (int)h0;                   // window rows inside the image: [r_lo, r_hi)
#pragma unroll
  for (int it = 0; it < kThinWinIters; ++it) {
    const int r = pl.win_row[it];
    v[it] = (p0 < P && r >= 0 && r >= r_lo && r < r_hi) ? __ldg(win0 + pl.src_rel[it]) : 0.0f;
  }
}
__device__ __forceinline__ void store_window(float* sImg, const float (&v)[kThinWinIters]) {
#pragma unroll
  for (int it = 0; it < kThinWinIters; ++it) sImg[threadIdx.x + it * kThinThreads] = v[it];
}

__device__ __forceinline__ void build_patches(const float* sImg, __nv_bfloat16* sU, const PatchPlan& pl) {
  if (pl.pp < 0) return;
  uint32_t pk[8];
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(sImg[pl.idx16[i]], sImg[pl.idx16[i + 1]]);
    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
  }
  uint4* dst = reinterpret_cast<uint4*>(sU + pl.pp * kThinKP + (threadIdx.x & 1) * 16);
  dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// ---- forward: out[p, co] = epilogue(sum_k U[p,k] Wm[co,k]) -------------------------------------------
// block = 8 warps; tile = 128 pixels x 128 output channels (loop over 128-channel slabs); warp w owns pixel
// rows [16w, 16w+16).  smem: sU [128][40], sW [128][40], sOut [8][16][136] (bf16).
template <bool GATE>
__global__ void __launch_bounds__(kThinThreads, 3)
thin_conv_fwd_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ Wm, int w_stride,
                     const float* __restrict__ bias, const int32_t* __restrict__ labels,
                     const uint32_t* __restrict__ bits, int words, int relu, __nv_bfloat16* __restrict__ out,
                     const ThinGeom g, long P, long tiles) {
  extern __shared__ __align__(16) uint8_t smem[];
  __nv_bfloat16* sU = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sW = sU + kThinTileP * kThinKP;
  __nv_bfloat16* sOut = sW + 128 * kThinKP;
  __shared__ float sImg[kThinWinFloats + 2];
  PatchPlan plan;
  make_plan(g, kThinTileP, -1, plan);
  if (threadIdx.x == 0) { sImg[kThinWinFloats] = 0.0f; sImg[kThinWinFloats + 1] = 1.0f; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* so = sOut + warp * 16 * 136;
  __shared__ float sBias[128];
  const bool single = g.Cout <= 128;                      // one weight slab: stage it (and the bias) once per block
  auto load_slab = [&](int c0) {
    for (int e = threadIdx.x; e < 128 * 4; e += blockDim.x) {         // weight slab: 128 rows x 32 halves
      const int r = e >> 2, q = e & 3;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (c0 + r < g.Cout) v = __ldg(reinterpret_cast<const uint4*>(Wm + (size_t)(c0 + r) * w_stride) + q);
      *reinterpret_cast<uint4*>(sW + r * kThinKP + q * 8) = v;
    }
    for (int e = threadIdx.x; e < 128; e += blockDim.x) sBias[e] = (bias != nullptr && c0 + e < g.Cout) ? __ldg(bias + c0 + e) : 0.0f;
  };
  if (single) load_slab(0);
  float win[kThinWinIters];
  fetch_window(img, g, (long)blockIdx.x * kThinTileP, P, kThinTileP, plan, win);
  for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long p0 = t * kThinTileP;
    int lab = 0;
    uint32_t mw0[4] = {~0u, ~0u, ~0u, ~0u};
    if constexpr (GATE) {                                // this tile's label and first codebook words: in flight early
      lab = __ldg(labels + g.dhw.div((uint32_t)p0));
#pragma unroll
      for (int q = 0; q < 4; ++q) mw0[q] = (32 * q < g.Cout) ? __ldg(bits + (size_t)lab * words + q) : 0u;
    }
    store_window(sImg, win);
    __syncthreads();                                     // window staged (and, first pass, the weight slab)
    build_patches(sImg, sU, plan);
    fetch_window(img, g, (t + gridDim.x) * kThinTileP, P, kThinTileP, plan, win);   // next tile's window: in flight during this tile
    for (int c0 = 0; c0 < g.Cout; c0 += 128) {
      if (!single) {
        __syncthreads();                                 // previous slab consumed
        load_slab(c0);
      }
      __syncthreads();                                   // sU (and the slab) ready
      // A fragments of this warp's 16 pixel rows, both k steps
      uint32_t a[2][4];
      {
        const int row = warp * 16 + (lane & 15), kofs = (lane >> 4) * 8;
        ldsm_x4(a[0], sU + row * kThinKP + kofs);
        ldsm_x4(a[1], sU + row * kThinKP + 16 + kofs);
      }
      const int r0 = lane >> 2, cq = (lane & 3) * 2;
      // a tile lies inside ONE image (checked on the host), so one label and one set of codebook words serve all
      // of its rows; they were requested at the top of the tile (mw0), long before this point
      uint32_t m0[4] = {~0u, ~0u, ~0u, ~0u};
      if constexpr (GATE) {
        if (c0 == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) m0[q] = mw0[q];
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) m0[q] = (c0 + 32 * q < g.Cout) ? __ldg(bits + (size_t)lab * words + (c0 >> 5) + q) : 0u;
        }
      }
      // two halves of 8 n-tiles: 32 accumulator registers live instead of 64 (three blocks per SM); unrolled, so
      // the codebook words are indexed statically
#pragma unroll
      for (int nh = 0; nh < 16; nh += 8) {
        float acc[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          uint32_t b[4];                                 // rows = 8 output channels, 4 k-octets
          ldsm_x4(b, sW + ((nh + n) * 8 + (lane & 7)) * kThinKP + (lane >> 3) * 8);
          mma_bf16(acc[n], a[0], b[0], b[1]);
          mma_bf16(acc[n], a[1], b[2], b[3]);
        }
        // epilogue on the fragment layout: thread holds rows r0 = lane/4 and r0 + 8, columns n*8 + (lane%4)*2 + {0,1}
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const int col = (nh + n) * 8 + cq;
          const float2 bb = *reinterpret_cast<const float2*>(sBias + col);
          float v00 = acc[n][0] + bb.x, v01 = acc[n][1] + bb.y, v10 = acc[n][2] + bb.x, v11 = acc[n][3] + bb.y;
          if (relu) { v00 = fmaxf(v00, 0.f); v01 = fmaxf(v01, 0.f); v10 = fmaxf(v10, 0.f); v11 = fmaxf(v11, 0.f); }
          if constexpr (GATE) {
            const uint32_t wa = m0[(nh + n) >> 2] >> (col & 31);
            v00 = (wa & 1u) ? v00 : 0.f; v01 = (wa & 2u) ? v01 : 0.f;
            v10 = (wa & 1u) ? v10 : 0.f; v11 = (wa & 2u) ? v11 : 0.f;
          }
          *reinterpret_cast<__nv_bfloat162*>(so + r0 * 136 + col) = __floats2bfloat162_rn(v00, v01);
          *reinterpret_cast<__nv_bfloat162*>(so + (r0 + 8) * 136 + col) = __floats2bfloat162_rn(v10, v11);
        }
      }
      __syncwarp();
      // 16 rows x 256 B: one instruction writes two full rows
      {
        const int row0 = lane >> 4, c16 = lane & 15;
        __nv_bfloat16* op = out + (p0 + warp * 16 + row0) * g.Cout + c0 + c16 * 8;
        const __nv_bfloat16* sp = so + row0 * 136 + c16 * 8;
        if (p0 + kThinTileP <= P && c0 + 128 <= g.Cout) {              // full tile: no per-row predicates
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(op + (size_t)i * 2 * g.Cout) = *reinterpret_cast<const uint4*>(sp + i * 2 * 136);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (p0 + warp * 16 + row0 + 2 * i < P && c0 + c16 * 8 < g.Cout)
              *reinterpret_cast<uint4*>(op + (size_t)i * 2 * g.Cout) = *reinterpret_cast<const uint4*>(sp + i * 2 * 136);
        }
      }
      __syncwarp();
    }
    __syncthreads();                                     // before the next tile overwrites sU
  }
}

// ---- weight gradient: D[co][k] = sum_p dY[p,co] U[p,k]; column `ones_col` of U is 1 -> D[co][ones_col] = dbias
// block = 8 warps, warp w owns output channels [16w, 16w+16) of the current 128-channel slab; tile = 64 pixels.
// partial[block][Cout][32] fp32, reduced over blocks by partial_reduce (fixed order).
__global__ void __launch_bounds__(kThinThreads, 3)
thin_conv_wgrad_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ dy, float* __restrict__ partial,
                       const ThinGeom g, long P, long tiles, int ones_col) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int DYP = 136;                              // pitch of the dY tile in halves (272 B)
  __nv_bfloat16* sU = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sDY = sU + kThinWP * kThinKP;
  __shared__ float sImg[kThinWinFloats + 2];
  PatchPlan plan;
  make_plan(g, kThinWP, ones_col, plan);
  if (threadIdx.x == 0) { sImg[kThinWinFloats] = 0.0f; sImg[kThinWinFloats + 1] = 1.0f; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < g.Cout; c0 += 128) {
    float acc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.0f;
    // dY tiles (64 rows x 16 units of 16 B, 4 units per thread) stream through a kWgStages-deep cp.async ring:
    // with one tile in flight per block the kernel sat at ~1.4 TB/s (bytes in flight = a fifth of the HBM
    // latency-bandwidth product); the ring keeps kWgStages - 1 tiles (48 KB per block) outstanding.
    auto issue_dy = [&](long t, int stage) {
      const long p0 = t * kThinWP;
      __nv_bfloat16* dst = sDY + stage * kThinWP * DYP;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = threadIdx.x + i * kThinThreads, pp = e >> 4, q = e & 15;
        const bool ok = t < tiles && p0 + pp < P && c0 + q * 8 < g.Cout;
        const void* src = ok ? (const void*)(reinterpret_cast<const uint4*>(dy + (p0 + pp) * g.Cout + c0) + q) : (const void*)dy;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + pp * DYP + q * 8);
        const int nbytes = ok ? 16 : 0;                  // src-size 0: the 16 bytes are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(nbytes) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    long tn = blockIdx.x;
#pragma unroll
    for (int s_ = 0; s_ < kWgStages - 1; ++s_, tn += gridDim.x) issue_dy(tn, s_);
    float win[kThinWinIters];
    fetch_window(img, g, (long)blockIdx.x * kThinWP, P, kThinWP, plan, win);
    int stage = 0;
    for (long t = blockIdx.x; t < tiles; t += gridDim.x, tn += gridDim.x) {
      __syncthreads();                                   // previous tile's sU and ring slot consumed
      issue_dy(tn, (stage + kWgStages - 1) % kWgStages); // refill the slot freed by the previous tile
      store_window(sImg, win);
      __syncthreads();
      build_patches(sImg, sU, plan);
      fetch_window(img, g, (t + gridDim.x) * kThinWP, P, kThinWP, plan, win);       // next tile's window
      asm volatile("cp.async.wait_group %0;" ::"n"(kWgStages - 1) : "memory");
      __syncthreads();
      const __nv_bfloat16* cur = sDY + stage * kThinWP * DYP;
#pragma unroll
      for (int ks = 0; ks < kThinWP / 16; ++ks) {
        // A = dY^T: m = output channel, k = pixel.  sDY is [pixel][channel] -> transposed 8x8 loads.
        uint32_t a[4];
        {
          const int prow = ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;        // matrices 0,1: k 0-7; 2,3: k 8-15
          const int ccol = warp * 16 + ((lane >> 3) & 1) * 8;                   // matrices 0,2: m 0-7; 1,3: m 8-15
          ldsm_x4_t(a, cur + prow * DYP + ccol);
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          uint32_t b[2];                                 // B[k = pixel][n = patch column]: sU is [pixel][column] -> transposed
          const int prow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x2_t(b, sU + prow * kThinKP + n * 8);
          mma_bf16(acc[n], a, b[0], b[1]);
        }
      }
      stage = (stage + 1) % kWgStages;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // fragment -> partial[block][c0 + 16*warp + row][n*8 + col]
    const int r0 = lane >> 2, cq = (lane & 3) * 2;
    float* pb = partial + ((size_t)blockIdx.x * g.Cout + c0 + warp * 16) * kThinK;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      if (c0 + warp * 16 + r0 < g.Cout)
        *reinterpret_cast<float2*>(pb + (size_t)r0 * kThinK + n * 8 + cq) = make_float2(acc[n][0], acc[n][1]);
      if (c0 + warp * 16 + r0 + 8 < g.Cout)
        *reinterpret_cast<float2*>(pb + (size_t)(r0 + 8) * kThinK + n * 8 + cq) = make_float2(acc[n][2], acc[n][3]);
    }
  }
}

static constexpr int kThinRedLanes = 16;
__global__ void __launch_bounds__(32 * kThinRedLanes)
thin_partial_reduce_kernel(const float* __restrict__ partial, int nblocks, size_t n, float* __restrict__ out) {
  __shared__ float sm1[kThinRedLanes][32];
  for (size_t i0 = (size_t)blockIdx.x * 32; i0 < n; i0 += (size_t)gridDim.x * 32)
    reduce_columns<float, kThinRedLanes>(partial, nblocks, n, n, i0, out, sm1);
}

static int thin_geom(ThinGeom* g, int N, int H, int W, int Ct, int R, int S, int Cout, int tile_px) {
  if (!(Ct >= 1 && Ct <= 4 && R == S && (R == 1 || R == 3) && R * S * Ct <= kThinK - 1))
    return fail(MCGM_E_UNSUPPORTED, "thin_conv: Ct=%d R=%d S=%d unsupported (R = S in {1,3}, R*S*Ct <= %d)", Ct, R, S, kThinK - 1);
  if (Cout % 8 || (long)N * H * W >= (1L << 31)) return fail(MCGM_E_UNSUPPORTED, "thin_conv: Cout=%d must be a multiple of 8, < 2^31 pixels", Cout);
  if (tile_px % W || (H * W) % tile_px || (tile_px / W + R - 1) * (W + R - 1) * Ct > kThinWinFloats)
    return fail(MCGM_E_UNSUPPORTED, "thin_conv: %d-pixel tiles must be whole rows of one %dx%d image", tile_px, H, W);
  g->N = N; g->H = H; g->W = W; g->Ct = Ct; g->R = R; g->S = S; g->pad = (R - 1) / 2; g->Cout = Cout;
  g->dw = PixDiv::make((uint32_t)W);
  g->dhw = PixDiv::make((uint32_t)(H * W));
  return 0;
}

// thin_umma.cu: the same forward on tcgen05 + TMA (0 launched, 1 not applicable, < 0 error)
int thin_conv_fwd_umma(const float* img, const void* w_bf16, const float* bias, const int32_t* labels, const uint32_t* bits,
                       int relu, void* out_bf16, int N, int H, int W, int Ct, int R, int S, int Cout, cudaStream_t st);

// ... and the weight gradient (-> number of CTA partials written, 0 = not applicable, < 0 error)
int thin_conv_wgrad_umma(const float* img, const void* dy_bf16, float* partial, int max_partials, int N, int H, int W, int Ct,
                         int R, int S, int Cout, cudaStream_t st);

static constexpr int kThinFwdSmem = (kThinTileP * kThinKP + 128 * kThinKP + 8 * 16 * 136) * 2;
static constexpr int kThinWgSmem = (kThinWP * kThinKP + kWgStages * kThinWP * 136) * 2;
static constexpr int kThinWgBlocks = 444;            // 3 per SM (registers): latency is hidden by occupancy + the register prefetch

}  // namespace mcgm

using namespace mcgm;

namespace mcgm { int g_thin_umma_wgrad = 1, g_thin_umma_fwd = 1; }      // developer A/B knobs (mcgm_debug_set 12 / 13)

extern "C" int mcgm_thin_conv_fwd(const float* img, const void* w_bf16, int w_row_stride, const float* bias,
                                  const int32_t* labels, const uint32_t* bits, int M, int relu, void* out_bf16, int N,
                                  int H, int W, int Ct, int R, int S, int Cout, void* stream) {
  MCGM_CHECK_ARG(img && w_bf16 && out_bf16 && N > 0 && H > 0 && W > 0, "thin_conv_fwd: bad arguments");
  MCGM_CHECK_ARG(w_row_stride >= kThinK && w_row_stride % 8 == 0, "thin_conv_fwd: weight rows must hold >= 32 bf16 columns, 16-byte aligned");
  MCGM_CHECK_ARG((labels == nullptr) == (bits == nullptr) && (!bits || (M > 0 && Cout % 32 == 0)), "thin_conv_fwd: MASK needs labels, bits and Cout %% 32 == 0");
  // weights padded to 64 columns (128-byte rows = the UMMA operand layout): the tcgen05 / TMA kernel of thin_umma.cu;
  // 32-column weights or shapes it does not cover: the mma.sync kernel below
  if (w_row_stride == 64 && g_thin_umma_fwd) {
    const int r = thin_conv_fwd_umma(img, w_bf16, bias, labels, bits, relu, out_bf16, N, H, W, Ct, R, S, Cout, (cudaStream_t)stream);
    if (r <= 0) return r;
  }
  ThinGeom g;
  if (int e = thin_geom(&g, N, H, W, Ct, R, S, Cout, kThinTileP)) return e;
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(thin_conv_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThinFwdSmem));
    MCGM_CUDA(cudaFuncSetAttribute(thin_conv_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kThinFwdSmem));
    configured = true;
  }
  const long P = (long)N * H * W, tiles = (P + kThinTileP - 1) / kThinTileP;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const long grid = tiles < (long)sms * 3 ? tiles : (long)sms * 3;      // three resident blocks per SM
  if (bits != nullptr)
    thin_conv_fwd_kernel<true><<<(int)grid, kThinThreads, kThinFwdSmem, (cudaStream_t)stream>>>(
        img, (const __nv_bfloat16*)w_bf16, w_row_stride, bias, labels, bits, (Cout + 31) / 32, relu, (__nv_bfloat16*)out_bf16, g, P, tiles);
  else
    thin_conv_fwd_kernel<false><<<(int)grid, kThinThreads, kThinFwdSmem, (cudaStream_t)stream>>>(
        img, (const __nv_bfloat16*)w_bf16, w_row_stride, bias, labels, bits, (Cout + 31) / 32, relu, (__nv_bfloat16*)out_bf16, g, P, tiles);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t mcgm_thin_conv_wgrad_workspace_bytes(int Cout) { return (size_t)kThinWgBlocks * Cout * kThinK * sizeof(float); }

extern "C" int mcgm_thin_conv_wgrad(const float* img, const void* dy_bf16, float* dwm, int N, int H, int W, int Ct, int R,
                                    int S, int Cout, void* workspace, void* stream) {
  MCGM_CHECK_ARG(img && dy_bf16 && dwm && workspace && N > 0 && H > 0 && W > 0, "thin_conv_wgrad: bad arguments");
  ThinGeom g;
  if (int e = thin_geom(&g, N, H, W, Ct, R, S, Cout, kThinWP)) return e;
  if (Cout % 16) return fail(MCGM_E_UNSUPPORTED, "thin_conv_wgrad: Cout=%d must be a multiple of 16", Cout);
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(thin_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kThinWgSmem));
    configured = true;
  }
  const long P = (long)N * H * W, tiles = (P + kThinWP - 1) / kThinWP;
  cudaStream_t st = (cudaStream_t)stream;
  // tcgen05 / TMA kernel (thin_umma.cu) where it applies; its CTA partials use the same [blocks][Cout][32] workspace
  int grid = g_thin_umma_wgrad ? thin_conv_wgrad_umma(img, dy_bf16, (float*)workspace, kThinWgBlocks, N, H, W, Ct, R, S, Cout, st) : 0;
  if (grid < 0) return grid;
  if (grid == 0) {
    grid = (int)(tiles < kThinWgBlocks ? tiles : kThinWgBlocks);
    thin_conv_wgrad_kernel<<<grid, kThinThreads, kThinWgSmem, st>>>(img, (const __nv_bfloat16*)dy_bf16, (float*)workspace, g, P, tiles,
                                                                   R * S * Ct);
    MCGM_LAUNCH_CHECK();
  }
  const size_t nout = (size_t)Cout * kThinK;
  thin_partial_reduce_kernel<<<(int)((nout + 31) / 32), 32 * kThinRedLanes, 0, st>>>((const float*)workspace, grid, nout, dwm);
  MCGM_LAUNCH_CHECK();
  return 0;
}
