// thin.cu — convolutions whose one side has <= 4 channels: the RGB side of
// Conv2d(3,128,3,1,1) / Conv2d(3,128,1,1,0) (src/models/mcgan.py:77,87), of
// Conv2d(256,3,3,1,1) (src/models/mcgan.py:59) and of the first / last layers of MCVAE
// (src/models/mcvae.py:41,93).  Only the THIN tensor (fp32 NHWC [N,H,W,Ct], Ct <= 4) is
// im2col'd / col2im'd here; the arithmetic runs as 1x1 GEMMs on the tcgen05 conv kernels, which
// read the wide tensor exactly once in every direction.
#include "common.cuh"

namespace mcgm {

static constexpr int kMaxCt = 4;

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
  // column -> (row offset, column offset, channel) tables: no division in the hot loop.  toff is the
  // element offset of a column's source relative to the centre pixel: for interior pixels (every tap in
  // bounds — 88 % of a 32x32 image) a value is one table read and one load, no bounds arithmetic.
  __shared__ signed char tdr[128], tds[128], tj[128];
  __shared__ int toff[128];
  const int T = R * S;
  for (int col = threadIdx.x; col < KU; col += blockDim.x) {
    if (col < T * Ct) {
      const int t = col / Ct, r = t / S;
      tdr[col] = (signed char)(sgn * (r - pad));
      tds[col] = (signed char)(sgn * (t - r * S - pad));
      tj[col] = (signed char)(col - t * Ct);
      toff[col] = (sgn * (r - pad) * W + sgn * (t - r * S - pad)) * Ct + (col - t * Ct);
    } else {
      tj[col] = -1;
      toff[col] = 0;
    }
  }
  __syncthreads();
  const int dr_a = sgn * (0 - pad), dr_b = sgn * (R - 1 - pad), ds_a = sgn * (0 - pad), ds_b = sgn * (S - 1 - pad);
  const int dr_lo = dr_a < dr_b ? dr_a : dr_b, dr_hi = dr_a < dr_b ? dr_b : dr_a;
  const int ds_lo = ds_a < ds_b ? ds_a : ds_b, ds_hi = ds_a < ds_b ? ds_b : ds_a;
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
      const int h0 = (int)h * stride, w0 = (int)w * stride;
      if (h0 + dr_lo >= 0 && h0 + dr_hi < H && w0 + ds_lo >= 0 && w0 + ds_hi < W) {
        const float* base = thin + ((size_t)(n * H + h0) * W + w0) * Ct;
        const int nv = T * Ct - (int)gidx * V;               // live columns of this group
#pragma unroll
        for (int i = 0; i < V; ++i)
          if (i < nv) o.v[i] = __ldg(base + toff[gidx * V + i]);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const int col = gidx * V + i, j = tj[col];
          const int hh = h0 + tdr[col], ww = w0 + tds[col];
          if (j >= 0 && hh >= 0 && hh < H && ww >= 0 && ww < W)
            o.v[i] = __ldg(thin + ((size_t)(n * H + hh) * W + ww) * Ct + j);
        }
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


// Tiled col2im: a CTA owns 256 consecutive pixels of the flattened (n, h, w) order.  The T rows those pixels gather from —
// the tile plus W + 1 pixels on either side — are read ONCE, coalesced (16 bytes per thread), into shared memory with a
// row pitch of CT + 1 floats, so the per-pixel tap reads (one row per thread, same column across the warp) hit 32
// different banks.  The row-per-thread global reads of thin_col2im_kernel touch a 32-byte sector for every 4 bytes they
// use per instruction: 70 us for the 67 MB of T at 512 x 32 x 32, against 10.4 us for one pass at HBM speed.
static constexpr int kC2iTile = 256;
__global__ void __launch_bounds__(kC2iTile)
thin_col2im_tiled_kernel(const float* __restrict__ Tm, int CT, float* __restrict__ thin, long total, int H, int W, int Ct,
                         int R, int S, const float* __restrict__ bias, int flags) {
  extern __shared__ float srows[];
  const int pad = (R - 1) / 2, halo = pad * (W + 1), pitch = CT + 1, rows = kC2iTile + 2 * halo, q4 = CT >> 2;
  const long p0 = (long)blockIdx.x * kC2iTile;
  for (int i = threadIdx.x; i < rows * q4; i += kC2iTile) {
    const int r = i / q4, c = (i - r * q4) << 2;
    const long pr = p0 - halo + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pr >= 0 && pr < total) v = __ldg(reinterpret_cast<const float4*>(Tm + pr * CT + c));
    float* d = srows + r * pitch + c;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const long p = p0 + threadIdx.x;
  if (p >= total) return;
  const int w = (int)(p % W), h = (int)((p / W) % H);
  float acc[kMaxCt];
#pragma unroll
  for (int j = 0; j < kMaxCt; ++j) acc[j] = (j < Ct && (flags & MCGM_EPI_BIAS)) ? __ldg(bias + j) : 0.0f;
  for (int r = 0; r < R; ++r) {
    const int hh = h + r - pad;
    if (hh < 0 || hh >= H) continue;
    for (int s = 0; s < S; ++s) {
      const int ww = w + s - pad;
      if (ww < 0 || ww >= W) continue;
      const float* tp = srows + ((int)threadIdx.x + halo + (r - pad) * W + (s - pad)) * pitch + (r * S + s) * Ct;
#pragma unroll
      for (int j = 0; j < kMaxCt; ++j)
        if (j < Ct) acc[j] += tp[j];
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxCt; ++j)
    if (j < Ct) thin[p * Ct + j] = (flags & MCGM_EPI_TANH) ? tanhf(acc[j]) : acc[j];
}


}  // namespace mcgm

using namespace mcgm;

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
  const int pad = (R - 1) / 2;
  const size_t tiled_smem = (size_t)(kC2iTile + 2 * pad * (W + 1)) * (CT + 1) * sizeof(float);
  if (R == S && (R & 1) && CT % 4 == 0 && tiled_smem <= 48 * 1024 && ((uintptr_t)T & 15) == 0) {
    const long tiles = (total + kC2iTile - 1) / kC2iTile;
    MCGM_CHECK_ARG(tiles < (1L << 31), "thin_col2im: tensor too large");
    thin_col2im_tiled_kernel<<<(int)tiles, kC2iTile, tiled_smem, (cudaStream_t)stream>>>(T, CT, thin, total, H, W, Ct, R, S, bias, flags);
    MCGM_LAUNCH_CHECK();
    return 0;
  }
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  thin_col2im_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(T, CT, thin, N, H, W, Ct, R, S, bias, flags);
  MCGM_LAUNCH_CHECK();
  return 0;
}
