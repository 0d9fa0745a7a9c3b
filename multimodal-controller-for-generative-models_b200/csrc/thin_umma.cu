// thin_umma.cu — the RGB-side forward convolutions of the discriminator on the Blackwell-native data path:
//   Conv2d(3,128,3,1,1) -> ReLU -> MC   and   Conv2d(3,128,1,1,0)      (src/models/mcgan.py:77-79, 87)
//
// These layers are HBM-bound (K = R*S*Ct <= 32: the 128-channel output, 134 MB at batch 512, is all that moves) and the
// mma.sync kernel of thin_mma.cu was bound by instruction issue instead (ncu r01: issue slots 66 %, 2.1 TB/s = 33 % of
// HBM): every warp built patches, ran warp MMAs and copied the tile out through registers.  Here the work is split over
// specialised warps and the tile movement is done by the copy engines:
//
//   warp 0   TMA producer   the fp32 halo window of a 128-pixel tile — (128/W + R-1) image rows x W*Ct floats — is ONE
//                           3-D TMA box; rows above / below the image are zero-filled by the TMA unit (= the padding)
//   warps 2-5 patch builders thread = pixel: reads its <= 27 window values from shared memory, converts to bf16 and writes
//                           its row of the patch matrix U[128][32] in the K-major 128B-swizzled UMMA layout
//   warp 1   MMA issuer     D[128 px, Cout] = U x Wt^T on tcgen05 (two K = 16 instructions), accumulator in TMEM
//                           (double-buffered); the weights Wt[Cout][64] (zero-padded columns) are loaded once per CTA
//   warps 6-9 epilogue      TMEM -> registers -> bias / ReLU / MC mask -> bf16 -> 128B-swizzled staging tile in shared
//                           memory -> cp.async.bulk.tensor store (128 pixels x 64 channels per store)
//
// Everything is double-buffered (window, patch matrix, accumulator, staging tile); the CTA is persistent over tiles.
// Algorithmic bytes: P*Ct*4 read + P*Cout*2 written.
#include "common.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace mcgm {

static constexpr int kTuThreads = 320;
static constexpr int kTuTile = 128;                       // pixels per tile = UMMA M
static constexpr int kTuABytes = kTuTile * 128;           // patch matrix: 128 rows of 128 B (64 bf16 columns, 32 used)
static constexpr int kTuWinBytes = 8192;                  // halo window buffer (floats: (128/W + 2) * W * Ct <= 1536 for W >= 8)
static constexpr int kTuSlabBytes = kTuTile * 128;        // staging tile: 128 pixels x 64 channels bf16
static constexpr int kTuMaxCout = 256;

struct ThinUParams {
  int N, H, W, Ct, R, pad, Cout;
  int TH;                        // image rows per tile (128 / W)
  int win_rows, row_floats;      // window: (TH + R - 1) rows of W*Ct floats
  int tiles, tiles_per_img;
  int relu, words;
  uint64_t desc_templ;
  uint32_t idesc;
  const float* bias;
  const int32_t* labels;
  const uint32_t* bits;
};

__device__ __forceinline__ void tu_sts16(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t tu_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Two CTAs per SM (<= 96 registers, <= 113 KB shared memory, 2 x Cout <= 256 TMEM columns each): a tile is a serial
// chain window -> patches -> MMA -> epilogue, and one CTA's double buffering alone leaves the copy engines idle
// between tiles; the second resident CTA fills those gaps.
__global__ void __launch_bounds__(kTuThreads, 2)
thin_conv_fwd_umma_kernel(const __grid_constant__ CUtensorMap tmap_img, const __grid_constant__ CUtensorMap tmap_w,
                          const __grid_constant__ CUtensorMap tmap_out, const ThinUParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  // layout: A[2] | B (Cout x 128 B) | staging[2] | window[2] | barriers | bias
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + 2 * kTuABytes;
  const uint32_t s_base = b_base + (uint32_t)p.Cout * 128u;
  const uint32_t w_base = s_base + 2 * kTuSlabBytes;
  const uint32_t bar = w_base + 2 * kTuWinBytes;
  auto h_full = [&](int b) { return bar + 8u * b; };
  auto h_empty = [&](int b) { return bar + 8u * (2 + b); };
  auto a_full = [&](int b) { return bar + 8u * (4 + b); };
  auto a_empty = [&](int b) { return bar + 8u * (6 + b); };
  auto t_full = [&](int b) { return bar + 8u * (8 + b); };
  auto t_empty = [&](int b) { return bar + 8u * (10 + b); };
  const uint32_t w_full = bar + 8u * 12;
  const uint32_t tmem_slot = bar + 8u * 13;
  float* sbias = reinterpret_cast<float*>(smem_raw + (bar - base) + 128);
  uint32_t* smask = reinterpret_cast<uint32_t*>(sbias + kTuMaxCout);       // per tile: AND-mask of every bf16 channel PAIR
  const float* swin = reinterpret_cast<const float*>(smem_raw + (w_base - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.Cout; i += kTuThreads) sbias[i] = p.bias ? __ldg(p.bias + i) : 0.0f;
  // the patch matrix's unused columns (32..63 of every row) are never read by the K = 32 MMAs, but keep them defined
  for (int i = threadIdx.x; i < 2 * kTuABytes / 16; i += kTuThreads) tu_sts16(a_base + 16u * i, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_img);
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_out);
    for (int b = 0; b < 2; ++b) {
      mbar_init(h_full(b), 1); mbar_init(h_empty(b), 4);
      mbar_init(a_full(b), 4); mbar_init(a_empty(b), 1);
      mbar_init(t_full(b), 1); mbar_init(t_empty(b), 128);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  const int tmem_cols = 2 * p.Cout <= 32 ? 32 : 2 * p.Cout <= 64 ? 64 : 2 * p.Cout <= 128 ? 128 : 2 * p.Cout <= 256 ? 256 : 512;
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer: weights once, then one halo window per tile =====
      mbar_arrive_expect_tx(w_full, (uint32_t)p.Cout * 128u);
      tma_load_2d(b_base, &tmap_w, w_full, 0, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int b = it & 1, ph = (it >> 1) & 1;
        const int n = t / p.tiles_per_img, h0 = (t - n * p.tiles_per_img) * p.TH;
        mbar_wait(h_empty(b), ph ^ 1);
        mbar_arrive_expect_tx(h_full(b), (uint32_t)(p.win_rows * p.row_floats * 4));
        tma_load_3d(w_base + b * kTuWinBytes, &tmap_img, h_full(b), 0, h0 - p.pad, n);    // rows outside the image: zeros
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== MMA issuer =====
      mbar_wait(w_full, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int b = it & 1, ph = (it >> 1) & 1;
        mbar_wait(a_full(b), ph);
        mbar_wait(t_empty(b), ph ^ 1);
        tc_fence_after();
        const uint64_t adesc = smem_desc(p.desc_templ, a_base + b * kTuABytes);
        const uint64_t bdesc = smem_desc(p.desc_templ, b_base);
        const uint32_t d = tmem_base + b * p.Cout;
        umma_f16(d, adesc, bdesc, p.idesc, 0u);
        umma_f16(d, adesc + 2, bdesc + 2, p.idesc, 1u);
        umma_commit(a_empty(b));
        umma_commit(t_full(b));
      }
    }
  } else if (warp < 6) {
    // ===== patch builders: thread = pixel of the tile =====
    const int m = (warp - 2) * 32 + lane;
    const int th = m / p.W, tw = m - th * p.W;
    const int KT = p.R * p.R * p.Ct;
    const int sw = m & 7;
    // window offsets of this pixel's patch columns: tile-invariant, computed once (per-tile divisions made the mma.sync
    // kernel issue-bound); -1 = zero (outside the image in w, or a padding column)
    int off[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      off[k] = -1;
      if (k < KT) {
        const int tap = k / p.Ct, j = k - tap * p.Ct, tr = tap / p.R, ts = tap - tr * p.R;
        const int col = tw + ts - p.pad;
        if (col >= 0 && col < p.W) off[k] = (th + tr) * p.row_floats + col * p.Ct + j;
      }
    }
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const int b = it & 1, ph = (it >> 1) & 1;
      mbar_wait(h_full(b), ph);
      const float* win = swin + b * (kTuWinBytes / 4);
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = off[k] >= 0 ? win[off[k]] : 0.0f;
      mbar_wait(a_empty(b), ph ^ 1);                       // the MMAs of two tiles ago have read this patch buffer
      const uint32_t row = a_base + b * kTuABytes + m * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tu_sts16(row + (((uint32_t)c ^ sw) << 4), tu_pack(v[8 * c], v[8 * c + 1]), tu_pack(v[8 * c + 2], v[8 * c + 3]),
                 tu_pack(v[8 * c + 4], v[8 * c + 5]), tu_pack(v[8 * c + 6], v[8 * c + 7]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(a_full(b)); mbar_arrive(h_empty(b)); }
    }
  } else {
    // ===== epilogue: warps 6..9 own TMEM lane quadrants 2, 3, 0, 1 =====
    // The epilogue is what this kernel's time is made of (ncu source view: its instructions, not a stall, carry the
    // samples), so it is kept to ~3 instructions per output value: bias add, branch-free ReLU (max with 0 or -inf),
    // pack, and ONE AND per channel pair for the MC mask — a 128-pixel tile lies inside one image, so the 0xFFFF /
    // 0xFFFF0000 masks of all channel pairs are tile-uniform: built once per tile in shared memory, read by broadcast.
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int et = (warp - 6) * 32 + lane;                  // thread index inside the epilogue group
    const int sw = m & 7;
    const bool leader_warp = (warp == 6);
    const int slabs = p.Cout / 64;
    const float lo = p.relu ? 0.0f : -INFINITY;
    const bool masked = p.bits != nullptr;
    int it = 0, sidx = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const int b = it & 1, ph = (it >> 1) & 1;
      if (masked && et < p.Cout / 2) {                        // (the previous tile's readers are past their last barrier)
        const int lab = __ldg(p.labels + t / p.tiles_per_img);
        const uint32_t w = __ldg(p.bits + (size_t)lab * p.words + (et >> 4)) >> ((2 * et) & 31);
        smask[et] = ((w & 1u) ? 0x0000ffffu : 0u) | ((w & 2u) ? 0xffff0000u : 0u);
      }
      mbar_wait(t_full(b), ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + b * p.Cout;
      for (int s = 0; s < slabs; ++s, ++sidx) {
        uint32_t acc[2][32];
        tmem_ld32(t_row + s * 64, acc[0]);
        tmem_ld32(t_row + s * 64 + 32, acc[1]);
        const uint32_t buf = s_base + (sidx & 1) * kTuSlabBytes;
        if (leader_warp && elect_one()) tma_store_wait_read_1();               // the store that used this staging buffer (two slabs ago) is done
        named_bar_sync(1, 128);                              // ... and this tile's mask table is complete
        tmem_ld_wait();
        if (s == slabs - 1) {
          tc_fence_before();
          mbar_arrive(t_empty(b));
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {                  // 4 columns per step: one 16-byte bias read
            const float4 b4 = *reinterpret_cast<const float4*>(sbias + s * 64 + 32 * k + 2 * i);
            o[i] = tu_pack(fmaxf(__uint_as_float(acc[k][2 * i]) + b4.x, lo), fmaxf(__uint_as_float(acc[k][2 * i + 1]) + b4.y, lo));
            o[i + 1] = tu_pack(fmaxf(__uint_as_float(acc[k][2 * i + 2]) + b4.z, lo), fmaxf(__uint_as_float(acc[k][2 * i + 3]) + b4.w, lo));
          }
          if (masked) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const uint4 m4 = *reinterpret_cast<const uint4*>(smask + s * 32 + 16 * k + i);
              o[i] &= m4.x; o[i + 1] &= m4.y; o[i + 2] &= m4.z; o[i + 3] &= m4.w;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            tu_sts16(buf + m * 128 + ((((uint32_t)(4 * k + u)) ^ sw) << 4), o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (leader_warp && elect_one()) {
          tma_store_2d(&tmap_out, buf, s * 64, t * kTuTile);
          tma_store_commit();
        }
      }
    }
    if (leader_warp && elect_one()) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

static int tu_fwd_smem_bytes(int Cout) { return 2 * kTuABytes + Cout * 128 + 2 * kTuSlabBytes + 2 * kTuWinBytes + 128 + kTuMaxCout * 4 + kTuMaxCout * 2; }

// -> 0 launched, 1 not applicable (the caller falls back to the mma.sync kernel), < 0 error
int thin_conv_fwd_umma(const float* img, const void* w_bf16, const float* bias, const int32_t* labels, const uint32_t* bits,
                       int relu, void* out_bf16, int N, int H, int W, int Ct, int R, int S, int Cout, cudaStream_t st) {
  const int K = R * S * Ct;
  if (!(R == S && (R == 1 || R == 3) && K <= 32 && Cout % 64 == 0 && Cout <= kTuMaxCout && W >= 8 && W % 4 == 0 &&
        kTuTile % W == 0 && (H * W) % kTuTile == 0 && W * Ct <= 256))
    return 1;
  ThinUParams p{};
  p.N = N; p.H = H; p.W = W; p.Ct = Ct; p.R = R; p.pad = (R - 1) / 2; p.Cout = Cout;
  p.TH = kTuTile / W;
  p.win_rows = p.TH + R - 1;
  p.row_floats = W * Ct;
  if (p.win_rows * p.row_floats * 4 > kTuWinBytes || p.win_rows > 256) return 1;
  p.tiles_per_img = H * W / kTuTile;
  p.tiles = N * p.tiles_per_img;
  p.relu = relu; p.words = (Cout + 31) / 32;
  p.bias = bias; p.labels = labels; p.bits = bits;
  p.desc_templ = make_smem_desc_hi_lbo(16, 1024, 1, 2);
  p.idesc = make_idesc_bf16(kTuTile, Cout, 0, 0, 0);
  CUtensorMap mi, mw, mo;
  {
    const uint64_t dims[3] = {(uint64_t)W * Ct, (uint64_t)H, (uint64_t)N};
    const uint32_t box[3] = {(uint32_t)(W * Ct), (uint32_t)p.win_rows, 1u};
    if (int e = make_tmap(&mi, img, 3, dims, box, 1, 1, 0)) return e;
  }
  {
    const uint64_t dims[2] = {64u, (uint64_t)Cout};
    const uint32_t box[2] = {64u, (uint32_t)Cout};
    if (int e = make_tmap(&mw, w_bf16, 2, dims, box, 1, 0, 2)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)N * H * W};
    const uint32_t box[2] = {64u, (uint32_t)kTuTile};
    if (int e = make_tmap(&mo, out_bf16, 2, dims, box, 1, 0, 2)) return e;
  }
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(thin_conv_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tu_fwd_smem_bytes(kTuMaxCout)));
    // the whole L1 / shared array as shared memory: two ~100 KB CTAs must be co-resident
    MCGM_CUDA(cudaFuncSetAttribute(thin_conv_fwd_umma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int resident = Cout <= 128 ? 2 : 1;                 // 2 x Cout TMEM columns and ~100 KB of shared memory per CTA
  const int grid = p.tiles < resident * sms ? p.tiles : resident * sms;
  thin_conv_fwd_umma_kernel<<<grid, kTuThreads, tu_fwd_smem_bytes(Cout), st>>>(mi, mw, mo, p);
  MCGM_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient of the same layers:  dWm[co][k] = sum_p dY[p, co] * U[p, k]   (k = (r*S+s)*Ct + j; column R*S*Ct of U is
// all ones, so the bias gradient falls out of the same product).  dY (bf16 NHWC, 134 MB at batch 512) is read exactly once:
//   warp 0   TMA producer   per 128-pixel tile: the halo window (3-D box) and the dY tile as 64-pixel x 64-channel boxes
//                           (128B swizzle) = MN-major UMMA operand straight from the NHWC tensor, as in conv_wgrad_kernel
//   warps 2-5 patch builders as in the forward kernel; U[128 px][64] (32 live columns) is the MN-major B operand
//   warp 1   MMA issuer     D[co, k] += dY^T U : eight K = 16-pixel tcgen05.mma per tile and 128-channel block; the fp32
//                           accumulator stays in TMEM over ALL tiles of the CTA (persistent) — no per-tile epilogue
//   warps 6-9 epilogue      once per CTA: TMEM -> partial[cta][Cout][32]; a fixed-order reduction over the CTAs follows
// ---------------------------------------------------------------------------------------------------------------
struct ThinWgParams {
  int N, H, W, Ct, R, pad, Cout;
  int TH, win_rows, row_floats, tiles, tiles_per_img, ones_col;
  uint64_t desc_templ;
  uint32_t idesc;
  float* partial;                // [gridDim][Cout][32]
};

__global__ void __launch_bounds__(kTuThreads, 2)
thin_conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmap_img, const __grid_constant__ CUtensorMap tmap_dy,
                            const ThinWgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) __trap();
  const int nco = p.Cout / 64;                              // 64-channel boxes per pixel half
  const uint32_t dy_stage_bytes = (uint32_t)(2 * nco) * 8192u;
  // layout: U[2] | dY[2] (2 * nco boxes of 8 KB each) | window[2] | barriers
  const uint32_t u_base = base;
  const uint32_t d_base = u_base + 2 * kTuABytes;
  const uint32_t w_base = d_base + 2 * dy_stage_bytes;
  const uint32_t bar = w_base + 2 * kTuWinBytes;
  auto h_full = [&](int b) { return bar + 8u * b; };
  auto h_empty = [&](int b) { return bar + 8u * (2 + b); };
  auto d_full = [&](int b) { return bar + 8u * (4 + b); };
  auto a_full = [&](int b) { return bar + 8u * (6 + b); };
  auto s_empty = [&](int b) { return bar + 8u * (8 + b); };
  const uint32_t t_full = bar + 8u * 10;
  const uint32_t tmem_slot = bar + 8u * 11;
  const float* swin = reinterpret_cast<const float*>(smem_raw + (w_base - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * kTuABytes / 16; i += kTuThreads) tu_sts16(u_base + 16u * i, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_img);
    tma_prefetch_desc(&tmap_dy);
    for (int b = 0; b < 2; ++b) {
      mbar_init(h_full(b), 1); mbar_init(h_empty(b), 4); mbar_init(d_full(b), 1);
      mbar_init(a_full(b), 4); mbar_init(s_empty(b), 1);
    }
    mbar_init(t_full, 1);
    fence_mbar_init();
  }
  const int tmem_cols = p.Cout <= 128 ? 64 : 128;           // 64 accumulator columns per 128-channel block
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int b = it & 1, ph = (it >> 1) & 1;
        const int n = t / p.tiles_per_img, h0 = (t - n * p.tiles_per_img) * p.TH;
        mbar_wait(h_empty(b), ph ^ 1);
        mbar_arrive_expect_tx(h_full(b), (uint32_t)(p.win_rows * p.row_floats * 4));
        tma_load_3d(w_base + b * kTuWinBytes, &tmap_img, h_full(b), 0, h0 - p.pad, n);
        mbar_wait(s_empty(b), ph ^ 1);                       // the MMAs of two tiles ago have consumed this stage
        mbar_arrive_expect_tx(d_full(b), dy_stage_bytes);
        const uint32_t ds = d_base + b * dy_stage_bytes;
        for (int h = 0; h < 2; ++h)
          for (int j = 0; j < nco; ++j)
            tma_load_2d(ds + (uint32_t)(h * nco + j) * 8192u, &tmap_dy, d_full(b), j * 64, t * kTuTile + h * 64);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int b = it & 1, ph = (it >> 1) & 1;
        mbar_wait(d_full(b), ph);
        mbar_wait(a_full(b), ph);
        tc_fence_after();
        const uint32_t ds = d_base + b * dy_stage_bytes;
        const uint32_t us = u_base + b * kTuABytes;
        for (int mt = 0; mt < p.Cout / 128; ++mt) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {                       // 16 pixel rows of 128 B per MMA
            const int h = k >> 2, kk = k & 3;
            const uint64_t adesc = smem_desc(p.desc_templ, ds + (uint32_t)(h * nco + mt * 2) * 8192u) + (uint64_t)(kk * 128);
            const uint64_t bdesc = smem_desc(p.desc_templ, us) + (uint64_t)(k * 128);
            umma_f16(tmem_base + mt * 64, adesc, bdesc, p.idesc, (uint32_t)((it | k) != 0));
          }
        }
        umma_commit(s_empty(b));
      }
      umma_commit(t_full);
    }
  } else if (warp < 6) {
    const int m = (warp - 2) * 32 + lane;
    const int th = m / p.W, tw = m - th * p.W;
    const int KT = p.R * p.R * p.Ct;
    const int sw = m & 7;
    int off[32];                                             // -1: zero, -2: one (the bias column)
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      off[k] = (k == p.ones_col) ? -2 : -1;
      if (k < KT) {
        const int tap = k / p.Ct, j = k - tap * p.Ct, tr = tap / p.R, ts = tap - tr * p.R;
        const int col = tw + ts - p.pad;
        if (col >= 0 && col < p.W) off[k] = (th + tr) * p.row_floats + col * p.Ct + j;
      }
    }
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const int b = it & 1, ph = (it >> 1) & 1;
      mbar_wait(h_full(b), ph);
      const float* win = swin + b * (kTuWinBytes / 4);
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = off[k] >= 0 ? win[off[k]] : (off[k] == -2 ? 1.0f : 0.0f);
      __syncwarp();
      if (lane == 0) mbar_arrive(h_empty(b));                // window consumed
      mbar_wait(s_empty(b), ph ^ 1);
      const uint32_t row = u_base + b * kTuABytes + m * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tu_sts16(row + (((uint32_t)c ^ sw) << 4), tu_pack(v[8 * c], v[8 * c + 1]), tu_pack(v[8 * c + 2], v[8 * c + 3]),
                 tu_pack(v[8 * c + 4], v[8 * c + 5]), tu_pack(v[8 * c + 6], v[8 * c + 7]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(b));
    }
  } else {
    const int quad = warp & 3;
    const int co = quad * 32 + lane;
    mbar_wait(t_full, 0);
    tc_fence_after();
    for (int mt = 0; mt < p.Cout / 128; ++mt) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + mt * 64, acc);
      tmem_ld_wait();
      float* dst = p.partial + ((size_t)blockIdx.x * p.Cout + mt * 128 + co) * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]),
                                                          __uint_as_float(acc[i + 2]), __uint_as_float(acc[i + 3]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

static int tu_wg_smem_bytes(int Cout) { return 2 * kTuABytes + 2 * (2 * (Cout / 64)) * 8192 + 2 * kTuWinBytes + 128; }

// -> number of CTA partials written to `partial` ([ctas][Cout][32] fp32), 0 if not applicable, < 0 on error
int thin_conv_wgrad_umma(const float* img, const void* dy_bf16, float* partial, int max_partials, int N, int H, int W, int Ct,
                         int R, int S, int Cout, cudaStream_t st) {
  const int K = R * S * Ct;
  if (!(R == S && (R == 1 || R == 3) && K <= 31 && (Cout == 128 || Cout == 256) && W >= 8 && W % 4 == 0 && kTuTile % W == 0 &&
        (H * W) % kTuTile == 0 && W * Ct <= 256))
    return 0;
  ThinWgParams p{};
  p.N = N; p.H = H; p.W = W; p.Ct = Ct; p.R = R; p.pad = (R - 1) / 2; p.Cout = Cout;
  p.TH = kTuTile / W;
  p.win_rows = p.TH + R - 1;
  p.row_floats = W * Ct;
  if (p.win_rows * p.row_floats * 4 > kTuWinBytes || p.win_rows > 256) return 0;
  p.tiles_per_img = H * W / kTuTile;
  p.tiles = N * p.tiles_per_img;
  p.ones_col = K;
  p.partial = partial;
  p.desc_templ = make_smem_desc_hi_lbo(8192, 1024, 1, 2);       // MN-major operands: 8 KB between 64-wide MN blocks
  p.idesc = make_idesc_bf16(128, 64, 1, 1, 0);
  CUtensorMap mi, md;
  {
    const uint64_t dims[3] = {(uint64_t)W * Ct, (uint64_t)H, (uint64_t)N};
    const uint32_t box[3] = {(uint32_t)(W * Ct), (uint32_t)p.win_rows, 1u};
    if (int e = make_tmap(&mi, img, 3, dims, box, 1, 1, 0)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)N * H * W};
    const uint32_t box[2] = {64u, 64u};
    if (int e = make_tmap(&md, dy_bf16, 2, dims, box, 1, 0, 2)) return e;
  }
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(thin_conv_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tu_wg_smem_bytes(kTuMaxCout)));
    MCGM_CUDA(cudaFuncSetAttribute(thin_conv_wgrad_umma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int resident = Cout <= 128 ? 2 : 1;
  int grid = p.tiles < resident * sms ? p.tiles : resident * sms;
  if (grid > max_partials) grid = max_partials;
  thin_conv_wgrad_umma_kernel<<<grid, kTuThreads, tu_wg_smem_bytes(Cout), st>>>(mi, md, p);
  MCGM_LAUNCH_CHECK();
  return grid;
}

}  // namespace mcgm
