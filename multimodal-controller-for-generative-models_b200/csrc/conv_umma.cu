// conv_umma.cu — implicit-GEMM convolution on 5th-gen tensor cores.
//
// Replaces the cudnn_convolution / convolution_backward calls behind
// nn.Conv2d in the MC block (src/models/mcgan.py:19,23,29,59,77-85,103-128).
//
//   forward / dgrad  : D[pixels, Cout] = sum_{tap, ci} X[pixel + tap, ci] * Wt[Cout, tap, ci]
//     A = a 128-pixel tile (TN images x TH rows x TW cols) x 64 input channels, fetched
//         per filter tap by ONE 4-D TMA box whose (w,h) origin is shifted by the tap;
//         out-of-bounds box elements are zero-filled by TMA == the conv zero padding.
//         No im2col buffer exists anywhere.
//     B = [Cout tile, 64] slab of the KRSC weight, 2-D TMA box.
//     Both land in smem as 128-byte rows with the 128B swizzle == the canonical
//     K-major UMMA layout; tcgen05.mma (M=128, N=Cout tile, K=16) accumulates in TMEM.
//   wgrad            : D[Cout, Cin] (per tap) = sum_pixels dY[pixel, co] * X[pixel + tap, ci]
//     the reduction dimension is the pixel index, which is the SLOW dimension of NHWC,
//     so both operands are fed as MN-major UMMA operands straight from the NHWC tensors.
//
// Kernel structure (both): 6 warps — warp 0 lane 0 = TMA producer, warp 1 lane 0 =
// MMA issuer, warps 2..5 = epilogue (one TMEM lane quadrant each).  Persistent
// over tiles, smem ring of kStages, double-buffered TMEM accumulator so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "reduce.cuh"
#include "tmap.cuh"
#include "umma.cuh"

namespace mcgm {

// ---------------------------------------------------------------------------
// tensor-map plumbing (driver entry point fetched at run time; no -lcuda)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d[5];
  uint32_t b[5];
  int rank, estride, f32, swizzle;
  bool operator==(const MapKey& o) const {
    if (ptr != o.ptr || rank != o.rank || estride != o.estride || f32 != o.f32 || swizzle != o.swizzle) return false;
    for (int i = 0; i < 5; ++i)
      if (d[i] != o.d[i] || b[i] != o.b[i]) return false;
    return true;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr) ^ (size_t)(k.rank + 16 * k.estride + 64 * k.f32 + 128 * k.swizzle) * 0x9e3779b97f4a7c15ull;
    for (int i = 0; i < 5; ++i) h = h * 1099511628211ull ^ (k.d[i] * 31 + k.b[i]);
    return h;
  }
};

// bf16 tensor, dims innermost-first, dense strides, 128B swizzle, zero OOB fill.  `estride` is the
// TMA traversal stride of the two spatial dims of a 4-D NHWC map (1, or 2 for stride-2 gathers): a
// box of b elements then lands b/estride elements in shared memory.
int make_tmap(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box, int estride, int f32,
              int swizzle) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{};
  key.ptr = ptr;
  key.rank = rank;
  key.estride = estride;
  key.f32 = f32;
  key.swizzle = swizzle;
  for (int i = 0; i < rank; ++i) { key.d[i] = dims[i]; key.b[i] = box[i]; }
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(MCGM_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // The encode is a DRIVER call and needs a current context on the calling thread.  An autograd
  // worker thread may reach this point before any runtime call bound one: bind the context of the
  // device that owns `ptr` once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
      MCGM_CUDA(cudaSetDevice(attr.device));
      MCGM_CUDA(cudaFree(0));
    }
    ctx_bound = true;
  }
  cuuint64_t gdim[4], gstride[3];
  cuuint32_t bdim[4], estr[4];
  uint64_t stride = f32 ? 4 : 2;  // bytes
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = (rank == 4 && (i == 1 || i == 2)) ? (cuuint32_t)estride : 1u;
    stride *= dims[i];
    if (i < rank - 1) gstride[i] = stride;
  }
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride,
                  bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MCGM_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu)",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0));
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
  }
  return 0;
}

// 5-D view of a bf16 NHWC tensor for MN-major operand tiles: dims (64 channels, W, H, N, C/64) — the channel BLOCK is the
// slowest box dimension, so ONE box of `blocks` channel blocks lands as `blocks` consecutive [pixels][64 ch] tiles with
// the 128B swizzle, exactly the MN-block sequence a UMMA descriptor walks with its leading-dimension offset.  One TMA
// operation per operand instead of one per 64 channels: the per-operation cost of cp.async.bulk.tensor is what bounded
// the weight gradient's ring.
int make_tmap_blocked(CUtensorMap* out, const void* ptr, int C, int W, int H, int N, int box_w, int box_h, int box_n,
                      int blocks, int estride) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{};
  key.ptr = ptr; key.rank = 5; key.estride = estride; key.f32 = 0; key.swizzle = 2;
  const uint64_t dims[5] = {64u, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(C / 64)};
  const uint32_t box[5] = {64u, (uint32_t)box_w, (uint32_t)box_h, (uint32_t)box_n, (uint32_t)blocks};
  for (int i = 0; i < 5; ++i) { key.d[i] = dims[i]; key.b[i] = box[i]; }
  key.d[0] = (uint64_t)C;                     // the channel count distinguishes views of one pointer
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(MCGM_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < 5; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = (i == 1 || i == 2) ? (cuuint32_t)estride : 1u; }
  gstride[0] = (cuuint64_t)C * 2;                       // w
  gstride[1] = (cuuint64_t)W * C * 2;                   // h
  gstride[2] = (cuuint64_t)H * W * C * 2;               // n
  gstride[3] = 128;                                     // channel block
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5u, const_cast<void*>(ptr), gdim, gstride, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MCGM_E_CUDA, "cuTensorMapEncodeTiled (blocked 5-D view) failed with CUresult %d (C %d W %d H %d N %d, box %d %d %d x %d blocks)",
                (int)r, C, W, H, N, box_w, box_h, box_n, blocks);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = *out;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// debug knobs (bring-up only): descriptor field overrides
// ---------------------------------------------------------------------------
struct DescKnobs {
  uint32_t k_lbo = 16, k_sbo = 1024;         // K-major operands
  uint32_t mn_lbo = 8192, mn_sbo = 1024;     // MN-major operands
  uint32_t version = 1, layout = 2;
  uint32_t mt2 = 1, mt2_256 = 0;             // forward conv: two M tiles per CTA (BN <= 128 / BN = 256)
  uint32_t nt2 = 1;                          // wgrad: two filter taps per CTA (Cin tile 128)
  uint32_t wg_splits = 0;                    // wgrad: force the split-K factor (0 = automatic)
  // wgrad: 5-D blocked tensor maps = one TMA operation per operand instead of one per 64 channels.  Parity green, time
  // unchanged (404 vs 406 us, 130 vs 132 us): the ring is not bound by the operation count.  Off by default.
  uint32_t wg_blocked = 0;
  uint32_t wg_bias_last = 1;                 // wgrad, NT = 2, odd tap count: dbias on the single-tap group (0: on group 0)
  uint32_t wg_pair = 1;                      // wgrad, NT = 2: one N = 2*BN MMA per k step for both taps (0: one MMA per tap)
  // staged forward epilogue: bit 2 = fetch the residual / ReLU-sign tile by TMA load (else through registers)
  uint32_t tma_epi = 7;
  // forward conv: CTAs per cluster sharing one weight slab by TMA multicast (1 = off, 2, 4).  Bit-identical results
  // (tests/gpu_checks.py:check_conv_cluster) but measured 0-2 % SLOWER (2) / 10 % slower (4): the kernel is bound by
  // shared-memory bandwidth (TMA writes + both MMA operands read from smem = 192 B/clk against 128), not by the L2 feed
  // that multicast relieves — DESIGN.md §7.  Kept as a tested knob, off by default.
  uint32_t cluster = 1;
  // TIMING DIAGNOSTICS ONLY (results are garbage): bit 0 skip the MMAs, 1 skip the activation loads, 2 skip the weight
  // loads, 3 skip the epilogue body.  tools/microbench.py uses them to time the pipeline's legs in isolation.
  uint32_t fwd_dbg = 0;
  // forward conv: CTA pairs (cta_group::2) for the 256-wide tiles / the 128-wide pair tiles (bit 0 / bit 1)
  uint32_t pair_mma = 3;
};
static DescKnobs g_knobs;

// ---------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------
static constexpr int kConvThreads = 320;     // forward: TMA warp + MMA warp + 8 epilogue warps (2 per TMEM lane quadrant)
static constexpr int kWgradThreads = 192;    // wgrad: TMA warp + MMA warp + 4 epilogue warps
static constexpr int kTileM = 128;
static constexpr int kBlockK = 64;                       // bf16 elements = 128 B rows
static constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KB

static constexpr int kMaxTaps = 16;
static constexpr int kMaxBiasBytes = 2048;           // forward: bias of up to 512 (padded) output channels in smem

// Geometry of the generalised forward conv.  M tiles are laid on a TILE GRID (Hg x Wg per image):
//   plain stride-1 conv   : grid = output = input, taps = R*S with offsets (r-pad, s-pad)
//   stride-2 conv (IS = 2): grid = output; input pixel of tap t = 2*grid + (dh,dw)   (TMA traversal stride 2)
//   phase conv  (PH = 4, OS = 2): grid = input; output pixel = 2*grid + (ph,pw); each phase has its own
//       2x2 taps and weight rows — nearest-up2 followed by conv3x3, and ConvTranspose2d(4,2,1).
struct ConvFwdParams {
  int N, Hg, Wg, Ho, Wo, Cin, Cout;
  int IS, OS, PH, T;                         // input traversal stride, output scale, phases, taps per phase
  int tf32, kblk;                            // fp32 operands read as TF32; elements per 128-byte k-block row (32 / 64)
  signed char tap_dh[4 * kMaxTaps], tap_dw[4 * kMaxTaps];   // [phase][tap]
  int TW, TH, TN, tiles_w, tiles_h;
  int n_tiles, m_super, total_tiles;
  int m_super_c, total_ctiles;               // the same counted in CLUSTER tiles (CL consecutive CTA tiles, one per CTA of a cluster)
  int kpt, num_kb;
  uint64_t desc_templ;
  uint32_t idesc;
  const float* bias;
  const __nv_bfloat16* res;
  const __nv_bfloat16* aux;
  const int32_t* labels;
  const uint32_t* bits;
  int words;
  int flags;
  int quad;                                  // 4 consecutive tile pixels never straddle a tile row (TW % 4 == 0)
  int tma_epi;                               // epilogue stages bf16 tiles in shared memory and writes them with TMA stores
  int side_tma;                              // ... and fetches the residual / ReLU-sign tile with a TMA load into the same buffer
  int dbg;                                   // timing diagnostics (DescKnobs::fwd_dbg), 0 in production
  int pair_mma;                              // launched as CTA pairs (cta_group::2 MMAs, M = 256)
  float* stats;                              // tma_epi only: [gridDim][2][Cout] per-CTA sum / sum of squares of the stored values
  void* out;
};

// MT = number of 128-pixel M tiles one CTA accumulates against the SAME weight slab: the B
// bytes fetched from L2 per MMA halve with MT = 2 (the kernel is L2->SM bound, not tensor bound).
// TMAEPI: the epilogue owns kEpiGroups 16 KB staging tiles (128 pixels x 64 channels bf16, 128B-swizzled = one TMA box)
// carved out of the operand ring: warps write their accumulator rows there (conflict-free thanks to the swizzle) and ONE
// thread per group stores the tile with cp.async.bulk.tensor — no per-thread global stores, no register transposes.
static constexpr int kSlabBytes = kTileM * 64 * 2;   // 16 KB
// Shared memory of the staged forward kernel: operand ring + staging tiles.  224 KB (the whole 227 KB a CTA may own, L1
// down to its minimum): the staged epilogue no longer lives on L1 (TMA stores / loads instead of row-strided global
// accesses), so the fourth pipeline stage is worth more than the cache — measured 3-7 % on every 3x3 shape against 192 KB.
#ifndef MCGM_FWD_RING_BYTES
#define MCGM_FWD_RING_BYTES 229376
#endif
// K extent of a pipeline stage (bf16 elements).  The operand ring is bound by bytes in flight over TMA latency (~1400
// cycles L2 -> shared memory): a stage's bytes stay allocated from the load's issue until the LAST MMA that reads them
// retires, so a 64-element stage (4 k-steps) holds its 48 KB for 512 tensor cycles longer than the data needs.  With
// 32-element stages (64-byte rows, SWIZZLE_64B, twice as many stages) the same ring turns over faster.
#ifndef MCGM_FWD_KB_128
#define MCGM_FWD_KB_128 64
#endif
#ifndef MCGM_FWD_KB_256
#define MCGM_FWD_KB_256 64
#endif
__host__ __device__ constexpr int fwd_kb(int BN, int MT, bool bf16_out, bool tf32, int CL) {
  return (!bf16_out || tf32 || CL != 1) ? 64 : (BN == 128 && MT == 2) ? MCGM_FWD_KB_128 : (BN == 256 && MT == 1) ? MCGM_FWD_KB_256 : 64;
}
template <int BN, int MT, bool TMAEPI = false, int KB = 64, bool PAIR = false>
struct FwdCfg {
  static constexpr int kKB = KB;
  static constexpr int kATile = kTileM * KB * 2;           // one M tile's k-block
  static constexpr int kBBytes = BN * KB * 2 / (PAIR ? 2 : 1);   // a CTA pair splits the weight rows between its two CTAs
  static constexpr int kStageBytes = MT * kATile + kBBytes;
  static constexpr int kEpiGroups = BN >= 128 ? 2 : 1;
  // Staging tiles per group.  The TMA store of a slab (and the side-slab load) queue behind the operand loads the producer
  // keeps in flight, a few thousand cycles; with one tile per group every slab waits for the previous slab's store to
  // drain before it may touch the buffer.  The 128-wide pair tile has only 2 x 64 MMA-cycles per k step to hide a slab
  // behind, so it alternates between TWO tiles (one ring stage less); the 256-wide tile hides the wait as it is.
#ifndef MCGM_FWD_STAGING_BUFS_128
#define MCGM_FWD_STAGING_BUFS_128 1
#endif
#ifndef MCGM_FWD_STAGING_BUFS_256
#define MCGM_FWD_STAGING_BUFS_256 1
#endif
  // CTA pairs hold half the weight rows per CTA: the 128-wide pair tile's second staging tile then costs no ring stage
#ifndef MCGM_FWD_STAGING_BUFS_PAIR_128
#define MCGM_FWD_STAGING_BUFS_PAIR_128 2
#endif
#ifndef MCGM_FWD_STAGING_BUFS_PAIR_256
#define MCGM_FWD_STAGING_BUFS_PAIR_256 1
#endif
  static constexpr int kStagingBufs =
      PAIR ? ((BN == 128 && MT == 2) ? MCGM_FWD_STAGING_BUFS_PAIR_128 : (BN == 256 ? MCGM_FWD_STAGING_BUFS_PAIR_256 : 1))
           : ((BN == 128 && MT == 2) ? MCGM_FWD_STAGING_BUFS_128 : (BN == 256 ? MCGM_FWD_STAGING_BUFS_256 : 1));
  static constexpr int kStagingBytes = TMAEPI ? kEpiGroups * kStagingBufs * kSlabBytes : 0;
  static constexpr int kRing = TMAEPI ? MCGM_FWD_RING_BYTES : 196608;
  static constexpr int kStages = ((kRing - kStagingBytes) / kStageBytes) > 8 ? 8 : ((kRing - kStagingBytes) / kStageBytes);
  static constexpr int kAccStages = (2 * MT * BN <= 512) ? 2 : 1;
  static constexpr int kTmemRaw = kAccStages * MT * BN;
  static constexpr int kTmemCols = kTmemRaw <= 32 ? 32 : kTmemRaw <= 64 ? 64 : kTmemRaw <= 128 ? 128 : kTmemRaw <= 256 ? 256 : 512;
  // No alignment slack: the dynamic window is declared 1024-byte aligned (the kernel traps if the base it
  // gets is not).  The total must stay under the 196 KB shared-memory carve-out (200704 B incl. 1 KB the
  // system reserves): one step more and L1 drops from 60 KB to 28 KB, which costs the epilogue's row-strided
  // loads and stores 10-40 % (measured on the 1x1 / thin / residual convs).
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 256 /*barriers*/ + kMaxBiasBytes;
  static_assert(kStages >= 2, "forward conv needs at least two pipeline stages");
  static_assert(kSmemBytes <= (TMAEPI ? MCGM_FWD_RING_BYTES : 196608) + 3072, "forward conv: shared-memory budget");
  static_assert(kSmemBytes <= 232448, "forward conv exceeds the 227 KB a CTA may own");
};

// 4x4 transpose of 16-byte units across the 4 lanes of a quad.  Before: lane l holds units 0..3 (64
// contiguous bytes) of ITS row.  After: lane l holds unit (l & 3) of the rows of lanes (l & ~3) + 0..3, so
// one 16-byte access per lane moves 64 contiguous bytes per row and a warp instruction touches 8 rows
// instead of 32.  The thread-per-row layout TMEM hands out would otherwise make every STG / LDG a 32-line
// scatter of half sectors, which is what bounded the epilogue (removing the stores alone sped the 256-wide
// 3x3 conv up 14 % and the 1x1 / thin convs 1.5-1.9x).  The transpose is its own inverse.
__device__ __forceinline__ void quad_transpose(uint4 (&u)[4], const int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
#pragma unroll
  for (int s = 0; s < 4; s += 2) {                         // unit pairs (0,1), (2,3) with lane ^ 1
    const uint4 send = b0 ? u[s] : u[s + 1];
    uint4 recv;
    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1); recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
    recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1); recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
    if (b0) u[s] = recv; else u[s + 1] = recv;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {                            // unit pairs (0,2), (1,3) with lane ^ 2
    const uint4 send = b1 ? u[s] : u[s + 2];
    uint4 recv;
    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 2); recv.y = __shfl_xor_sync(0xffffffffu, send.y, 2);
    recv.z = __shfl_xor_sync(0xffffffffu, send.z, 2); recv.w = __shfl_xor_sync(0xffffffffu, send.w, 2);
    if (b1) u[s] = recv; else u[s + 2] = recv;
  }
}

// One chunk (CW <= 32 consecutive output channels of one pixel) of the fused epilogue.  Executed by ALL
// lanes of the warp (quad shuffles inside); `row_ok` only predicates the global accesses of this row.
//   acc    fp32 bit patterns from TMEM
//   sbias  shared-memory bias of these columns (zero-padded; staged once per CTA)
//   mbits  codebook bits of these columns for the sample of this pixel, fetched one chunk ahead
//   sd     bf16 pipeline: the residual row of these columns (RES / RES_UP2), else the ReLU-sign row
//          (POSMASK), fetched one chunk ahead — quad-transposed when QUAD.  With both flags the sign row is
//          read in place; the fp32 parity pipeline (TF32) reads everything in place.
//   QUAD   bf16 rows move as quad-transposed 64-byte runs (needs CW == 32 and quads that do not straddle a
//          tile row: p.quad)
template <int CW, bool GENERIC_STORE, bool QUAD, typename OutT, bool TF32>
__device__ __forceinline__ void epilogue_chunk(const ConvFwdParams& p, const int flags, uint32_t (&acc)[32],
                                               const int pix, const int rpix, const int col0, const bool row_ok,
                                               const float* sbias, const uint32_t mbits, uint4 (&sd)[4],
                                               const int lane) {
  float v[32];
#pragma unroll
  for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(acc[i]);
  if (flags & MCGM_EPI_BIAS) {
#pragma unroll
    for (int i = 0; i < CW; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(sbias + i);      // warp-uniform address: broadcast
      v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
    }
  }
  const bool res_on = flags & (MCGM_EPI_RES | MCGM_EPI_RES_UP2);
  if constexpr (QUAD) {
    if (res_on || (flags & MCGM_EPI_POSMASK)) quad_transpose(sd, lane);   // back to one row per lane
  }
  if (res_on) {
    if constexpr (TF32) {                                  // fp32 pipeline: residual is fp32
      if (row_ok) {
        const float* rp = reinterpret_cast<const float*>(p.res) + (long)rpix * p.Cout + col0;
#pragma unroll
        for (int i = 0; i < CW; i += 4) {
          Vec<float> r;
          r.load(rp + i);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[i + j] += r.v[j];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < CW; i += 8) {
        float r[8];
        unpack16<__nv_bfloat16>(sd[i / 8], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i + j] += r[j];
      }
    }
  }
  if (flags & MCGM_EPI_RELU) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if (flags & MCGM_EPI_MASK) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = ((mbits >> i) & 1u) ? v[i] : 0.0f;
  }
  if (flags & MCGM_EPI_POSMASK) {
    if constexpr (TF32) {
      if (row_ok) {
        const float* ap = reinterpret_cast<const float*>(p.aux) + (long)pix * p.Cout + col0;
#pragma unroll
        for (int i = 0; i < CW; i += 4) {
          Vec<float> r;
          r.load(ap + i);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[i + j] = (r.v[j] > 0.0f) ? v[i + j] : 0.0f;
        }
      }
    } else if (res_on) {                                   // rare: both side tensors — sign row read in place
      if (row_ok) {
        const __nv_bfloat16* ap = p.aux + (long)pix * p.Cout + col0;
#pragma unroll
        for (int i = 0; i < CW; i += 8) {
          Vec<__nv_bfloat16> r;
          r.load(ap + i);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i + j] = (r.v[j] > 0.0f) ? v[i + j] : 0.0f;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < CW; i += 8) {
        float r[8];
        unpack16<__nv_bfloat16>(sd[i / 8], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i + j] = (r[j] > 0.0f) ? v[i + j] : 0.0f;
      }
    }
  }
  if constexpr (sizeof(OutT) == 4) {                       // tanh: fp32 outputs only (checked on the host)
    if (flags & MCGM_EPI_TANH) {
#pragma unroll
      for (int i = 0; i < CW; ++i) v[i] = tanhf(v[i]);
    }
  }
  if constexpr (QUAD && sizeof(OutT) == 2) {               // bf16 rows as quad-transposed 64-byte runs
    uint4 o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      o[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    quad_transpose(o, lane);
    const uint32_t okb = __ballot_sync(0xffffffffu, row_ok);
    const int base = lane & ~3;
    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + col0 + (lane & 3) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pk = __shfl_sync(0xffffffffu, pix, base + k);
      if ((okb >> (base + k)) & 1u) *reinterpret_cast<uint4*>(op + (long)pk * p.Cout) = o[k];
    }
  } else if constexpr (!GENERIC_STORE) {                   // Cout is a multiple of 32
    if (row_ok) {
      OutT* op = reinterpret_cast<OutT*>(p.out) + (long)pix * p.Cout + col0;
      constexpr int V = Vec<OutT>::N;
#pragma unroll
      for (int i = 0; i < CW; i += V) {
        Vec<OutT> o;
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = v[i + j];
        o.store(op + i);
      }
    }
  } else {
    if (row_ok) {
      OutT* op = reinterpret_cast<OutT*>(p.out) + (long)pix * p.Cout + col0;
#pragma unroll
      for (int i = 0; i < CW; ++i) {
        if (col0 + i < p.Cout) {
          if constexpr (sizeof(OutT) == 4) op[i] = v[i];
          else op[i] = __float2bfloat16(v[i]);
        }
      }
    }
  }
}

// bf16 side row (residual / ReLU-sign source) of one chunk: CW/8 independent 16-byte loads.  QUAD: lane l
// reads unit (l & 3) of the rows of its quad (64 contiguous bytes per row and instruction); the consumer
// transposes back.  All lanes of the warp take part.
template <int CW, bool QUAD>
__device__ __forceinline__ void side_load(const __nv_bfloat16* base, const int pix, const int Cout, const int col0,
                                          const bool on, uint4 (&r)[4], const int lane) {
  if constexpr (QUAD) {
    const uint32_t okb = __ballot_sync(0xffffffffu, on);
    const int b = lane & ~3;
    const __nv_bfloat16* q = base + col0 + (lane & 3) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pk = __shfl_sync(0xffffffffu, pix, b + k);
      if ((okb >> (b + k)) & 1u) r[k] = __ldg(reinterpret_cast<const uint4*>(q + (long)pk * Cout));
    }
  } else {
    if (on) {
      const uint4* q = reinterpret_cast<const uint4*>(base + (long)pix * Cout + col0);
#pragma unroll
      for (int i = 0; i < (CW >= 8 ? CW / 8 : 0); ++i) r[i] = __ldg(q + i);
    }
  }
}

// One 32-column chunk of the fused epilogue for the shared-memory staged path: values in, packed bf16 out.
//   acc    fp32 bit patterns from TMEM (row = this lane's pixel)
//   sd     the residual row (RES / RES_UP2) or the ReLU-sign row (POSMASK) of these 32 columns, one row per lane
__device__ __forceinline__ void epilogue_chunk_pack(const int flags, const uint32_t (&acc)[32], const float* sbias,
                                                    const uint32_t mbits, const uint4 (&sd)[4], const bool row_ok,
                                                    uint4 (&o)[4]) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  if (flags & MCGM_EPI_BIAS) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(sbias + i);      // warp-uniform address: broadcast
      v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
    }
  }
  if (flags & (MCGM_EPI_RES | MCGM_EPI_RES_UP2)) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      float r[8];
      unpack16<__nv_bfloat16>(sd[i / 8], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i + j] += r[j];
    }
  }
  if (flags & MCGM_EPI_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if (flags & MCGM_EPI_MASK) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = ((mbits >> i) & 1u) ? v[i] : 0.0f;
  }
  if (flags & MCGM_EPI_POSMASK) {                          // never together with a residual on this path (host)
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      float r[8];
      unpack16<__nv_bfloat16>(sd[i / 8], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i + j] = (r[j] > 0.0f) ? v[i + j] : 0.0f;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
      w[j] = row_ok ? *reinterpret_cast<uint32_t*>(&h) : 0u;     // rows outside the tensor: zeros (never stored; keep the statistics clean)
    }
    o[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds4(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// CL > 1: thread-block clusters of CL CTAs work on CL consecutive M tiles against the SAME weight slab.  Each CTA fetches
// 1/CL of every weight k-block and MULTICASTS it to all CTAs of the cluster, so one L2 read feeds CL SMs: the kernel is
// bound by the L2 -> SM operand feed (48 KB per 512 tensor-core cycles and SM), and the weights are 2/3 of those bytes
// for the 256-wide layers.  A pipeline stage is free when EVERY CTA of the cluster has consumed it (the peers' multicasts
// land in it), so every MMA warp signals the stage's empty barrier in all CTAs (tcgen05.commit ... multicast::cluster).
//
// PAIR (CL = 2): the two CTAs form a CTA PAIR and run cta_group::2 MMAs — M = 256: each CTA's M tile(s) against ONE weight
// slab of which each CTA holds (and loads) half the rows.  Per SM and k-block that is 1/6 (128-wide pair tiles) to 1/3
// (256-wide tiles) fewer bytes through TMA and shared memory than two independent CTAs, and the tensor core reads B
// once for both.  The even CTA issues every MMA; stage / accumulator hand-offs cross the pair:
//   full[s]   leader's barrier; armed by the leader's producer with BOTH CTAs' bytes, completed by both CTAs' loads
//   empty[s]  each CTA's own; the leader's tcgen05.commit multicasts the arrival to both
//   tfull[a]  each CTA's own; multicast commit
//   tempty[a] leader's barrier; the epilogue threads of BOTH CTAs arrive on it
template <int BN, int MT, typename OutT, bool TF32, int CL = 1, bool PAIR = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_side,
                const ConvFwdParams p) {
  constexpr bool kTmaEpi = (sizeof(OutT) == 2) && !TF32 && BN >= 64;    // bf16 tiles of >= 64 channels: staged + TMA-stored
  static_assert(!PAIR || (CL == 2 && kTmaEpi), "CTA pairs: clusters of two, bf16 staged epilogue");
  constexpr int KB = fwd_kb(BN, MT, sizeof(OutT) == 2, TF32, CL);
  using Cfg = FwdCfg<BN, MT, kTmaEpi, KB, PAIR>;
  constexpr int kElems = TF32 ? 32 : KB;                 // operand elements per k-block row (128 bytes; 64 with KB = 32)
  constexpr int kABytes = Cfg::kATile;                   // shadows the 64-element constant
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();                       // 128B-swizzled TMA tiles need a 1024-byte aligned base
  const uint32_t staging_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = staging_base + Cfg::kStagingBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tfull[2], tempty[2], tmem ptr (4 B + pad), side[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  auto side_bar = [&](int g) { return bar_base + 8u * (2 * Cfg::kStages + 5 + g); };
  // bias of all n_tiles * BN (zero-padded) output channels, after the 256 B barrier block
  float* sbias = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 256);
  for (int i = threadIdx.x; i < p.n_tiles * BN; i += kConvThreads)
    sbias[i] = ((p.flags & MCGM_EPI_BIAS) && i < p.Cout) ? __ldg(p.bias + i) : 0.0f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;         // this CTA's rank inside its cluster
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;         // cluster index / number of clusters
  constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (kTmaEpi && p.tma_epi) {
      tma_prefetch_desc(&tmap_out);
      if (p.side_tma) tma_prefetch_desc(&tmap_side);
    }
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), PAIR ? 1 : CL); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), (kConvThreads - 64) * (PAIR ? 2 : 1));
      mbar_init(side_bar(s), 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();                // the peers' barriers are initialised before anyone signals them
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // a CTA tile = MT consecutive 128-pixel M tiles x one BN-wide slab of output channels; a cluster tile = CL of them that
  // differ in the M index only.  CTA tile of cluster tile tc: (nt, ms = msc * CL + crank, ph); ms past the last tile is a
  // dummy (every coordinate out of range: zero-filled loads, clipped stores) that keeps the cluster in lock step.
  const int tiles_per_img_block = p.tiles_w * p.tiles_h;
  auto tile_of = [&](int tc, int& nt, int& ms, int& ph) {
    nt = tc % p.n_tiles;
    ms = ((tc / p.n_tiles) % p.m_super_c) * CL + crank;
    ph = tc / (p.n_tiles * p.m_super_c);
  };

  if (warp == 0) {
    if (elect_one()) {
      // ===== TMA producer =====
      int stage = 0, phase = 0;
      for (int tc = cid; tc < p.total_ctiles; tc += ncl) {
        int nt, ms, ph;
        tile_of(tc, nt, ms, ph);
        const int mt0 = ms * MT;
        const int wrow0 = ph * p.Cout + nt * BN;             // weight rows of this phase / channel slab
        int w0[MT], h0[MT], n0[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int mt = mt0 + j;
          w0[j] = (mt % p.tiles_w) * p.TW * p.IS;
          h0[j] = ((mt / p.tiles_w) % p.tiles_h) * p.TH * p.IS;
          n0[j] = (mt / tiles_per_img_block) * p.TN;       // past the batch for a ragged last pair: TMA zero-fills
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int dh = p.tap_dh[ph * kMaxTaps + tap], dw = p.tap_dw[ph * kMaxTaps + tap];
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          if (CL == 1 && p.dbg) {                              // timing diagnostics: leave out one or both operands
            mbar_arrive_expect_tx(full_bar(stage), ((p.dbg & 2) ? 0 : MT * kABytes) + ((p.dbg & 4) ? 0 : Cfg::kBBytes));
            if (!(p.dbg & 2))
              for (int j = 0; j < MT; ++j)
                tma_load_4d(sa + j * kABytes, &tmap_x, full_bar(stage), cb * kElems, w0[j] + dw, h0[j] + dh, n0[j]);
            if (!(p.dbg & 4)) tma_load_2d(sa + MT * kABytes, &tmap_w, full_bar(stage), tap * p.Cin + cb * kElems, wrow0);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            if (++cb == p.kpt) { cb = 0; ++tap; }
            continue;
          }
          if constexpr (PAIR) {
            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
#pragma unroll
            for (int j = 0; j < MT; ++j)
              tma_load_4d_2sm(sa + j * kABytes, &tmap_x, full_bar(stage), cb * kElems, w0[j] + dw, h0[j] + dh, n0[j]);
            tma_load_2d_2sm(sa + MT * kABytes, &tmap_w, full_bar(stage), tap * p.Cin + cb * kElems, wrow0 + crank * (BN / 2));
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            if (++cb == p.kpt) { cb = 0; ++tap; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
#pragma unroll
          for (int j = 0; j < MT; ++j)
            tma_load_4d(sa + j * kABytes, &tmap_x, full_bar(stage), cb * kElems, w0[j] + dw, h0[j] + dh, n0[j]);
          if constexpr (CL == 1) {
            tma_load_2d(sa + MT * kABytes, &tmap_w, full_bar(stage), tap * p.Cin + cb * kElems, wrow0);
          } else {                                           // my 1/CL of the weight rows, to every CTA of the cluster
            tma_load_2d_mc(sa + MT * kABytes + crank * (Cfg::kBBytes / CL), &tmap_w, full_bar(stage),
                           tap * p.Cin + cb * kElems, wrow0 + crank * (BN / CL), kClusterMask);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          if (++cb == p.kpt) { cb = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    if ((!PAIR || crank == 0) && elect_one()) {
      // ===== MMA issuer (CTA pairs: the leader CTA's, for both) =====
      int stage = 0, phase = 0, it = 0;
      for (int tc = cid; tc < p.total_ctiles; tc += ncl, ++it) {
        const int as = it % Cfg::kAccStages, aphase = (it / Cfg::kAccStages) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (MT * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (CL == 1 && (p.dbg & 1)) {                        // timing diagnostics: consume the stage without computing
            mbar_arrive(empty_bar(stage));
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t bdesc = smem_desc(p.desc_templ, sa + MT * kABytes);
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            const uint64_t adesc = smem_desc(p.desc_templ, sa + j * kABytes);
#pragma unroll
            for (int k = 0; k < KB / 16; ++k) {
              if constexpr (PAIR) umma_f16_2sm(d_tmem + j * BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (kb | k) != 0);
              else umma_any(TF32, d_tmem + j * BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (kb | k) != 0);
            }
          }
          if constexpr (PAIR) umma_commit_2sm(empty_bar(stage));    // frees the stage in both CTAs
          else if constexpr (CL == 1) umma_commit(empty_bar(stage));
          else umma_commit_mc(empty_bar(stage), kClusterMask);      // the stage is free for the peers' multicasts too
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (CL == 1 && (p.dbg & 1)) mbar_arrive(tfull_bar(as));
        else if constexpr (PAIR) umma_commit_2sm(tfull_bar(as));
        else umma_commit(tfull_bar(as));
      }
    }
  } else {
    // ===== epilogue: 8 warps.  A warp may only touch TMEM lanes [32*(warp%4), +32); the two warps of a
    // quadrant split the accumulator columns.
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int tw = m % p.TW, th = (m / p.TW) % p.TH, tn = m / (p.TW * p.TH);
    const int flags = p.flags;
    const bool res_on = flags & (MCGM_EPI_RES | MCGM_EPI_RES_UP2);

    // ---------------------------------------------------------------------------------------------------------
    // (A) bf16 tiles of >= 64 channels: shared-memory staged, TMA-stored.
    // The accumulator of a CTA tile is cut into SLABS of 128 pixels x 64 channels = 16 KB = one TMA box with the
    // 128B swizzle.  The warps of a GROUP (BN >= 128: the four warps of one column half; BN = 64: all eight) fill a
    // slab: every thread converts its pixel row (fused bias / residual / ReLU / MC mask / ReLU-sign mask) and writes
    // it as 16-byte units at unit ^ (row & 7) — the swizzle makes the 32 rows of a warp hit all banks exactly once.
    // One thread then issues cp.async.bulk.tensor (global <- shared): no per-thread global stores, no register
    // transposes, out-of-tensor rows clipped by the TMA unit.  The residual (or ReLU-sign source) slab is fetched BY
    // TMA INTO THE SAME BUFFER first (side_tma) and each thread reads back exactly the units it overwrites.
    // Optional batch statistics of the stored values (the BatchNorm that consumes this tensor): every warp sums the
    // 32 rows it has just written, two columns per lane, conflict-free 4-byte reads — no shuffles.
    // ---------------------------------------------------------------------------------------------------------
    auto tile_loop_tma = [&](auto quad_tag) {
      constexpr bool QUAD = decltype(quad_tag)::value;       // register-path side rows travel quad-transposed
      constexpr int G = Cfg::kEpiGroups;
      constexpr int GT = 256 / G;                            // threads per group
      constexpr int CPS = BN >= 128 ? 2 : 1;                 // 32-column chunks of one warp inside a slab
      constexpr int SPG = BN >= 128 ? (BN / 2) / 64 : 1;     // slabs per group and M tile
      const int grp = BN >= 128 ? half : 0;
      const bool leader_warp = (quad == 0) && (BN >= 128 || half == 0);   // its elected lane issues the TMA traffic of the group
      constexpr int NB = Cfg::kStagingBufs;                  // staging tiles per group, used alternately
      const uint32_t buf0 = staging_base + grp * NB * kSlabBytes;
      const uint32_t sbar = side_bar(grp);
      const int sw = m & 7;
      const int cw0 = BN >= 128 ? 0 : half * 32;             // this warp's first column inside the slab
      const bool side_tma = p.side_tma;
      const bool side_reg = !side_tma && (res_on || (flags & MCGM_EPI_POSMASK));
      const __nv_bfloat16* side = res_on ? p.res : p.aux;
      const bool stats_on = p.stats != nullptr;
      float st_sum[SPG][2], st_sq[SPG][2];
#pragma unroll
      for (int k = 0; k < SPG; ++k) { st_sum[k][0] = st_sum[k][1] = st_sq[k][0] = st_sq[k][1] = 0.0f; }
      uint32_t side_phase = 0;
      int it = 0;
      for (int tc = cid; tc < p.total_ctiles; tc += ncl, ++it) {
        const int as = it % Cfg::kAccStages, aphase = (it / Cfg::kAccStages) & 1;
        int nt, ms, ph;
        tile_of(tc, nt, ms, ph);
        const int mt0 = ms * MT;
        int pix[MT], rpix[MT], lab[MT], cw[MT], ch[MT], cn[MT];
        bool valid[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int mt = mt0 + j;
          const int w0 = (mt % p.tiles_w) * p.TW, h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
          cn[j] = (mt / tiles_per_img_block) * p.TN;
          cw[j] = w0 * p.OS + (ph & 1);                        // TMA coordinates of the output box of this M tile
          ch[j] = h0 * p.OS + (ph >> 1);
          const int wg = w0 + tw, hg = h0 + th, n = cn[j] + tn;
          valid[j] = (n < p.N) && (hg < p.Hg) && (wg < p.Wg);
          const int h = hg * p.OS + (ph >> 1), w = wg * p.OS + (ph & 1);
          pix[j] = (n * p.Ho + h) * p.Wo + w;
          rpix[j] = (flags & MCGM_EPI_RES_UP2) ? (n * (p.Ho >> 1) + (h >> 1)) * (p.Wo >> 1) + (w >> 1) : pix[j];
          lab[j] = ((flags & MCGM_EPI_MASK) && valid[j]) ? __ldg(p.labels + n) : 0;
        }
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (MT * BN);
        bool waited = false;
        if (p.dbg & 8) {                                       // timing diagnostics: release the accumulator untouched
          mbar_wait(tfull_bar(as), aphase);
          tc_fence_before();
          if constexpr (PAIR) mbar_arrive_leader(tempty_bar(as));
          else mbar_arrive(tempty_bar(as));
          continue;
        }
        // rolled over the M tiles (the unrolled body of both tiles does not fit the instruction cache: a third of the
        // epilogue's stall samples were instruction fetches), unrolled over the slabs of one tile
#pragma unroll 1
        for (int j = 0; j < MT; ++j) {
#pragma unroll
        for (int sl = 0; sl < SPG; ++sl) {
          const int s = j * SPG + sl;
          const int c_tile = (BN >= 128 ? half * (BN / 2) + sl * 64 : 0);       // slab's first column inside the CTA tile
          const int c_slab = nt * BN + c_tile;                                  // ... and inside the tensor
          const bool ok = (MT == 2 && j) ? valid[MT - 1] : valid[0];
          const uint32_t buf = buf0 + (s % NB) * kSlabBytes;
          const uint32_t row_addr = buf + m * 128;
          // everything that does not depend on the accumulator: side rows through registers, mask words
          uint4 sd[CPS][4];
          uint32_t mwd[CPS];
#pragma unroll
          for (int k = 0; k < CPS; ++k) {
            const int col = c_slab + cw0 + 32 * k;
            if (side_reg)
              side_load<32, QUAD>(side, res_on ? ((MT == 2 && j) ? rpix[MT - 1] : rpix[0]) : ((MT == 2 && j) ? pix[MT - 1] : pix[0]),
                                  p.Cout, col, ok, sd[k], lane);
            mwd[k] = 0u;
            if ((flags & MCGM_EPI_MASK) && ok)
              mwd[k] = __ldg(p.bits + (size_t)((MT == 2 && j) ? lab[MT - 1] : lab[0]) * p.words + (col >> 5));
          }
          // the staging buffer is free once the previous store has been read out of it; then the side slab may land in it
          if (leader_warp && elect_one()) {
            if constexpr (NB == 2) tma_store_wait_read_1();    // the store before last used this tile
            else tma_store_wait_read();
            if (side_tma) {
              mbar_arrive_expect_tx(sbar, kSlabBytes);
              tma_load_4d(buf, &tmap_side, sbar, c_slab, (MT == 2 && j) ? cw[MT - 1] : cw[0], (MT == 2 && j) ? ch[MT - 1] : ch[0],
                          (MT == 2 && j) ? cn[MT - 1] : cn[0]);
            }
          }
          if (!waited) {                                       // accumulator of this CTA tile ready?
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            waited = true;
          }
          uint32_t acc[CPS][32];
#pragma unroll
          for (int k = 0; k < CPS; ++k) tmem_ld32(t_row + j * BN + c_tile + cw0 + 32 * k, acc[k]);
          if (side_tma) mbar_wait(sbar, side_phase);
          else named_bar_sync(1 + grp, GT);
          side_phase ^= 1u;
          tmem_ld_wait();
          if (s == MT * SPG - 1) {                             // accumulator drained: the MMA warp may refill it
            tc_fence_before();
            if constexpr (PAIR) mbar_arrive_leader(tempty_bar(as));      // the leader's MMA warp refills both CTAs' TMEM
            else mbar_arrive(tempty_bar(as));
          }
#pragma unroll
          for (int k = 0; k < CPS; ++k) {
            const int u0 = (cw0 + 32 * k) >> 3;                // first 16-byte unit of this chunk inside the 128-byte row
            if (side_tma) {
#pragma unroll
              for (int u = 0; u < 4; ++u) sd[k][u] = lds16(row_addr + (((u0 + u) ^ sw) << 4));
            } else if (QUAD && side_reg) {
              quad_transpose(sd[k], lane);                     // back to one row per lane
            }
            uint4 o[4];
            epilogue_chunk_pack(flags, acc[k], sbias + c_slab + cw0 + 32 * k, mwd[k], sd[k], ok, o);
#pragma unroll
            for (int u = 0; u < 4; ++u) sts16(row_addr + (((u0 + u) ^ sw) << 4), o[u]);
          }
          if (stats_on) {                                      // column sums of the 32 rows this warp has just written
            __syncwarp();
            const int col = cw0 + 2 * lane;                    // column pair inside the slab
            if (col < cw0 + 32 * CPS) {
              const uint32_t base = buf + (uint32_t)(quad * 32) * 128 + ((col & 7) << 1);
              const int unit = col >> 3;
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                const uint32_t wv = lds4(base + r * 128 + ((unit ^ (r & 7)) << 4));
                const float a = __uint_as_float(wv << 16), b = __uint_as_float(wv & 0xffff0000u);
                s0 += a; s1 += b; q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
              }
              st_sum[sl][0] += s0; st_sum[sl][1] += s1; st_sq[sl][0] += q0; st_sq[sl][1] += q1;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + grp, GT);
          if (leader_warp && elect_one()) {
            tma_store_4d(&tmap_out, buf, c_slab, (MT == 2 && j) ? cw[MT - 1] : cw[0], (MT == 2 && j) ? ch[MT - 1] : ch[0],
                         (MT == 2 && j) ? cn[MT - 1] : cn[0]);
            tma_store_commit();
          }
        }
        }
      }
      if (leader_warp && elect_one()) tma_store_wait_all();
      if (stats_on) {
        // one statistics slot per CTA: the four warp rows are summed through the (now idle) staging buffers in a fixed
        // order — the finalize kernel reads 4x fewer partials
        float* sm = reinterpret_cast<float*>(smem_raw + (staging_base - smem_base));       // [4 quads][2][BN] floats <= 8 KB
        named_bar_sync(3, 256);                                // every group's last store has left the staging buffers
        const int col = cw0 + 2 * lane;
        const bool mine = col < cw0 + 32 * CPS;
        if (mine) {
#pragma unroll
          for (int sl = 0; sl < SPG; ++sl) {
            const int c = (BN >= 128 ? half * (BN / 2) + sl * 64 : 0) + col;      // n_tiles == 1 (host)
            sm[(quad * 2 + 0) * BN + c] = st_sum[sl][0]; sm[(quad * 2 + 0) * BN + c + 1] = st_sum[sl][1];
            sm[(quad * 2 + 1) * BN + c] = st_sq[sl][0];  sm[(quad * 2 + 1) * BN + c + 1] = st_sq[sl][1];
          }
        }
        named_bar_sync(3, 256);
        if (mine && quad == 0) {
          float* sp = p.stats + (size_t)blockIdx.x * 2 * p.Cout;
#pragma unroll
          for (int sl = 0; sl < SPG; ++sl) {
            const int c = (BN >= 128 ? half * (BN / 2) + sl * 64 : 0) + col;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              sp[c + e] = (sm[0 * BN + c + e] + sm[2 * BN + c + e]) + (sm[4 * BN + c + e] + sm[6 * BN + c + e]);
              sp[p.Cout + c + e] = (sm[1 * BN + c + e] + sm[3 * BN + c + e]) + (sm[5 * BN + c + e] + sm[7 * BN + c + e]);
            }
          }
        }
      }
    };

    // ---------------------------------------------------------------------------------------------------------
    // (B) every other output (fp32, TF32 pipeline, narrow tiles): per-thread stores.  The epilogue is latency-, not
    // issue-bound, so nothing it needs is fetched after the accumulator is ready: the bias sits in shared memory, the
    // labels of the tile's samples plus the first codebook word and residual / ReLU-sign row are requested BEFORE the
    // wait on the MMA, and inside the tile the TMEM load, codebook word and side row of chunk q+1 are in flight while
    // chunk q is processed.  Two chunk bodies (even / odd buffers) in a rolled loop keep the code inside the
    // instruction cache.  bf16 rows travel quad-transposed (see quad_transpose).
    // ---------------------------------------------------------------------------------------------------------
    [[maybe_unused]] constexpr int CW = BN < 32 ? BN : 32;   // columns per TMEM load
    constexpr int HALF = (BN / 2 >= CW) ? BN / 2 : BN;       // columns per warp (narrow tiles: warp half 0 only)
    constexpr int NCH = HALF / CW;
    constexpr int Q = MT * NCH;                              // chunks per CTA tile and warp: 1, 2 or 4
    static_assert(Q == 1 || Q % 2 == 0, "chunk pipeline handles one chunk or pairs");
    constexpr bool kQuadOk = CW == 32 && !TF32;
    const bool works = half * HALF < BN;
    const __nv_bfloat16* side = TF32 ? nullptr : (res_on ? p.res : ((flags & MCGM_EPI_POSMASK) ? p.aux : nullptr));
    auto tile_loop = [&](auto quad_tag) {
      constexpr bool QUAD = decltype(quad_tag)::value;
      int it = 0;
      for (int tc = cid; tc < p.total_ctiles; tc += ncl, ++it) {
        const int as = it % Cfg::kAccStages, aphase = (it / Cfg::kAccStages) & 1;
        int nt, ms, ph;
        tile_of(tc, nt, ms, ph);
        const int mt0 = ms * MT;
        const int cl0 = nt * BN + half * HALF;                 // first output channel of this warp
        int pix[MT], rpix[MT], lab[MT];                        // pixel indices fit 31 bits (checked on the host)
        bool valid[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int mt = mt0 + j;
          const int wg = (mt % p.tiles_w) * p.TW + tw;
          const int hg = ((mt / p.tiles_w) % p.tiles_h) * p.TH + th;
          const int n = (mt / tiles_per_img_block) * p.TN + tn;
          valid[j] = works && (n < p.N) && (hg < p.Hg) && (wg < p.Wg);
          const int h = hg * p.OS + (ph >> 1), w = wg * p.OS + (ph & 1);     // output pixel (phase = (ph_h, ph_w))
          pix[j] = (n * p.Ho + h) * p.Wo + w;
          rpix[j] = (flags & MCGM_EPI_RES_UP2) ? (n * (p.Ho >> 1) + (h >> 1)) * (p.Wo >> 1) + (w >> 1) : pix[j];
          lab[j] = ((flags & MCGM_EPI_MASK) && valid[j]) ? __ldg(p.labels + n) : 0;
        }
        // everything chunk q needs besides the accumulator
        auto fetch_side = [&](int q, uint4 (&sd)[4], uint32_t& mwd) {
          const int j = q / NCH, col = cl0 + (q % NCH) * CW;
          const bool ok = ((MT == 2 && j) ? valid[MT - 1] : valid[0]) && col < p.Cout;
          if (side != nullptr)
            side_load<CW, QUAD>(side, res_on ? ((MT == 2 && j) ? rpix[MT - 1] : rpix[0]) : ((MT == 2 && j) ? pix[MT - 1] : pix[0]),
                                p.Cout, col, ok, sd, lane);
          mwd = 0u;
          if ((flags & MCGM_EPI_MASK) && ok)                   // MASK implies Cout % 32 == 0 (CW == 32)
            mwd = __ldg(p.bits + (size_t)((MT == 2 && j) ? lab[MT - 1] : lab[0]) * p.words + (col >> 5));
        };
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + as * (MT * BN) + half * HALF;
        auto fetch_acc = [&](int q, uint32_t (&a)[32]) {
          const uint32_t ta = t_row + (q / NCH) * BN + (q % NCH) * CW;
          if constexpr (CW == 32) tmem_ld32(ta, a);
          else tmem_ld16(ta, a);
        };
        auto run = [&](int q, uint32_t (&a)[32], uint4 (&sd)[4], uint32_t mwd) {
          const int j = q / NCH, col = cl0 + (q % NCH) * CW;
          if (col < p.Cout)                                    // warp-uniform
            epilogue_chunk<CW, BN == 16, QUAD, OutT, TF32>(
                p, flags, a, (MT == 2 && j) ? pix[MT - 1] : pix[0], (MT == 2 && j) ? rpix[MT - 1] : rpix[0], col,
                (MT == 2 && j) ? valid[MT - 1] : valid[0], sbias + col, mwd, sd, lane);
        };
        uint32_t acc[2][32];
        uint4 sd[2][4];
        uint32_t mwd[2];
        if (works) fetch_side(0, sd[0], mwd[0]);
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        if (works) {
          fetch_acc(0, acc[0]);
          if constexpr (Q == 1) {
            tmem_ld_wait();
            run(0, acc[0], sd[0], mwd[0]);
          } else {
#pragma unroll 1
            for (int q = 0; q < Q; q += 2) {
              tmem_ld_wait();
              fetch_acc(q + 1, acc[1]);
              fetch_side(q + 1, sd[1], mwd[1]);
              run(q, acc[0], sd[0], mwd[0]);
              tmem_ld_wait();
              if (q + 2 < Q) {
                fetch_acc(q + 2, acc[0]);
                fetch_side(q + 2, sd[0], mwd[0]);
              }
              run(q + 1, acc[1], sd[1], mwd[1]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(as));
      }
    };
    if constexpr (kTmaEpi) {
#ifdef MCGM_KEEP_OLD_EPI          // A/B build only: also compile the per-thread-store epilogue into the staged instantiations
      if (!p.tma_epi) {
        if (p.quad) tile_loop(std::true_type{});
        else tile_loop(std::false_type{});
      } else
#endif
      if (p.quad) tile_loop_tma(std::true_type{});
      else tile_loop_tma(std::false_type{});
    } else if constexpr (kQuadOk) {
      if (p.quad) tile_loop(std::true_type{});
      else tile_loop(std::false_type{});
    } else {
      tile_loop(std::false_type{});
    }
  }

  tc_fence_before();
  if constexpr (CL > 1) cluster_sync_all();                // no CTA leaves while a peer may still signal its barriers
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// wgrad kernel: one (cout tile, cin tile, tap, k-split) work item per CTA
// ---------------------------------------------------------------------------
// Generalised weight gradient: k-blocks are 64-pixel boxes of a GRID (Hg x Wg per image);
//   dY pixel of tap t = AS*grid + (a_dh,a_dw)[t],  X pixel of tap t = BS*grid + (b_dh,b_dw)[t]
//   stride-1 conv : AS = BS = 1, a = 0, b = (r-pad, s-pad)
//   4x4 stride-2  : grid = dY grid, AS = 1, a = 0; BS = 2, b = (p-1, q-1)            (16 taps)
//   phase conv    : grid = X grid,  AS = 2, a = (p-1, q-1); BS = 1, b = 0            (16 taps)
static constexpr int kWgMaxTaps = 32;     // explicit-tap weight gradients (PixelCNN's 4x7 vertical stack: 28 taps)
struct WgradParams {
  int N, Cin, Cout;
  int AS, BS;
  signed char a_dh[kWgMaxTaps], a_dw[kWgMaxTaps], b_dh[kWgMaxTaps], b_dw[kWgMaxTaps];
  int KW, KH, KN, kb_w, kb_h;           // a k-block is a KN x KH x KW box of 64 grid pixels
  int total_kb, splits, kb_per_split;
  int cout_tiles, cin_tiles, taps, tap_groups;
  uint64_t desc_templ;
  uint32_t idesc, idesc_bias;
  int blocked_maps;                     // the tensor maps are the 5-D blocked views (make_tmap_blocked): one load per operand
  uint32_t idesc_pair;                  // NT = 2: one N = 2*BN instruction covers both taps (their X tiles are adjacent in smem)
  float* partial;                       // [splits][Cout*R*S*Cin]
  float* bias_partial;                  // [splits][bias_slots][Cout] or nullptr
  int bias_taps[4], bias_slots;         // taps whose dY pixels tile dY exactly once (stride-2 dY: 4 parity classes)
  int bias_spread;                      // bias MMAs dealt round-robin over ALL tap groups (bias_slots = tap_groups)
  // grouped launch: `groups` independent problems of identical geometry share one grid (CTA -> group by index)
  int groups, ctas_per_group;
  size_t group_partial_stride, group_bias_stride;      // floats between the groups' partial / bias-partial blocks
};

static constexpr int kWgMaxGroups = 4;
struct WgMaps {
  CUtensorMap dy[kWgMaxGroups];
  CUtensorMap x[kWgMaxGroups];
};

// NT = number of filter taps one CTA accumulates against the SAME dY tile (A operand):
// the dY bytes fetched per MMA halve with NT = 2.
// KB = pixels per k-block (ring stage): 64, or 32 (bf16 only).  The rows of the MN-major operand tiles are pixels, so a
// 32-pixel stage has half the 128-byte rows of a 64-pixel one — same TMA efficiency, twice the stages, and a stage's bytes
// are released after 2 instead of 4 k-steps (the ring is bound by bytes in flight over TMA latency).
#ifndef MCGM_WG_KB
#define MCGM_WG_KB 64
#endif
template <int BN, int NT, int KB = 64>
struct WgCfg {
  static constexpr int kKB = KB;
  static constexpr int kBlkBytes = KB * 128;           // one 64-wide MN block of a stage: KB pixel rows of 128 B
  static constexpr int kATile = kTileM * KB * 2;
  static constexpr int kBBytes = BN * KB * 2;
  static constexpr int kStageBytes = kATile + NT * kBBytes;
  static constexpr int kBiasCol = NT * BN;             // 16 extra accumulator columns: dY^T x ones = bias gradient
  static constexpr int kTmemRaw = NT * BN + 16;
  static constexpr int kTmemCols = kTmemRaw <= 32 ? 32 : kTmemRaw <= 64 ? 64 : kTmemRaw <= 128 ? 128 : kTmemRaw <= 256 ? 256 : 512;
  static constexpr int kOnesBytes = 8192;              // an all-ones [64 k][64 mn] bf16 operand tile
  // The whole 227 KB a CTA may own: the kernel never touches L1 (TMA operands, float4 partial stores), and the operand
  // ring is bound by bytes in flight over TMA latency — the fourth 48 KB stage is what the 192 KB budget left out.
#ifndef MCGM_WG_RING_BYTES
#define MCGM_WG_RING_BYTES 229376
#endif
  static constexpr int kRing = MCGM_WG_RING_BYTES - kOnesBytes - 1280;
  static constexpr int kStages = (kRing / kStageBytes) > 8 ? 8 : (kRing / kStageBytes);
  static constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 + 256;
  static_assert(kSmemBytes <= 232448, "wgrad exceeds the 227 KB a CTA may own");
};

template <int BN, int NT, int KB = 64>
__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgMaps maps, const WgradParams pin) {
  using Cfg = WgCfg<BN, NT, KB>;
  constexpr int kABytes = Cfg::kATile;                 // shadows the 64-pixel constant
  // Several independent weight gradients of identical geometry (the layers of one backward pass that share a
  // shape) can share one launch: alone, a small-spatial layer fills 148 SMs only through ~29-way split-K, whose
  // partials then dominate its cost; two to four of them side by side need a quarter of the splits each.
  const int grp = blockIdx.x / pin.ctas_per_group;
  const CUtensorMap& tmap_dy = maps.dy[grp];
  const CUtensorMap& tmap_x = maps.x[grp];
  const WgradParams& p = pin;
  float* const partial = pin.partial + (size_t)grp * pin.group_partial_stride;
  float* const bias_partial = pin.bias_partial ? pin.bias_partial + (size_t)grp * pin.group_bias_stride : nullptr;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ones_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = ones_base + Cfg::kOnesBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * Cfg::kStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int id = blockIdx.x % p.ctas_per_group;
  const int split = id % p.splits; id /= p.splits;
  const int tap0 = (id % p.tap_groups) * NT; id /= p.tap_groups;
  const int cin_t = id % p.cin_tiles;
  const int cout_t = id / p.cin_tiles;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, p.total_kb);
  const int ntaps = min(NT, p.taps - tap0);          // the last group of a 3x3 kernel holds one tap
  const int co0 = cout_t * kTileM, ci0 = cin_t * BN;
  // the CTAs of tap group 0 / cin tile 0 also accumulate dbias = dY^T x ones (one extra N=16 MMA per k step)
  int bias_slot = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < p.bias_slots && p.bias_taps[i] == tap0) bias_slot = i;
  // bias_spread: EVERY tap group takes the bias MMAs of the k-blocks with kb % tap_groups == its index (slot = group):
  // a few percent on every CTA instead of +44 % on the CTAs of one group, which then set the time of the one-wave launch
  const int tgroup = tap0 / NT;
  if (p.bias_spread) bias_slot = tgroup;
  const bool do_bias = bias_partial != nullptr && bias_slot >= 0 && cin_t == 0;
  // k-blocks of this CTA that carry a bias MMA (spread mode: a stride-tap_groups subset, possibly empty)
  const int bias_first = p.bias_spread ? kb0 + ((tgroup - kb0 % p.tap_groups) + p.tap_groups) % p.tap_groups : kb0;
  const bool bias_any = bias_first < kb1;
  if (do_bias && warp >= 2) {
    const uint32_t ones2 = 0x3f803f80u;              // two bf16 1.0
    for (int i = (warp - 2) * 32 + lane; i < Cfg::kOnesBytes / 16; i += 128)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ones_base + 16u * i), "r"(ones2) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to tcgen05.mma
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      int dr[NT], ds[NT];                        // X-side tap offsets of this CTA's taps (registers: loops are unrolled)
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        const int tap = min(tap0 + q, p.taps - 1);
        dr[q] = p.b_dh[tap];
        ds[q] = p.b_dw[tap];
      }
      const int ar = p.a_dh[tap0], ac = p.a_dw[tap0];   // dY-side offset (non-zero only when NT == 1)
      int bw = kb0 % p.kb_w, bh = (kb0 / p.kb_w) % p.kb_h, bn = kb0 / (p.kb_w * p.kb_h);
      const uint32_t tx_bytes = kABytes + ntaps * Cfg::kBBytes;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int w0 = bw * p.KW, h0 = bh * p.KH, n0 = bn * p.KN;
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(stage), tx_bytes);
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        if (p.blocked_maps) {                    // one operation per operand: all its 64-channel blocks in one 5-D box
          tma_load_5d(sa, &tmap_dy, full_bar(stage), 0, w0 * p.AS + ac, h0 * p.AS + ar, n0, co0 / 64);
#pragma unroll
          for (int q = 0; q < NT; ++q)
            if (q < ntaps)
              tma_load_5d(sa + kABytes + q * Cfg::kBBytes, &tmap_x, full_bar(stage), 0, w0 * p.BS + ds[q], h0 * p.BS + dr[q], n0, ci0 / 64);
        } else {
#pragma unroll
        for (int j = 0; j < kTileM / 64; ++j)
          tma_load_4d(sa + j * Cfg::kBlkBytes, &tmap_dy, full_bar(stage), co0 + 64 * j, w0 * p.AS + ac, h0 * p.AS + ar, n0);
#pragma unroll
        for (int q = 0; q < NT; ++q) {
          if (q < ntaps) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sa + kABytes + q * Cfg::kBBytes + j * Cfg::kBlkBytes, &tmap_x, full_bar(stage), ci0 + 64 * j,
                          w0 * p.BS + ds[q], h0 * p.BS + dr[q], n0);
          }
        }
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        if (++bw == p.kb_w) { bw = 0; if (++bh == p.kb_h) { bh = 0; ++bn; } }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
        const uint64_t adesc = smem_desc(p.desc_templ, sa);
        const uint32_t acc_on = kb > kb0;
        if (NT == 2 && ntaps == 2 && p.idesc_pair != 0u) {
          // Both taps in ONE instruction: their X tiles sit back to back in the stage, i.e. 2*BN/64 MN-blocks at the
          // uniform 8 KB stride the descriptor's LBO describes = a valid N = 2*BN operand.  dY (the A operand) is then
          // read from shared memory once per k step instead of once per tap: the kernel is bound by shared-memory
          // bandwidth (TMA writes + both MMA operands), and this removes a quarter of the operand reads.
          const uint64_t bdesc = smem_desc(p.desc_templ, sa + kABytes);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)
            umma_f16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), p.idesc_pair, acc_on | (uint32_t)(k != 0));
        } else
#pragma unroll
        for (int q = 0; q < NT; ++q) {
          if (q < ntaps) {
            const uint64_t bdesc = smem_desc(p.desc_templ, sa + kABytes + q * Cfg::kBBytes);
#pragma unroll
            for (int k = 0; k < KB / 16; ++k)   // 16 pixel rows of 128 B per MMA
              umma_f16(tmem_base + q * BN, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), p.idesc,
                       acc_on | (uint32_t)(k != 0));
          }
        }
        if (do_bias && (!p.bias_spread || kb % p.tap_groups == tgroup)) {
          const uint64_t odesc = smem_desc(p.desc_templ, ones_base);
          const uint32_t bias_acc = kb > bias_first;
#pragma unroll
          for (int k = 0; k < KB / 16; ++k)
            umma_f16(tmem_base + Cfg::kBiasCol, adesc + (uint64_t)(k * 128), odesc, p.idesc_bias,
                     bias_acc | (uint32_t)(k != 0));
        }
        umma_commit(empty_bar(stage));
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int quad = warp & 3;
    const int co = co0 + quad * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const size_t per_split = (size_t)p.Cout * p.taps * p.Cin;
    if (do_bias) {
      uint32_t acc[32];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + Cfg::kBiasCol, acc);
      tmem_ld_wait();
      if (co < p.Cout)
        bias_partial[((size_t)split * p.bias_slots + bias_slot) * p.Cout + co] = bias_any ? __uint_as_float(acc[0]) : 0.0f;
    }
    for (int q = 0; q < ntaps; ++q) {
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + q * BN;
      float* dst = partial + (size_t)split * per_split + ((size_t)co * p.taps + tap0 + q) * p.Cin + ci0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t acc[32];
        tmem_ld32(t_row + c, acc);
        tmem_ld_wait();
        if (co < p.Cout) {
          const bool have = kb1 > kb0;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + c + i) =
                have ? make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                   __uint_as_float(acc[i + 3]))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Split-K reduction: dw[i] = sum_s partial[s][i], dbias[c] = sum_r bias_partial[r][c].
// Block = 32 consecutive outputs x 8 split lanes (one warp each): a warp reads 512 contiguous bytes of one
// split, lane w of a column sums splits w, w+8, ... with four loads in flight, and the eight lane sums are
// combined through shared memory in lane order — a fixed order, hence deterministic.  (One thread per output
// walking all splits was a chain of up to 148 dependent L2 round trips on 8..144 blocks: 13-16 us a call.)
struct WgReduceOut {
  float* dw[kWgMaxGroups];
  float* dbias[kWgMaxGroups];
};

__global__ void __launch_bounds__(32 * kRedLanes)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, size_t n4 /*float4 count*/, const WgReduceOut out,
                    const float* __restrict__ bias_partial, int bias_rows, int Cout, int wblocks,
                    size_t group_partial_stride, size_t group_bias_stride) {
  __shared__ float4 sm4[kRedLanes][32];
  const int g = blockIdx.y;
  partial += (size_t)g * group_partial_stride;
  if ((int)blockIdx.x < wblocks) {
    for (size_t i0 = (size_t)blockIdx.x * 32; i0 < n4; i0 += (size_t)wblocks * 32)
      reduce_columns<float4>(reinterpret_cast<const float4*>(partial), splits, n4, n4, i0, reinterpret_cast<float4*>(out.dw[g]), sm4);
  } else {
    float (*sm1)[32] = reinterpret_cast<float (*)[32]>(&sm4[0][0]);
    reduce_columns<float>(bias_partial + (size_t)g * group_bias_stride, bias_rows, (size_t)Cout, (size_t)Cout,
                          (size_t)(blockIdx.x - wblocks) * 32, out.dbias[g], sm1);
  }
}

// ---------------------------------------------------------------------------
// weight layout kernels
// ---------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int K, int C, int R, int S,
                                   const float* __restrict__ inv_sigma, __nv_bfloat16* __restrict__ krsc,
                                   __nv_bfloat16* __restrict__ crsk_flip) {
  const long total = (long)K * C * R * S;
  const float sc = inv_sigma ? __ldg(inv_sigma) : 1.0f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    // i indexes the KRSC destination so the bf16 writes are coalesced
    long t = i;
    const int c = (int)(t % C); t /= C;
    const int s = (int)(t % S); t /= S;
    const int r = (int)(t % R);
    const int k = (int)(t / R);
    const float v = w[(((long)k * C + c) * R + r) * S + s] * sc;
    const __nv_bfloat16 b = __float2bfloat16(v);
    if (krsc) krsc[i] = b;
    if (crsk_flip) crsk_flip[(((long)c * R + (R - 1 - r)) * S + (S - 1 - s)) * K + k] = b;
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dw_krsc, int K, int C, int R, int S,
                                    float* __restrict__ dw_oihw, int accumulate) {
  const long total = (long)K * C * R * S;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int c = (int)(t % C); t /= C;
    const int s = (int)(t % S); t /= S;
    const int r = (int)(t % R);
    const int k = (int)(t / R);
    const long o = (((long)k * C + c) * R + r) * S + s;
    dw_oihw[o] = accumulate ? dw_oihw[o] + dw_krsc[i] : dw_krsc[i];
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static bool pick_pixel_box(int H, int W, int npix, int* bw, int* bh, int* bn) {
  // a box of `npix` pixels: as many columns as fit, then rows, then images
  int tw = W < npix ? W : npix;
  if (npix % tw) return false;
  int rem = npix / tw;
  int th = H < rem ? H : rem;
  if (rem % th) return false;
  int tn = rem / th;
  if (W % tw || H % th) return false;
  *bw = tw; *bh = th; *bn = tn;
  return true;
}

struct FwdMaps {
  CUtensorMap x, w, out, side;          // out / side: only read by kernels launched with p.tma_epi / p.side_tma
};

template <int BN, int MT, typename OutT, bool TF32 = false, int CL = 1, bool PAIR = false>
static int launch_fwd(const FwdMaps& maps, ConvFwdParams p, cudaStream_t st) {
  constexpr bool kTmaEpi = (sizeof(OutT) == 2) && !TF32 && BN >= 64;
  using Cfg = FwdCfg<BN, MT, kTmaEpi, fwd_kb(BN, MT, sizeof(OutT) == 2, TF32, CL), PAIR>;
  auto kern = conv_fwd_kernel<BN, MT, OutT, TF32, CL, PAIR>;
  if (!kTmaEpi) p.tma_epi = p.side_tma = 0;
  static int max_clusters = 0;         // co-resident clusters of this kernel on this device (CL > 1)
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (CL > 1) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(CL * 148); q.blockDim = dim3(kConvThreads); q.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute a{};
      a.id = cudaLaunchAttributeClusterDimension;
      a.val.clusterDim.x = CL; a.val.clusterDim.y = 1; a.val.clusterDim.z = 1;
      q.attrs = &a; q.numAttrs = 1;
      MCGM_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q));
      if (max_clusters < 1) return fail(MCGM_E_CUDA, "no cluster of %d CTAs fits this device", CL);
    }
    configured = true;
  }
  int sms = sm_count();
  if (sms <= 0) return fail(MCGM_E_CUDA, "cannot query SM count");
  p.m_super = (p.m_super + MT - 1) / MT;           // in: M tiles per phase; out: CTA tiles per phase
  p.total_tiles = p.m_super * p.n_tiles * p.PH;
  p.m_super_c = (p.m_super + CL - 1) / CL;
  p.total_ctiles = p.m_super_c * p.n_tiles * p.PH;
  if (CL == 1) {
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    kern<<<grid, kConvThreads, Cfg::kSmemBytes, st>>>(maps.x, maps.w, maps.out, maps.side, p);
    MCGM_LAUNCH_CHECK();
    return 0;
  }
  const int nclusters = p.total_ctiles < max_clusters ? p.total_ctiles : max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * CL); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  MCGM_CUDA(cudaLaunchKernelEx(&cfg, kern, maps.x, maps.w, maps.out, maps.side, p));
  return 0;
}

template <int BN, int MT>
static int launch_fwd_t(const FwdMaps& maps, const ConvFwdParams& p, bool f32, cudaStream_t st, int cl = 1) {
  if (p.tf32) return launch_fwd<BN, MT, float, true>(maps, p, st);      // fp32 parity pipeline
  if (f32) return launch_fwd<BN, MT, float>(maps, p, st);
  if constexpr ((BN == 256 && MT == 1) || (BN == 128 && MT == 2)) {     // the shapes that carry the flops: cluster variants
    if (cl == 2 && p.pair_mma) return launch_fwd<BN, MT, __nv_bfloat16, false, 2, true>(maps, p, st);
    if (cl == 2) return launch_fwd<BN, MT, __nv_bfloat16, false, 2>(maps, p, st);
    if (cl == 4) return launch_fwd<BN, MT, __nv_bfloat16, false, 4>(maps, p, st);
  }
  return launch_fwd<BN, MT, __nv_bfloat16>(maps, p, st);
}

template <int BN, int NT, int KB = 64>
static int launch_wgrad(const WgMaps& maps, const WgradParams& p, int grid, cudaStream_t st) {
  using Cfg = WgCfg<BN, NT, KB>;
  auto kern = conv_wgrad_kernel<BN, NT, KB>;
  static bool configured = false;
  if (!configured) {
    MCGM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  kern<<<grid, kWgradThreads, Cfg::kSmemBytes, st>>>(maps, p);
  MCGM_LAUNCH_CHECK();
  return 0;
}

static constexpr int kWgMaxBiasSlots = 16;     // bias-gradient partial rows per split the workspace is sized for

struct WgradPlan {
  int BN, NT, tap_groups, cout_tiles, cin_tiles, taps, splits, kb_per_split, total_kb, KW, KH, KN;
  int Hg, Wg, AS, BS;           // grid dims and traversal strides of dY / X
  bool a_shift;                 // dY offsets differ per tap (phase conv): no tap pairing
  size_t per_split;  // floats
};

// H, W are the dims of X
static int plan_wgrad(int kind, int tf32, int N, int H, int W, int Cin, int Cout, int R, int S, WgradPlan* pl, int groups = 1,
                      int n_taps = 0) {
  if (Cin % 64 || Cout % 64) return fail(MCGM_E_UNSUPPORTED, "wgrad: Cin=%d and Cout=%d must be multiples of 64", Cin, Cout);
  pl->AS = pl->BS = 1;
  pl->a_shift = false;
  if (kind == MCGM_CONV_S1) {
    if (!((R == 1 && S == 1) || (R == 3 && S == 3))) return fail(MCGM_E_UNSUPPORTED, "wgrad: only 1x1 and 3x3 stride-1 kernels");
    pl->Hg = H; pl->Wg = W; pl->taps = R * S;
  } else if (kind == MCGM_CONV_S2K4) {
    if (H % 2 || W % 2) return fail(MCGM_E_ARG, "wgrad: stride-2 conv needs even H, W");
    pl->Hg = H / 2; pl->Wg = W / 2; pl->taps = 16; pl->BS = 2;
  } else if (kind == MCGM_CONV_PHASE2) {
    pl->Hg = H; pl->Wg = W; pl->taps = 16; pl->AS = 2; pl->a_shift = true;
  } else if (kind == MCGM_CONV_TAPS) {
    if (n_taps < 1 || n_taps > kWgMaxTaps) return fail(MCGM_E_UNSUPPORTED, "wgrad: 1..%d explicit taps", kWgMaxTaps);
    pl->Hg = H; pl->Wg = W; pl->taps = n_taps;
  } else {
    return fail(MCGM_E_ARG, "wgrad: unknown kind %d", kind);
  }
  // k-block = 64 grid pixels of bf16 or 32 of fp32: its operand tiles are the same bytes either way
  if (!pick_pixel_box(pl->Hg, pl->Wg, (tf32 || MCGM_WG_KB == 32) ? 32 : 64, &pl->KW, &pl->KH, &pl->KN))
    return fail(MCGM_E_UNSUPPORTED, "wgrad: grid H=%d W=%d cannot be tiled into k-block boxes", pl->Hg, pl->Wg);
  pl->BN = (Cin % 256 == 0) ? 256 : (Cin % 128 == 0 ? 128 : 64);
  pl->cin_tiles = Cin / pl->BN;
  pl->cout_tiles = (Cout + kTileM - 1) / kTileM;
  const int n_boxes = (N + pl->KN - 1) / pl->KN;
  pl->total_kb = n_boxes * (pl->Hg / pl->KH) * (pl->Wg / pl->KW);
  pl->NT = (g_knobs.nt2 && pl->taps > 1 && pl->BN <= 128 && !pl->a_shift) ? 2 : 1;
  pl->tap_groups = (pl->taps + pl->NT - 1) / pl->NT;
  const int tiles = pl->cout_tiles * pl->cin_tiles * pl->tap_groups;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int splits = sms / (tiles * groups);             // at most one wave of CTAs (all groups of a grouped launch together)
  if (g_knobs.wg_splits) splits = (int)g_knobs.wg_splits;
  if (splits > pl->total_kb) splits = pl->total_kb;
  // one wave of CTAs; the cap only binds for single-tile problems (thin / 1x1 weight gradients), whose
  // partials are small — capping at 64 left 84 SMs idle on an HBM-bound reduction
  if (splits > sms) splits = sms;
  if (splits < 1) splits = 1;
  pl->kb_per_split = (pl->total_kb + splits - 1) / splits;
  pl->splits = (pl->total_kb + pl->kb_per_split - 1) / pl->kb_per_split;
  pl->per_split = (size_t)Cout * pl->taps * Cin;
  return 0;
}

}  // namespace mcgm

using namespace mcgm;

namespace mcgm { extern int g_thin_umma_wgrad, g_thin_umma_fwd; }

extern "C" int mcgm_debug_set(int key, uint64_t value) {
  switch (key) {
    case 0: g_knobs.k_lbo = (uint32_t)value; break;
    case 1: g_knobs.k_sbo = (uint32_t)value; break;
    case 2: g_knobs.mn_lbo = (uint32_t)value; break;
    case 3: g_knobs.mn_sbo = (uint32_t)value; break;
    case 4: g_knobs.version = (uint32_t)value; break;
    case 5: g_knobs.layout = (uint32_t)value; break;
    case 6: g_knobs.mt2 = (uint32_t)value; break;
    case 7: g_knobs.mt2_256 = (uint32_t)value; break;
    case 8: g_knobs.nt2 = (uint32_t)value; break;
    case 9: g_knobs.wg_splits = (uint32_t)value; break;
    case 10: g_knobs.tma_epi = (uint32_t)value; break;
    case 11: g_knobs.cluster = (uint32_t)value; break;
    case 14: g_knobs.wg_pair = (uint32_t)value; break;
    case 15: g_knobs.fwd_dbg = (uint32_t)value; break;
    case 16: g_knobs.pair_mma = (uint32_t)value; break;
    case 17: g_knobs.wg_blocked = (uint32_t)value; break;
    case 18: g_knobs.wg_bias_last = (uint32_t)value; break;
    case 12: g_thin_umma_wgrad = (int)value; break;       // thin (RGB-side) convs: tcgen05 / TMA kernels (1) or mma.sync kernels (0)
    case 13: g_thin_umma_fwd = (int)value; break;
    default: return fail(MCGM_E_ARG, "debug_set: unknown key %d", key);
  }
  return 0;
}

extern "C" int mcgm_pack_weight(const float* w_oihw, int K, int C, int R, int S, const float* inv_sigma,
                                void* w_krsc_bf16, void* w_crsk_flip_bf16, void* stream) {
  MCGM_CHECK_ARG(w_oihw && (w_krsc_bf16 || w_crsk_flip_bf16) && K > 0 && C > 0 && R > 0 && S > 0, "pack_weight: bad arguments");
  const long total = (long)K * C * R * S;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_weight_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, K, C, R, S, inv_sigma,
                                                                   (__nv_bfloat16*)w_krsc_bf16,
                                                                   (__nv_bfloat16*)w_crsk_flip_bf16);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_unpack_wgrad(const float* dw_krsc, int K, int C, int R, int S, float* dw_oihw, int accumulate,
                                 void* stream) {
  MCGM_CHECK_ARG(dw_krsc && dw_oihw && K > 0 && C > 0 && R > 0 && S > 0, "unpack_wgrad: bad arguments");
  const long total = (long)K * C * R * S;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  unpack_wgrad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dw_krsc, K, C, R, S, dw_oihw, accumulate);
  MCGM_LAUNCH_CHECK();
  return 0;
}

// kind: MCGM_CONV_S1 (stride-1 RxS, R=S in {1,3}), MCGM_CONV_S2K4 (4x4 stride 2 pad 1: out = in/2),
// MCGM_CONV_PHASE2 (4 output phases x 2x2 taps: out = 2*in).  H, W are the INPUT dims.
// in_dtype MCGM_BF16: x, w, res, aux bf16.  MCGM_F32: all fp32, operands read by the tensor cores as
// TF32 (kind::tf32, K = 8 per instruction) — the fp32 parity pipeline; out_dtype must be fp32.
static int conv_fwd_impl(int kind, int in_dtype, const void* x, const void* w, void* out, int out_dtype, int N, int H, int W,
                         int Cin, int Cout, int R, int S, const float* bias, const void* res, const void* aux,
                         const int32_t* labels, const uint32_t* bits, int M, int flags, void* stream,
                         float* stats = nullptr, int* stat_slots_host = nullptr,
                         const signed char* taps_dh = nullptr, const signed char* taps_dw = nullptr, int n_taps = 0) {
  MCGM_CHECK_ARG(x && w && out && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv_fwd: bad arguments");
  MCGM_CHECK_ARG(out_dtype == MCGM_BF16 || out_dtype == MCGM_F32, "conv_fwd: out_dtype %d unsupported", out_dtype);
  MCGM_CHECK_ARG(in_dtype == MCGM_BF16 || in_dtype == MCGM_F32, "conv_fwd: in_dtype %d unsupported", in_dtype);
  const int tf32 = in_dtype == MCGM_F32;
  const int kblk = tf32 ? 32 : 64;
  if (tf32) MCGM_CHECK_ARG(out_dtype == MCGM_F32, "conv_fwd: fp32 operands need an fp32 output");
  if (Cin % kblk) return fail(MCGM_E_UNSUPPORTED, "conv_fwd: Cin=%d must be a multiple of %d (use the thin kernels)", Cin, kblk);
  if (flags & MCGM_EPI_BIAS) MCGM_CHECK_ARG(bias, "conv_fwd: BIAS flag without bias");
  if (flags & MCGM_EPI_TANH) MCGM_CHECK_ARG(out_dtype == MCGM_F32, "conv_fwd: TANH needs an fp32 output");
  if (flags & (MCGM_EPI_RES | MCGM_EPI_RES_UP2)) MCGM_CHECK_ARG(res && (Cout % 8 == 0), "conv_fwd: RES flag needs res and Cout %% 8 == 0");
  if (flags & MCGM_EPI_MASK) MCGM_CHECK_ARG(labels && bits && M > 0 && Cout % 32 == 0, "conv_fwd: MASK needs labels, bits and Cout %% 32 == 0");
  if (flags & MCGM_EPI_POSMASK) MCGM_CHECK_ARG(aux && (Cout % 8 == 0), "conv_fwd: POSMASK needs aux and Cout %% 8 == 0");

  ConvFwdParams p{};
  p.N = N; p.Cin = Cin; p.Cout = Cout;
  p.IS = 1; p.OS = 1; p.PH = 1;
  p.tf32 = tf32; p.kblk = kblk;
  if (kind == MCGM_CONV_S1) {
    if (!((R == 1 && S == 1) || (R == 3 && S == 3))) return fail(MCGM_E_UNSUPPORTED, "conv_fwd: only 1x1 and 3x3 stride-1 kernels (got %dx%d)", R, S);
    p.Hg = H; p.Wg = W; p.Ho = H; p.Wo = W; p.T = R * S;
    for (int t = 0; t < p.T; ++t) { p.tap_dh[t] = (signed char)(t / S - (R - 1) / 2); p.tap_dw[t] = (signed char)(t % S - (S - 1) / 2); }
  } else if (kind == MCGM_CONV_S2K4) {
    MCGM_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "conv_fwd: stride-2 conv needs even H, W");
    p.Hg = H / 2; p.Wg = W / 2; p.Ho = H / 2; p.Wo = W / 2; p.T = 16; p.IS = 2;
    for (int t = 0; t < 16; ++t) { p.tap_dh[t] = (signed char)(t / 4 - 1); p.tap_dw[t] = (signed char)(t % 4 - 1); }
  } else if (kind == MCGM_CONV_PHASE2) {
    p.Hg = H; p.Wg = W; p.Ho = 2 * H; p.Wo = 2 * W; p.T = 4; p.OS = 2; p.PH = 4;
    for (int ph = 0; ph < 4; ++ph)
      for (int t = 0; t < 4; ++t) {
        p.tap_dh[ph * kMaxTaps + t] = (signed char)(t / 2 - 1 + (ph >> 1));
        p.tap_dw[ph * kMaxTaps + t] = (signed char)(t % 2 - 1 + (ph & 1));
      }
  } else if (kind == MCGM_CONV_TAPS) {
    // stride-1 conv with an explicit list of (dh, dw) input offsets per tap: the causal / cropped stacks of the
    // gated PixelCNN (src/models/mcpixelcnn.py:31-36, 48-52) — weight [Cout, n_taps, Cin]
    MCGM_CHECK_ARG(taps_dh && taps_dw && n_taps >= 1 && n_taps <= 4 * kMaxTaps, "conv_fwd: TAPS needs 1..%d tap offsets", 4 * kMaxTaps);
    p.Hg = H; p.Wg = W; p.Ho = H; p.Wo = W; p.T = n_taps;
    for (int t = 0; t < n_taps; ++t) { p.tap_dh[t] = taps_dh[t]; p.tap_dw[t] = taps_dw[t]; }
  } else {
    return fail(MCGM_E_ARG, "conv_fwd: unknown kind %d", kind);
  }
  if (flags & MCGM_EPI_RES_UP2) MCGM_CHECK_ARG(p.Ho % 2 == 0 && p.Wo % 2 == 0, "conv_fwd: RES_UP2 needs even output H, W");
  if (!pick_pixel_box(p.Hg, p.Wg, kTileM, &p.TW, &p.TH, &p.TN))
    return fail(MCGM_E_UNSUPPORTED, "conv_fwd: tile grid H=%d W=%d cannot be tiled into 128-pixel boxes", p.Hg, p.Wg);
  if (p.TW * p.IS > 256 || p.TH * p.IS > 256) return fail(MCGM_E_UNSUPPORTED, "conv_fwd: TMA box too large for stride %d", p.IS);
  p.tiles_w = p.Wg / p.TW;
  p.tiles_h = p.Hg / p.TH;
  p.quad = (p.TW % 4 == 0);
  if ((long)N * p.Ho * p.Wo >= (1l << 31)) return fail(MCGM_E_UNSUPPORTED, "conv_fwd: more than 2^31 output pixels");
  const int tiles_nb = (N + p.TN - 1) / p.TN;
  int BN;
  if (Cout % 256 == 0) BN = 256;
  else if (Cout % 128 == 0) BN = 128;
  else if (Cout % 64 == 0) BN = 64;
  else if (Cout == 32) BN = 32;
  else if (Cout <= 16) BN = 16;
  else return fail(MCGM_E_UNSUPPORTED, "conv_fwd: Cout=%d unsupported (multiple of 64, 32, or <= 16)", Cout);
  // phase conv: the weight rows of phase ph start at ph*Cout.  With Cout < 16 (RGB output of a
  // ConvTranspose2d) the 16-row box also covers the next phases' rows; those accumulator columns are
  // never stored (the epilogue stops at Cout).
  if (p.PH > 1 && Cout % BN && Cout > 16) return fail(MCGM_E_UNSUPPORTED, "conv_fwd: phase conv needs Cout %% %d == 0", BN);
  p.n_tiles = (Cout + BN - 1) / BN;
  if (p.n_tiles * BN * 4 > kMaxBiasBytes) return fail(MCGM_E_UNSUPPORTED, "conv_fwd: Cout=%d exceeds the %d-channel bias staging buffer", Cout, kMaxBiasBytes / 4);
  const int m_tiles = p.tiles_w * p.tiles_h * tiles_nb;
  p.kpt = Cin / kblk;
  p.num_kb = p.T * p.kpt;
  p.desc_templ = make_smem_desc_hi_lbo(g_knobs.k_lbo, g_knobs.k_sbo, g_knobs.version, g_knobs.layout);
  p.idesc = make_idesc_bf16(kTileM, BN, 0, 0, tf32);
  p.bias = bias; p.res = (const __nv_bfloat16*)res; p.aux = (const __nv_bfloat16*)aux;
  p.labels = labels; p.bits = bits; p.words = (Cout + 31) / 32; p.flags = flags; p.out = out;

  cudaStream_t st = (cudaStream_t)stream;
  const bool f32 = out_dtype == MCGM_F32;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  // two M tiles per CTA (halves the weight bytes per MMA) whenever single tiles would need > 1 wave
  const bool pair = g_knobs.mt2 && (m_tiles * p.n_tiles * p.PH > sms);
  // clusters with multicast weights for the layers that carry the flops, when every SM gets several tiles anyway
  int cl = 1;
  if (!f32 && !tf32 && ((BN == 256 && !(pair && g_knobs.mt2_256)) || (BN == 128 && pair)) &&
      (g_knobs.cluster == 2 || g_knobs.cluster == 4) && (long)m_tiles * p.n_tiles * p.PH >= 4L * sms)
    cl = (int)g_knobs.cluster;
  // CTA pairs for the tiles that carry the flops (ragged tile counts are padded with a dummy tile, see tile_of)
  const bool mt2_sel = (BN == 256) ? (pair && g_knobs.mt2_256) : pair;
  if (!f32 && !tf32 && cl == 1 && m_tiles >= 2 * (mt2_sel ? 2 : 1) &&
      ((BN == 256 && !mt2_sel && (g_knobs.pair_mma & 1u)) || (BN == 128 && mt2_sel && (g_knobs.pair_mma & 2u)))) {
    cl = 2;
    p.pair_mma = 1;
    p.idesc = make_idesc_bf16(2 * kTileM, BN, 0, 0, 0);
  }
  FwdMaps maps{};
  // the k-block of the instantiation launch_fwd_t will pick (fwd_kb): 64-byte rows travel with SWIZZLE_64B
  const int mt_sel = (BN == 256) ? ((pair && g_knobs.mt2_256) ? 2 : 1) : (pair ? 2 : 1);
  const int kb_sel = fwd_kb(BN, mt_sel, !f32, tf32, cl);
  int op_swizzle = 2;
  if (kb_sel == 32) {
    op_swizzle = 1;
    p.kblk = 32; p.kpt = Cin / 32; p.num_kb = p.T * p.kpt;
    p.desc_templ = make_smem_desc_hi_lbo(g_knobs.k_lbo, 512, g_knobs.version, 4);    // SBO = 8 rows x 64 B, layout SWIZZLE_64B
  }
  const int kblk_sel = (kb_sel == 32) ? 32 : kblk;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint32_t box[4] = {(uint32_t)kblk_sel, (uint32_t)(p.TW * p.IS), (uint32_t)(p.TH * p.IS), (uint32_t)p.TN};
    if (int e = make_tmap(&maps.x, x, 4, dims, box, p.IS, tf32, op_swizzle)) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)p.T * Cin, (uint64_t)p.PH * Cout};
    const uint32_t box[2] = {(uint32_t)kblk_sel, (uint32_t)(BN / cl)};      // a cluster CTA fetches BN / cl weight rows
    if (int e = make_tmap(&maps.w, w, 2, dims, box, 1, tf32, op_swizzle)) return e;
  }
  // Staged epilogue (bf16 tiles of >= 64 channels): the output tile of an M tile is the box (64 ch, TW*OS, TH*OS, TN) at
  // (c, w0*OS + pw, h0*OS + ph, n0) of the NHWC tensor, traversal stride OS — the store-side twin of the operand loads.
  // The residual / ReLU-sign tensor has the output's geometry, so (same resolution only) its tile is loaded by the same box.
  const bool both_sides = (flags & (MCGM_EPI_RES | MCGM_EPI_RES_UP2)) && (flags & MCGM_EPI_POSMASK);
  const bool tma_ok = !f32 && !tf32 && BN >= 64;
  if (tma_ok && both_sides)
    return fail(MCGM_E_UNSUPPORTED, "conv_fwd: a residual together with POSMASK is only available with an fp32 output");
  if (tma_ok) {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)N};
    const uint32_t box[4] = {64u, (uint32_t)(p.TW * p.OS), (uint32_t)(p.TH * p.OS), (uint32_t)p.TN};
    if (int e = make_tmap(&maps.out, out, 4, dims, box, p.OS, 0)) return e;
    p.tma_epi = 1;
    const void* side = (flags & MCGM_EPI_RES) ? res : ((flags & MCGM_EPI_POSMASK) ? aux : nullptr);
    if (side != nullptr && p.OS == 1 && (g_knobs.tma_epi & 4u)) {
      if (int e = make_tmap(&maps.side, side, 4, dims, box, 1, 0)) return e;
      p.side_tma = 1;
    }
  }
  p.dbg = (int)g_knobs.fwd_dbg;
  p.m_super = m_tiles;           // launch_fwd turns this into ceil(m_tiles / MT)
  if (stats != nullptr) {
    // batch statistics of the STORED output for the BatchNorm that consumes it, accumulated by the staged epilogue
    if (!(p.tma_epi && p.n_tiles == 1))
      return fail(MCGM_E_UNSUPPORTED, "conv_fwd_bnstats: needs a bf16 output with Cout in {64, 128, 256} (one channel tile, staged epilogue)");
    MCGM_CHECK_ARG(stat_slots_host, "conv_fwd_bnstats: stat_slots_host is NULL");
    // one slot for every CTA the device can hold; CTAs that are not launched leave zeros
    *stat_slots_host = sms;
    MCGM_CUDA(cudaMemsetAsync(stats, 0, (size_t)sms * 2 * Cout * sizeof(float), st));
    p.stats = stats;
  }
  switch (BN) {
    case 256: return (pair && g_knobs.mt2_256) ? launch_fwd_t<256, 2>(maps, p, f32, st) : launch_fwd_t<256, 1>(maps, p, f32, st, cl);
    case 128: return pair ? launch_fwd_t<128, 2>(maps, p, f32, st, cl) : launch_fwd_t<128, 1>(maps, p, f32, st);
    case 64: return pair ? launch_fwd_t<64, 2>(maps, p, f32, st) : launch_fwd_t<64, 1>(maps, p, f32, st);
    case 32: return launch_fwd_t<32, 1>(maps, p, f32, st);
    default: return launch_fwd_t<16, 1>(maps, p, f32, st);
  }
}

extern "C" int mcgm_conv_fwd(const void* x, const void* w, void* out, int out_dtype, int N, int H, int W, int Cin,
                             int Cout, int R, int S, const float* bias, const void* res, const void* aux,
                             const int32_t* labels, const uint32_t* bits, int M, int flags, void* stream) {
  return conv_fwd_impl(MCGM_CONV_S1, MCGM_BF16, x, w, out, out_dtype, N, H, W, Cin, Cout, R, S, bias, res, aux, labels, bits, M, flags, stream);
}

extern "C" int mcgm_conv_fwd_ex(int kind, int in_dtype, const void* x, const void* w, void* out, int out_dtype,
                                int N, int H, int W, int Cin, int Cout, int R, int S, const float* bias,
                                const void* res, const void* aux, const int32_t* labels, const uint32_t* bits,
                                int M, int flags, void* stream) {
  return conv_fwd_impl(kind, in_dtype, x, w, out, out_dtype, N, H, W, Cin, Cout, R, S, bias, res, aux, labels, bits, M, flags, stream);
}

extern "C" int mcgm_conv_fwd_taps(const void* x, const void* w, void* out, int out_dtype, int N, int H, int W, int Cin,
                                  int Cout, int n_taps, const signed char* taps_dh_host, const signed char* taps_dw_host,
                                  const float* bias, const void* res, const void* aux, const int32_t* labels,
                                  const uint32_t* bits, int M, int flags, void* stream) {
  return conv_fwd_impl(MCGM_CONV_TAPS, MCGM_BF16, x, w, out, out_dtype, N, H, W, Cin, Cout, 0, 0, bias, res, aux, labels, bits, M,
                       flags, stream, nullptr, nullptr, taps_dh_host, taps_dw_host, n_taps);
}

extern "C" size_t mcgm_conv_fwd_bnstats_bytes(int Cout) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  return (size_t)4 * sms * 2 * Cout * sizeof(float);
}

extern "C" int mcgm_conv_fwd_bnstats(int kind, const void* x, const void* w, void* out, int N, int H, int W, int Cin,
                                     int Cout, int R, int S, const float* bias, const void* res, int flags,
                                     float* stats_partial, int* stat_slots_host, void* stream) {
  MCGM_CHECK_ARG(stats_partial && stat_slots_host, "conv_fwd_bnstats: stats_partial / stat_slots_host are NULL");
  return conv_fwd_impl(kind, MCGM_BF16, x, w, out, MCGM_BF16, N, H, W, Cin, Cout, R, S, bias, res, nullptr, nullptr, nullptr,
                       0, flags, stream, stats_partial, stat_slots_host);
}

static size_t wgrad_ws_bytes(int kind, int in_dtype, int N, int H, int W, int Cin, int Cout, int R, int S) {
  WgradPlan pl;
  if (plan_wgrad(kind, in_dtype == MCGM_F32, N, H, W, Cin, Cout, R, S, &pl)) return 0;
  size_t bytes = (size_t)pl.splits * pl.per_split * sizeof(float);
  bytes += (size_t)pl.splits * kWgMaxBiasSlots * Cout * sizeof(float);  // bias-gradient partials (4 parity slots, or one per tap group)
  return bytes;
}

static int conv_wgrad_grouped_impl(int kind, int in_dtype, int n, const void* const* xs, const void* const* dys,
                                   float* const* dws, float* const* dbs, int N, int H, int W, int Cin, int Cout, int R, int S,
                                   void* workspace, void* stream, const signed char* taps_dh = nullptr,
                                   const signed char* taps_dw = nullptr, int n_taps = 0) {
  MCGM_CHECK_ARG(n >= 1 && n <= kWgMaxGroups && xs && dys && dws && workspace && N > 0, "conv_wgrad: bad arguments (1..%d groups)", kWgMaxGroups);
  // MN-major TF32 operands need the 128B/32B-atom swizzle on both the TMA and the UMMA side; with the plain
  // 128B swizzle used here the tensor cores return zeros (measured).  The fp32 parity pipeline therefore
  // computes weight gradients from bf16x3 operand pieces on this (bf16) kernel instead.
  if (in_dtype != MCGM_BF16) return fail(MCGM_E_UNSUPPORTED, "conv_wgrad: only bf16 operands (fp32 parity mode uses bf16x3 pieces)");
  const int tf32 = 0;
  const bool want_bias = dbs != nullptr && dbs[0] != nullptr;
  for (int g = 0; g < n; ++g)
    MCGM_CHECK_ARG(xs[g] && dys[g] && dws[g] && (!want_bias || dbs[g]), "conv_wgrad: group %d has a NULL pointer", g);
  WgradPlan pl;
  if (int e = plan_wgrad(kind, tf32, N, H, W, Cin, Cout, R, S, &pl, n, n_taps)) return e;
  if (kind == MCGM_CONV_TAPS) MCGM_CHECK_ARG(taps_dh && taps_dw, "conv_wgrad: TAPS needs the tap offsets");
  WgradParams p{};
  p.N = N; p.Cin = Cin; p.Cout = Cout; p.AS = pl.AS; p.BS = pl.BS;
  for (int t = 0; t < pl.taps; ++t) {
    if (kind == MCGM_CONV_S1) { p.b_dh[t] = (signed char)(t / S - (R - 1) / 2); p.b_dw[t] = (signed char)(t % S - (S - 1) / 2); }
    else if (kind == MCGM_CONV_S2K4) { p.b_dh[t] = (signed char)(t / 4 - 1); p.b_dw[t] = (signed char)(t % 4 - 1); }
    else if (kind == MCGM_CONV_TAPS) { p.b_dh[t] = taps_dh[t]; p.b_dw[t] = taps_dw[t]; }
    else { p.a_dh[t] = (signed char)(t / 4 - 1); p.a_dw[t] = (signed char)(t % 4 - 1); }
  }
  p.KW = pl.KW; p.KH = pl.KH; p.KN = pl.KN; p.kb_w = pl.Wg / pl.KW; p.kb_h = pl.Hg / pl.KH;
  p.total_kb = pl.total_kb; p.splits = pl.splits; p.kb_per_split = pl.kb_per_split;
  p.cout_tiles = pl.cout_tiles; p.cin_tiles = pl.cin_tiles; p.taps = pl.taps; p.tap_groups = pl.tap_groups;
  p.blocked_maps = g_knobs.wg_blocked ? 1 : 0;
  constexpr bool kb32 = (MCGM_WG_KB == 32);            // bf16 k-blocks of 32 pixels: MN blocks are 4 KB apart
  p.desc_templ = make_smem_desc_hi_lbo((tf32 || kb32) ? 4096 : g_knobs.mn_lbo, g_knobs.mn_sbo, g_knobs.version, g_knobs.layout);
  p.idesc = make_idesc_bf16(kTileM, pl.BN, 1, 1, tf32);
  p.idesc_pair = (g_knobs.wg_pair && 2 * pl.BN <= 256) ? make_idesc_bf16(kTileM, 2 * pl.BN, 1, 1, tf32) : 0u;
  p.idesc_bias = make_idesc_bf16(kTileM, 16, 1, 1, tf32);
  p.bias_slots = 1; p.bias_taps[0] = 0;
  // The bias gradient (one extra N = 16 MMA per k step) rides on the tap group with the LEAST work: with two taps per CTA
  // a 3x3 kernel has four two-tap groups and one single-tap group, whose CTAs issue half the MMAs.  On group 0 the extra
  // MMAs made those CTAs the critical path of the one-wave launch (+29 us at 128->128 32x32, knob 18 = 0 restores it).
  if (g_knobs.wg_bias_last && pl.NT == 2 && (pl.taps & 1)) p.bias_taps[0] = pl.taps - 1;
  // One tap per CTA (256-wide layers): every group has the same work, so the bias MMAs are dealt round-robin over all of
  // them (k-block kb -> group kb % tap_groups): 576 -> ~420 us at 256->256 32x32 instead of +44 % on group 0's CTAs.
  if (g_knobs.wg_bias_last && pl.NT == 1 && pl.taps > 1 && !pl.a_shift && pl.tap_groups <= kWgMaxBiasSlots) {
    p.bias_spread = 1;
    p.bias_slots = pl.tap_groups;
  }
  if (kind == MCGM_CONV_PHASE2) {          // dY is read with stride 2: the four offsets in {0,1}^2 cover it once
    p.bias_slots = 4; p.bias_taps[0] = 5; p.bias_taps[1] = 6; p.bias_taps[2] = 9; p.bias_taps[3] = 10;
  }
  // workspace: [groups][splits][per_split] weight partials, then [groups][splits * bias_slots][Cout] bias partials
  p.groups = n;
  p.ctas_per_group = pl.cout_tiles * pl.cin_tiles * pl.tap_groups * pl.splits;
  p.group_partial_stride = (size_t)pl.splits * pl.per_split;
  p.group_bias_stride = (size_t)pl.splits * p.bias_slots * Cout;
  p.partial = (float*)workspace;
  p.bias_partial = want_bias ? p.partial + (size_t)n * p.group_partial_stride : nullptr;

  // dims of dY and X: the grid scaled by the traversal strides
  const int Hdy = pl.Hg * pl.AS, Wdy = pl.Wg * pl.AS;
  WgMaps maps{};
  WgReduceOut out{};
  for (int g = 0; g < n; ++g) {
    if (p.blocked_maps) {
      if (int e = make_tmap_blocked(&maps.dy[g], dys[g], Cout, Wdy, Hdy, N, pl.KW * pl.AS, pl.KH * pl.AS, pl.KN, kTileM / 64, pl.AS)) return e;
      if (int e = make_tmap_blocked(&maps.x[g], xs[g], Cin, W, H, N, pl.KW * pl.BS, pl.KH * pl.BS, pl.KN, pl.BN / 64, pl.BS)) return e;
    } else {
      {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Wdy, (uint64_t)Hdy, (uint64_t)N};
        const uint32_t box[4] = {64u, (uint32_t)(pl.KW * pl.AS), (uint32_t)(pl.KH * pl.AS), (uint32_t)pl.KN};
        if (int e = make_tmap(&maps.dy[g], dys[g], 4, dims, box, pl.AS, tf32)) return e;
      }
      {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint32_t box[4] = {64u, (uint32_t)(pl.KW * pl.BS), (uint32_t)(pl.KH * pl.BS), (uint32_t)pl.KN};
        if (int e = make_tmap(&maps.x[g], xs[g], 4, dims, box, pl.BS, tf32)) return e;
      }
    }
    out.dw[g] = dws[g];
    out.dbias[g] = want_bias ? dbs[g] : nullptr;
  }
  for (int g = n; g < kWgMaxGroups; ++g) { maps.dy[g] = maps.dy[0]; maps.x[g] = maps.x[0]; }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = p.ctas_per_group * n;
  int e;
  if (kb32 && !tf32) {
    switch (pl.BN) {
      case 256: e = launch_wgrad<256, 1, 32>(maps, p, grid, st); break;
      case 128: e = pl.NT == 2 ? launch_wgrad<128, 2, 32>(maps, p, grid, st) : launch_wgrad<128, 1, 32>(maps, p, grid, st); break;
      default: e = pl.NT == 2 ? launch_wgrad<64, 2, 32>(maps, p, grid, st) : launch_wgrad<64, 1, 32>(maps, p, grid, st); break;
    }
  } else
  switch (pl.BN) {
    case 256: e = launch_wgrad<256, 1>(maps, p, grid, st); break;
    case 128: e = pl.NT == 2 ? launch_wgrad<128, 2>(maps, p, grid, st) : launch_wgrad<128, 1>(maps, p, grid, st); break;
    default: e = pl.NT == 2 ? launch_wgrad<64, 2>(maps, p, grid, st) : launch_wgrad<64, 1>(maps, p, grid, st); break;
  }
  if (e) return e;
  const size_t n4 = pl.per_split / 4;
  long wblocks = (long)((n4 + 31) / 32);
  if (wblocks > 148 * 16) wblocks = 148 * 16;
  const int bblocks = want_bias ? (Cout + 31) / 32 : 0;
  wgrad_reduce_kernel<<<dim3((unsigned)(wblocks + bblocks), n), 32 * kRedLanes, 0, st>>>(
      p.partial, pl.splits, n4, out, p.bias_partial, pl.splits * p.bias_slots, Cout, (int)wblocks, p.group_partial_stride,
      p.group_bias_stride);
  MCGM_LAUNCH_CHECK();
  return 0;
}

static int conv_wgrad_impl(int kind, int in_dtype, const void* x, const void* dy, float* dw, float* dbias, int N, int H,
                           int W, int Cin, int Cout, int R, int S, void* workspace, void* stream) {
  MCGM_CHECK_ARG(x && dy && dw && workspace && N > 0, "conv_wgrad: bad arguments");
  return conv_wgrad_grouped_impl(kind, in_dtype, 1, &x, &dy, &dw, dbias ? &dbias : nullptr, N, H, W, Cin, Cout, R, S, workspace, stream);
}

// Grouped form: n <= 4 independent weight gradients of IDENTICAL geometry in one launch pair (host arrays of
// device pointers).  Workspace: mcgm_conv_wgrad_grouped_workspace_bytes.
extern "C" size_t mcgm_conv_wgrad_grouped_workspace_bytes(int n, int N, int H, int W, int Cin, int Cout, int R, int S) {
  WgradPlan pl;
  if (n < 1 || n > kWgMaxGroups || plan_wgrad(MCGM_CONV_S1, 0, N, H, W, Cin, Cout, R, S, &pl, n)) return 0;
  return ((size_t)n * pl.splits * pl.per_split + (size_t)n * pl.splits * kWgMaxBiasSlots * Cout) * sizeof(float);
}
extern "C" int mcgm_conv_wgrad_grouped(int n, const void* const* x_host, const void* const* dy_host, float* const* dw_host,
                                       float* const* dbias_host, int N, int H, int W, int Cin, int Cout, int R, int S,
                                       void* workspace, void* stream) {
  return conv_wgrad_grouped_impl(MCGM_CONV_S1, MCGM_BF16, n, x_host, dy_host, dw_host, dbias_host, N, H, W, Cin, Cout, R, S,
                                 workspace, stream);
}

// Explicit-tap weight gradient: dw fp32 [Cout, n_taps, Cin], x pixel of tap t = dy pixel + (dh, dw)[t].
extern "C" size_t mcgm_conv_wgrad_taps_workspace_bytes(int N, int H, int W, int Cin, int Cout, int n_taps) {
  WgradPlan pl;
  if (plan_wgrad(MCGM_CONV_TAPS, 0, N, H, W, Cin, Cout, 0, 0, &pl, 1, n_taps)) return 0;
  return ((size_t)pl.splits * pl.per_split + (size_t)pl.splits * kWgMaxBiasSlots * Cout) * sizeof(float);
}
extern "C" int mcgm_conv_wgrad_taps(const void* x, const void* dy, float* dw, float* dbias, int N, int H, int W, int Cin,
                                    int Cout, int n_taps, const signed char* taps_dh_host, const signed char* taps_dw_host,
                                    void* workspace, void* stream) {
  MCGM_CHECK_ARG(x && dy && dw && workspace && N > 0, "conv_wgrad_taps: bad arguments");
  return conv_wgrad_grouped_impl(MCGM_CONV_TAPS, MCGM_BF16, 1, &x, &dy, &dw, dbias ? &dbias : nullptr, N, H, W, Cin, Cout, 0, 0,
                                 workspace, stream, taps_dh_host, taps_dw_host, n_taps);
}

extern "C" size_t mcgm_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S) {
  return wgrad_ws_bytes(MCGM_CONV_S1, MCGM_BF16, N, H, W, Cin, Cout, R, S);
}
extern "C" int mcgm_conv_wgrad(const void* x, const void* dy, float* dw, float* dbias, int N, int H, int W, int Cin,
                               int Cout, int R, int S, void* workspace, void* stream) {
  return conv_wgrad_impl(MCGM_CONV_S1, MCGM_BF16, x, dy, dw, dbias, N, H, W, Cin, Cout, R, S, workspace, stream);
}
extern "C" size_t mcgm_conv_wgrad_ex_workspace_bytes(int kind, int in_dtype, int N, int H, int W, int Cin, int Cout,
                                                     int R, int S) {
  return wgrad_ws_bytes(kind, in_dtype, N, H, W, Cin, Cout, R, S);
}
extern "C" int mcgm_conv_wgrad_ex(int kind, int in_dtype, const void* x, const void* dy, float* dw, float* dbias,
                                  int N, int H, int W, int Cin, int Cout, int R, int S, void* workspace,
                                  void* stream) {
  return conv_wgrad_impl(kind, in_dtype, x, dy, dw, dbias, N, H, W, Cin, Cout, R, S, workspace, stream);
}
