/*
 * mcgm.h — C ABI of libmcgm.so, the B200 (sm_100a) implementation of the
 * MultimodalController-gated conv block hot path.
 *
 * The reference (diaoenmao/Multimodal-Controller-for-Generative-Models) has no
 * FFI of its own: its boundary is the Python nn.Module API
 * (src/modules/modules.py:49-85, src/models/mcgan.py:193-203).  Every entry
 * point below replaces the PyTorch library calls that one reference line (or a
 * run of lines) dispatches to; the citation sits on each declaration.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch / C++ types.
 *   - every data pointer is a DEVICE pointer unless the name ends in _host.
 *   - activations are NHWC ("channels last"): x[((n*H + h)*W + w)*C + c].
 *   - conv weights are KRSC: w[((k*R + r)*S + s)*C + c]  (k = out channel).
 *   - `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = ok, negative = error (see MCGM_E_*); the message is
 *     available from mcgm_last_error() (thread-local, valid until next call).
 *   - nothing here allocates device memory; callers pass workspaces whose size
 *     comes from the *_workspace_bytes() helpers.
 *   - dtype codes: MCGM_F32 = 0, MCGM_BF16 = 1.
 */
#ifndef MCGM_H_
#define MCGM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCGM_F32 0
#define MCGM_BF16 1

/* convolution kinds of the *_ex entry points */
#define MCGM_CONV_S1 0        /* stride 1, R x S in {1x1, 3x3}, pad (R-1)/2                         */
#define MCGM_CONV_S2K4 1      /* 4x4 kernel, stride 2, pad 1: output = input / 2  (mcvae.py:41-49)   */
#define MCGM_CONV_PHASE2 2    /* 4 output phases x 2x2 taps: output = 2 x input.  Nearest-up2 -> Conv3x3
                                 (mcgan.py:17-19) and ConvTranspose2d(4,2,1) (mcvae.py:87-92)          */
#define MCGM_CONV_TAPS 3      /* stride 1, explicit (dh, dw) input offset per tap: the causal / cropped
                                 vertical and horizontal stacks of the gated PixelCNN (mcpixelcnn.py:31-52) */

#define MCGM_E_ARG (-1)      /* bad argument / unsupported shape */
#define MCGM_E_CUDA (-2)     /* CUDA runtime / driver error */
#define MCGM_E_UNSUPPORTED (-3)

/* gate flags */
#define MCGM_GATE_BN 1        /* y = a*x+b with a = gamma*rstd, b = beta - mean*a */
#define MCGM_GATE_RELU 2      /* ReLU before the mask */
#define MCGM_GATE_UP2 4       /* nearest 2x upsample on write (fwd) / 2x2 sum on read (bwd) */
#define MCGM_GATE_MASK 8      /* multiply by codebook[label[n], c] */
#define MCGM_GATE_AFFINE 16   /* with GATE_BN: (mean, rstd, gamma, beta) are fixed parameters (Glow's ActNorm,
                                 mcglow.py:24-55: y = scale*(x+loc) -> mean=-loc, rstd=1, gamma=scale, beta=0);
                                 backward returns dgamma = sum g*(x-mean)*rstd, dbeta = sum g, dx = gamma*rstd*g */

/* conv epilogue flags */
#define MCGM_EPI_BIAS 1
#define MCGM_EPI_RES 2        /* out += res[n,h,w,:]           (same resolution)   */
#define MCGM_EPI_RES_UP2 4    /* out += res[n,h/2,w/2,:]       (nearest-upsampled) */
#define MCGM_EPI_RELU 8
#define MCGM_EPI_MASK 16      /* out *= codebook[label[n], c]  */
#define MCGM_EPI_TANH 32
#define MCGM_EPI_POSMASK 64   /* out *= (aux[n,h,w,c] > 0)  — ReLU backward fused in dgrad */

const char* mcgm_last_error(void);
int mcgm_version(void);
/* number of SMs of the current device (grid sizing), or negative error */
int mcgm_sm_count(void);

/* ---- codebook ------------------------------------------------------------
 * Replaces `indicator.matmul(self.codebook)` (src/modules/modules.py:73): the
 * {0,1} fp32 codebook [M,C] is bit-packed once to [M, ceil(C/32)] uint32 words
 * (bit c%32 of word c/32); the mask of sample n is row label[n]. */
int mcgm_pack_codebook(const float* codebook, int M, int C, uint32_t* bits, void* stream);

/* ---- MultimodalController.forward (src/modules/modules.py:71-76) ----------
 * y[n,p,c] = x[n,p,c] * code(n,c), P = positions per sample (H*W, or 1 for 2-D x).
 * nchw = 1: x is NCHW-contiguous (the reference's layout), i.e. y[n,c,p] = x[n,c,p]*code(n,c).
 * code comes from (labels,bits) when labels != NULL, else from dense fp32 code[N,C]
 * (the general `indicator @ codebook` case).  Backward is the same call on dy. */
int mcgm_mc_forward(const void* x, void* y, int dtype, int nchw, int N, int P, int C,
                    const int32_t* labels, const uint32_t* bits, int M,
                    const float* code, void* stream);

/* ---- BatchNorm2d statistics (train mode; src/models/mcgan.py:15,20,56 →
 * torch native_batch_norm).  x is [P_total, C].  Writes mean[C], rstd[C]
 * (biased variance, eps inside the sqrt) and, when running_mean != NULL, the
 * momentum update with the UNBIASED variance.  workspace: see below. */
size_t mcgm_bn_stats_workspace_bytes(int C);
int mcgm_bn_stats(const void* x, int dtype, long P_total, int C,
                  float* mean, float* rstd,
                  float* running_mean, float* running_var, float momentum, float eps,
                  void* workspace, void* stream);

/* ---- fused gate forward: BN-apply → ReLU → (2x nearest upsample) → MC mask
 * (src/models/mcgan.py:15-18, 20-22, 56-58; D's ReLU→MC :104-105 with flags
 * without MCGM_GATE_BN).  x [N,H,W,C]; y [N,H',W',C] (H'=2H with UP2);
 * y_raw (optional, may be NULL) receives mask*x at the INPUT resolution — the
 * shortcut branch `mc_1(up(x))` of GenResBlock (mcgan.py:26-30) commuted with
 * the upsample, or DisResBlock's raw-gated shortcut (mcgan.py:111-115). */
int mcgm_gate_fwd(const void* x, void* y, void* y_raw, int dtype,
                  int N, int H, int W, int C,
                  const int32_t* labels, const uint32_t* bits, int M,
                  const float* mean, const float* rstd,
                  const float* gamma, const float* beta,
                  int flags, void* stream);

/* ---- fused gate backward.  dy [N,H',W',C]; d_raw (optional) is the gradient
 * that arrived at y_raw.  With MCGM_GATE_BN it runs the two-pass BN backward
 * (reduce: dgamma,dbeta; apply: dx) — both passes are launched by this call.
 * workspace: mcgm_gate_bwd_workspace_bytes(C). */
size_t mcgm_gate_bwd_workspace_bytes(int C);
int mcgm_gate_bwd(const void* dy, const void* d_raw, const void* x, void* dx, int dtype,
                  int N, int H, int W, int C,
                  const int32_t* labels, const uint32_t* bits, int M,
                  const float* mean, const float* rstd,
                  const float* gamma, const float* beta,
                  float* dgamma, float* dbeta,
                  int flags, void* workspace, void* stream);

/* ---- weight packing: fp32 OIHW master [K,C,R,S] (state_dict layout) →
 * bf16 KRSC for the forward conv and bf16 CRSK with flipped taps for dgrad
 * (either output may be NULL).  `inv_sigma` (device scalar, may be NULL)
 * multiplies the weight: the W/sigma of spectral norm (src/models/utils.py:17-21
 * → torch.nn.utils.spectral_norm compute_weight). */
int mcgm_pack_weight(const float* w_oihw, int K, int C, int R, int S,
                     const float* inv_sigma, void* w_krsc_bf16, void* w_crsk_flip_bf16,
                     void* stream);
/* fp32 KRSC gradient → fp32 OIHW (+= if accumulate) */
int mcgm_unpack_wgrad(const float* dw_krsc, int K, int C, int R, int S,
                      float* dw_oihw, int accumulate, void* stream);

/* ---- spectral normalisation, batched over layers (src/models/utils.py:17-21 -> torch legacy
 * spectral_norm compute_weight).  For layer i: W_host[i] fp32 [K_i, CRS_i] (weight_orig viewed as a
 * matrix), u_host[i] [K_i], v_host[i] [CRS_i] (the weight_u / weight_v buffers, updated IN PLACE when
 * do_iter).  state (device) receives per layer, at element offset state_off_host[i], the snapshot
 * [1/sigma | u | v] that forward packing and backward use.  The *_host arrays are HOST arrays of
 * device pointers / sizes; at most 16 layers per call. */
int mcgm_sn_power_iter(int n_layers, const float* const* W_host, const int* K_host, const int* CRS_host,
                       float* const* u_host, float* const* v_host, float* state,
                       const long* state_off_host, int do_iter, float eps, void* stream);
/* dW = (dW_bar - <dW_bar, W_bar> u v^T) / sigma.  dW_bar is [K][row_stride] fp32 whose rows hold
 * R*S*C entries in (r,s,c) order (krsc = 1, the layout of mcgm_conv_wgrad) or (c,r,s) order. */
size_t mcgm_sn_weight_bwd_workspace_bytes(void);
int mcgm_sn_weight_bwd(const float* dwbar, long row_stride, int krsc, const float* W_oihw,
                       const float* sn_state, int K, int C, int R, int S, float* dW_oihw,
                       void* workspace, void* stream);

/* Batched form for one whole backward pass (n <= 32 items; the *_host arrays are HOST arrays of device pointers /
 * sizes, read during the call): item i is (dwbar_i, row_stride_i, krsc_i, W_i, state_i, K_i, C_i, R_i*S_i) as in
 * mcgm_sn_weight_bwd, and its result is ADDED to grad_i (fp32 OIHW, the parameter's gradient).  Items that share
 * grad_i (the real- and fake-batch passes of one layer) are applied by the same threads in order: two launches in
 * total, deterministic. */
/* Batched mcgm_pack_weight / mcgm_pack_weight_padded for all spectral-normed conv layers of one forward (n <= 16;
 * *_host = HOST arrays): layer i is W_i (fp32 OIHW) scaled by inv_sigma_i[0] (device scalar, may be NULL).
 * mode_i = 0: out0_i = bf16 KRSC, out1_i = bf16 CRSK with flipped taps; mode_i = 1: out0_i = bf16
 * [K][row_stride_i] rows in (r,s,c) order, zero padded.  One launch. */
int mcgm_pack_weight_batched(int n, const float* const* W_host, const float* const* inv_sigma_host, const int* K_host,
                             const int* C_host, const int* R_host, const int* S_host, const int* mode_host,
                             const int* row_stride_host, void* const* out0_host, void* const* out1_host,
                             void* stream);

size_t mcgm_sn_weight_bwd_batched_workspace_bytes(void);
int mcgm_sn_weight_bwd_batched(int n, const float* const* dwbar_host, const long* row_stride_host, const int* krsc_host,
                               const float* const* W_host, const float* const* state_host, const int* K_host,
                               const int* C_host, const int* RS_host, float* const* grad_host, void* workspace,
                               void* stream);
/* bf16 [K][row_stride] <- W[k,c,r,s] * inv_sigma with columns in (r,s,c) order, zero padded: the
 * weight matrix of the im2col'd thin convolution (row_stride = 64). */
int mcgm_pack_weight_padded(const float* w_oihw, int K, int C, int R, int S, const float* inv_sigma,
                            void* out_bf16, int row_stride, void* stream);

/* ---- implicit-GEMM convolution, tcgen05 + TMA (replaces cudnn_convolution at
 * src/models/mcgan.py:19,23,29,59,77-85,103-128).  stride 1, pad (R-1)/2,
 * R=S in {1,3}; x bf16 NHWC [N,H,W,Cin] with Cin % 64 == 0; w bf16 KRSC
 * [Cout,R,S,Cin]; out NHWC [N,H,W,Cout], out_dtype bf16 or f32.
 * dgrad is the same call with the flipped CRSK weight from mcgm_pack_weight.
 * Epilogue (flags MCGM_EPI_*): +bias[Cout] (fp32), +res (bf16 NHWC, same-res
 * or half-res), ReLU, MC mask by (labels,bits), tanh, (aux>0) masking. */
int mcgm_conv_fwd(const void* x, const void* w, void* out, int out_dtype,
                  int N, int H, int W, int Cin, int Cout, int R, int S,
                  const float* bias, const void* res, const void* aux,
                  const int32_t* labels, const uint32_t* bits, int M,
                  int flags, void* stream);

/* Generalised geometry on the same kernel (H, W are INPUT dims):
 *   MCGM_CONV_S2K4  : w bf16 [Cout,4,4,Cin]; out[n,i,j] = sum_{p,q} x[n,2i+p-1,2j+q-1] . w[:,p,q,:]
 *                     (TMA traversal stride 2 gathers the strided pixels; no im2col, no space-to-depth)
 *   MCGM_CONV_PHASE2: w bf16 [4 phases (ph*2+pw)][Cout][2][2][Cin];
 *                     out[n,2i+ph,2j+pw] = sum_{a,b} x[n,i+a-1+ph,j+b-1+pw] . w[ph*2+pw,:,a,b,:]
 *                     — 2.25x fewer MACs than convolving the materialised upsampled tensor. */
int mcgm_conv_fwd_ex(int kind, int in_dtype, const void* x, const void* w, void* out, int out_dtype,
                     int N, int H, int W, int Cin, int Cout, int R, int S,
                     const float* bias, const void* res, const void* aux,
                     const int32_t* labels, const uint32_t* bits, int M,
                     int flags, void* stream);
/* in_dtype = MCGM_F32 is the fp32 PARITY pipeline: x, w (KRSC), res, aux and out are fp32 and the
 * tensor cores read the operands as TF32 (kind::tf32).  Feeding it tf32-split operands three times
 * (big*big + big*small + small*big, accumulated through the fp32 residual input) gives fp32-accurate
 * convolutions ("3xTF32"); see mcgm_split_tf32.  R, S are used by MCGM_CONV_S1 only. */

/* ---- weight gradient, tcgen05 + TMA with MN-major operands straight from the
 * NHWC tensors: dw[k,r,s,c] = sum_{n,h,w} dy[n,h,w,k] * x[n,h+r-pad,w+s-pad,c].
 * x [N,H,W,Cin] bf16, dy [N,H,W,Cout] bf16 (Cin,Cout % 64 == 0), dw fp32 KRSC.
 * Deterministic split-K: partials in workspace, fixed-order reduction.
 * dbias (optional) = sum over pixels of dy, accumulated on the tensor cores in the same
 * kernel (dY^T x ones) — dy is read from HBM once. */
size_t mcgm_conv_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S);
int mcgm_conv_wgrad(const void* x, const void* dy, float* dw, float* dbias,
                    int N, int H, int W, int Cin, int Cout, int R, int S,
                    void* workspace, void* stream);

/* Generalised geometry (H, W = dims of X; dw is fp32 [Cout][16 taps (p*4+q)][Cin]):
 *   MCGM_CONV_S2K4  : dy [N,H/2,W/2,Cout]; dw[k,p,q,c] = sum dy[n,i,j,k] * x[n,2i+p-1,2j+q-1,c]
 *   MCGM_CONV_PHASE2: dy [N,2H,2W,Cout];   dw[k,p,q,c] = sum dy[n,2i+p-1,2j+q-1,k] * x[n,i,j,c]
 *                     (phase ph, tap a of the forward weight is p = 3 - 2a - ph; same for q). */
/* Grouped form: n <= 4 independent stride-1 weight gradients of IDENTICAL geometry (the layers of one backward
 * pass that share a shape) in one launch pair.  x_host / dy_host / dw_host / dbias_host are HOST arrays of n device
 * pointers (dbias_host may be NULL).  Alone, a small-spatial layer fills the machine only through ~29-way split-K
 * whose partials dominate its cost; side by side each group needs a fraction of the splits. */
size_t mcgm_conv_wgrad_grouped_workspace_bytes(int n, int N, int H, int W, int Cin, int Cout, int R, int S);
int mcgm_conv_wgrad_grouped(int n, const void* const* x_host, const void* const* dy_host, float* const* dw_host,
                            float* const* dbias_host, int N, int H, int W, int Cin, int Cout, int R, int S,
                            void* workspace, void* stream);
size_t mcgm_conv_wgrad_ex_workspace_bytes(int kind, int in_dtype, int N, int H, int W, int Cin, int Cout,
                                          int R, int S);
int mcgm_conv_wgrad_ex(int kind, int in_dtype, const void* x, const void* dy, float* dw, float* dbias,
                       int N, int H, int W, int Cin, int Cout, int R, int S, void* workspace, void* stream);

/* ---- explicit-tap convolution (MCGM_CONV_TAPS) on the same tcgen05 kernels: w bf16 [Cout][n_taps][Cin],
 *   out[n,i,j] = sum_t x[n, i + dh[t], j + dw[t]] . w[:, t, :]      (out-of-image taps read zeros)
 * serves MCGatedMaskedConv2d (src/models/mcpixelcnn.py:23-61): a (k/2+1) x k vertical stack with padding
 * (k/2, k/2) cropped to the first H rows is taps dh in [-k/2, 0], dw in [-k/2, k/2]; the 1 x (k/2+1) horizontal
 * stack cropped to the first W columns is dh = 0, dw in [-k/2, 0]; the data gradient is the same call with negated
 * offsets and the transposed weight.  taps_*_host are HOST arrays (n_taps <= 64 forward, <= 32 weight gradient).
 * dw of the weight gradient is fp32 [Cout][n_taps][Cin]. */
int mcgm_conv_fwd_taps(const void* x, const void* w, void* out, int out_dtype, int N, int H, int W, int Cin, int Cout,
                       int n_taps, const signed char* taps_dh_host, const signed char* taps_dw_host,
                       const float* bias, const void* res, const void* aux, const int32_t* labels,
                       const uint32_t* bits, int M, int flags, void* stream);
size_t mcgm_conv_wgrad_taps_workspace_bytes(int N, int H, int W, int Cin, int Cout, int n_taps);
int mcgm_conv_wgrad_taps(const void* x, const void* dy, float* dw, float* dbias, int N, int H, int W, int Cin, int Cout,
                         int n_taps, const signed char* taps_dh_host, const signed char* taps_dw_host,
                         void* workspace, void* stream);

/* ---- MC-gated activation of the gated PixelCNN (MCGatedActivation, src/models/mcpixelcnn.py:9-20):
 *   x, y = h.chunk(2, dim=1);  out = codebook[label] * relu(bn(x)) * sigmoid(y)
 * h bf16 NHWC [N,H,W,2C] (x = channels [0,C), y = [C,2C) of each pixel row), out bf16 [N,H,W,C], HW = H*W.
 * mcgm_sgate_stats: train-mode batch statistics of the x half (+ running-stat update: unbiased var, momentum).
 * mcgm_sgate_bwd: dh = [dx | dy] bf16 [N,H,W,2C], dgamma / dbeta fp32 [C]; batch_stats = 1 when (mean, rstd)
 * are this batch's statistics (training), 0 for running statistics.  workspace: mcgm_sgate_workspace_bytes(C). */
size_t mcgm_sgate_workspace_bytes(int C);
int mcgm_sgate_stats(const void* h, long P, int C, float* mean, float* rstd, float* running_mean, float* running_var,
                     float momentum, float eps, void* workspace, void* stream);
int mcgm_sgate_fwd(const void* h, void* out, int N, int HW, int C, const int32_t* labels, const uint32_t* bits, int M,
                   const float* mean, const float* rstd, const float* gamma, const float* beta, void* stream);
int mcgm_sgate_bwd(const void* dout, const void* h, void* dh, int N, int HW, int C, const int32_t* labels,
                   const uint32_t* bits, int M, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, float* dgamma, float* dbeta, int batch_stats, void* workspace, void* stream);

/* ---- thin convolutions on the tensor cores: im2col / col2im of the THIN tensor only.
 *   im2col : U[p, t*Ct + j] = thin[stride*p +/- off(t), j]  bf16 [N,Ho,Wo,KU], zero padded; off(t) =
 *            (r - pad, s - pad); flip: -off (stride 1 only).  stride 2 / 4x4 / pad 1 serves
 *            Conv2d(3,64,4,2,1) (mcvae.py:41) and the data gradient of ConvTranspose2d(64,3,4,2,1).
 *   col2im : thin[p, j] = bias[j] + sum_t T[p + off(t), t*Ct + j]  (T fp32 [N,H,W,CT]) -> tanh
 * With U, Conv2d(3,Cw) forward is mcgm_conv_fwd 1x1 (Cin = KU), its weight gradient is
 * mcgm_conv_wgrad 1x1; Conv2d(Cw,3) forward is a 1x1 mcgm_conv_fwd to T followed by col2im.
 * The wide tensor is read exactly once in every direction. */
int mcgm_thin_im2col(const float* thin, void* U, int out_dtype, int N, int H, int W, int Ct, int R, int S,
                     int stride, int pad, int flip, int KU, void* stream);
int mcgm_thin_col2im(const float* T, int CT, float* thin, int N, int H, int W, int Ct, int R, int S,
                     const float* bias, int flags, void* stream);

/* ---- thin convolutions WITHOUT an im2col buffer: Conv2d(Ct<=4 -> Cout, RxS in {1x1, 3x3}, stride 1, same padding)
 * on an fp32 NHWC image (src/models/mcgan.py:77-79 Conv2d(3,128,3,1,1) -> ReLU -> MC, :87 Conv2d(3,128,1,1,0)).
 * The patch matrix (<= 32 columns, k = (r*S + s)*Ct + j) lives in shared memory only; warp-level bf16 MMAs with
 * fp32 accumulation; the wide tensor is written (forward) / read (weight gradient) exactly once.
 *   w_bf16: [Cout][w_row_stride] bf16, first R*S*Ct columns of a row = the filter in (r,s,c) order
 *           (mcgm_pack_weight_padded with row_stride >= 32), the rest zero.
 *   forward epilogue: + bias, ReLU (relu != 0), x codebook[label[n], c] (labels/bits non-NULL), bf16 NHWC out.
 *   wgrad: dwm fp32 [Cout][32]: columns < R*S*Ct = dW in (r,s,c) order, column R*S*Ct = dbias (the patch matrix
 *          carries a column of ones), remaining columns zero.  Deterministic two-level reduction. */
int mcgm_thin_conv_fwd(const float* img, const void* w_bf16, int w_row_stride, const float* bias,
                       const int32_t* labels, const uint32_t* bits, int M, int relu, void* out_bf16, int N, int H,
                       int W, int Ct, int R, int S, int Cout, void* stream);
size_t mcgm_thin_conv_wgrad_workspace_bytes(int Cout);
int mcgm_thin_conv_wgrad(const float* img, const void* dy_bf16, float* dwm, int N, int H, int W, int Ct, int R, int S,
                         int Cout, void* workspace, void* stream);

/* ---- operand splitting for the fp32 parity pipeline (cfg['compute_dtype'] = 'fp32'):
 *   split_tf32   : x = big + small (+ O(2^-21 |x|)), both exactly TF32-representable — three
 *                  kind::tf32 convolutions big*big + big*small + small*big recover an fp32 product
 *   split_bf16x3 : x = p1 + p2 + p3, bf16 pieces — six bf16 weight-gradient GEMMs do the same. */
int mcgm_split_tf32(const float* x, float* big, float* small, long n, void* stream);
int mcgm_split_bf16x3(const float* x, void* p1, void* p2, void* p3, long n, void* stream);

/* ---- BatchNorm statistics from the producing conv (src/models/mcgan.py:15,20,56: BatchNorm2d of a tensor that
 * a Conv2d has just written).  mcgm_conv_fwd_bnstats = mcgm_conv_fwd_ex on bf16 with a linear epilogue (bias,
 * residual) whose epilogue also accumulates, per CTA warp row, the per-channel sum and sum of squares of the
 * values it stores: stats_partial [slots][2][Cout] with *stat_slots_host slots (<= 4 * SM count; buffer size from
 * mcgm_conv_fwd_bnstats_bytes).  mcgm_bn_finalize turns such partials (or mcgm_bn_stats' own) into mean / rstd
 * and the running statistics, in fp64 and a fixed order — the BN gate then needs no statistics pass over the
 * tensor.  Supported: Cout = 256. */
size_t mcgm_conv_fwd_bnstats_bytes(int Cout);
int mcgm_conv_fwd_bnstats(int kind, const void* x, const void* w, void* out, int N, int H, int W, int Cin, int Cout,
                          int R, int S, const float* bias, const void* res, int flags, float* stats_partial,
                          int* stat_slots_host, void* stream);
int mcgm_bn_finalize(const float* partial, int nblocks, int C, long count, float* mean, float* rstd,
                     float* running_mean, float* running_var, float momentum, float eps, void* stream);

/* ---- Glow flow step, few-channel part (fp32 NHWC [P = N*H*W, C], C a multiple of 4, C <= 64)
 * front = ActNorm(C) -> InvConv2dLU (src/models/mcglow.py:24-55, 76-116):
 *     y = W (scale * (x + loc)),  W = [C_out][C_in] row-major (calc_weight(), mcglow.py:104-110);
 *     xa (optional) = bf16 y[:, :C/2] zero-padded to xa_ch channels — the coupling network's input in the
 *     operand layout of mcgm_conv_fwd.
 *   mcgm_flow_front_bwd: dx = scale * W^T g with g = dy + [dxa, 0]; G_out [C][C+1]:
 *     G[o][i] = sum_p g[p,o] (x[p,i] + loc[i]) (i < C), G[o][C] = sum_p g[p,o]; the caller forms
 *     dW = G[:, :C] * scale, dscale = sum_o W * G[:, :C], dloc = scale * (W^T G[:, C]).
 * tail = AffineCoupling around its network (mcglow.py:153-164) with ZeroConv2d's scale (mcglow.py:119-130):
 *     h = net * exp(3 zscale); affine: s = sigmoid(h[:C/2] + 2), out = [y[:C/2], (y[C/2:] + h[C/2:]) s],
 *     logdet[n] = sum log s; additive (affine = 0): out = [y[:C/2], y[C/2:] + h[:C/2]].
 *     net is fp32 NHWC with net_ch >= C (affine) / C/2 channels per pixel.
 *   mcgm_flow_tail_bwd: dy, dnet (all net_ch channels written) and dz_partial [N][C or C/2] whose sum over
 *     n is dzscale. */
int mcgm_flow_front_fwd(const float* x, const float* loc, const float* scale, const float* W, float* y,
                        void* xa_bf16, long P, int C, int xa_ch, void* stream);
size_t mcgm_flow_front_bwd_workspace_bytes(int C);
int mcgm_flow_front_bwd(const float* dy, const void* dxa_bf16, const float* x, const float* loc, const float* scale,
                        const float* W, float* dx, float* G_out, long P, int C, int xa_ch, void* workspace,
                        void* stream);
int mcgm_flow_tail_fwd(const float* y, const float* net, const float* zscale, float* out, float* logdet, int N,
                       int HW, int C, int net_ch, int affine, void* stream);
int mcgm_flow_tail_bwd(const float* dout, const float* dlogdet, const float* y, const float* net,
                       const float* zscale, float* dy, float* dnet, float* dz_partial, int N, int HW, int C,
                       int net_ch, int affine, void* stream);

/* The sampling direction of the coupling (AffineCoupling.reverse, src/models/mcglow.py:166-175):
 * y = [out_a, out_b / sigmoid(net_s * exp(3 zscale_s) + 2) - net_t * exp(3 zscale_t)]   (additive: out_b - net * exp(3 zscale)).
 * The rest of Flow.reverse (inverse 1x1 conv :114-116, ActNorm.reverse :54-55) is mcgm_flow_front_fwd with the
 * parameters W' = diag(1/scale) W^-1, scale' = 1, loc' = -W (scale * loc) — see models/mcglow.py. */
int mcgm_flow_tail_rev(const float* out, const float* net, const float* zscale, float* y, int N, int HW, int C,
                       int net_ch, int affine, void* stream);

/* ---- optimiser step on flat fp32 buffers (optim.Adam src/train_gan.py:223-236, eps 1e-8, no amsgrad;
 * clip_grad_norm_ + Adam src/train_vae.py:110-111, src/train_glow.py:120-121).  n must be a multiple of 4,
 * buffers 16-byte aligned.  `step` is a DEVICE int32 (number of updates applied so far; advanced by the
 * call), so the update is CUDA-graph capturable.  The gradient is read as
 * grad_scale * (grad_scale_dev ? grad_scale_dev[0] : 1) * grad — the data-parallel 1/world factor and the
 * clip coefficient cost no pass of their own.
 * mcgm_grad_clip_coef: coef_out[0] = min(1, max_norm / (||grad||_2 + 1e-6)), coef_out[1] = ||grad||_2
 * (deterministic two-level reduction; no host synchronisation). */
size_t mcgm_grad_clip_workspace_bytes(void);
int mcgm_grad_clip_coef(const float* grad, long n, float max_norm, float* coef_out, void* workspace, void* stream);
int mcgm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int* step,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   const float* grad_scale_dev, void* stream);
/* Same update with the hyper-parameters read from DEVICE memory: hyper_dev = [lr, beta1, beta2, eps,
 * weight_decay, grad_scale] (6 floats).  A CUDA graph that captured this call follows later changes of
 * optimizer.param_groups[0]['lr'] (ReduceLROnPlateau, src/train_vae.py:35,78-81) because the host stages the
 * six values into that buffer before each replay. */
int mcgm_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int* step,
                       const float* hyper_dev, const float* grad_scale_dev, void* stream);

/* ---- pooling / resampling helpers on bf16 NHWC (AvgPool2d(2) mcgan.py:81,88;
 * Upsample backward = 2x2 sum).  avgpool2: y[n,h,w,c] = scale * sum_{2x2} x;
 * upsample2_add: y[n,h,w,c] = a[n,h,w,c] + b[n,h/2,w/2,c] (b may be NULL → copy-upsample of a at half res is not needed). */
int mcgm_pool2(const void* x, void* y, int dtype, int N, int H, int W, int C, float scale, void* stream);
int mcgm_unpool2(const void* dy, void* dx, int dtype, int N, int H, int W, int C, float scale, void* stream);

/* ---- debug / bring-up / same-box A-B knobs.  The product path never sets them; the defaults are the measured best.
 *   0/1   K-major LBO / SBO bytes, 2/3 MN-major LBO / SBO bytes, 4 descriptor version, 5 layout code (bring-up)
 *   6     forward conv: two M tiles per CTA for BN <= 128 (1)          7   ... for BN = 256 (0)
 *   8     weight gradient: two taps per CTA (1)                         9   force the split-K factor (0 = automatic)
 *   10    staged epilogue: bit 0 on, bit 2 residual / ReLU-sign tile by TMA load (7)
 *   11    forward conv: multicast clusters of 1 / 2 / 4 CTAs (1 = off; bit-identical, no gain)
 *   12/13 RGB-side convs: tcgen05 kernels for the weight gradient / the forward (1; 0 = the mma.sync kernels)
 *   14    weight gradient: one N = 2*BN MMA per tap pair (1)
 *   15    forward conv TIMING DIAGNOSTICS, results are garbage: bit 0 skip MMAs, 1 skip activation loads,
 *         2 skip weight loads, 3 skip the epilogue body (tools/microbench.py legs)
 *   16    forward conv as CTA pairs / cta_group::2 MMAs: bit 0 the 256-wide tiles, bit 1 the 128-wide pair tiles (3)
 *   17    weight gradient: blocked 5-D tensor maps, one TMA operation per operand (0; parity green, no gain)
 *   18    weight gradient: bias-gradient MMAs on the least-loaded tap group / dealt over all groups (1)
 * Every variant is parity-tested (tests/gpu_checks.py: check_conv_variants, check_conv_cluster). */
int mcgm_debug_set(int key, uint64_t value);

#ifdef __cplusplus
}
#endif
#endif /* MCGM_H_ */
