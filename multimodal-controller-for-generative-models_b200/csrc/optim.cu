// optim.cu — the optimiser side of the training step: Adam, gradient-norm
// clipping and the data-parallel 1/world scaling over ONE flat fp32 buffer per network.
//
// Reference: optim.Adam(lr, betas, weight_decay) src/train_gan.py:223-236 (eps 1e-8, no amsgrad) and
// clip_grad_norm_(parameters, 1) + Adam(3e-4) src/train_vae.py:110-111 / src/train_glow.py:120-121.
// torch runs those as ~10 multi-tensor launches per step plus (clip) a host round trip; here a step is
// one HBM-bound pass: 4 reads (p, g, m, v) + 3 writes per element, the step counter lives on the device
// (CUDA-graph capturable) and the clip coefficient is read from device memory by the update kernel.
#include "common.cuh"

namespace mcgm {

static constexpr int kOptThreads = 256;
static constexpr int kNormBlocks = 296;          // 2 per SM; the partials are summed in fixed order

// sum of squares of a flat fp32 buffer -> one partial per block (deterministic two-level reduction)
__global__ void __launch_bounds__(kOptThreads)
sumsq_partial_kernel(const float* __restrict__ g, long n4, float* __restrict__ partial) {
  float acc = 0.0f;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  __shared__ float sm[kOptThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) t += sm[w];
    partial[blockIdx.x] = t;
  }
}

// total_norm = sqrt(sum partials) (fp64, fixed order); coef = min(1, max_norm / (total_norm + 1e-6)) —
// torch.nn.utils.clip_grad_norm_ semantics.  out[0] = coef, out[1] = total_norm.
__global__ void clip_coef_kernel(const float* __restrict__ partial, int nblocks, float max_norm, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)partial[b];
  const float norm = (float)sqrt(s);
  const float coef = max_norm / (norm + 1e-6f);
  out[0] = coef < 1.0f ? coef : 1.0f;
  out[1] = norm;
}

// One Adam update of a flat parameter buffer (torch.optim.Adam, amsgrad=False):
//   g' = gscale * g (+ wd * p);  m += (g' - m)(1 - b1);  v = b2 v + (1 - b2) g'^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),   t = step[0] + 1
// gscale = grad_scale * (gscale_dev ? gscale_dev[0] : 1): the data-parallel 1/world factor and the clip
// coefficient ride along instead of costing their own passes over the gradient.
__global__ void __launch_bounds__(kOptThreads)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long n4, const int* __restrict__ step, float lr, float beta1, float beta2, float eps,
                 float weight_decay, float grad_scale, const float* __restrict__ gscale_dev,
                 const float* __restrict__ hyper_dev) {
  if (hyper_dev != nullptr) {     // [lr, beta1, beta2, eps, weight_decay, grad_scale] in device memory: a captured
    lr = __ldg(hyper_dev);        // graph replays with whatever the host last staged there (lr schedulers)
    beta1 = __ldg(hyper_dev + 1); beta2 = __ldg(hyper_dev + 2); eps = __ldg(hyper_dev + 3);
    weight_decay = __ldg(hyper_dev + 4); grad_scale = __ldg(hyper_dev + 5);
  }
  const double t = (double)(__ldg(step) + 1);
  const float bc1 = (float)(1.0 - pow((double)beta1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
  const float step_size = lr / bc1;
  const float gs = grad_scale * (gscale_dev ? __ldg(gscale_dev) : 1.0f);
  const float w1 = 1.0f - beta1, w2 = 1.0f - beta2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = __ldg(g4 + i);
    float pe[4] = {pp.x, pp.y, pp.z, pp.w}, me[4] = {mm.x, mm.y, mm.z, mm.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
    const float ge[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = ge[k] * gs;
      if (weight_decay != 0.0f) gk = fmaf(weight_decay, pe[k], gk);
      me[k] = fmaf(gk - me[k], w1, me[k]);
      ve[k] = fmaf(ve[k], beta2, w2 * gk * gk);
      const float denom = sqrtf(ve[k]) / bc2_sqrt + eps;
      pe[k] = fmaf(-step_size, me[k] / denom, pe[k]);
    }
    p4[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    m4[i] = make_float4(me[0], me[1], me[2], me[3]);
    v4[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
  }
}

__global__ void adam_advance_kernel(int* step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) step[0] += 1;
}

static int opt_grid(long n4) {
  long blocks = (n4 + kOptThreads - 1) / kOptThreads;
  const long cap = 148 * 8;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace mcgm

using namespace mcgm;

extern "C" size_t mcgm_grad_clip_workspace_bytes(void) { return (size_t)kNormBlocks * sizeof(float); }

extern "C" int mcgm_grad_clip_coef(const float* grad, long n, float max_norm, float* coef_out, void* workspace,
                                   void* stream) {
  MCGM_CHECK_ARG(grad && coef_out && workspace && n >= 0 && n % 4 == 0, "grad_clip_coef: bad arguments (n must be a multiple of 4)");
  MCGM_CHECK_ARG(((uintptr_t)grad & 15) == 0, "grad_clip_coef: the gradient buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  sumsq_partial_kernel<<<kNormBlocks, kOptThreads, 0, st>>>(grad, n / 4, (float*)workspace);
  MCGM_LAUNCH_CHECK();
  clip_coef_kernel<<<1, 32, 0, st>>>((const float*)workspace, kNormBlocks, max_norm, coef_out);
  MCGM_LAUNCH_CHECK();
  return 0;
}

static int adam_step_impl(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int* step,
                          float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                          const float* grad_scale_dev, const float* hyper_dev, void* stream) {
  MCGM_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step && n >= 0 && n % 4 == 0,
                 "adam_step: bad arguments (n must be a multiple of 4)");
  MCGM_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                 "adam_step: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  adam_step_kernel<<<opt_grid(n / 4), kOptThreads, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n / 4, step, lr, beta1, beta2,
                                                          eps, weight_decay, grad_scale, grad_scale_dev, hyper_dev);
  MCGM_LAUNCH_CHECK();
  adam_advance_kernel<<<1, 32, 0, st>>>(step);
  MCGM_LAUNCH_CHECK();
  return 0;
}

extern "C" int mcgm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int* step,
                              float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                              const float* grad_scale_dev, void* stream) {
  return adam_step_impl(param, grad, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, weight_decay, grad_scale,
                        grad_scale_dev, nullptr, stream);
}

extern "C" int mcgm_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int* step,
                                  const float* hyper_dev, const float* grad_scale_dev, void* stream) {
  MCGM_CHECK_ARG(hyper_dev, "adam_step_dev: hyper_dev is NULL");
  return adam_step_impl(param, grad, exp_avg, exp_avg_sq, n, step, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f, grad_scale_dev, hyper_dev,
                        stream);
}
