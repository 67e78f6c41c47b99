// Fused Adam over a flat fp32 parameter bucket (torch.optim.Adam semantics without weight decay / amsgrad; the
// reference's optimiser setting is Adam(lr 2e-4, betas (0.5, 0.999)): miscc/config.py:38, pretrain_DAMSM.py:270).
// One pass over (param, grad, m, v); grad_scale folds the 1/world_size of the gradient all-reduce average.
#include "common.cuh"

namespace tgsr {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale) {
  const float step = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float gi = g[i] * gscale;
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// same update with the step count read from device memory (bias corrections computed in-kernel), so that the whole
// training step can be captured in a CUDA graph and replayed
__global__ void __launch_bounds__(256)
adam_devstep_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                    float lr, float b1, float b2, float eps, const int* __restrict__ step, float gscale) {
  const double t = (double)step[0];        // bias corrections in double, as the host path (and torch.optim.Adam) compute them
  const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float st = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float gi = g[i] * gscale;
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    p[i] -= st * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

}  // namespace tgsr

/* step: 1-based step count (bias corrections 1 - beta^step computed on the host in double). */
extern "C" int tgsr_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                             float beta1, float beta2, float eps, float bias_corr1, float bias_corr2_sqrt, float grad_scale,
                             tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(param); TGSR_REQUIRE_PTR(grad); TGSR_REQUIRE_PTR(exp_avg); TGSR_REQUIRE_PTR(exp_avg_sq);
  TGSR_REQUIRE(n > 0 && bias_corr1 > 0.f && bias_corr2_sqrt > 0.f, TGSR_ERR_BAD_SHAPE);
  const size_t sms = g_sm_count > 0 ? g_sm_count : 148;
  size_t blocks = (n + 255) / 256;
  if (blocks > sms * 8) blocks = sms * 8;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                 bias_corr1, bias_corr2_sqrt, grad_scale);
  return launch_status();
}

/* Graph-capturable variant: `step` is a device int32 holding the 1-based step count (the caller increments it on the stream). */
extern "C" int tgsr_adam_devstep_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                                     float beta1, float beta2, float eps, const int* step, float grad_scale,
                                     tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(param); TGSR_REQUIRE_PTR(grad); TGSR_REQUIRE_PTR(exp_avg); TGSR_REQUIRE_PTR(exp_avg_sq); TGSR_REQUIRE_PTR(step);
  TGSR_REQUIRE(n > 0, TGSR_ERR_BAD_SHAPE);
  const size_t sms = g_sm_count > 0 ? g_sm_count : 148;
  size_t blocks = (n + 255) / 256;
  if (blocks > sms * 8) blocks = sms * 8;
  adam_devstep_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                         eps, step, grad_scale);
  return launch_status();
}
