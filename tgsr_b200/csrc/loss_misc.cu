// Small losses of miscc/losses.py that stand alone (the matching losses are in damsm.cu):
//   KL_loss(mu, logvar)  (losses.py:806-810): -0.5 * mean(1 + logvar - mu^2 - exp(logvar))
//   cosine_similarity(x1, x2, dim, eps) (losses.py:12-18): sum(x1*x2) / max(|x1|*|x2|, eps) along one axis.
// Forward and backward each; reductions in double inside a CTA / warp, one kernel per call, no atomics.
#include "common.cuh"

namespace tgsr {
namespace {

__global__ void __launch_bounds__(256)
kl_loss_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, float* __restrict__ out, size_t n) {
  __shared__ double red[8];
  double acc = 0.0;
  for (size_t i = threadIdx.x; i < n; i += 256) {
    const float m = mu[i], lv = logvar[i];
    // element order of the reference: ((mu^2 + exp(logvar)) * -1 + 1) + logvar, each step rounded to fp32
    const float e = __fadd_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(m, m), expf(lv)), -1.0f), 1.0f), lv);
    acc += (double)e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] = (float)(t / (double)n) * -0.5f;
  }
}

__global__ void __launch_bounds__(256)
kl_loss_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, const float* __restrict__ g,
                   float* __restrict__ d_mu, float* __restrict__ d_logvar, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float s = -0.5f * g[0] / (float)n;                  // d loss / d element
  d_mu[i] = s * (-2.0f * mu[i]);
  d_logvar[i] = s * (1.0f - expf(logvar[i]));
}

// One warp per output: element (o, d, i) of an [outer, D, inner] view lives at (o*D + d)*inner + i.
__global__ void __launch_bounds__(256)
cosine_kernel(const float* __restrict__ x1, const float* __restrict__ x2, float* __restrict__ out, float* __restrict__ stats,
              size_t n_out, int D, int inner, float eps) {
  const size_t w = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (w >= n_out) return;
  const int lane = threadIdx.x & 31;
  const size_t o = w / inner, i = w - o * inner;
  const size_t base = o * D * (size_t)inner + i;
  double s12 = 0.0, s11 = 0.0, s22 = 0.0;
  for (int d = lane; d < D; d += 32) {
    const float a = x1[base + (size_t)d * inner], b = x2[base + (size_t)d * inner];
    s12 += (double)a * b; s11 += (double)a * a; s22 += (double)b * b;
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) {
    s12 += __shfl_xor_sync(0xffffffffu, s12, k); s11 += __shfl_xor_sync(0xffffffffu, s11, k);
    s22 += __shfl_xor_sync(0xffffffffu, s22, k);
  }
  if (lane == 0) {
    const float w1 = (float)sqrt(s11), w2 = (float)sqrt(s22);
    const float den = fmaxf(w1 * w2, eps);
    out[w] = (float)s12 / den;
    if (stats != nullptr) { stats[3 * w] = (float)s12; stats[3 * w + 1] = w1; stats[3 * w + 2] = w2; }
  }
}

// c = s12 / max(w1*w2, eps).  Unclamped: dc/dx1 = x2/(w1 w2) - c x1 / w1^2 (and symmetrically); clamped: dc/dx1 = x2 / eps.
__global__ void __launch_bounds__(256)
cosine_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ stats,
                  const float* __restrict__ g, float* __restrict__ d1, float* __restrict__ d2, size_t n_out, int D, int inner,
                  float eps) {
  const size_t w = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (w >= n_out) return;
  const int lane = threadIdx.x & 31;
  const size_t o = w / inner, i = w - o * inner;
  const size_t base = o * D * (size_t)inner + i;
  const float s12 = stats[3 * w], w1 = stats[3 * w + 1], w2 = stats[3 * w + 2], gg = g[w];
  const float prod = w1 * w2;
  const bool clamped = !(prod > eps);
  const float inv = 1.0f / (clamped ? eps : prod);
  const float c1 = clamped ? 0.0f : s12 * inv / (w1 * w1), c2 = clamped ? 0.0f : s12 * inv / (w2 * w2);
  for (int d = lane; d < D; d += 32) {
    const size_t e = base + (size_t)d * inner;
    const float a = x1[e], b = x2[e];
    d1[e] = gg * (b * inv - c1 * a);
    d2[e] = gg * (a * inv - c2 * b);
  }
}

}  // namespace
}  // namespace tgsr

using namespace tgsr;

extern "C" int tgsr_kl_loss_f32(const float* mu, const float* logvar, float* out, size_t n, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(mu); TGSR_REQUIRE_PTR(logvar); TGSR_REQUIRE_PTR(out);
  TGSR_REQUIRE(n > 0, TGSR_ERR_BAD_SHAPE);
  kl_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mu, logvar, out, n);
  return launch_status();
}

extern "C" int tgsr_kl_loss_bwd_f32(const float* mu, const float* logvar, const float* g, float* d_mu, float* d_logvar, size_t n,
                                    tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(mu); TGSR_REQUIRE_PTR(logvar); TGSR_REQUIRE_PTR(g); TGSR_REQUIRE_PTR(d_mu); TGSR_REQUIRE_PTR(d_logvar);
  TGSR_REQUIRE(n > 0, TGSR_ERR_BAD_SHAPE);
  kl_loss_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mu, logvar, g, d_mu, d_logvar, n);
  return launch_status();
}

extern "C" int tgsr_cosine_similarity_f32(const float* x1, const float* x2, float* out, float* stats, size_t outer, int D,
                                          int inner, float eps, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x1); TGSR_REQUIRE_PTR(x2); TGSR_REQUIRE_PTR(out);
  TGSR_REQUIRE(outer > 0 && D > 0 && inner > 0, TGSR_ERR_BAD_SHAPE);
  const size_t n_out = outer * (size_t)inner;
  cosine_kernel<<<(unsigned)((n_out * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x1, x2, out, stats, n_out, D, inner, eps);
  return launch_status();
}

extern "C" int tgsr_cosine_similarity_bwd_f32(const float* x1, const float* x2, const float* stats, const float* g, float* d1,
                                              float* d2, size_t outer, int D, int inner, float eps, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x1); TGSR_REQUIRE_PTR(x2); TGSR_REQUIRE_PTR(stats); TGSR_REQUIRE_PTR(g); TGSR_REQUIRE_PTR(d1); TGSR_REQUIRE_PTR(d2);
  TGSR_REQUIRE(outer > 0 && D > 0 && inner > 0, TGSR_ERR_BAD_SHAPE);
  const size_t n_out = outer * (size_t)inner;
  cosine_bwd_kernel<<<(unsigned)((n_out * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x1, x2, stats, g, d1, d2, n_out, D,
                                                                                            inner, eps);
  return launch_status();
}
