// Small fp32 Linear (+bias) (+folded BatchNorm1d) + GLU, and the CA_NET reparametrisation.
// reference: util.py:379-386 (CA_NET.encode), util.py:388-395 (reparametrize),
//            util.py:417-420 + :432 (INIT_STAGE_G.fc).
#include "common.cuh"

namespace tgsr {

// One warp per output pair (b, n): z_a = row n, z_g = row n + N/2; warp-shuffle reduction over K.
__global__ void linear_glu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                  const float* __restrict__ bias, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float* __restrict__ y, int B, int K, int N) {
  const int half = N / 2;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= B * half) return;
  const int b = warp / half, n = warp % half;
  const float* xr = x + (size_t)b * K;
  const float* wa = w + (size_t)n * K;
  const float* wg = w + (size_t)(n + half) * K;
  float a = 0.f, g = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float xv = __ldg(xr + k);
    a = fmaf(xv, __ldg(wa + k), a);
    g = fmaf(xv, __ldg(wg + k), g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    g += __shfl_xor_sync(0xffffffffu, g, o);
  }
  if (lane == 0) {
    if (bias) { a += bias[n]; g += bias[n + half]; }
    if (scale) { a = fmaf(a, scale[n], shift[n]); g = fmaf(g, scale[n + half], shift[n + half]); }
    y[(size_t)b * half + n] = a * (1.0f / (1.0f + expf(-g)));
  }
}

__global__ void ca_reparam_kernel(const float* __restrict__ mulogvar, const float* __restrict__ eps,
                                  float* __restrict__ c_code, int B, int c_dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * c_dim) return;
  const int b = i / c_dim, c = i % c_dim;
  const float mu = mulogvar[(size_t)b * 2 * c_dim + c], lv = mulogvar[(size_t)b * 2 * c_dim + c_dim + c];
  c_code[i] = fmaf(eps[i], expf(0.5f * lv), mu);
}

}  // namespace tgsr

extern "C" int tgsr_linear_glu_f32(const float* x, const float* w, const float* bias, const float* scale,
                                   const float* shift, float* y, int B, int K, int N, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(B > 0 && K > 0 && N > 0 && N % 2 == 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((scale == nullptr) == (shift == nullptr), TGSR_ERR_NULL);
  const long long warps = (long long)B * (N / 2);
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  TGSR_REQUIRE(blocks < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  linear_glu_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(x, w, bias, scale, shift, y, B, K, N);
  return launch_status();
}

extern "C" int tgsr_ca_reparam_f32(const float* mulogvar, const float* eps, float* c_code, int B, int c_dim,
                                   tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(mulogvar); TGSR_REQUIRE_PTR(eps); TGSR_REQUIRE_PTR(c_code);
  TGSR_REQUIRE(B > 0 && c_dim > 0, TGSR_ERR_BAD_SHAPE);
  ca_reparam_kernel<<<ceil_div(B * c_dim, 256), 256, 0, (cudaStream_t)stream>>>(mulogvar, eps, c_code, B, c_dim);
  return launch_status();
}
