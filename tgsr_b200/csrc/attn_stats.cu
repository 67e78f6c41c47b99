// Reductions over the word axis of the attention maps that the reference computes on the host after copying the maps back
// (SURVEY.md section 8(f) rank 4):
//   * miscc/utils.py:122-125  attn_max = attn.max(dim=1, keepdim=True)   -- the per-pixel strongest word response; the same
//     max-map weighs the pixels in weight_MSE (miscc/losses.py:796);
//   * miscc/utils.py:389-392  conf_score[j] = sum(one_map * (one_map > 2*thresh)) for the max-map (j = 0) and every word map.
// One pass over attn [B, T, Q] (fp32): grid (ceil(Q / (256*4)), B); each thread owns 4 adjacent pixels (float4 loads), keeps
// the running max / first arg-max over the T maps and the thresholded sums; per-CTA partial sums are combined in double
// precision with one atomic per (CTA, map) so the result does not depend on the launch shape beyond fp64 rounding.
#include "common.cuh"

namespace tgsr {
namespace {

constexpr int ST_THREADS = 256;

__global__ void __launch_bounds__(ST_THREADS)
attn_word_stats_kernel(const float* __restrict__ attn, float* __restrict__ maxv, uint8_t* __restrict__ argmax,
                       double* __restrict__ conf, int T, int Q, float thresh2, double replicas) {
  __shared__ float red[ST_THREADS / 32];
  const int b = blockIdx.y;
  const int q = (blockIdx.x * ST_THREADS + threadIdx.x) * 4;
  const bool live = q < Q;
  const float* base = attn + (size_t)b * T * Q + q;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  uint32_t am = 0;                                     // four byte lanes of arg-max
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = 0; t <= T; ++t) {
    // t < T: word map t;  t == T: the max map itself (its thresholded sum goes to conf[b, 0])
    float4 v = m;
    if (t < T) {
      v = live ? __ldg(reinterpret_cast<const float4*>(base + (size_t)t * Q)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (v.x > m.x) { m.x = v.x; am = (am & 0xffffff00u) | (uint32_t)t; }
      if (v.y > m.y) { m.y = v.y; am = (am & 0xffff00ffu) | ((uint32_t)t << 8); }
      if (v.z > m.z) { m.z = v.z; am = (am & 0xff00ffffu) | ((uint32_t)t << 16); }
      if (v.w > m.w) { m.w = v.w; am = (am & 0x00ffffffu) | ((uint32_t)t << 24); }
    }
    float s = 0.f;
    if (live) s = (v.x > thresh2 ? v.x : 0.f) + (v.y > thresh2 ? v.y : 0.f) + (v.z > thresh2 ? v.z : 0.f) + (v.w > thresh2 ? v.w : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < ST_THREADS / 32; ++w) tot += (double)red[w];
      if (tot != 0.0) atomicAdd(conf + (size_t)b * (T + 1) + (t < T ? t + 1 : 0), tot * replicas);
    }
    __syncthreads();
  }
  if (live) {
    *reinterpret_cast<float4*>(maxv + (size_t)b * Q + q) = m;
    *reinterpret_cast<uint32_t*>(argmax + (size_t)b * Q + q) = am;
  }
}

}  // namespace
}  // namespace tgsr

extern "C" int tgsr_attn_word_stats_f32(const float* attn, float* maxv, uint8_t* argmax, double* conf, int B, int T, int Q,
                                        float thresh, int replicas, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(attn); TGSR_REQUIRE_PTR(maxv); TGSR_REQUIRE_PTR(argmax); TGSR_REQUIRE_PTR(conf);
  TGSR_REQUIRE(B > 0 && T > 0 && T <= 255 && Q > 0 && Q % 4 == 0 && replicas > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((reinterpret_cast<uintptr_t>(attn) & 15) == 0 && (reinterpret_cast<uintptr_t>(maxv) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(argmax) & 3) == 0, TGSR_ERR_BAD_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(conf, 0, sizeof(double) * (size_t)B * (T + 1), st));
  attn_word_stats_kernel<<<dim3((unsigned)ceil_div(Q, ST_THREADS * 4), (unsigned)B), ST_THREADS, 0, st>>>(
      attn, maxv, argmax, conf, T, Q, 2.0f * thresh, (double)replicas);
  return launch_status();
}
