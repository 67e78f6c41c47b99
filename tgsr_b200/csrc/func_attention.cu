// func_attention forward (DAMSM double-softmax attention), fp32.
// reference: GlobalAttention.py:33-74.
//   A[s,t]  = sum_d ctx[d,s] q[d,t]            (:51)
//   A       = softmax_t(A)                      (:54)   per region, over words
//   A2[t,s] = softmax_s(gamma1 * A[s,t])        (:59-63) per word, over regions
//   wctx[d,t] = sum_s ctx[d,s] A2[t,s]          (:70)
// One CTA per batch element; the whole [S x Tq] map lives in shared memory, so the reference's
// three transposed copies and two standalone softmax launches disappear.
#include "common.cuh"
#include <math_constants.h>

namespace tgsr {

constexpr int FA_THREADS = 256;

template <int TPAD>
__global__ void __launch_bounds__(FA_THREADS)
func_attn_fwd_kernel(const float* __restrict__ query, const float* __restrict__ context, float gamma1,
                     float* __restrict__ wctx, float* __restrict__ attn, int ndf, int Tq, int S) {
  extern __shared__ __align__(16) float fsm[];
  float* q_s = fsm;                    // [ndf][TPAD]
  float* a_s = q_s + ndf * TPAD;       // [TPAD][S+1]  (word-major so the region softmax is conflict-free)
  const int SP = S + 1;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* q = query + (size_t)b * ndf * Tq;
  const float* ctx = context + (size_t)b * ndf * S;

  for (int i = tid; i < ndf * TPAD; i += FA_THREADS) {
    const int d = i / TPAD, t = i % TPAD;
    q_s[i] = t < Tq ? __ldg(q + (size_t)d * Tq + t) : 0.f;
  }
  __syncthreads();
  const float4* q4 = reinterpret_cast<const float4*>(q_s);

  // pass 1: one region per thread, logits over words in registers, softmax over words
  for (int s = tid; s < S; s += FA_THREADS) {
    float l[TPAD];
#pragma unroll
    for (int t = 0; t < TPAD; ++t) l[t] = 0.f;
#pragma unroll 4
    for (int d = 0; d < ndf; ++d) {
      const float x = __ldg(ctx + (size_t)d * S + s);
#pragma unroll
      for (int t4 = 0; t4 < TPAD / 4; ++t4) {
        const float4 v = q4[d * (TPAD / 4) + t4];
        l[4 * t4 + 0] = fmaf(x, v.x, l[4 * t4 + 0]);
        l[4 * t4 + 1] = fmaf(x, v.y, l[4 * t4 + 1]);
        l[4 * t4 + 2] = fmaf(x, v.z, l[4 * t4 + 2]);
        l[4 * t4 + 3] = fmaf(x, v.w, l[4 * t4 + 3]);
      }
    }
    float m = -CUDART_INF_F;
#pragma unroll
    for (int t = 0; t < TPAD; ++t) if (t < Tq) m = fmaxf(m, l[t]);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < TPAD; ++t) { l[t] = t < Tq ? expf(l[t] - m) : 0.f; sum += l[t]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < TPAD; ++t) if (t < Tq) a_s[t * SP + s] = l[t] * inv * gamma1;
  }
  __syncthreads();

  // pass 2: one warp per word, softmax over regions
  const int warp = tid / 32, lane = tid % 32, nwarp = FA_THREADS / 32;
  for (int t = warp; t < Tq; t += nwarp) {
    float* row = a_s + t * SP;
    float m = -CUDART_INF_F;
    for (int s = lane; s < S; s += 32) m = fmaxf(m, row[s]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int s = lane; s < S; s += 32) { const float e = expf(row[s] - m); row[s] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float* out = attn + ((size_t)b * Tq + t) * S;
    for (int s = lane; s < S; s += 32) { const float v = row[s] * inv; row[s] = v; out[s] = v; }
  }
  __syncthreads();

  // pass 3: one warp per feature row d: wctx[d,t] = sum_s ctx[d,s] A2[t,s]
  for (int d = warp; d < ndf; d += nwarp) {
    float acc[TPAD];
#pragma unroll
    for (int t = 0; t < TPAD; ++t) acc[t] = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float x = __ldg(ctx + (size_t)d * S + s);
#pragma unroll
      for (int t = 0; t < TPAD; ++t) if (t < Tq) acc[t] = fmaf(x, a_s[t * SP + s], acc[t]);
    }
#pragma unroll
    for (int t = 0; t < TPAD; ++t) {
      if (t < Tq) {
        float v = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wctx[((size_t)b * ndf + d) * Tq + t] = v;
      }
    }
  }
}

template <int TPAD>
static int launch_fa(const float* query, const float* context, float gamma1, float* wctx, float* attn, int B,
                     int ndf, int Tq, int S, cudaStream_t st) {
  const size_t smem = ((size_t)ndf * TPAD + (size_t)TPAD * (S + 1)) * sizeof(float);
  if (smem > 200 * 1024) return TGSR_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(func_attn_fwd_kernel<TPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return cuda_fail(e);
  func_attn_fwd_kernel<TPAD><<<B, FA_THREADS, smem, st>>>(query, context, gamma1, wctx, attn, ndf, Tq, S);
  return launch_status();
}

}  // namespace tgsr

extern "C" int tgsr_func_attn_fwd_f32(const float* query, const float* context, float gamma1, float* wctx,
                                      float* attn, int B, int ndf, int Tq, int S, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(query); TGSR_REQUIRE_PTR(context); TGSR_REQUIRE_PTR(wctx); TGSR_REQUIRE_PTR(attn);
  TGSR_REQUIRE(B > 0 && ndf > 0 && Tq > 0 && S > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(Tq <= 48 && S <= 1024 && ndf <= 1024, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  if (Tq <= 20) return launch_fa<20>(query, context, gamma1, wctx, attn, B, ndf, Tq, S, st);
  if (Tq <= 32) return launch_fa<32>(query, context, gamma1, wctx, attn, B, ndf, Tq, S, st);
  return launch_fa<48>(query, context, gamma1, wctx, attn, B, ndf, Tq, S, st);
}
