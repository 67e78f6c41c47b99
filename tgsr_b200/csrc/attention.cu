// Fused word-region attention, fp32 NCHW I/O  (reference: GlobalAttention.py:87-130).
//
// One kernel does what the reference does in 9 launches: the 1x1 word projection
// S = W.ctx[b] (per CTA, in shared memory; ctx[b] staged with one TMA bulk copy), the logits
// L[q,t] = sum_i x[i,q] S[i,t], the caption mask (with the reference's row-order quirk), the
// softmax over words in registers, and the weighted sum out[i,q] = sum_t S[i,t] P[q,t].
// HBM traffic is the algorithmic minimum: x read once, out and attn written once.
//
// Thread mapping: one thread owns PIX consecutive pixels (vectorised, coalesced NCHW loads and
// stores); logits live in registers; S is read from shared memory as broadcast float4.
#include "common.cuh"
#include <math_constants.h>

namespace tgsr {

constexpr int ATT_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// TMA 1D bulk copy global -> shared, completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int N> struct VecF;
template <> struct VecF<1> { using type = float; };
template <> struct VecF<2> { using type = float2; };
template <> struct VecF<4> { using type = float4; };

template <int PIX>
__device__ __forceinline__ void load_pix(const float* p, float (&v)[PIX]) {
  using V = typename VecF<PIX>::type;
  V t = __ldg(reinterpret_cast<const V*>(p));
  const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
  for (int k = 0; k < PIX; ++k) v[k] = f[k];
}
template <int PIX>
__device__ __forceinline__ void store_pix(float* p, const float (&v)[PIX]) {
  using V = typename VecF<PIX>::type;
  V t;
  float* f = reinterpret_cast<float*>(&t);
#pragma unroll
  for (int k = 0; k < PIX; ++k) f[k] = v[k];
  __stcs(reinterpret_cast<V*>(p), t);   // streaming store: written once, not re-read by this kernel
}

// smem layout (dynamic): [ S: idf*TPAD f32 | ctx: cdf*T f32 | maskbits: B u64 | mbarrier u64 ]
template <int TPAD, int PIX>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_f32_kernel(const float* __restrict__ input, const float* __restrict__ context,
                    const float* __restrict__ w_ctx, const uint8_t* __restrict__ mask,
                    float* __restrict__ wctx, float* __restrict__ attn, int B, int idf, int cdf, int Q, int T,
                    int chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* S = reinterpret_cast<float*>(smem_raw);
  float* ctx_s = S + idf * TPAD;
  unsigned long long* maskbits = reinterpret_cast<unsigned long long*>(ctx_s + cdf * T);
  uint64_t* bar = reinterpret_cast<uint64_t*>(maskbits + (mask ? B : 0));

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int q_begin = blockIdx.x * chunk;
  const int q_end = min(Q, q_begin + chunk);

  // ---- prologue: stage ctx[b] with one TMA bulk copy, build caption bitmasks, project ----
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(cdf * T * sizeof(float));
    mbar_expect_tx(bar, bytes);
    bulk_g2s(ctx_s, context + (size_t)b * cdf * T, bytes, bar);
  }
  if (mask) {
    for (int c = tid; c < B; c += ATT_THREADS) {
      unsigned long long bits = 0ull;
      for (int t = 0; t < T; ++t) bits |= (unsigned long long)(mask[c * T + t] != 0) << t;
      maskbits[c] = bits;
    }
  }
  mbar_wait(bar, 0);
  for (int idx = tid; idx < idf * TPAD; idx += ATT_THREADS) {
    const int i = idx / TPAD, t = idx - i * TPAD;
    float acc = 0.f;
    if (t < T) {
      const float* wr = w_ctx + (size_t)i * cdf;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int c = 0; c < cdf; c += 4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + c));
        a0 = fmaf(w4.x, ctx_s[(c + 0) * T + t], a0);
        a1 = fmaf(w4.y, ctx_s[(c + 1) * T + t], a1);
        a2 = fmaf(w4.z, ctx_s[(c + 2) * T + t], a2);
        a3 = fmaf(w4.w, ctx_s[(c + 3) * T + t], a3);
      }
      acc = (a0 + a1) + (a2 + a3);
    }
    S[idx] = acc;
  }
  __syncthreads();

  const float4* S4 = reinterpret_cast<const float4*>(S);
  const size_t in_base = (size_t)b * idf * Q;
  const size_t at_base = (size_t)b * T * Q;

  for (int q = q_begin + tid * PIX; q < q_end; q += ATT_THREADS * PIX) {
    float l[PIX][TPAD];
#pragma unroll
    for (int p = 0; p < PIX; ++p)
#pragma unroll
      for (int t = 0; t < TPAD; ++t) l[p][t] = 0.f;

    // logits: L[q,t] = sum_i x[i,q] * S[i,t]          (GlobalAttention.py:107)
#pragma unroll 4
    for (int i = 0; i < idf; ++i) {
      float x[PIX];
      load_pix<PIX>(input + in_base + (size_t)i * Q + q, x);
#pragma unroll
      for (int t4 = 0; t4 < TPAD / 4; ++t4) {
        const float4 s = S4[i * (TPAD / 4) + t4];
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
          l[p][4 * t4 + 0] = fmaf(x[p], s.x, l[p][4 * t4 + 0]);
          l[p][4 * t4 + 1] = fmaf(x[p], s.y, l[p][4 * t4 + 1]);
          l[p][4 * t4 + 2] = fmaf(x[p], s.z, l[p][4 * t4 + 2]);
          l[p][4 * t4 + 3] = fmaf(x[p], s.w, l[p][4 * t4 + 3]);
        }
      }
    }

    // caption mask (GlobalAttention.py:109-116): row r = b*Q+q is masked by caption r % B.
#pragma unroll
    for (int p = 0; p < PIX; ++p) {
      unsigned long long bits = 0ull;
      if (mask) bits = maskbits[(b * Q + q + p) % B];
      float m = -CUDART_INF_F;
#pragma unroll
      for (int t = 0; t < TPAD; ++t) {
        if (t >= T || ((bits >> t) & 1ull)) l[p][t] = -CUDART_INF_F;
        m = fmaxf(m, l[p][t]);
      }
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < TPAD; ++t) {
        const float e = (t < T) ? expf(l[p][t] - m) : 0.f;   // exp(-inf)=0 -> masked words exactly 0
        l[p][t] = e;
        sum += e;
      }
      const float inv = 1.0f / sum;                          // all-masked row -> NaN, as torch
#pragma unroll
      for (int t = 0; t < TPAD; ++t) l[p][t] *= inv;
    }

    // attn[b,t,q]  (GlobalAttention.py:118-122,127)
#pragma unroll
    for (int t = 0; t < TPAD; ++t) {
      if (t < T) {
        float v[PIX];
#pragma unroll
        for (int p = 0; p < PIX; ++p) v[p] = l[p][t];
        store_pix<PIX>(attn + at_base + (size_t)t * Q + q, v);
      }
    }
    // weightedContext[b,i,q] = sum_t S[i,t] P[q,t]   (GlobalAttention.py:126)
#pragma unroll 2
    for (int i = 0; i < idf; ++i) {
      float o[PIX];
#pragma unroll
      for (int p = 0; p < PIX; ++p) o[p] = 0.f;
#pragma unroll
      for (int t4 = 0; t4 < TPAD / 4; ++t4) {
        const float4 s = S4[i * (TPAD / 4) + t4];
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
          o[p] = fmaf(s.x, l[p][4 * t4 + 0], o[p]);
          o[p] = fmaf(s.y, l[p][4 * t4 + 1], o[p]);
          o[p] = fmaf(s.z, l[p][4 * t4 + 2], o[p]);
          o[p] = fmaf(s.w, l[p][4 * t4 + 3], o[p]);
        }
      }
      store_pix<PIX>(wctx + in_base + (size_t)i * Q + q, o);
    }
  }
}

static size_t attn_smem_bytes(int idf, int cdf, int T, int TPAD, int B, bool has_mask) {
  return (size_t)idf * TPAD * 4 + (size_t)cdf * T * 4 + (has_mask ? (size_t)B * 8 : 0) + 16;
}

template <int TPAD, int PIX>
static int launch_attn(const float* input, const float* context, const float* w_ctx, const uint8_t* mask,
                       float* wctx, float* attn, int B, int idf, int cdf, int Q, int T, cudaStream_t st) {
  const size_t smem = attn_smem_bytes(idf, cdf, T, TPAD, B, mask != nullptr);
  if (smem > 200 * 1024) return TGSR_ERR_UNSUPPORTED;
  // pixels per CTA: aim for >= 4 CTAs per SM so the per-CTA projection prologue stays amortised.
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  const int unit = ATT_THREADS * PIX;
  long long total_units = (long long)B * ceil_div(Q, unit);
  int loops = (int)((total_units + 4LL * sms - 1) / (4LL * sms));
  if (loops < 1) loops = 1;
  if (loops > 8) loops = 8;
  const int chunk = unit * loops;
  dim3 grid(ceil_div(Q, chunk), B);
  attn_fwd_f32_kernel<TPAD, PIX><<<grid, ATT_THREADS, smem, st>>>(input, context, w_ctx, mask, wctx, attn, B, idf,
                                                                 cdf, Q, T, chunk);
  return launch_status();
}

template <int TPAD, int PIX>
static int set_attr() {
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_f32_kernel<TPAD, PIX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

int init_attention() {
  int s;
  if ((s = set_attr<20, 4>()) != TGSR_OK) return s;
  if ((s = set_attr<20, 1>()) != TGSR_OK) return s;
  if ((s = set_attr<32, 2>()) != TGSR_OK) return s;
  if ((s = set_attr<32, 1>()) != TGSR_OK) return s;
  if ((s = set_attr<48, 2>()) != TGSR_OK) return s;
  if ((s = set_attr<48, 1>()) != TGSR_OK) return s;
  return TGSR_OK;
}

}  // namespace tgsr

extern "C" int tgsr_attn_fwd_f32(const float* input, const float* context, const float* w_ctx, const uint8_t* mask,
                                 float* wctx, float* attn, int B, int idf, int cdf, int Q, int T,
                                 tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(input); TGSR_REQUIRE_PTR(context); TGSR_REQUIRE_PTR(w_ctx);
  TGSR_REQUIRE_PTR(wctx); TGSR_REQUIRE_PTR(attn);
  TGSR_REQUIRE(B > 0 && idf > 0 && cdf > 0 && Q > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(g_sm_count > 0, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE(T <= 48 && idf <= 256 && cdf <= 512 && (cdf % 4) == 0 && B <= 4096, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE((long long)B * Q < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  if (T <= 20) {
    if (Q % 4 == 0) return launch_attn<20, 4>(input, context, w_ctx, mask, wctx, attn, B, idf, cdf, Q, T, st);
    return launch_attn<20, 1>(input, context, w_ctx, mask, wctx, attn, B, idf, cdf, Q, T, st);
  } else if (T <= 32) {
    if (Q % 2 == 0) return launch_attn<32, 2>(input, context, w_ctx, mask, wctx, attn, B, idf, cdf, Q, T, st);
    return launch_attn<32, 1>(input, context, w_ctx, mask, wctx, attn, B, idf, cdf, Q, T, st);
  } else {
    if (Q % 2 == 0) return launch_attn<48, 2>(input, context, w_ctx, mask, wctx, attn, B, idf, cdf, Q, T, st);
    return launch_attn<48, 1>(input, context, w_ctx, mask, wctx, attn, B, idf, cdf, Q, T, st);
  }
}
