// Backward of the fused word-region attention (GlobalAttention.py:87-130), fp32 NCHW.
//
// forward:  S = W.ctx[b] (idf x T);  L[q,t] = sum_i x[i,q] S[i,t];  P = softmax_t(mask(L));  out[i,q] = sum_t S[i,t] P[q,t]
// backward, given dOut [B,idf,Q] (and optionally dAttn [B,T,Q] for the returned maps):
//   dP[q,t]  = sum_i dOut[i,q] S[i,t]  (+ dAttn[t,q])
//   dL[q,t]  = P[q,t] (dP[q,t] - sum_t' P[q,t'] dP[q,t'])             masked words have P = 0 -> dL = 0
//   dx[i,q]  = sum_t dL[q,t] S[i,t]
//   dS[i,t]  = sum_q (dOut[i,q] P[q,t] + x[i,q] dL[q,t])              (reduction over pixels)
//   dW[i,c]  = sum_{b,t} dS[b,i,t] ctx[b,c,t]
// The reference obtains these from autograd over 9 kernels; here: one pixel-streaming kernel (dx and the dS
// partial sums of a 256-pixel chunk) and one tiny kernel for dW.  The text encoder is frozen in the training
// step, so no gradient flows to the word embeddings.
#include "common.cuh"

namespace tgsr {

constexpr int AB_PIX = 256;

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem: S [idf][TP] | Ps [256][TP+1] | dLs [256][TP+1] | dSs [idf][TP]
template <int TP>
__global__ void __launch_bounds__(AB_PIX)
attn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ S, const float* __restrict__ attn,
                const float* __restrict__ dout, const float* __restrict__ dattn, float* __restrict__ dx,
                float* __restrict__ dS, int idf, int Q, int T) {
  extern __shared__ __align__(16) float sm[];
  float* Ss = sm;
  float* Ps = Ss + idf * TP;
  float* dLs = Ps + AB_PIX * (TP + 1);
  float* dSs = dLs + AB_PIX * (TP + 1);                            // [idf][TP] partial dS of this CTA
  const int tid = threadIdx.x, b = blockIdx.y, q = blockIdx.x * AB_PIX + tid;
  const bool live = q < Q;
  for (int i = tid; i < idf * TP; i += AB_PIX) Ss[i] = __ldg(S + (size_t)b * idf * TP + i);
  __syncthreads();
  const float* xb = x + (size_t)b * idf * Q;
  const float* db = dout + (size_t)b * idf * Q;
  float P[TP], dP[TP];
#pragma unroll
  for (int t = 0; t < TP; ++t) {
    P[t] = (live && t < T) ? __ldg(attn + ((size_t)b * T + t) * Q + q) : 0.f;
    dP[t] = (live && t < T && dattn) ? __ldg(dattn + ((size_t)b * T + t) * Q + q) : 0.f;
  }
  for (int i = 0; i < idf; ++i) {
    const float d = live ? __ldg(db + (size_t)i * Q + q) : 0.f;
#pragma unroll
    for (int t = 0; t < TP; ++t) dP[t] = fmaf(d, Ss[i * TP + t], dP[t]);
  }
  float dot = 0.f;
#pragma unroll
  for (int t = 0; t < TP; ++t) dot = fmaf(P[t], dP[t], dot);
#pragma unroll
  for (int t = 0; t < TP; ++t) {
    dP[t] = P[t] * (dP[t] - dot);                                 // now dL
    Ps[tid * (TP + 1) + t] = P[t];
    dLs[tid * (TP + 1) + t] = dP[t];
  }
  if (live) {
    for (int i = 0; i < idf; ++i) {
      float v = 0.f;
#pragma unroll
      for (int t = 0; t < TP; ++t) v = fmaf(dP[t], Ss[i * TP + t], v);
      dx[(size_t)b * idf * Q + (size_t)i * Q + q] = v;
    }
  }
  __syncthreads();
  // dS partial sums of this chunk: D[i][t] = sum_q dOut[i][q] P[q][t] + x[i][q] dL[q][t], a [idf x 256] . [256 x TP] product
  // per term -> warp-level tensor cores (3xTF32 m16n8k8: fp32-grade products, fp32 accumulate).  Warp w owns the 32
  // pixels [32w, 32w+32) as its K range; A fragments come straight from global memory (dOut was just read by this CTA,
  // x is read here for the only time), B fragments from the staged P / dL rows (row stride TP+1: conflict-free).  The
  // eight partial products meet in shared memory, then one atomic per (i, t) and CTA goes to dS.
  const int q0 = blockIdx.x * AB_PIX;
  {
    constexpr int NT = TP / 8;
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
    for (int i = tid; i < idf * TP; i += AB_PIX) dSs[i] = 0.f;
    __syncthreads();
    const int kb = warp * 32;
    for (int m0 = 0; m0 < idf; m0 += 16) {
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
      const int r0 = m0 + g, r1 = m0 + g + 8;
#pragma unroll
      for (int term = 0; term < 2; ++term) {
        const float* A = term == 0 ? db : xb;
        const float* Bs = term == 0 ? Ps : dLs;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int p0 = kb + ks * 8 + tig, p1 = p0 + 4;
          const bool l0 = q0 + p0 < Q, l1 = q0 + p1 < Q;
          // 3xTF32 (hi*hi + lo*hi + hi*lo): these sums cancel heavily, single-pass TF32 leaves 1e-3-level errors
          float av[4];
          av[0] = (l0 && r0 < idf) ? __ldg(A + (size_t)r0 * Q + q0 + p0) : 0.f;
          av[1] = (l0 && r1 < idf) ? __ldg(A + (size_t)r1 * Q + q0 + p0) : 0.f;
          av[2] = (l1 && r0 < idf) ? __ldg(A + (size_t)r0 * Q + q0 + p1) : 0.f;
          av[3] = (l1 && r1 < idf) ? __ldg(A + (size_t)r1 * Q + q0 + p1) : 0.f;
          uint32_t ah[4], al[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) { ah[k] = f2tf32(av[k]); al[k] = f2tf32(av[k] - __uint_as_float(ah[k])); }
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const float bv0 = Bs[p0 * (TP + 1) + nt * 8 + g], bv1 = Bs[p1 * (TP + 1) + nt * 8 + g];
            const uint32_t bh0 = f2tf32(bv0), bh1 = f2tf32(bv1);
            const uint32_t bl0 = f2tf32(bv0 - __uint_as_float(bh0)), bl1 = f2tf32(bv1 - __uint_as_float(bh1));
            mma_tf32_1688(acc[nt], ah, bh0, bh1);
            mma_tf32_1688(acc[nt], al, bh0, bh1);
            mma_tf32_1688(acc[nt], ah, bl0, bl1);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int t0 = nt * 8 + 2 * tig;
        if (r0 < idf) { atomicAdd(dSs + r0 * TP + t0, acc[nt][0]); atomicAdd(dSs + r0 * TP + t0 + 1, acc[nt][1]); }
        if (r1 < idf) { atomicAdd(dSs + r1 * TP + t0, acc[nt][2]); atomicAdd(dSs + r1 * TP + t0 + 1, acc[nt][3]); }
      }
    }
    __syncthreads();
    for (int o = tid; o < idf * TP; o += AB_PIX) {
      const int t = o % TP;
      if (t < T) atomicAdd(dS + (size_t)b * idf * TP + o, dSs[o]);
    }
  }
}

// dW[i][c] = sum_{b,t} dS[b][i][t] ctx[b][c][t]
__global__ void __launch_bounds__(256)
attn_dw_kernel(const float* __restrict__ dS, const float* __restrict__ context, float* __restrict__ dW, int B, int idf,
               int cdf, int T, int TP) {
  const int o = blockIdx.x * 256 + threadIdx.x;
  if (o >= idf * cdf) return;
  const int i = o / cdf, c = o - i * cdf;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* s = dS + ((size_t)b * idf + i) * TP;
    const float* x = context + ((size_t)b * cdf + c) * T;
    for (int t = 0; t < T; ++t) acc = fmaf(s[t], __ldg(x + t), acc);
  }
  dW[o] = acc;
}

template <int TP> static int set_bwd_attr() {
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}
int init_attention_bwd() {
  int s;
  if ((s = set_bwd_attr<24>()) != TGSR_OK) return s;
  if ((s = set_bwd_attr<32>()) != TGSR_OK) return s;
  return set_bwd_attr<48>();
}

// attention_f32_mma.cu
size_t attn_f32_ws_floats(int B, int idf, int T);
int attn_project_f32_launch(const float* context, const float* w_ctx, float* S, int B, int idf, int cdf, int T, cudaStream_t st);

}  // namespace tgsr

using namespace tgsr;

/* ws: 2 * tgsr_attn_f32_ws_floats(B, idf, T) floats (S and dS). */
extern "C" int tgsr_attn_bwd_f32(const float* input, const float* context, const float* w_ctx, const float* attn,
                                 const float* d_wctx, const float* d_attn, float* ws, float* d_input, float* d_w_ctx,
                                 int B, int idf, int cdf, int Q, int T, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(input); TGSR_REQUIRE_PTR(context); TGSR_REQUIRE_PTR(w_ctx); TGSR_REQUIRE_PTR(attn); TGSR_REQUIRE_PTR(d_wctx);
  TGSR_REQUIRE_PTR(ws); TGSR_REQUIRE_PTR(d_input); TGSR_REQUIRE_PTR(d_w_ctx);
  TGSR_REQUIRE(B > 0 && idf > 0 && cdf > 0 && Q > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(T <= 48 && idf % 8 == 0 && idf <= 256 && cdf % 4 == 0 && B <= 65535, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  const int TP = T <= 24 ? 24 : (T <= 32 ? 32 : 48);
  const size_t nws = attn_f32_ws_floats(B, idf, T);
  float* S = ws;
  float* dS = ws + nws;
  int rc = attn_project_f32_launch(context, w_ctx, S, B, idf, cdf, T, st);
  if (rc != TGSR_OK) return rc;
  TGSR_CUDA(cudaMemsetAsync(dS, 0, nws * sizeof(float), st));
  dim3 grid(ceil_div(Q, AB_PIX), B);
#define TGSR_LAUNCH_AB(TPV)                                                                                      \
  attn_bwd_kernel<TPV><<<grid, AB_PIX, ((size_t)2 * idf * TPV + 2 * AB_PIX * (TPV + 1)) * 4, st>>>(input, S, attn, d_wctx, d_attn, \
                                                                                            d_input, dS, idf, Q, T)
  if (TP == 24) TGSR_LAUNCH_AB(24); else if (TP == 32) TGSR_LAUNCH_AB(32); else TGSR_LAUNCH_AB(48);
#undef TGSR_LAUNCH_AB
  rc = launch_status();
  if (rc != TGSR_OK) return rc;
  attn_dw_kernel<<<ceil_div(idf * cdf, 256), 256, 0, st>>>(dS, context, d_w_ctx, B, idf, cdf, T, TP);
  return launch_status();
}
