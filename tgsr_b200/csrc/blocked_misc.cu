// Small kernels of the bf16 blocked-layout pipeline that are not GEMM-shaped:
//   * conv_in: conv3x3(3 -> 2C) + folded BN + GLU from the NCHW fp32 LR image into C blocked bf16
//     channels (INIT_STAGE_GImgup.im2f util.py:741-744 / NetG_high*.convin model.py:228); K = 27,
//     3.5 MFLOP per image -- CUDA cores, fp32 math.
//   * attention on blocked activations (GlobalAttention.py:87-130): same fused algorithm as
//     attention.cu, reading h_code from planes of the stage's concat buffer and writing c_code
//     into the next planes of the SAME buffer (torch.cat of util.py:773,817 disappears).
#include "common.cuh"
#include "fmt16.cuh"
#include <math_constants.h>

namespace tgsr {

typedef uint16_t h16;

// ---------------------------------------------------------------- conv_in
// thread = (pixel, output plane of 8 channels): 8 GLU outputs = 16 pre-activation channels.
constexpr int CI_MAXC = 64;   // max output channels C (weights [2C][3][3][3] in smem)
template <int FMT>
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ x, const float* __restrict__ xsub, const float* __restrict__ w,
               const float* __restrict__ scale, const float* __restrict__ shift, h16* __restrict__ y,
               int B, int H, int W, int C, int yP, int yp0) {
  __shared__ float ws[2 * CI_MAXC * 27];
  __shared__ float sc[2 * CI_MAXC], sh[2 * CI_MAXC];
  for (int i = threadIdx.x; i < 2 * C * 27; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) { sc[i] = scale ? scale[i] : 1.f; sh[i] = shift ? shift[i] : 0.f; }
  __syncthreads();
  const int HW = H * W, planes = C / 8;
  const size_t total = (size_t)B * planes * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW), pl = (int)((i / HW) % planes), b = (int)(i / ((size_t)HW * planes));
    const int py = pix / W, px = pix % W;
    float in[27];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py + ky - 1, xx = px + kx - 1;
          float v = 0.f;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const size_t o = (((size_t)b * 3 + c) * H + yy) * W + xx;
            v = __ldg(x + o);
            if (xsub) v -= __ldg(xsub + o);
          }
          in[c * 9 + ky * 3 + kx] = v;
        }
    float o8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ca = pl * 8 + e, cg = C + ca;
      float a = 0.f, g = 0.f;
#pragma unroll
      for (int k = 0; k < 27; ++k) { a = fmaf(in[k], ws[ca * 27 + k], a); g = fmaf(in[k], ws[cg * 27 + k], g); }
      a = fmaf(a, sc[ca], sh[ca]);
      g = fmaf(g, sc[cg], sh[cg]);
      o8[e] = a * (1.0f / (1.0f + __expf(-g)));
    }
    *reinterpret_cast<uint4*>(y + (((size_t)b * yP + yp0 + pl) * HW + pix) * 8) = pack8_16<FMT>(o8);
  }
}

// ---------------------------------------------------------------- attention on blocked planes
constexpr int AB_THREADS = 256;
constexpr int AB_IDF = 32;
constexpr int AB_TPAD = 20;

// smem: S [32][20] f32 | ctx [cdf][T] f32 | maskbits [B] u64 | mbarrier
template <int PIX, int FMT>
__global__ void __launch_bounds__(AB_THREADS)
attn_blocked_kernel(const h16* __restrict__ xin, int xP, int xp0, const float* __restrict__ context,
                    const float* __restrict__ w_ctx, const uint8_t* __restrict__ mask,
                    h16* __restrict__ yout, int yP, int yp0, float* __restrict__ attn, int B, int cdf, int Q,
                    int T, int chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* S = reinterpret_cast<float*>(smem_raw);
  float* ctx_s = S + AB_IDF * AB_TPAD;
  unsigned long long* maskbits = reinterpret_cast<unsigned long long*>(ctx_s + cdf * T);
  const int tid = threadIdx.x, b = blockIdx.y;
  const int q_begin = blockIdx.x * chunk, q_end = min(Q, q_begin + chunk);

  for (int i = tid; i < cdf * T; i += AB_THREADS) ctx_s[i] = __ldg(context + (size_t)b * cdf * T + i);
  if (mask) {
    for (int c = tid; c < B; c += AB_THREADS) {
      unsigned long long bits = 0ull;
      for (int t = 0; t < T; ++t) bits |= (unsigned long long)(mask[c * T + t] != 0) << t;
      maskbits[c] = bits;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < AB_IDF * AB_TPAD; idx += AB_THREADS) {
    const int i = idx / AB_TPAD, t = idx - i * AB_TPAD;
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
  const size_t plane = (size_t)Q * 8;
  const h16* xb = xin + ((size_t)b * xP + xp0) * plane;
  h16* yb = yout + ((size_t)b * yP + yp0) * plane;
  float* ab = attn + (size_t)b * T * Q;

  for (int q = q_begin + tid * PIX; q < q_end; q += AB_THREADS * PIX) {
    float l[PIX][AB_TPAD];
#pragma unroll
    for (int p = 0; p < PIX; ++p)
#pragma unroll
      for (int t = 0; t < AB_TPAD; ++t) l[p][t] = 0.f;
    uint4 raw[AB_IDF / 8][PIX];
#pragma unroll
    for (int pl = 0; pl < AB_IDF / 8; ++pl)
#pragma unroll
      for (int p = 0; p < PIX; ++p)
        raw[pl][p] = __ldg(reinterpret_cast<const uint4*>(xb + pl * plane + (size_t)(q + p) * 8));
#pragma unroll
    for (int pl = 0; pl < AB_IDF / 8; ++pl) {
      float x[PIX][8];
#pragma unroll
      for (int p = 0; p < PIX; ++p) unpack8_16<FMT>(raw[pl][p], x[p]);
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int t4 = 0; t4 < AB_TPAD / 4; ++t4) {
          const float4 s = S4[(pl * 8 + e) * (AB_TPAD / 4) + t4];
#pragma unroll
          for (int p = 0; p < PIX; ++p) {
            l[p][4 * t4 + 0] = fmaf(x[p][e], s.x, l[p][4 * t4 + 0]);
            l[p][4 * t4 + 1] = fmaf(x[p][e], s.y, l[p][4 * t4 + 1]);
            l[p][4 * t4 + 2] = fmaf(x[p][e], s.z, l[p][4 * t4 + 2]);
            l[p][4 * t4 + 3] = fmaf(x[p][e], s.w, l[p][4 * t4 + 3]);
          }
        }
    }
#pragma unroll
    for (int p = 0; p < PIX; ++p) {
      unsigned long long bits = 0ull;
      if (mask) bits = maskbits[(b * Q + q + p) % B];        // GlobalAttention.py:109-112 row-order quirk
      float m = -CUDART_INF_F;
#pragma unroll
      for (int t = 0; t < AB_TPAD; ++t) {
        if (t >= T || ((bits >> t) & 1ull)) l[p][t] = -CUDART_INF_F;
        m = fmaxf(m, l[p][t]);
      }
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < AB_TPAD; ++t) { const float e = (t < T) ? __expf(l[p][t] - m) : 0.f; l[p][t] = e; sum += e; }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int t = 0; t < AB_TPAD; ++t) l[p][t] *= inv;
    }
#pragma unroll
    for (int t = 0; t < AB_TPAD; ++t) {
      if (t < T) {
        if (PIX == 2) __stcs(reinterpret_cast<float2*>(ab + (size_t)t * Q + q), make_float2(l[0][t], l[PIX - 1][t]));
        else __stcs(ab + (size_t)t * Q + q, l[0][t]);
      }
    }
#pragma unroll
    for (int pl = 0; pl < AB_IDF / 8; ++pl) {
#pragma unroll
      for (int p = 0; p < PIX; ++p) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = pl * 8 + e;
          float acc = 0.f;
#pragma unroll
          for (int t4 = 0; t4 < AB_TPAD / 4; ++t4) {
            const float4 s = S4[i * (AB_TPAD / 4) + t4];
            acc = fmaf(s.x, l[p][4 * t4 + 0], acc); acc = fmaf(s.y, l[p][4 * t4 + 1], acc);
            acc = fmaf(s.z, l[p][4 * t4 + 2], acc); acc = fmaf(s.w, l[p][4 * t4 + 3], acc);
          }
          o[e] = acc;
        }
        *reinterpret_cast<uint4*>(yb + pl * plane + (size_t)(q + p) * 8) = pack8_16<FMT>(o);
      }
    }
  }
}

int init_blocked_misc() {
  cudaError_t e = cudaFuncSetAttribute(attn_blocked_kernel<1, TGSR_FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if (e != cudaSuccess) return cuda_fail(e);
  e = cudaFuncSetAttribute(attn_blocked_kernel<1, TGSR_FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

}  // namespace tgsr

extern "C" int tgsr_conv_in_blocked(const float* x, const float* x_sub, const float* w, const float* scale,
                                    const float* shift, void* y, int yP, int yp0, int B, int H, int W, int C,
                                    int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && yp0 >= 0 && yp0 + C / 8 <= yP, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((scale == nullptr) == (shift == nullptr), TGSR_ERR_NULL);
  TGSR_REQUIRE(C <= CI_MAXC, TGSR_ERR_UNSUPPORTED);
  const size_t total = (size_t)B * (C / 8) * H * W;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  if (fmt == TGSR_FMT_F16)
    conv_in_kernel<TGSR_FMT_F16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_sub, w, scale, shift, reinterpret_cast<h16*>(y), B, H, W, C, yP, yp0);
  else
    conv_in_kernel<TGSR_FMT_BF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_sub, w, scale, shift, reinterpret_cast<h16*>(y), B, H, W, C, yP, yp0);
  return launch_status();
}

extern "C" int tgsr_attn_fwd_blocked(const void* x, int xP, int xp0, const float* context, const float* w_ctx,
                                     const uint8_t* mask, void* y, int yP, int yp0, float* attn, int B, int idf,
                                     int cdf, int Q, int T, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(context); TGSR_REQUIRE_PTR(w_ctx); TGSR_REQUIRE_PTR(y); TGSR_REQUIRE_PTR(attn);
  TGSR_REQUIRE(B > 0 && idf > 0 && cdf > 0 && Q > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(g_sm_count > 0, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE(idf == AB_IDF && T <= AB_TPAD && cdf <= 512 && cdf % 4 == 0 && B <= 4096, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(xp0 >= 0 && xp0 + idf / 8 <= xP && yp0 >= 0 && yp0 + idf / 8 <= yP, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * Q < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  const size_t smem = (size_t)AB_IDF * AB_TPAD * 4 + (size_t)cdf * T * 4 + (mask ? (size_t)B * 8 : 0) + 16;
  const int PIX = 1;   // PIX=2 needs 252 registers (1 CTA/SM); one pixel per thread keeps 3 CTAs/SM resident
  const int unit = AB_THREADS * PIX;
  long long total_units = (long long)B * ceil_div(Q, unit);
  int loops = (int)((total_units + 4LL * g_sm_count - 1) / (4LL * g_sm_count));
  loops = loops < 1 ? 1 : (loops > 8 ? 8 : loops);
  const int chunk = unit * loops;
  dim3 grid(ceil_div(Q, chunk), B);
  cudaStream_t st = (cudaStream_t)stream;
  auto xi = reinterpret_cast<const h16*>(x);
  auto yo = reinterpret_cast<h16*>(y);
  (void)PIX;
  if (fmt == TGSR_FMT_F16)
    attn_blocked_kernel<1, TGSR_FMT_F16><<<grid, AB_THREADS, smem, st>>>(xi, xP, xp0, context, w_ctx, mask, yo, yP, yp0, attn, B, cdf, Q, T, chunk);
  else
    attn_blocked_kernel<1, TGSR_FMT_BF16><<<grid, AB_THREADS, smem, st>>>(xi, xP, xp0, context, w_ctx, mask, yo, yP, yp0, attn, B, cdf, Q, T, chunk);
  return launch_status();
}
