// Training-mode building blocks of the generator on BLOCKED 16-bit activations [B][P][H][W][8] (the layout of the tcgen05
// conv kernels, conv_tc.cu): batch-statistics BatchNorm forward / backward with the fused GLU and residual add, the
// nearest x2 upsample and its 2x2 sum-pool adjoint.  With these the bf16 training step chains conv (tcgen05) -> BN
// -> GLU -> conv ... without ever leaving the blocked layout; statistics, normalisation and gradients are computed in
// fp32 (sums in fp64), activations are rounded to 16 bits once per layer.
//
// reference: nn.BatchNorm2d in train mode (util.py:64, :78, :115-122, model.py:228-232), GLU util.py:45-53, ResBlock skip
// util.py:126-130, nn.Upsample(2) util.py:76.  The fp32 NCHW twins of these kernels live in train_ops.cu.
#include "common.cuh"
#include "fmt16.cuh"

namespace tgsr {

typedef uint16_t h16;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- BatchNorm statistics (train forward)
// grid (P, SPLIT): CTA (p, s) sums the 8 channels of plane p over pixel chunks s, s + SPLIT, ...; every thread keeps
// 8 + 8 fp32 partials over at most 64 pixels before folding them into fp64, a CTA adds 16 doubles to ws with atomics.
// ws [2 * 8P] = (sum, sum of squares) per channel; bn16_stats_final turns them into mean / invstd (+ running stats).
// Fold per-thread fp32 partials (at most 64 elements each) into the CTA's fp64 accumulators in shared memory: warp
// shuffle reduction in fp32, lane 0 adds in fp64.  Keeping the fp64 state out of the register file is what lets two CTAs
// share an SM with several 16-byte loads in flight per thread (these kernels are latency-bound otherwise).
template <int NA>
__device__ __forceinline__ void fold_partials(float (&s)[NA], double (*sacc)[NA], int warp, int lane) {
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    float v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sacc[warp][k] += (double)v;
    s[k] = 0.f;
  }
}

template <int FMT>
__global__ void __launch_bounds__(256, 2)
bn16_stats_partial_kernel(const h16* __restrict__ z, double* __restrict__ ws, int B, int P, int HW, FastDiv dHW) {
  __shared__ double sacc[8][16];
  const int p = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 128) sacc[threadIdx.x >> 4][threadIdx.x & 15] = 0.0;
  __syncthreads();
  const uint32_t n_items = (uint32_t)B * HW;                               // < 2^31 (checked by the launcher)
  const uint32_t nchunk = (n_items + 64 * 256 - 1) / (64 * 256);           // chunk = 64 pixels per thread
  float sq[16];                                                            // sums | sums of squares
#pragma unroll
  for (int e = 0; e < 16; ++e) sq[e] = 0.f;
  for (uint32_t ch = blockIdx.y; ch < nchunk; ch += gridDim.y) {
    for (int it = 0; it < 64; it += 8) {
      uint4 u[8];
      bool ok[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {                                        // eight independent 16-byte loads in flight
        const uint32_t i = (ch * 64 + it + k) * 256 + threadIdx.x;         // (b, pixel) index, coalesced across the CTA
        ok[k] = i < n_items;
        uint32_t b, pix;
        dHW.divmod(ok[k] ? i : 0u, b, pix);
        u[k] = __ldg(reinterpret_cast<const uint4*>(z + (((size_t)b * P + p) * (size_t)HW + pix) * 8));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (ok[k]) {
          float v[8];
          unpack8_16<FMT>(u[k], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) { sq[e] += v[e]; sq[8 + e] = fmaf(v[e], v[e], sq[8 + e]); }
        }
      }
    }
    fold_partials<16>(sq, sacc, warp, lane);
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sacc[w][threadIdx.x];
    const int e = threadIdx.x & 7, which = threadIdx.x >> 3;
    atomicAdd(ws + 2 * (p * 8 + e) + which, t);
  }
}

__global__ void bn16_stats_final_kernel(const double* __restrict__ ws, float* __restrict__ mean, float* __restrict__ invstd,
                                        float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                        float eps, int C, double n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = ws[2 * c] / n;
  double var = ws[2 * c + 1] / n - m * m;                       // biased variance (normalisation)
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// ---------------------------------------------------------------- normalise (+GLU) (+residual)
// thread = (image, OUTPUT plane, pixel): 8 output channels.  GLU: output plane p pairs value plane p with gate plane
// p + Pf (Pf = output planes).  coef[c] = (gamma*invstd, beta - mean*gamma*invstd) staged in shared memory.
template <int FMT, bool GLU>
__global__ void __launch_bounds__(256)
bn16_act_kernel(const h16* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, const h16* __restrict__ residual,
                h16* __restrict__ y, int B, int P, int HW, FastDiv dHW, FastDiv dPf) {
  extern __shared__ float coef[];                               // [2][C]
  const int C = P * 8;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float a = gamma[c] * invstd[c];
    coef[c] = a; coef[C + c] = beta[c] - mean[c] * a;
  }
  __syncthreads();
  const int Pf = GLU ? P / 2 : P;
  const uint32_t total = (uint32_t)B * Pf * HW;                 // < 2^31 (checked by the launcher)
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    uint32_t pix, pf, b, t;
    dHW.divmod(i, t, pix); dPf.divmod(t, b, pf);
    float a[8], o[8];
    unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(z + (((size_t)b * P + pf) * (size_t)HW + pix) * 8)), a);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = fmaf(a[e], coef[pf * 8 + e], coef[C + pf * 8 + e]);
    if (GLU) {
      float g[8];
      unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(z + (((size_t)b * P + pf + Pf) * (size_t)HW + pix) * 8)), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gg = fmaf(g[e], coef[(pf + Pf) * 8 + e], coef[C + (pf + Pf) * 8 + e]);
        o[e] *= __fdividef(1.0f, 1.0f + __expf(-gg));
      }
    }
    if (residual) {
      float r[8];
      unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(residual + (size_t)i * 8)), r);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += r[e];
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8_16<FMT>(o);
  }
}

// ---------------------------------------------------------------- backward
// du = dL/d(pre-activation u = gamma*xhat + beta).  none: du = dy.  GLU (a = u_c, g = u_{c+Cf}, y = a*s(g)):
// du_a = dy*s, du_g = dy*a*s*(1-s).  Stage 1 accumulates S1 = sum du, S2 = sum du*xhat per channel (fp64 atomics);
// stage 2 writes dz = gamma*invstd*(du - S1/N - xhat*S2/N).  dgamma = S2, dbeta = S1.
template <int FMT, bool GLU>
__device__ __forceinline__ void bn16_du(const h16* __restrict__ z, const h16* __restrict__ dy, const float* coef, int C, int P,
                                        int Pf, size_t b, size_t pf, size_t pix, int HW, float (&xa)[8], float (&xg)[8],
                                        float (&dua)[8], float (&dug)[8]) {
  float d[8];
  unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(dy + ((b * Pf + pf) * (size_t)HW + pix) * 8)), d);
  unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(z + ((b * P + pf) * (size_t)HW + pix) * 8)), xa);
  if (GLU) {
    unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(z + ((b * P + pf + Pf) * (size_t)HW + pix) * 8)), xg);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ca = (int)pf * 8 + e, cg = ca + Pf * 8;
      const float a = fmaf(xa[e], coef[ca], coef[C + ca]);
      const float g = fmaf(xg[e], coef[cg], coef[C + cg]);
      const float s = __fdividef(1.0f, 1.0f + __expf(-g));
      dua[e] = d[e] * s;
      dug[e] = d[e] * a * s * (1.0f - s);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) { dua[e] = d[e]; dug[e] = 0.f; xg[e] = 0.f; }
  }
}

template <int FMT, bool GLU>
__global__ void __launch_bounds__(256, 2)
bn16_bwd_reduce_kernel(const h16* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta, const h16* __restrict__ dy,
                       double* __restrict__ ws, int B, int P, int HW, FastDiv dHW) {
  extern __shared__ float coef[];
  constexpr int NA = GLU ? 32 : 16;                             // (S1 | raw S2) x (value | gate) x 8 channels
  __shared__ double sacc[8][NA];
  const int C = P * 8;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float a = gamma[c] * invstd[c];
    coef[c] = a; coef[C + c] = beta[c] - mean[c] * a;
  }
  for (int k = threadIdx.x; k < 8 * NA; k += 256) sacc[k / NA][k % NA] = 0.0;
  __syncthreads();
  const int Pf = GLU ? P / 2 : P;
  const int pf = blockIdx.x;                                     // grid (Pf, SPLIT)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) s[k] = 0.f;
  const uint32_t n_items = (uint32_t)B * HW;                     // < 2^31 (checked by the launcher)
  const uint32_t nchunk = (n_items + 16 * 256 - 1) / (16 * 256);
  constexpr int NLD = GLU ? 3 : 2;                               // 16-byte loads per element: dy, z value plane (, z gate plane)
  for (uint32_t ch = blockIdx.y; ch < nchunk; ch += gridDim.y) {
    for (int it = 0; it < 16; it += 4) {
      // all loads of four elements first (8-12 independent 16-byte requests per thread), then the arithmetic
      uint4 u[4][NLD];
      bool ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t i = (ch * 16 + it + k) * 256 + threadIdx.x;
        ok[k] = i < n_items;
        uint32_t b, pix;
        dHW.divmod(ok[k] ? i : 0u, b, pix);
        u[k][0] = __ldg(reinterpret_cast<const uint4*>(dy + (((size_t)b * Pf + pf) * (size_t)HW + pix) * 8));
        u[k][1] = __ldg(reinterpret_cast<const uint4*>(z + (((size_t)b * P + pf) * (size_t)HW + pix) * 8));
        if (GLU) u[k][NLD - 1] = __ldg(reinterpret_cast<const uint4*>(z + (((size_t)b * P + pf + Pf) * (size_t)HW + pix) * 8));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        float d[8], xa[8], xg[8];
        unpack8_16<FMT>(u[k][0], d);
        unpack8_16<FMT>(u[k][1], xa);
        if (GLU) unpack8_16<FMT>(u[k][NLD - 1], xg);
        // S2 = sum du * xhat = invstd * (sum du * x - mean * sum du): the raw moment costs one FMA per element here, the
        // affine correction is applied once per CTA below
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (GLU) {
            const int ca = pf * 8 + e, cg = ca + Pf * 8;
            const float a = fmaf(xa[e], coef[ca], coef[C + ca]);
            const float gg = fmaf(xg[e], coef[cg], coef[C + cg]);
            const float sg = __fdividef(1.0f, 1.0f + __expf(-gg));
            const float dua = d[e] * sg, dug = d[e] * a * sg * (1.0f - sg);
            s[e] += dua;
            s[8 + e] = fmaf(dua, xa[e], s[8 + e]);
            s[16 + e] += dug;
            s[24 + e] = fmaf(dug, xg[e], s[24 + e]);
          } else {
            s[e] += d[e];
            s[8 + e] = fmaf(d[e], xa[e], s[8 + e]);
          }
        }
      }
    }
    fold_partials<NA>(s, sacc, warp, lane);                      // fp32 over 16 elements per thread, fp64 from there on
  }
  __syncthreads();
  if (threadIdx.x < NA) {
    const int k = threadIdx.x, e = k & 7, which = (k >> 3) & 1, gate = k >> 4;     // (S1 | S2) x (value | gate)
    double t = 0.0, t1 = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { t += sacc[w][k]; t1 += sacc[w][k & ~8]; }
    const int c = (pf + (gate ? Pf : 0)) * 8 + e;
    if (which) t = (t - (double)mean[c] * t1) * (double)invstd[c];
    atomicAdd(ws + 2 * c + which, t);
  }
}

__global__ void bn16_bwd_final_kernel(const double* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (float)ws[2 * c];
  dgamma[c] = (float)ws[2 * c + 1];
}

template <int FMT, bool GLU>
__global__ void __launch_bounds__(256)
bn16_bwd_apply_kernel(const h16* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, const h16* __restrict__ dy,
                      const float* __restrict__ dgamma, const float* __restrict__ dbeta, h16* __restrict__ dz, int B, int P,
                      int HW, float inv_n, FastDiv dHW, FastDiv dPf) {
  extern __shared__ float coef[];                               // [2][C] affine + [4][C]: mean, invstd, S1/N, S2/N
  const int C = P * 8;
  float* mu = coef + 2 * C; float* rs = mu + C; float* s1 = rs + C; float* s2 = s1 + C;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float a = gamma[c] * invstd[c];
    coef[c] = a; coef[C + c] = beta[c] - mean[c] * a;
    mu[c] = mean[c]; rs[c] = invstd[c]; s1[c] = dbeta[c] * inv_n; s2[c] = dgamma[c] * inv_n;
  }
  __syncthreads();
  const int Pf = GLU ? P / 2 : P;
  const uint32_t total = (uint32_t)B * Pf * HW;                 // < 2^31 (checked by the launcher)
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    uint32_t pix, pf, b, t;
    dHW.divmod(i, t, pix); dPf.divmod(t, b, pf);
    float xa[8], xg[8], dua[8], dug[8], o[8];
    bn16_du<FMT, GLU>(z, dy, coef, C, P, Pf, b, pf, pix, HW, xa, xg, dua, dug);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (int)pf * 8 + e;
      const float xh = (xa[e] - mu[c]) * rs[c];
      o[e] = coef[c] * (dua[e] - s1[c] - xh * s2[c]);
    }
    *reinterpret_cast<uint4*>(dz + (((size_t)b * P + pf) * (size_t)HW + pix) * 8) = pack8_16<FMT>(o);
    if (GLU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = (int)(pf + Pf) * 8 + e;
        const float xh = (xg[e] - mu[c]) * rs[c];
        o[e] = coef[c] * (dug[e] - s1[c] - xh * s2[c]);
      }
      *reinterpret_cast<uint4*>(dz + (((size_t)b * P + pf + Pf) * (size_t)HW + pix) * 8) = pack8_16<FMT>(o);
    }
  }
}

// ---------------------------------------------------------------- nearest x2 upsample and its adjoint
__global__ void __launch_bounds__(256)
upsample2x_blocked_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint32_t planes, int H, int W, FastDiv dWo,
                          FastDiv dHo) {
  const uint32_t total = planes * 4u * H * W;                    // < 2^31 (checked by the launcher)
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    uint32_t ox, oy, pl, t;
    dWo.divmod(i, t, ox); dHo.divmod(t, pl, oy);
    y[i] = __ldg(x + ((size_t)pl * H + (oy >> 1)) * W + (ox >> 1));
  }
}

template <int FMT>
__global__ void __launch_bounds__(256)
sumpool2x2_blocked_kernel(const h16* __restrict__ x, h16* __restrict__ y, uint32_t planes, int H, int W, FastDiv dW,
                          FastDiv dH) {
  const uint32_t total = planes * H * W;                         // output pixels; < 2^29 (checked by the launcher)
  const int Wi = 2 * W;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    uint32_t ox, oy, pl, t;
    dW.divmod(i, t, ox); dH.divmod(t, pl, oy);
    const h16* p = x + (((size_t)pl * 2 * H + 2 * oy) * (size_t)Wi + 2 * ox) * 8;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(p)), a);
    unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(p + 8)), b);
    unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(p + (size_t)Wi * 8)), c);
    unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(p + (size_t)Wi * 8 + 8)), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (a[e] + b[e]) + (c[e] + d[e]);
    *reinterpret_cast<uint4*>(y + (size_t)i * 8) = pack8_16<FMT>(o);
  }
}

static inline int ew_blocks16(size_t total) {
  const size_t b = (total + 255) / 256;
  const size_t cap = (size_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}
static inline int split16(int P, size_t n_items, int per_block) {
  // enough CTAs to fill the machine (~4 per SM) without leaving a CTA fewer than one chunk
  const size_t chunks = (n_items + per_block - 1) / per_block;
  size_t want = (size_t)(g_sm_count > 0 ? g_sm_count : 148) * 8 / (P > 0 ? P : 1);
  if (want < 1) want = 1;
  return (int)(want < chunks ? want : (chunks ? chunks : 1));
}

}  // namespace tgsr

#define TGSR16_FMT(fmt) TGSR_REQUIRE((fmt) == TGSR_FMT_BF16 || (fmt) == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED)

extern "C" int tgsr_bn16_stats(const void* z, float* mean, float* invstd, float* running_mean, float* running_var,
                               float momentum, float eps, double* ws, int B, int C, int HW, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(z); TGSR_REQUIRE_PTR(mean); TGSR_REQUIRE_PTR(invstd); TGSR_REQUIRE_PTR(ws);
  TGSR16_FMT(fmt);
  TGSR_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && HW > 0 && (long long)B * HW < (1ll << 31) - 64 * 256, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((running_mean == nullptr) == (running_var == nullptr), TGSR_ERR_NULL);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  const int P = C / 8;
  const dim3 grid(P, split16(P, (size_t)B * HW, 64 * 256));
  if (fmt == TGSR_FMT_F16) bn16_stats_partial_kernel<TGSR_FMT_F16><<<grid, 256, 0, st>>>(reinterpret_cast<const h16*>(z), ws, B, P, HW, FastDiv((uint32_t)HW));
  else bn16_stats_partial_kernel<TGSR_FMT_BF16><<<grid, 256, 0, st>>>(reinterpret_cast<const h16*>(z), ws, B, P, HW, FastDiv((uint32_t)HW));
  bn16_stats_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, mean, invstd, running_mean, running_var, momentum, eps, C,
                                                            (double)B * HW);
  return launch_status(2);
}

extern "C" int tgsr_bn16_act(const void* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                             const void* residual, void* y, int B, int C, int HW, int act, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(z); TGSR_REQUIRE_PTR(mean); TGSR_REQUIRE_PTR(invstd); TGSR_REQUIRE_PTR(gamma); TGSR_REQUIRE_PTR(beta); TGSR_REQUIRE_PTR(y);
  TGSR16_FMT(fmt);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_GLU, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && HW > 0 && (act != TGSR_ACT_GLU || C % 16 == 0) && C <= 1024, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(!(residual && act == TGSR_ACT_GLU), TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE((long long)B * (C / 8) * HW < (1ll << 31) - (1 << 24), TGSR_ERR_BAD_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  const int P = C / 8, Pf = act == TGSR_ACT_GLU ? P / 2 : P;
  const int blocks = ew_blocks16((size_t)B * Pf * HW);
  const size_t sm = sizeof(float) * 2 * C;
  const h16* zz = reinterpret_cast<const h16*>(z); const h16* rr = reinterpret_cast<const h16*>(residual); h16* yy = reinterpret_cast<h16*>(y);
#define TGSR16_LAUNCH(F, G) bn16_act_kernel<F, G><<<blocks, 256, sm, st>>>(zz, mean, invstd, gamma, beta, rr, yy, B, P, HW, FastDiv((uint32_t)HW), FastDiv((uint32_t)Pf))
  if (fmt == TGSR_FMT_F16) { if (act == TGSR_ACT_GLU) TGSR16_LAUNCH(TGSR_FMT_F16, true); else TGSR16_LAUNCH(TGSR_FMT_F16, false); }
  else { if (act == TGSR_ACT_GLU) TGSR16_LAUNCH(TGSR_FMT_BF16, true); else TGSR16_LAUNCH(TGSR_FMT_BF16, false); }
#undef TGSR16_LAUNCH
  return launch_status();
}

extern "C" int tgsr_bn16_act_bwd(const void* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                 const void* dy, void* dz, float* dgamma, float* dbeta, double* ws, int B, int C, int HW,
                                 int act, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(z); TGSR_REQUIRE_PTR(mean); TGSR_REQUIRE_PTR(invstd); TGSR_REQUIRE_PTR(gamma); TGSR_REQUIRE_PTR(beta);
  TGSR_REQUIRE_PTR(dy); TGSR_REQUIRE_PTR(dz); TGSR_REQUIRE_PTR(dgamma); TGSR_REQUIRE_PTR(dbeta); TGSR_REQUIRE_PTR(ws);
  TGSR16_FMT(fmt);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_GLU, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && HW > 0 && (act != TGSR_ACT_GLU || C % 16 == 0) && C <= 1024, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * (C / 8) * HW < (1ll << 31) - (1 << 24), TGSR_ERR_BAD_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  const int P = C / 8, Pf = act == TGSR_ACT_GLU ? P / 2 : P;
  const FastDiv dHW((uint32_t)HW), dPf((uint32_t)Pf);
  const dim3 rgrid(Pf, split16(Pf, (size_t)B * HW, 16 * 256));
  const int blocks = ew_blocks16((size_t)B * Pf * HW);
  const float inv_n = (float)(1.0 / ((double)B * HW));
  const h16* zz = reinterpret_cast<const h16*>(z); const h16* dd = reinterpret_cast<const h16*>(dy); h16* oo = reinterpret_cast<h16*>(dz);
  const size_t sm2 = sizeof(float) * 2 * C, sm6 = sizeof(float) * 6 * C;
#define TGSR16_BWD(F, G)                                                                                               \
  do {                                                                                                                 \
    bn16_bwd_reduce_kernel<F, G><<<rgrid, 256, sm2, st>>>(zz, mean, invstd, gamma, beta, dd, ws, B, P, HW, dHW);             \
    bn16_bwd_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, dgamma, dbeta, C);                                     \
    bn16_bwd_apply_kernel<F, G><<<blocks, 256, sm6, st>>>(zz, mean, invstd, gamma, beta, dd, dgamma, dbeta, oo, B, P, HW, inv_n, dHW, dPf); \
  } while (0)
  if (fmt == TGSR_FMT_F16) { if (act == TGSR_ACT_GLU) TGSR16_BWD(TGSR_FMT_F16, true); else TGSR16_BWD(TGSR_FMT_F16, false); }
  else { if (act == TGSR_ACT_GLU) TGSR16_BWD(TGSR_FMT_BF16, true); else TGSR16_BWD(TGSR_FMT_BF16, false); }
#undef TGSR16_BWD
  return launch_status(3);
}

extern "C" int tgsr_upsample2x_blocked(const void* x, void* y, int B, int P, int H, int W, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(B > 0 && P > 0 && H > 0 && W > 0 && (long long)B * P * 4 * H * W < (1ll << 31) - (1 << 24), TGSR_ERR_BAD_SHAPE);
  const size_t planes = (size_t)B * P;
  upsample2x_blocked_kernel<<<ew_blocks16(planes * 4 * H * W), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), (uint32_t)planes, H, W, FastDiv(2u * W), FastDiv(2u * H));
  return launch_status();
}

extern "C" int tgsr_sumpool2x2_blocked(const void* x, void* y, int B, int P, int H, int W, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(y);
  TGSR16_FMT(fmt);
  TGSR_REQUIRE(B > 0 && P > 0 && H > 0 && W > 0 && (long long)B * P * H * W < (1ll << 29), TGSR_ERR_BAD_SHAPE);
  const size_t planes = (size_t)B * P;
  const int blocks = ew_blocks16(planes * H * W);
  const FastDiv dW((uint32_t)W), dH((uint32_t)H);
  if (fmt == TGSR_FMT_F16)
    sumpool2x2_blocked_kernel<TGSR_FMT_F16><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const h16*>(x), reinterpret_cast<h16*>(y), (uint32_t)planes, H, W, dW, dH);
  else
    sumpool2x2_blocked_kernel<TGSR_FMT_BF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const h16*>(x), reinterpret_cast<h16*>(y), (uint32_t)planes, H, W, dW, dH);
  return launch_status();
}
