// Training-mode building blocks of the generator (fp32 NCHW): batch-statistics BatchNorm forward / backward with
// the fused GLU and residual add, the pieces of the conv3x3 / conv5x5 backward that the forward kernel cannot
// provide (weight transform for the data gradient, 2x2 sum-pool for the folded nearest upsample, weight
// gradient), and the tanh image head backward.
//
// reference: nn.BatchNorm2d in train mode (util.py:64, :78, :115-122, model.py:228-232), GLU util.py:45-53,
// ResBlock skip util.py:126-130, nn.Upsample(2) util.py:76, conv_output + Tanh and `one*y + a*SRb`
// model.py:223-226,280.  The reference gets these gradients from autograd/cuDNN; here they are explicit kernels.
#include "common.cuh"

namespace tgsr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of up to 4 doubles (256 threads); result valid in every thread
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N]) {
  __shared__ double red[N][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) red[k][warp] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[k][w];
    v[k] = s;
  }
  __syncthreads();
}

// ---------------------------------------------------------------- BatchNorm statistics (train forward)
// grid (C, SPLIT): CTA (c, s) sums channel c over images s, s+SPLIT, ... (fp32 partials per image folded into double
// accumulators) and adds its two sums to ws[2c], ws[2c+1] (double atomics); bn_stats_final turns them into mean / invstd
// and updates the running statistics.  One CTA per channel left most of the SMs idle and took ~1 ms per layer.
__global__ void __launch_bounds__(256)
bn_stats_partial_kernel(const float* __restrict__ x, double* __restrict__ ws, int B, int C, int HW) {
  const int c = blockIdx.x;
  double acc[2] = {0.0, 0.0};
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const float* p = x + ((size_t)b * C + c) * HW;
    float s = 0.f, q = 0.f;
    for (int i = threadIdx.x; i < HW; i += 256) { const float v = __ldg(p + i); s += v; q = fmaf(v, v, q); }
    acc[0] += s; acc[1] += q;
  }
  block_sum<2>(acc);
  if (threadIdx.x == 0) { atomicAdd(ws + 2 * c, acc[0]); atomicAdd(ws + 2 * c + 1, acc[1]); }
}

__global__ void bn_stats_final_kernel(const double* __restrict__ ws, float* __restrict__ mean, float* __restrict__ invstd,
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
// y[b,c] = act(gamma*(x-mean)*invstd + beta) (+ residual); GLU: C_out = C/2, y = z[c] * sigmoid(z[c + C/2]).
template <bool GLU>
__global__ void __launch_bounds__(256)
bn_act_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
              const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ residual,
              float* __restrict__ y, int C, int HW, size_t total) {
  const int Cf = GLU ? C / 2 : C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int p = (int)(i % HW), c = (int)((i / HW) % Cf);
    const size_t b = i / ((size_t)HW * Cf);
    const size_t xi = (b * C + c) * HW + p;
    float z = fmaf((__ldg(x + xi) - mean[c]) * invstd[c], gamma[c], beta[c]);
    if (GLU) {
      const int cg = c + Cf;
      const float g = fmaf((__ldg(x + xi + (size_t)Cf * HW) - mean[cg]) * invstd[cg], gamma[cg], beta[cg]);
      z = z * (1.0f / (1.0f + expf(-g)));
    }
    if (residual) z += __ldg(residual + i);
    y[i] = z;
  }
}

// ---------------------------------------------------------------- backward, stage 1: per-channel sums
// dz = dL/dz of the normalised pre-activation z; S1 = sum dz, S2 = sum dz * xhat.  One CTA per output channel
// (GLU: the CTA owns the value channel c and its gate channel c + C/2).  Writes dgamma = S2, dbeta = S1.
template <bool GLU>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ dy,
                     double* __restrict__ ws, int B, int C, int HW) {
  const int Cf = GLU ? C / 2 : C;
  const int c = blockIdx.x, cg = c + Cf;
  const float m0 = mean[c], r0 = invstd[c], g0 = gamma[c], b0 = beta[c];
  const float m1 = GLU ? mean[cg] : 0.f, r1 = GLU ? invstd[cg] : 0.f, g1 = GLU ? gamma[cg] : 0.f, b1 = GLU ? beta[cg] : 0.f;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = blockIdx.y; b < B; b += gridDim.y) {             // grid (Cf, SPLIT): images s, s+SPLIT, ...
    const float* xa = x + ((size_t)b * C + c) * HW;
    const float* xg = x + ((size_t)b * C + cg) * HW;
    const float* d = dy + ((size_t)b * Cf + c) * HW;
    float s1a = 0.f, s2a = 0.f, s1g = 0.f, s2g = 0.f;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float xh = (__ldg(xa + i) - m0) * r0;
      float dz = __ldg(d + i);
      if (GLU) {
        const float xhg = (__ldg(xg + i) - m1) * r1;
        const float za = fmaf(xh, g0, b0), zg = fmaf(xhg, g1, b1);
        const float sg = 1.0f / (1.0f + __expf(-zg));
        const float dzg = dz * za * sg * (1.0f - sg);
        dz *= sg;
        s1g += dzg; s2g = fmaf(dzg, xhg, s2g);
      }
      s1a += dz; s2a = fmaf(dz, xh, s2a);
    }
    acc[0] += s1a; acc[1] += s2a; acc[2] += s1g; acc[3] += s2g;
  }
  block_sum<4>(acc);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 2 * c, acc[0]); atomicAdd(ws + 2 * c + 1, acc[1]);
    if (GLU) { atomicAdd(ws + 2 * cg, acc[2]); atomicAdd(ws + 2 * cg + 1, acc[3]); }
  }
}

__global__ void bn_bwd_final_kernel(const double* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (float)ws[2 * c];
  dgamma[c] = (float)ws[2 * c + 1];
}

// ---------------------------------------------------------------- backward, stage 2: dx
// dx = gamma * invstd * (dz - S1/N - xhat * S2/N), dz recomputed from dy as in stage 1.
template <bool GLU>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ dy,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dx, int C,
                    int HW, size_t total, float inv_n) {
  const int Cf = GLU ? C / 2 : C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int p = (int)(i % HW), c = (int)((i / HW) % Cf);
    const size_t b = i / ((size_t)HW * Cf);
    const size_t xi = (b * C + c) * HW + p;
    const float xh = (__ldg(x + xi) - mean[c]) * invstd[c];
    float dz = __ldg(dy + i);
    if (GLU) {
      const int cg = c + Cf;
      const size_t xgi = xi + (size_t)Cf * HW;
      const float xhg = (__ldg(x + xgi) - mean[cg]) * invstd[cg];
      const float za = fmaf(xh, gamma[c], beta[c]), zg = fmaf(xhg, gamma[cg], beta[cg]);
      const float sg = 1.0f / (1.0f + __expf(-zg));
      const float dzg = dz * za * sg * (1.0f - sg);
      dz *= sg;
      dx[xgi] = gamma[cg] * invstd[cg] * (dzg - dbeta[cg] * inv_n - xhg * dgamma[cg] * inv_n);
    }
    dx[xi] = gamma[c] * invstd[c] * (dz - dbeta[c] * inv_n - xh * dgamma[c] * inv_n);
  }
}

// ---------------------------------------------------------------- conv backward helpers
// wt[ci][co][k-1-ky][k-1-kx] = w[co][ci][ky][kx]: the data gradient of a stride-1 same-padding conv is the same
// conv of dy with wt.
__global__ void conv_weight_flip_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int K) {
  const int n = Cout * Cin * K * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int kx = i % K, ky = (i / K) % K, ci = (i / (K * K)) % Cin, co = i / (K * K * Cin);
    wt[(((size_t)ci * Cout + co) * K + (K - 1 - ky)) * K + (K - 1 - kx)] = w[i];
  }
}

// backward of nn.Upsample(scale_factor=2, nearest): y[b,c,h,w] = sum of the 2x2 block of x [B,C,2H,2W]
__global__ void __launch_bounds__(256)
sumpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, size_t total) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const size_t bc = i / ((size_t)W * H);
    const float2* r0 = reinterpret_cast<const float2*>(x + (bc * 2 * H + 2 * h) * 2 * W) + w;
    const float2* r1 = reinterpret_cast<const float2*>(x + (bc * 2 * H + 2 * h + 1) * 2 * W) + w;
    const float2 a = __ldg(r0), b = __ldg(r1);
    y[i] = (a.x + a.y) + (b.x + b.y);
  }
}

// ---------------------------------------------------------------- conv weight gradient
// dw[co][ci][ky][kx] = sum_{b,y,x} dy[b,co,y,x] * xin[b,ci,y+ky-R,x+kx-R]   (xin = nearest-upsampled x if up2x)
// grid (co blocks of 8, ci blocks of 8, splits); each CTA walks its share of (image, 16x16 tile) pairs, stages the
// dy tile and the x halo tile in shared memory and accumulates an 8x8xKxK block in registers: thread (co, ci, quarter)
// sums K*K taps over a quarter of the tile's pixels; partial blocks meet in shared memory, then one atomicAdd per
// output and CTA.
constexpr int WG_TILE = 16;
template <int K>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int B, int Cin, int H,
                  int W, int Cout, int up2x) {
  constexpr int R = K / 2, TS = WG_TILE + K - 1, KK = K * K;
  __shared__ float xs[8][TS][TS + 1];
  __shared__ float ds[8][WG_TILE][WG_TILE + 1];
  __shared__ float red[4][64][KK];
  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const int tiles_x = ceil_div(Wo, WG_TILE), tiles_y = ceil_div(Ho, WG_TILE);
  const int n_work = B * tiles_x * tiles_y;
  const int co0 = blockIdx.x * 8, ci0 = blockIdx.y * 8;
  const int tid = threadIdx.x, pair = tid & 63, quarter = tid >> 6, co = pair >> 3, ci = pair & 7;
  float acc[KK];
#pragma unroll
  for (int k = 0; k < KK; ++k) acc[k] = 0.f;
  for (int wk = blockIdx.z; wk < n_work; wk += gridDim.z) {
    const int b = wk / (tiles_x * tiles_y), t = wk % (tiles_x * tiles_y);
    const int ty0 = (t / tiles_x) * WG_TILE, tx0 = (t % tiles_x) * WG_TILE;
    __syncthreads();
    for (int idx = tid; idx < 8 * TS * TS; idx += 256) {
      const int c = idx / (TS * TS), r = idx % (TS * TS), yy = r / TS, xx = r % TS;
      const int gy = ty0 + yy - R, gx = tx0 + xx - R;
      float v = 0.f;
      if (ci0 + c < Cin && gy >= 0 && gy < Ho && gx >= 0 && gx < Wo) {
        const int sy = up2x ? (gy >> 1) : gy, sx = up2x ? (gx >> 1) : gx;
        v = __ldg(x + (((size_t)b * Cin + ci0 + c) * H + sy) * W + sx);
      }
      xs[c][yy][xx] = v;
    }
    for (int idx = tid; idx < 8 * WG_TILE * WG_TILE; idx += 256) {
      const int c = idx / (WG_TILE * WG_TILE), r = idx % (WG_TILE * WG_TILE), yy = r / WG_TILE, xx = r % WG_TILE;
      const int gy = ty0 + yy, gx = tx0 + xx;
      float v = 0.f;
      if (co0 + c < Cout && gy < Ho && gx < Wo) v = __ldg(dy + (((size_t)b * Cout + co0 + c) * Ho + gy) * Wo + gx);
      ds[c][yy][xx] = v;
    }
    __syncthreads();
    for (int yy = quarter * 4; yy < quarter * 4 + 4; ++yy) {
#pragma unroll 4
      for (int xx = 0; xx < WG_TILE; ++xx) {
        const float d = ds[co][yy][xx];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) acc[ky * K + kx] = fmaf(d, xs[ci][yy + ky][xx + kx], acc[ky * K + kx]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KK; ++k) red[quarter][pair][k] = acc[k];
  __syncthreads();
  for (int idx = tid; idx < 64 * KK; idx += 256) {
    const int pr = idx / KK, k = idx % KK, o = pr >> 3, i = pr & 7;
    if (co0 + o < Cout && ci0 + i < Cin) {
      const float v = (red[0][pr][k] + red[1][pr][k]) + (red[2][pr][k] + red[3][pr][k]);
      atomicAdd(dw + ((size_t)(co0 + o) * Cin + ci0 + i) * KK + k, v);
    }
  }
}

// ---------------------------------------------------------------- tanh image head backward
// forward (model.py:224,280): y = one * tanh(conv) + a * srb.  Given y, srb, dy:
//   t = (y - a*srb)/one;  dconv = dy * one * (1 - t^2);  dsrb = a * dy (written if dsrb != NULL).
__global__ void __launch_bounds__(256)
tanh_head_bwd_kernel(const float* __restrict__ y, const float* __restrict__ srb, const float* __restrict__ dy, float one,
                     float a, float* __restrict__ dconv, float* __restrict__ dsrb, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float g = __ldg(dy + i);
    const float t = (__ldg(y + i) - (srb ? a * __ldg(srb + i) : 0.f)) / one;
    dconv[i] = g * one * (1.0f - t * t);
    if (dsrb) dsrb[i] = a * g;
  }
}

static inline unsigned ew_blocks(size_t total) {
  const size_t sms = g_sm_count > 0 ? g_sm_count : 148;
  size_t b = (total + 255) / 256;
  if (b > sms * 16) b = sms * 16;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace tgsr

using namespace tgsr;

static inline int bn_split(int channels, int B) {
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int s = ceil_div(4 * sms, channels);
  return s < 1 ? 1 : (s > B ? B : s);
}

extern "C" int tgsr_bn_stats_f32(const float* x, float* mean, float* invstd, float* running_mean, float* running_var,
                                 float momentum, float eps, double* ws, int B, int C, int HW, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(mean); TGSR_REQUIRE_PTR(invstd); TGSR_REQUIRE_PTR(ws);
  TGSR_REQUIRE((running_mean == nullptr) == (running_var == nullptr), TGSR_ERR_NULL);
  TGSR_REQUIRE(B > 0 && C > 0 && HW > 0 && C <= 65535, TGSR_ERR_BAD_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(ws, 0, (size_t)2 * C * sizeof(double), st));
  bn_stats_partial_kernel<<<dim3(C, bn_split(C, B)), 256, 0, st>>>(x, ws, B, C, HW);
  bn_stats_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, mean, invstd, running_mean, running_var, momentum, eps, C,
                                                         (double)B * HW);
  return launch_status(2);
}

extern "C" int tgsr_bn_act_f32(const float* x, const float* mean, const float* invstd, const float* gamma,
                               const float* beta, const float* residual, float* y, int B, int C, int HW, int act,
                               tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(mean); TGSR_REQUIRE_PTR(invstd); TGSR_REQUIRE_PTR(gamma); TGSR_REQUIRE_PTR(beta);
  TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(B > 0 && C > 0 && HW > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_GLU, TGSR_ERR_UNSUPPORTED);
  if (act == TGSR_ACT_GLU) TGSR_REQUIRE(C % 2 == 0, TGSR_ERR_BAD_SHAPE);
  const size_t total = (size_t)B * (act == TGSR_ACT_GLU ? C / 2 : C) * HW;
  cudaStream_t st = (cudaStream_t)stream;
  if (act == TGSR_ACT_GLU) bn_act_kernel<true><<<ew_blocks(total), 256, 0, st>>>(x, mean, invstd, gamma, beta, residual, y, C, HW, total);
  else bn_act_kernel<false><<<ew_blocks(total), 256, 0, st>>>(x, mean, invstd, gamma, beta, residual, y, C, HW, total);
  return launch_status();
}

extern "C" int tgsr_bn_act_bwd_f32(const float* x, const float* mean, const float* invstd, const float* gamma,
                                   const float* beta, const float* dy, float* dx, float* dgamma, float* dbeta, double* ws,
                                   int B, int C, int HW, int act, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(mean); TGSR_REQUIRE_PTR(invstd); TGSR_REQUIRE_PTR(gamma); TGSR_REQUIRE_PTR(beta);
  TGSR_REQUIRE_PTR(dy); TGSR_REQUIRE_PTR(dx); TGSR_REQUIRE_PTR(dgamma); TGSR_REQUIRE_PTR(dbeta); TGSR_REQUIRE_PTR(ws);
  TGSR_REQUIRE(B > 0 && C > 0 && HW > 0 && C <= 65535, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_GLU, TGSR_ERR_UNSUPPORTED);
  if (act == TGSR_ACT_GLU) TGSR_REQUIRE(C % 2 == 0, TGSR_ERR_BAD_SHAPE);
  const int Cf = act == TGSR_ACT_GLU ? C / 2 : C;
  const size_t total = (size_t)B * Cf * HW;
  const float inv_n = 1.0f / ((float)B * (float)HW);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(ws, 0, (size_t)2 * C * sizeof(double), st));
  const dim3 rgrid(Cf, bn_split(Cf, B));
  if (act == TGSR_ACT_GLU) {
    bn_bwd_reduce_kernel<true><<<rgrid, 256, 0, st>>>(x, mean, invstd, gamma, beta, dy, ws, B, C, HW);
    bn_bwd_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, dgamma, dbeta, C);
    bn_bwd_apply_kernel<true><<<ew_blocks(total), 256, 0, st>>>(x, mean, invstd, gamma, beta, dy, dgamma, dbeta, dx, C, HW, total, inv_n);
  } else {
    bn_bwd_reduce_kernel<false><<<rgrid, 256, 0, st>>>(x, mean, invstd, gamma, beta, dy, ws, B, C, HW);
    bn_bwd_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, dgamma, dbeta, C);
    bn_bwd_apply_kernel<false><<<ew_blocks(total), 256, 0, st>>>(x, mean, invstd, gamma, beta, dy, dgamma, dbeta, dx, C, HW, total, inv_n);
  }
  return launch_status(3);
}

extern "C" int tgsr_conv_weight_flip_f32(const float* w, float* wt, int Cout, int Cin, int ksize, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(wt);
  TGSR_REQUIRE(Cout > 0 && Cin > 0 && ksize > 0, TGSR_ERR_BAD_SHAPE);
  const int n = Cout * Cin * ksize * ksize;
  conv_weight_flip_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, wt, Cout, Cin, ksize);
  return launch_status();
}

extern "C" int tgsr_sumpool2x2_f32(const float* x, float* y, int BC, int H, int W, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(BC > 0 && H > 0 && W > 0, TGSR_ERR_BAD_SHAPE);
  const size_t total = (size_t)BC * H * W;
  sumpool2x2_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, y, H, W, total);
  return launch_status();
}

extern "C" int tgsr_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, int B, int Cin, int H, int W, int Cout,
                                     int ksize, int up2x, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(dy); TGSR_REQUIRE_PTR(dw);
  TGSR_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(ksize == 3 || ksize == 5, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cout * Cin * ksize * ksize * sizeof(float), st));
  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const int n_work = B * ceil_div(Wo, WG_TILE) * ceil_div(Ho, WG_TILE);
  const int blocks_xy = ceil_div(Cout, 8) * ceil_div(Cin, 8);
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  int splits = ceil_div(4 * sms, blocks_xy);
  splits = splits < 1 ? 1 : (splits > n_work ? n_work : splits);
  if (splits > 65535) splits = 65535;
  dim3 grid(ceil_div(Cout, 8), ceil_div(Cin, 8), splits);
  if (ksize == 3) conv_wgrad_kernel<3><<<grid, 256, 0, st>>>(x, dy, dw, B, Cin, H, W, Cout, up2x);
  else conv_wgrad_kernel<5><<<grid, 256, 0, st>>>(x, dy, dw, B, Cin, H, W, Cout, up2x);
  return launch_status();
}

extern "C" int tgsr_tanh_head_bwd_f32(const float* y, const float* srb, const float* dy, float one, float a, float* dconv,
                                      float* dsrb, size_t n, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(y); TGSR_REQUIRE_PTR(dy); TGSR_REQUIRE_PTR(dconv);
  TGSR_REQUIRE(n > 0 && one != 0.0f, TGSR_ERR_BAD_SHAPE);
  tanh_head_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(y, srb, dy, one, a, dconv, dsrb, n);
  return launch_status();
}
