// fp32 NCHW direct convolution with the fused epilogues of the TGSR generator blocks.
// This is the exact-precision path (any Cin/Cout, 3x3 or 5x5, optional nearest-x2 upsample of
// the input); the bf16 tcgen05 implicit-GEMM path in conv_tc.cu is the throughput path.
//
// reference: util.py:62-69 (conv3x3/conv5x5), nn.BatchNorm2d eval folded to scale/shift,
// util.py:45-53 (GLU), util.py:126-130 (residual add), util.py:76 (nearest upsample),
// util.py:899-901 / model.py:224 (tanh heads), model.py:280 (one*y + a*SRb).
#include "common.cuh"

namespace tgsr {

constexpr int CD_TILE = 16;   // 16x16 output pixels per CTA, one per thread
constexpr int CD_COB = 8;     // final output channels per CTA
constexpr int CD_CCH = 8;     // input-channel chunk staged in shared memory

template <int K, bool GLU>
__global__ void __launch_bounds__(CD_TILE * CD_TILE)
conv2d_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ residual,
                  const float* __restrict__ addend, float out_scale, float addend_scale, float* __restrict__ y,
                  int Cin, int H, int W, int Cout, int up2x, int act_tanh) {
  constexpr int R = K / 2;
  constexpr int TS = CD_TILE + K - 1;
  constexpr int NH = GLU ? 2 : 1;
  __shared__ float tile[CD_CCH][TS][TS + 1];
  __shared__ float wsm[NH][CD_COB][CD_CCH][K * K];

  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const int tiles_x = ceil_div(Wo, CD_TILE);
  const int tx0 = (blockIdx.x % tiles_x) * CD_TILE, ty0 = (blockIdx.x / tiles_x) * CD_TILE;
  const int co0 = blockIdx.y * CD_COB;
  const int b = blockIdx.z;
  const int Cf = GLU ? Cout / 2 : Cout;          // channels written
  const int tid = threadIdx.x, tx = tid % CD_TILE, ty = tid / CD_TILE;

  float acc[NH][CD_COB];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int o = 0; o < CD_COB; ++o) acc[h][o] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CD_CCH) {
    __syncthreads();
    // stage the input halo tile (in output coordinates; nearest upsample folded into the index)
    for (int idx = tid; idx < CD_CCH * TS * TS; idx += CD_TILE * CD_TILE) {
      const int c = idx / (TS * TS), r = idx % (TS * TS), yy = r / TS, xx = r % TS;
      const int gy = ty0 + yy - R, gx = tx0 + xx - R;
      // branch-free (clamped address + select): the loads of the staging loop stay independent and overlap
      const bool ok = c0 + c < Cin && gy >= 0 && gy < Ho && gx >= 0 && gx < Wo;
      const int cy = ok ? gy : 0, cx = ok ? gx : 0, cc = ok ? c0 + c : 0;
      const int sy = up2x ? (cy >> 1) : cy, sx = up2x ? (cx >> 1) : cx;
      const float v = __ldg(x + (((size_t)b * Cin + cc) * H + sy) * W + sx);
      tile[c][yy][xx] = ok ? v : 0.f;
    }
    for (int idx = tid; idx < NH * CD_COB * CD_CCH * K * K; idx += CD_TILE * CD_TILE) {
      const int kk = idx % (K * K), c = (idx / (K * K)) % CD_CCH, o = (idx / (K * K * CD_CCH)) % CD_COB,
                h = idx / (K * K * CD_CCH * CD_COB);
      const int co = co0 + o + h * Cf;
      const bool okw = co0 + o < Cf && c0 + c < Cin;
      const float v = __ldg(w + (okw ? ((size_t)co * Cin + c0 + c) * K * K + kk : 0));
      wsm[h][o][c][kk] = okw ? v : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int c = 0; c < CD_CCH; ++c) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float v = tile[c][ty + ky][tx + kx];
#pragma unroll
          for (int h = 0; h < NH; ++h)
#pragma unroll
            for (int o = 0; o < CD_COB; ++o) acc[h][o] = fmaf(v, wsm[h][o][c][ky * K + kx], acc[h][o]);
        }
    }
  }

  const int oy = ty0 + ty, ox = tx0 + tx;
  if (oy >= Ho || ox >= Wo) return;
#pragma unroll
  for (int o = 0; o < CD_COB; ++o) {
    const int c = co0 + o;
    if (c >= Cf) break;
    float z = acc[0][o];
    if (scale) z = fmaf(z, __ldg(scale + c), __ldg(shift + c));
    if (GLU) {
      float g = acc[NH - 1][o];
      if (scale) g = fmaf(g, __ldg(scale + c + Cf), __ldg(shift + c + Cf));
      z = z * (1.0f / (1.0f + expf(-g)));
    }
    if (act_tanh) z = tanhf(z);
    const size_t oi = (((size_t)b * Cf + c) * Ho + oy) * Wo + ox;
    if (residual) z += __ldg(residual + oi);
    z *= out_scale;
    if (addend) z = fmaf(addend_scale, __ldg(addend + oi), z);
    y[oi] = z;
  }
}

__global__ void bn_fold_kernel(const float* g, const float* b, const float* m, const float* v, float eps,
                               float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // same operation order as ATen's eval batch_norm: invstd = 1/sqrt(var+eps)
  const float s = g[c] * (1.0f / sqrtf(v[c] + eps));
  scale[c] = s;
  shift[c] = b[c] - m[c] * s;
}

}  // namespace tgsr

extern "C" int tgsr_conv2d_f32(const float* x, const float* w, const float* scale, const float* shift,
                               const float* residual, const float* addend, float out_scale, float addend_scale,
                               float* y, int B, int Cin, int H, int W, int Cout, int ksize, int up2x, int act,
                               tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((scale == nullptr) == (shift == nullptr), TGSR_ERR_NULL);
  TGSR_REQUIRE(ksize == 3 || ksize == 5, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_GLU || act == TGSR_ACT_TANH, TGSR_ERR_UNSUPPORTED);
  if (act == TGSR_ACT_GLU) TGSR_REQUIRE(Cout % 2 == 0, TGSR_ERR_BAD_SHAPE);   // util.py:51 assert
  TGSR_REQUIRE(B <= 65535, TGSR_ERR_UNSUPPORTED);
  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const int Cf = act == TGSR_ACT_GLU ? Cout / 2 : Cout;
  dim3 grid(ceil_div(Wo, CD_TILE) * ceil_div(Ho, CD_TILE), ceil_div(Cf, CD_COB), B);
  dim3 block(CD_TILE * CD_TILE);
  cudaStream_t st = (cudaStream_t)stream;
  const int tanh_f = act == TGSR_ACT_TANH;
#define TGSR_LAUNCH_CD(K, G)                                                                                     \
  conv2d_f32_kernel<K, G><<<grid, block, 0, st>>>(x, w, scale, shift, residual, addend, out_scale, addend_scale, \
                                                  y, Cin, H, W, Cout, up2x, tanh_f)
  if (ksize == 3) {
    if (act == TGSR_ACT_GLU) TGSR_LAUNCH_CD(3, true); else TGSR_LAUNCH_CD(3, false);
  } else {
    if (act == TGSR_ACT_GLU) TGSR_LAUNCH_CD(5, true); else TGSR_LAUNCH_CD(5, false);
  }
#undef TGSR_LAUNCH_CD
  return launch_status();
}

extern "C" int tgsr_bn_fold_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                                float eps, float* scale, float* shift, int C, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(gamma); TGSR_REQUIRE_PTR(beta); TGSR_REQUIRE_PTR(mean); TGSR_REQUIRE_PTR(var);
  TGSR_REQUIRE_PTR(scale); TGSR_REQUIRE_PTR(shift);
  TGSR_REQUIRE(C > 0, TGSR_ERR_BAD_SHAPE);
  bn_fold_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, scale, shift, C);
  return launch_status();
}
