// Small kernels of the bf16 blocked-layout pipeline that are not GEMM-shaped:
//   * conv_in: conv3x3(3 -> 2C) + folded BN + GLU from the NCHW fp32 LR image into C blocked bf16
//     channels (INIT_STAGE_GImgup.im2f util.py:741-744 / NetG_high*.convin model.py:228); K = 27,
//     3.5 MFLOP per image -- CUDA cores, fp32 math.
//   (attention on blocked activations lives in attention_mma.cu)
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
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the attention kernel that follows starts its prologue early
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
          // branch-free: clamped address + select, so the 27 (54) loads are independent and in flight together
          const int yy = py + ky - 1, xx = px + kx - 1;
          const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
          const size_t o = (((size_t)b * 3 + c) * H + (ok ? yy : py)) * W + (ok ? xx : px);
          float v = __ldg(x + o);
          if (xsub) v -= __ldg(xsub + o);
          in[c * 9 + ky * 3 + kx] = ok ? v : 0.f;
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

int init_blocked_misc() { return TGSR_OK; }

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
