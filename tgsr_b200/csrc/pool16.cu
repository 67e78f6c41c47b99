// 3x3 / stride-2 max pooling (no padding, floor mode) on channels-last 16-bit tensors with an explicit backward: the four
// pooling layers of the frozen CNN_ENCODER (util.py:262-368: F.max_pool2d(x, 3, 2) in the Inception-v3 stem and in
// Mixed_6a / Mixed_7a) inside the bf16 training step, where the library kernels cost 1.4 ms per step (forward + an
// atomics-based backward).  Both directions are gather kernels: one thread = 8 channels (16 bytes) of one pixel.
//   forward : y[b,oy,ox,c] = max over the window, idx = position of the FIRST maximum (ky*3 + kx, torch's tie rule)
//   backward: gx[b,iy,ix,c] = sum of gy over the (at most four) windows that contain the pixel AND selected it.
#include "common.cuh"
#include "fmt16.cuh"

namespace tgsr {

typedef uint16_t h16;

template <int FMT>
__global__ void __launch_bounds__(256)
maxpool3s2_fwd_kernel(const h16* __restrict__ x, h16* __restrict__ y, uint8_t* __restrict__ idx, int H, int W, int C8, int Ho,
                      int Wo, uint32_t total, FastDiv dC8, FastDiv dWo, FastDiv dHo) {
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    uint32_t c8, ox, oy, b, t;
    dC8.divmod(i, t, c8); dWo.divmod(t, t, ox); dHo.divmod(t, b, oy);
    const h16* xp = x + ((((size_t)b * H + 2 * oy) * W + 2 * ox) * C8 + c8) * 8;
    float best[8];
    uint32_t arg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float v[8];
        unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(xp + ((size_t)ky * W + kx) * C8 * 8)), v);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v[e] > best[e] || (ky == 0 && kx == 0)) { best[e] = v[e]; arg[e] = ky * 3 + kx; }
      }
    *reinterpret_cast<uint4*>(y + (size_t)i * 8) = pack8_16<FMT>(best);
    uint2 a;
    a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2*>(idx + (size_t)i * 8) = a;
  }
}

template <int FMT>
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const h16* __restrict__ gy, const uint8_t* __restrict__ idx, h16* __restrict__ gx, int H, int W, int C8,
                      int Ho, int Wo, uint32_t total, FastDiv dC8, FastDiv dW, FastDiv dH) {
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    uint32_t c8, ix, iy, b, t;
    dC8.divmod(i, t, c8); dW.divmod(t, t, ix); dH.divmod(t, b, iy);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    const int oy_lo = iy >= 2 ? (int)((iy - 1) >> 1) : 0, oy_hi = min((int)(iy >> 1), Ho - 1);
    const int ox_lo = ix >= 2 ? (int)((ix - 1) >> 1) : 0, ox_hi = min((int)(ix >> 1), Wo - 1);
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const uint32_t want = (uint32_t)((iy - 2 * oy) * 3 + (ix - 2 * ox));
        const size_t o = ((((size_t)b * Ho + oy) * Wo + ox) * C8 + c8) * 8;
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(idx + o));
        float g[8];
        unpack8_16<FMT>(__ldg(reinterpret_cast<const uint4*>(gy + o)), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t sel = ((e < 4 ? a.x : a.y) >> (8 * (e & 3))) & 0xffu;
          if (sel == want) acc[e] += g[e];
        }
      }
    *reinterpret_cast<uint4*>(gx + (size_t)i * 8) = pack8_16<FMT>(acc);
  }
}

static inline int pool_blocks(size_t total) {
  const size_t b = (total + 255) / 256, cap = (size_t)(g_sm_count > 0 ? g_sm_count : 148) * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace tgsr

extern "C" int tgsr_maxpool3s2_nhwc16(const void* x, void* y, uint8_t* idx, int B, int H, int W, int C, int fmt,
                                      tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(y); TGSR_REQUIRE_PTR(idx);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && H >= 3 && W >= 3 && C > 0 && C % 8 == 0, TGSR_ERR_BAD_SHAPE);
  const int Ho = (H - 3) / 2 + 1, Wo = (W - 3) / 2 + 1, C8 = C / 8;
  const long long total = (long long)B * Ho * Wo * C8;
  TGSR_REQUIRE(total < (1ll << 31) - (1 << 24), TGSR_ERR_BAD_SHAPE);
  const h16* xx = reinterpret_cast<const h16*>(x); h16* yy = reinterpret_cast<h16*>(y);
  const FastDiv dC8((uint32_t)C8), dWo((uint32_t)Wo), dHo((uint32_t)Ho);
  if (fmt == TGSR_FMT_F16)
    maxpool3s2_fwd_kernel<TGSR_FMT_F16><<<pool_blocks((size_t)total), 256, 0, (cudaStream_t)stream>>>(xx, yy, idx, H, W, C8, Ho, Wo, (uint32_t)total, dC8, dWo, dHo);
  else
    maxpool3s2_fwd_kernel<TGSR_FMT_BF16><<<pool_blocks((size_t)total), 256, 0, (cudaStream_t)stream>>>(xx, yy, idx, H, W, C8, Ho, Wo, (uint32_t)total, dC8, dWo, dHo);
  return launch_status();
}

extern "C" int tgsr_maxpool3s2_nhwc16_bwd(const void* gy, const uint8_t* idx, void* gx, int B, int H, int W, int C, int fmt,
                                          tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(gy); TGSR_REQUIRE_PTR(idx); TGSR_REQUIRE_PTR(gx);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && H >= 3 && W >= 3 && C > 0 && C % 8 == 0, TGSR_ERR_BAD_SHAPE);
  const int Ho = (H - 3) / 2 + 1, Wo = (W - 3) / 2 + 1, C8 = C / 8;
  const long long total = (long long)B * H * W * C8;
  TGSR_REQUIRE(total < (1ll << 31) - (1 << 24), TGSR_ERR_BAD_SHAPE);
  const h16* gg = reinterpret_cast<const h16*>(gy); h16* xx = reinterpret_cast<h16*>(gx);
  const FastDiv dC8((uint32_t)C8), dW((uint32_t)W), dH((uint32_t)H);
  if (fmt == TGSR_FMT_F16)
    maxpool3s2_bwd_kernel<TGSR_FMT_F16><<<pool_blocks((size_t)total), 256, 0, (cudaStream_t)stream>>>(gg, idx, xx, H, W, C8, Ho, Wo, (uint32_t)total, dC8, dW, dH);
  else
    maxpool3s2_bwd_kernel<TGSR_FMT_BF16><<<pool_blocks((size_t)total), 256, 0, (cudaStream_t)stream>>>(gg, idx, xx, H, W, C8, Ho, Wo, (uint32_t)total, dC8, dW, dH);
  return launch_status();
}

// ReLU backward for the frozen encoder's fused conv-bias-ReLU: out = gy where y > 0 else 0, channels-last 16-bit.  gy may be a
// CHANNEL SLICE of a wider channels-last tensor (the gradient of a torch.cat along C): its pixels are gy_stride elements
// apart.  The library kernel for that case is a scalar strided loop (0.4 ms per step); here every access is a 16-byte vector.
namespace tgsr {
template <int FMT>
__global__ void __launch_bounds__(256)
relu_bwd_nhwc16_kernel(const h16* __restrict__ gy, size_t gy_stride, const h16* __restrict__ y, h16* __restrict__ out, int C8,
                       uint32_t total, FastDiv dC8) {
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    uint32_t pix, c8;
    dC8.divmod(i, pix, c8);
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(gy + (size_t)pix * gy_stride + c8 * 8));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(y + (size_t)i * 8));
    // 16-bit floats: positive <=> sign bit clear and not zero (a NaN activation passes the gradient, as torch's y > 0 does not;
    // the encoder's activations are finite)
    uint4 o;
    auto keep = [](uint32_t gv, uint32_t yv) {
      const uint32_t lo = ((yv & 0x8000u) == 0u && (yv & 0x7fffu) != 0u) ? 0xffffu : 0u;
      const uint32_t hi = ((yv & 0x80000000u) == 0u && (yv & 0x7fff0000u) != 0u) ? 0xffff0000u : 0u;
      return gv & (lo | hi);
    };
    o.x = keep(g.x, v.x); o.y = keep(g.y, v.y); o.z = keep(g.z, v.z); o.w = keep(g.w, v.w);
    *reinterpret_cast<uint4*>(out + (size_t)i * 8) = o;
  }
}
}  // namespace tgsr

extern "C" int tgsr_relu_bwd_nhwc16(const void* gy, long long gy_pixel_stride, const void* y, void* out, long long npix, int C,
                                    int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(gy); TGSR_REQUIRE_PTR(y); TGSR_REQUIRE_PTR(out);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(npix > 0 && C > 0 && C % 8 == 0 && gy_pixel_stride >= C && gy_pixel_stride % 8 == 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((reinterpret_cast<uintptr_t>(gy) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(out) & 15) == 0, TGSR_ERR_BAD_SHAPE);
  const long long total = npix * (C / 8);
  TGSR_REQUIRE(total < (1ll << 31) - (1 << 24), TGSR_ERR_BAD_SHAPE);
  const h16* gg = reinterpret_cast<const h16*>(gy); const h16* yy = reinterpret_cast<const h16*>(y); h16* oo = reinterpret_cast<h16*>(out);
  (void)fmt;   // the sign test is the same for bf16 and fp16
  relu_bwd_nhwc16_kernel<TGSR_FMT_BF16><<<pool_blocks((size_t)total), 256, 0, (cudaStream_t)stream>>>(gg, (size_t)gy_pixel_stride, yy, oo, C / 8,
                                                                                                 (uint32_t)total, FastDiv((uint32_t)(C / 8)));
  return launch_status();
}
