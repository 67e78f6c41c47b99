// Edge conversions of the eval flow that are not part of a network layer.
//   * uint8 quantisation of the SR image, trainer_objective.py:153-155:
//     im = round(max(0, min(255, (x + 1.0) * 127.5))).astype(uint8)  (numpy: fp32 arithmetic, round-half-even).
//     Done on the device so that the device->host copy of a step's result moves 1 byte per sample.
#include "common.cuh"

namespace tgsr {

__device__ __forceinline__ unsigned int quant_u8(float x) {
  const float v = __fmul_rn(__fadd_rn(x, 1.0f), 127.5f);      // no FMA contraction: bit-identical to numpy
  return (unsigned int)rintf(fmaxf(0.0f, fminf(255.0f, v)));  // rintf: round-half-even like np.round
}

__global__ void __launch_bounds__(256)
quantize_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, size_t n4, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const unsigned int p = quant_u8(v.x) | (quant_u8(v.y) << 8) | (quant_u8(v.z) << 16) | (quant_u8(v.w) << 24);
    reinterpret_cast<unsigned int*>(y)[i] = p;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {             // tail (n not a multiple of 4)
    const size_t i = n4 * 4 + threadIdx.x;
    y[i] = (uint8_t)quant_u8(x[i]);
  }
}

}  // namespace tgsr

extern "C" int tgsr_quantize_u8(const float* x, uint8_t* y, size_t n, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(n > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 3) == 0, TGSR_ERR_BAD_SHAPE);
  const size_t n4 = n / 4;
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > (size_t)sms * 8) blocks = (size_t)sms * 8;
  if (blocks < 1) blocks = 1;
  quantize_u8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, n4, n);
  return launch_status();
}
