// 16-bit operand formats of the tensor-core path.  Both are tcgen05 kind::f16 operands with fp32
// accumulation (same instruction, same rate, same bytes); they differ in mantissa/exponent split.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/tgsr.h"

namespace tgsr {

template <int FMT> struct Fmt16;

template <> struct Fmt16<TGSR_FMT_BF16> {
  static constexpr uint32_t kUmmaFormat = 1;   // UMMA F16F32Format::BF16
  __device__ __forceinline__ static uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static float2 unpack2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
  __device__ __forceinline__ static uint16_t cvt(float a) {
    __nv_bfloat16 h = __float2bfloat16_rn(a);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __device__ __forceinline__ static float cvtf(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
};

template <> struct Fmt16<TGSR_FMT_F16> {
  static constexpr uint32_t kUmmaFormat = 0;   // UMMA F16F32Format::F16
  __device__ __forceinline__ static uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static float2 unpack2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
  __device__ __forceinline__ static uint16_t cvt(float a) {
    __half h = __float2half_rn(a);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __device__ __forceinline__ static float cvtf(float a) { return __half2float(__float2half_rn(a)); }
};

template <int FMT>
__device__ __forceinline__ uint4 pack8_16(const float (&v)[8]) {
  uint4 r;
  r.x = Fmt16<FMT>::pack2(v[0], v[1]); r.y = Fmt16<FMT>::pack2(v[2], v[3]);
  r.z = Fmt16<FMT>::pack2(v[4], v[5]); r.w = Fmt16<FMT>::pack2(v[6], v[7]);
  return r;
}
template <int FMT>
__device__ __forceinline__ void unpack8_16(const uint4& u, float* v) {
  float2 f;
  f = Fmt16<FMT>::unpack2(u.x); v[0] = f.x; v[1] = f.y;
  f = Fmt16<FMT>::unpack2(u.y); v[2] = f.x; v[3] = f.y;
  f = Fmt16<FMT>::unpack2(u.z); v[4] = f.x; v[5] = f.y;
  f = Fmt16<FMT>::unpack2(u.w); v[6] = f.x; v[7] = f.y;
}

}  // namespace tgsr
