// Shared helpers for libtgsr_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tgsr.h"

namespace tgsr {

extern thread_local int g_last_cuda_error;
extern int g_sm_count;      // 0 until tgsr_init
extern int g_device;
extern unsigned long long g_launches;   // kernels enqueued by this library (bench.py's gpu_launches)

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return TGSR_ERR_CUDA;
}
// Check the launch that was just enqueued (no sync).
inline int launch_status(int n_kernels = 1) {
  g_launches += (unsigned long long)n_kernels;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

#define TGSR_REQUIRE_PTR(p) do { if ((p) == nullptr) return TGSR_ERR_NULL; } while (0)
#define TGSR_REQUIRE(cond, code) do { if (!(cond)) return (code); } while (0)
#define TGSR_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return tgsr::cuda_fail(e__); } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// n / d and n % d for n < 2^31 with one 64-bit multiply: m = ceil(2^p / d), p = 31 + ceil(log2 d); the product error
// n * (m*d - 2^p) stays below 2^p for every n < 2^31, so the quotient is exact (elementwise kernels decompose their
// linear index with it instead of 64-bit divisions, which cost more than the arithmetic they index).
struct FastDiv {
  uint32_t d, m, p;
  FastDiv() : d(1), m(1u << 31), p(31) {}
  explicit FastDiv(uint32_t dd) : d(dd) {
    uint32_t s = 0;
    while ((1ull << s) < dd) ++s;
    p = 31 + s;
    m = (uint32_t)(((1ull << p) + dd - 1) / dd);
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return (uint32_t)(((uint64_t)n * m) >> p); }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

}  // namespace tgsr
