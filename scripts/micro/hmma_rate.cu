// Microbenchmark: issue rate of the warp-level tensor-core path (mma.sync -> HMMA) on B200.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND>
__global__ void k(float* out, int iters) {
  float d[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 5, b1 = 11;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND> void run(const char* name, double mac_per_mma, int warps) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  k<KIND><<<148, warps * 32>>>(out, 100);
  cudaEventRecord(s);
  k<KIND><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double mmas = 148.0 * warps * iters * 8;
  printf("%-22s warps/SM %2d: %.1f TFLOP/s  (%.2f ns per MMA per SM)\n", name, warps, 2 * mmas * mac_per_mma / (ms * 1e-3) / 1e12,
         ms * 1e6 / (warps * (double)iters * 8));
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("m16n8k16 bf16", 2048, w);
    run<1>("m16n8k8 bf16", 1024, w);
    run<2>("m16n8k8 tf32", 1024, w);
    run<3>("m16n8k4 tf32", 512, w);
  }
  return 0;
}
