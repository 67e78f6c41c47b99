// Data gradient of the 3-output-channel image heads (GET_IMAGE_G[_noAct].img util.py:894-919, conv_output model.py:223-226)
// written straight into the blocked 16-bit layout of the feature gradient:
//   dh[b][ci][y][x] = sum over (co < 3, ky, kx) of dz[b][co][y + r - ky][x + r - kx] * w[co][ci][ky][kx]     (zero padding)
// (autograd / cuDNN in the reference).  A GEMM with M = pixels, N = 32 input channels and K = 3*k*k (27 or 75, padded to
// 32 / 80) on the warp-level tensor-core path: the A fragment is an im2col gather of the fp32 dz tile in shared memory
// (rounded to 16 bits on the way), the B fragments (the flipped filter) live in registers for the whole kernel.  One warp =
// one row of 32 pixels; the D fragment of channel tile j is exactly plane j of the blocked layout (4 bytes per lane, 16 bytes
// per quad, 128 contiguous bytes per 8 pixels).
#include "common.cuh"
#include "fmt16.cuh"

namespace tgsr {

typedef uint16_t h16;
constexpr int HD_ROWS = 8, HD_COLS = 32, HD_CIN = 32;

template <int FMT> struct HdMma;
template <> struct HdMma<TGSR_FMT_BF16> {
  __device__ __forceinline__ static void k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct HdMma<TGSR_FMT_F16> {
  __device__ __forceinline__ static void k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};

template <int K, int FMT>
__global__ void __launch_bounds__(HD_ROWS * 32)
head_dgrad16_kernel(const float* __restrict__ dz, const float* __restrict__ w, h16* __restrict__ dh, int hP, int hp0, int B, int H,
                    int W, int tiles_x, int tiles_y, int ntiles) {
  constexpr int R = K / 2, KK = K * K, KTOT = 3 * KK, KSTEPS = (KTOT + 15) / 16;
  constexpr int PH = HD_ROWS + K - 1, PW = HD_COLS + K - 1, PLANE = PH * PW;
  __shared__ float zs[3 * PLANE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;

  // k = co*KK + ky*K + kx  ->  offset of dz[co][. + (K-1-ky)][. + (K-1-kx)] inside the staged tile; the lane's four k per K-step
  int koff[KSTEPS][4];
  uint32_t bfr[KSTEPS][4][2];                                   // [K-step][channel tile][b0, b1]
#pragma unroll
  for (int s = 0; s < KSTEPS; ++s) {
    float wv[4][4];                                             // [k slot][channel tile]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 16 * s + 2 * tig + (j & 1) + (j >> 1) * 8;
      const bool on = k < KTOT;
      const int co = on ? k / KK : 0, t = on ? k - co * KK : 0, ky = t / K, kx = t - ky * K;
      koff[s][j] = on ? co * PLANE + (K - 1 - ky) * PW + (K - 1 - kx) : 0;
#pragma unroll
      for (int n = 0; n < 4; ++n) wv[j][n] = on ? __ldg(w + ((size_t)co * HD_CIN + 8 * n + g) * KK + t) : 0.f;
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      bfr[s][n][0] = Fmt16<FMT>::pack2(wv[0][n], wv[1][n]);
      bfr[s][n][1] = Fmt16<FMT>::pack2(wv[2][n], wv[3][n]);
    }
  }
  const size_t plane = (size_t)H * W;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, q = tile / tiles_x, ty = q % tiles_y, b = q / tiles_y;
    const int y0 = ty * HD_ROWS, x0 = tx * HD_COLS;
    __syncthreads();
    for (int e = tid; e < 3 * PLANE; e += HD_ROWS * 32) {
      const int pc = e % PW, t = e / PW, pr = t % PH, co = t / PH;
      const int yy = y0 - R + pr, xx = x0 - R + pc;
      zs[e] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(dz + (((size_t)b * 3 + co) * H + yy) * W + xx) : 0.f;
    }
    __syncthreads();
    const int y = y0 + warp;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float acc[4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
      const float* zr = zs + warp * PW + half * 16 + g;         // pixel rows g and g + 8 of this 16-pixel group
#pragma unroll
      for (int s = 0; s < KSTEPS; ++s) {
        uint32_t a[4];
        a[0] = Fmt16<FMT>::pack2(zr[koff[s][0]], zr[koff[s][1]]);
        a[1] = Fmt16<FMT>::pack2(zr[koff[s][0] + 8], zr[koff[s][1] + 8]);
        a[2] = Fmt16<FMT>::pack2(zr[koff[s][2]], zr[koff[s][3]]);
        a[3] = Fmt16<FMT>::pack2(zr[koff[s][2] + 8], zr[koff[s][3] + 8]);
#pragma unroll
        for (int n = 0; n < 4; ++n) HdMma<FMT>::k16(acc[n], a, bfr[s][n][0], bfr[s][n][1]);
      }
      if (y < H) {
        const int xa = x0 + half * 16 + g, xb = xa + 8;
        h16* o = dh + (((size_t)b * hP + hp0) * plane + (size_t)y * W) * 8 + 2 * tig;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          if (xa < W) *reinterpret_cast<uint32_t*>(o + ((size_t)n * plane + xa) * 8) = Fmt16<FMT>::pack2(acc[n][0], acc[n][1]);
          if (xb < W) *reinterpret_cast<uint32_t*>(o + ((size_t)n * plane + xb) * 8) = Fmt16<FMT>::pack2(acc[n][2], acc[n][3]);
        }
      }
    }
  }
}

}  // namespace tgsr

extern "C" int tgsr_head_dgrad16(const float* dz, const float* w, void* dh, int hP, int hp0, int B, int Cin, int H, int W, int ksize,
                                 int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(dz); TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(dh);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(Cin == HD_CIN && (ksize == 3 || ksize == 5), TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0 && hp0 >= 0 && hP >= hp0 + Cin / 8, TGSR_ERR_BAD_SHAPE);
  const int tiles_x = ceil_div(W, HD_COLS), tiles_y = ceil_div(H, HD_ROWS);
  const long long nt = (long long)B * tiles_x * tiles_y;
  TGSR_REQUIRE(nt < (1ll << 31), TGSR_ERR_BAD_SHAPE);
  const int ntiles = (int)nt;
  const int cap = (g_sm_count > 0 ? g_sm_count : 148) * 8;
  const int grid = ntiles < cap ? ntiles : cap;
  h16* oo = reinterpret_cast<h16*>(dh);
  cudaStream_t st = (cudaStream_t)stream;
#define TGSR_HD_LAUNCH(KK_, F) head_dgrad16_kernel<KK_, F><<<grid, HD_ROWS * 32, 0, st>>>(dz, w, oo, hP, hp0, B, H, W, tiles_x, tiles_y, ntiles)
  if (ksize == 3) { if (fmt == TGSR_FMT_F16) TGSR_HD_LAUNCH(3, TGSR_FMT_F16); else TGSR_HD_LAUNCH(3, TGSR_FMT_BF16); }
  else { if (fmt == TGSR_FMT_F16) TGSR_HD_LAUNCH(5, TGSR_FMT_F16); else TGSR_HD_LAUNCH(5, TGSR_FMT_BF16); }
#undef TGSR_HD_LAUNCH
  return launch_status();
}
