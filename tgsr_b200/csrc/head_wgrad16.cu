// Weight gradient of the 3-output-channel image heads (GET_IMAGE_G[_noAct].img util.py:894-919, conv_output model.py:223-226)
// straight from the blocked 16-bit features:
//   dW[co][ci][ky][kx] = sum over (b, y, x) of dz[b][co][y][x] * h[b][ci][y + ky - r][x + kx - r]      (zero padding, r = k/2)
// (autograd / cuDNN in the reference).  Cout = 3 makes this a poor fit for the tcgen05 weight-gradient kernel (M = 128 rows,
// 3 used), so it runs on the warp-level tensor-core path: per filter tap  D[m = ci (16)][n = 8] += A[m][k = 16 pixels] * B[k][n].
//   A = h^T: 8 consecutive pixels of one 8-channel plane are one 8x8 16-bit matrix with 16-byte rows in the blocked layout,
//       so ldmatrix.x4.trans on (plane 2c, plane 2c+1) x (pixels 0-7, 8-15) IS the A fragment; the tap is an address shift.
//   B = dz: columns 0..2 hold the 16-bit rounding of dz[co], columns 3..5 the rounding residual (N = 8 has room for both),
//       so dz enters with ~16 mantissa bits at no extra MMA; the two partial sums meet in the final atomics.
// A CTA owns 16 x 32 output pixels at a time; warp (kx, c) owns filter column kx and input channels 16c..16c+15 and walks the
// patch rows: one ldmatrix of patch row R serves the K taps ky (output row R - ky), whose B fragments slide through registers.
// Accumulators stay in registers over all tiles of a CTA; the epilogue adds them to dW with fp32 atomics.
#include "common.cuh"
#include "fmt16.cuh"
#include "ptx_sm100.cuh"

namespace tgsr {

typedef uint16_t h16;
constexpr int HW_TH = 16, HW_TW = 32, HW_CIN = 32;

template <int FMT> struct HwMma;
template <> struct HwMma<TGSR_FMT_BF16> {
  __device__ __forceinline__ static void k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct HwMma<TGSR_FMT_F16> {
  __device__ __forceinline__ static void k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// smem: patch [4 planes][PH][PW] x 16 B | bfrag [HW_TH][2 chunks][32 lanes] uint2
template <int K, int FMT>
__global__ void __launch_bounds__(K * 2 * 32)
head_wgrad16_kernel(const h16* __restrict__ x, int xP, int xp0, const float* __restrict__ dz, float* __restrict__ dw,
                    int B, int H, int W, int tiles_x, int tiles_y, int ntiles) {
  constexpr int R = K / 2, PH = HW_TH + K - 1, PW = HW_TW + K - 1, NPL = HW_CIN / 8, THREADS = K * 2 * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* patch = reinterpret_cast<uint4*>(smem_raw);
  uint2* bfrag = reinterpret_cast<uint2*>(smem_raw + NPL * PH * PW * 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int kx = warp >> 1, cg = warp & 1;                      // filter column, input-channel group of 16
  const size_t plane = (size_t)H * W;

  float acc[K][4];
#pragma unroll
  for (int ky = 0; ky < K; ++ky) { acc[ky][0] = acc[ky][1] = acc[ky][2] = acc[ky][3] = 0.f; }

  // ldmatrix row address of this lane inside the patch: matrix j = lane/8: plane 2cg + (j & 1), pixels 8(j >> 1) + lane % 8
  const uint32_t a_base = ptx::smem_u32(patch) +
                          (uint32_t)((((2 * cg + ((lane >> 3) & 1)) * PH) * PW + kx + ((lane >> 4) << 3) + (lane & 7)) * 16);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, q = tile / tiles_x, ty = q % tiles_y, b = q / tiles_y;
    const int y0 = ty * HW_TH, x0 = tx * HW_TW;
    __syncthreads();                                            // the previous tile's operands are no longer in use
    // ---- halo patch of h (zero outside the image)
    const h16* xb = x + ((size_t)b * xP + xp0) * plane * 8;
    for (int e = tid; e < NPL * PH * PW; e += THREADS) {
      const int pc = e % PW, t = e / PW, pr = t % PH, pl = t / PH;
      const int yy = y0 - R + pr, xx = x0 - R + pc;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = __ldg(reinterpret_cast<const uint4*>(xb + ((size_t)pl * plane + (size_t)yy * W + xx) * 8));
      patch[e] = v;
    }
    // ---- B fragments of dz: lane (g, tig) of (row, chunk) holds n = g: co = g % 3 (hi for g < 3, residual for g < 6),
    //      k = 2tig, 2tig+1 (b0) and 2tig+8, 2tig+9 (b1) of the 16-pixel chunk
    for (int e = tid; e < HW_TH * 2 * 32; e += THREADS) {
      const int l = e & 31, ch = (e >> 5) & 1, r = e >> 6, gg = l >> 2, tg = l & 3;
      uint2 v = make_uint2(0u, 0u);
      const int yy = y0 + r;
      if (gg < 6 && yy < H) {
        const float* zp = dz + (((size_t)b * 3 + (gg % 3)) * H + yy) * W;
        float f[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int xx = x0 + ch * 16 + 2 * tg + (i & 1) + (i >> 1) * 8;
          f[i] = xx < W ? __ldg(zp + xx) : 0.f;
          if (gg >= 3) f[i] -= Fmt16<FMT>::cvtf(f[i]);
        }
        v.x = Fmt16<FMT>::pack2(f[0], f[1]); v.y = Fmt16<FMT>::pack2(f[2], f[3]);
      }
      bfrag[e] = v;
    }
    __syncthreads();
    // ---- walk the patch rows; the B fragments of the last K output rows slide through registers
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint2 bw[K];
#pragma unroll
      for (int i = 0; i < K; ++i) bw[i] = make_uint2(0u, 0u);
#pragma unroll
      for (int pr = 0; pr < PH; ++pr) {
        // window: bw[ky] = fragment of output row pr - ky
#pragma unroll
        for (int i = K - 1; i > 0; --i) bw[i] = bw[i - 1];
        bw[0] = pr < HW_TH ? bfrag[(pr * 2 + ch) * 32 + lane] : make_uint2(0u, 0u);
        uint32_t a[4];
        ldmatrix_x4_trans(a, a_base + (uint32_t)((pr * PW + ch * 16) * 16));
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
          if (pr - ky >= 0 && pr - ky < HW_TH) HwMma<FMT>::k16(acc[ky], a, bw[ky].x, bw[ky].y);
      }
    }
  }
  // ---- D fragment: rows ci = 16cg + g (+8), columns n = 2tig, 2tig+1; n < 3: co = n, 3 <= n < 6: residual of co = n - 3
#pragma unroll
  for (int ky = 0; ky < K; ++ky)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = 2 * tig + (i & 1), ci = 16 * cg + g + (i >> 1) * 8;
      if (n < 6) atomicAdd(dw + (((size_t)(n % 3) * HW_CIN + ci) * K + ky) * K + kx, acc[ky][i]);
    }
}

}  // namespace tgsr

extern "C" int tgsr_head_wgrad16(const void* x, int xP, int xp0, const float* dz, float* dw, int B, int Cin, int H, int W,
                                 int ksize, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(dz); TGSR_REQUIRE_PTR(dw);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(Cin == HW_CIN && (ksize == 3 || ksize == 5), TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0 && xP >= xp0 + Cin / 8 && xp0 >= 0, TGSR_ERR_BAD_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 3 * Cin * ksize * ksize, st));
  const int tiles_x = ceil_div(W, HW_TW), tiles_y = ceil_div(H, HW_TH);
  const long long nt = (long long)B * tiles_x * tiles_y;
  TGSR_REQUIRE(nt < (1ll << 31), TGSR_ERR_BAD_SHAPE);
  const int ntiles = (int)nt;
  const h16* xx = reinterpret_cast<const h16*>(x);
#define TGSR_HW_LAUNCH(KK, F)                                                                                          \
  do {                                                                                                                 \
    constexpr int PHh = HW_TH + KK - 1, PWw = HW_TW + KK - 1;                                                          \
    constexpr size_t sm = (size_t)(HW_CIN / 8) * PHh * PWw * 16 + HW_TH * 2 * 32 * 8;                                  \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      TGSR_CUDA(cudaFuncSetAttribute(head_wgrad16_kernel<KK, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    const int per_sm = (int)((220 * 1024) / (sm + 1024)) > 4 ? 4 : (int)((220 * 1024) / (sm + 1024));                  \
    const int grid = ntiles < (g_sm_count > 0 ? g_sm_count : 148) * per_sm ? ntiles : (g_sm_count > 0 ? g_sm_count : 148) * per_sm;                                      \
    head_wgrad16_kernel<KK, F><<<grid, KK * 2 * 32, sm, st>>>(xx, xP, xp0, dz, dw, B, H, W, tiles_x, tiles_y, ntiles);   \
  } while (0)
  if (ksize == 3) { if (fmt == TGSR_FMT_F16) TGSR_HW_LAUNCH(3, TGSR_FMT_F16); else TGSR_HW_LAUNCH(3, TGSR_FMT_BF16); }
  else { if (fmt == TGSR_FMT_F16) TGSR_HW_LAUNCH(5, TGSR_FMT_F16); else TGSR_HW_LAUNCH(5, TGSR_FMT_BF16); }
#undef TGSR_HW_LAUNCH
  return launch_status();
}
