// Weight gradient of the 3-output-channel image heads (GET_IMAGE_G[_noAct].img util.py:894-919, conv_output model.py:223-226)
// straight from the blocked 16-bit features:
//   dW[co][ci][ky][kx] = sum over (b, y, x) of dz[b][co][y][x] * h[b][ci][y + ky - r][x + kx - r]      (zero padding, r = k/2)
// (autograd / cuDNN in the reference).  Cout = 3 makes this a poor fit for the tcgen05 weight-gradient kernel (M = 128 rows,
// 3 used), so it runs on the warp-level tensor-core path: per filter tap  D[m = ci (16)][n = 8] += A[m][k = 16 pixels] * B[k][n].
//   A = h^T: 8 consecutive pixels of one 8-channel plane are one 8x8 16-bit matrix with 16-byte rows in the blocked layout,
//       so ldmatrix.x4.trans on (plane 2c, plane 2c+1) x (pixels 0-7, 8-15) IS the A fragment; the tap is an address shift.
//   B = dz: columns 0..2 hold the 16-bit rounding of dz[co], columns 3..5 the rounding residual (N = 8 has room for both),
//       so dz enters with ~16 mantissa bits at no extra MMA; the two partial sums meet in the final atomics.
// A CTA owns 16 x 32 output pixels at a time (double-buffered through cp.async); warp (kx, c) owns filter column kx and input channels 16c..16c+15 and walks the
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

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;                                  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

// smem, two stages: patch [4 planes][PH][PW] x 16 B | dz tile [3][HW_TH][HW_TW] fp32.  Every byte arrives through cp.async
// (zero-filled outside the image), so the loads of tile n+1 are in flight while tile n is multiplied.
template <int K, int FMT>
__global__ void __launch_bounds__(K * 2 * 32)
head_wgrad16_kernel(const h16* __restrict__ x, int xP, int xp0, const float* __restrict__ dz, float* __restrict__ dw,
                    int B, int H, int W, int tiles_x, int tiles_y, int ntiles) {
  constexpr int R = K / 2, PH = HW_TH + K - 1, PW = HW_TW + K - 1, NPL = HW_CIN / 8, THREADS = K * 2 * 32;
  constexpr int PATCH_B = NPL * PH * PW * 16, DZ_B = 3 * HW_TH * HW_TW * 4, STAGE_B = PATCH_B + DZ_B;
  constexpr int BFRAG_OFF = 2 * STAGE_B;                         // B fragments of the current tile: [HW_TH][2 chunks][32 lanes] uint2
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint2* bfrag = reinterpret_cast<uint2*>(smem_raw + BFRAG_OFF);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int kx = warp >> 1, cg = warp & 1;                      // filter column, input-channel group of 16
  const size_t plane = (size_t)H * W;
  const uint32_t smem0 = ptx::smem_u32(smem_raw);

  float acc[K][4];
#pragma unroll
  for (int ky = 0; ky < K; ++ky) { acc[ky][0] = acc[ky][1] = acc[ky][2] = acc[ky][3] = 0.f; }

  // ldmatrix row address of this lane inside a patch: matrix j = lane/8: plane 2cg + (j & 1), pixels 8(j >> 1) + lane % 8
  const uint32_t a_off = (uint32_t)((((2 * cg + ((lane >> 3) & 1)) * PH) * PW + kx + ((lane >> 4) << 3) + (lane & 7)) * 16);

  auto prefetch = [&](int tile, int stage) {
    const int tx = tile % tiles_x, q = tile / tiles_x, ty = q % tiles_y, b = q / tiles_y;
    const int y0 = ty * HW_TH, x0 = tx * HW_TW;
    const uint32_t pbase = smem0 + stage * STAGE_B;
    const h16* xb = x + ((size_t)b * xP + xp0) * plane * 8;
    for (int e = tid; e < NPL * PH * PW; e += THREADS) {
      const int pc = e % PW, t = e / PW, pr = t % PH, pl = t / PH;
      const int yy = y0 - R + pr, xx = x0 - R + pc;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      cp_async16_zfill(pbase + e * 16, xb + ((size_t)pl * plane + (size_t)(ok ? yy : 0) * W + (ok ? xx : 0)) * 8, ok);
    }
    for (int e = tid; e < 3 * HW_TH * (HW_TW / 4); e += THREADS) {             // 16-byte chunks = 4 pixels (W % 4 == 0)
      const int c4 = e % (HW_TW / 4), t = e / (HW_TW / 4), r = t % HW_TH, co = t / HW_TH;
      const int yy = y0 + r, xx = x0 + 4 * c4;
      const bool ok = yy < H && xx < W;
      cp_async16_zfill(pbase + PATCH_B + e * 16, dz + (((size_t)b * 3 + co) * H + (ok ? yy : 0)) * W + (ok ? xx : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int stage = 0;
  if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < ntiles) {
      prefetch(next, stage ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                                            // this tile's bytes are visible to every warp
    const uint32_t a_base = smem0 + stage * STAGE_B + a_off;
    // ---- B fragments of dz, built ONCE per tile for all warps: lane (g, tig) of (row, chunk) holds n = g: channel g % 3
    //      (the 16-bit value for g < 3, the rounding residual for g < 6, zero above), k = 2tig, 2tig+1 (b0) and 2tig+8, 2tig+9 (b1)
    {
      const float* zt = reinterpret_cast<const float*>(smem_raw + stage * STAGE_B + PATCH_B);
      for (int e = tid; e < HW_TH * 2 * 32; e += THREADS) {
        const int l = e & 31, ch = (e >> 5) & 1, r = e >> 6, gg = l >> 2, tg = l & 3;
        uint2 v = make_uint2(0u, 0u);
        if (gg < 6) {
          const float* zp = zt + (gg % 3) * (HW_TH * HW_TW) + r * HW_TW + ch * 16 + 2 * tg;
          float2 f0 = *reinterpret_cast<const float2*>(zp), f1 = *reinterpret_cast<const float2*>(zp + 8);
          if (gg >= 3) {
            f0.x -= Fmt16<FMT>::cvtf(f0.x); f0.y -= Fmt16<FMT>::cvtf(f0.y);
            f1.x -= Fmt16<FMT>::cvtf(f1.x); f1.y -= Fmt16<FMT>::cvtf(f1.y);
          }
          v = make_uint2(Fmt16<FMT>::pack2(f0.x, f0.y), Fmt16<FMT>::pack2(f1.x, f1.y));
        }
        bfrag[e] = v;
      }
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
    __syncthreads();                                            // everyone is done with this stage before it is refilled
    stage ^= 1;
  }
  // ---- D fragment: rows ci = 16cg + g (+8), columns n = 2tig, 2tig+1; n < 3: co = n, 3 <= n < 6: residual part of co = n - 3.
  // Value and residual columns of one co sit in different lanes of a quad: they are added through shuffles, so each
  // (co, ci, tap) costs ONE atomic per CTA.
  const int qbase = lane & ~3;
#pragma unroll
  for (int ky = 0; ky < K; ++ky)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float v0 = acc[ky][2 * r], v1 = acc[ky][2 * r + 1];
      const float t1v1 = __shfl_sync(0xffffffffu, v1, qbase + 1);
      const float t2v0 = __shfl_sync(0xffffffffu, v0, qbase + 2);
      const float t2v1 = __shfl_sync(0xffffffffu, v1, qbase + 2);
      const int ci = 16 * cg + g + r * 8;
      float* o = dw + ((size_t)ci * K + ky) * K + kx;              // + co * HW_CIN * K * K
      if (tig == 0) {
        atomicAdd(o, v0 + t1v1);                                   // co 0: n = 0 and n = 3
        atomicAdd(o + (size_t)HW_CIN * K * K, v1 + t2v0);          // co 1: n = 1 and n = 4
      } else if (tig == 1) {
        atomicAdd(o + (size_t)2 * HW_CIN * K * K, v0 + t2v1);      // co 2: n = 2 and n = 5
      }
    }
}

}  // namespace tgsr

extern "C" int tgsr_head_wgrad16(const void* x, int xP, int xp0, const float* dz, float* dw, int B, int Cin, int H, int W,
                                 int ksize, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(dz); TGSR_REQUIRE_PTR(dw);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(Cin == HW_CIN && (ksize == 3 || ksize == 5), TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0 && W % 4 == 0 && xP >= xp0 + Cin / 8 && xp0 >= 0, TGSR_ERR_BAD_SHAPE);
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
    constexpr size_t sm = 2 * ((size_t)(HW_CIN / 8) * PHh * PWw * 16 + 3 * HW_TH * HW_TW * 4) + HW_TH * 2 * 32 * 8;   /* two stages + B fragments */                                  \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      TGSR_CUDA(cudaFuncSetAttribute(head_wgrad16_kernel<KK, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    const int per_sm = (int)((220 * 1024) / (sm + 1024)) > 4 ? 4 : (int)((220 * 1024) / (sm + 1024));                  \
    const int cap = (g_sm_count > 0 ? g_sm_count : 148) * per_sm;                                                      \
    int grid = (ntiles + 7) / 8;                 /* >= 8 tiles per CTA: every CTA ends with 3 * Cin * k * k atomics */   \
    if (grid > cap) grid = cap;                                                                                        \
    head_wgrad16_kernel<KK, F><<<grid, KK * 2 * 32, sm, st>>>(xx, xP, xp0, dz, dw, B, H, W, tiles_x, tiles_y, ntiles);   \
  } while (0)
  if (ksize == 3) { if (fmt == TGSR_FMT_F16) TGSR_HW_LAUNCH(3, TGSR_FMT_F16); else TGSR_HW_LAUNCH(3, TGSR_FMT_BF16); }
  else { if (fmt == TGSR_FMT_F16) TGSR_HW_LAUNCH(5, TGSR_FMT_F16); else TGSR_HW_LAUNCH(5, TGSR_FMT_BF16); }
#undef TGSR_HW_LAUNCH
  return launch_status();
}
