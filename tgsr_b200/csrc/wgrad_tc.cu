// tcgen05 weight gradient of conv3x3 on blocked 16-bit activations:
//   dW[co][ci][ky][kx] = sum over (b, y, x) of dz[b][co][y][x] * x[b][ci][y+ky-1][x+kx-1]        (zero padding)
// (the reference gets it from autograd / cuDNN for every conv3x3 of util.py:62-65 in the cfg-5 training step).
//
// GEMM view per filter tap: D_tap[M = co][N = ci] += A[M][K] * B[K][N] with K = pixels.  Both operands come straight
// from the blocked layout [plane][pixel][8 ch] as MN-MAJOR UMMA operands: 8 consecutive pixels of one plane are one
// 128-byte core matrix (8 K-rows x 8 channels), so
//   A (dz tile, 16 x 8 pixels):  SBO = plane stride (next 8 channels), LBO = 128 B (next 8 pixels = next tile row);
//   B (x halo patch, 18 x 10):   SBO = plane stride, LBO = row pitch (160 B); the TAP is only a start-address shift,
// exactly as in the forward kernel (conv_tc.cu).  One MMA = 16 pixels (two tile rows); a tile = 8 K-steps x taps MMAs.
// M is always 128 (rows >= Cout read whatever follows the dz tile in shared memory; their accumulator rows are never
// read).  The nine accumulators [128 x Cin] fp32 need 9*Cin TMEM columns: Cin = 32 fits (288); Cin = 64 (576 > 512) is
// split into two tap groups handled by alternating CTAs (each tile is then fetched by two CTAs; L2 absorbs it).
// Accumulation stays in TMEM over all tiles of a CTA; the epilogue adds the CTA's partial dW to global memory with fp32
// atomics (dW is zeroed by the launcher).
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..5 epilogue.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "fmt16.cuh"
#include <cstdlib>

namespace tgsr {

using namespace ptx;
typedef uint16_t h16;

constexpr int WG_TH = 16, WG_TW = 8;
constexpr int WG_THREADS = 192;

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnW g_encode_w = nullptr;

struct WgParams {
  float* ws;                      // [9 taps][Cin][Cout] fp32 partial sums, accumulated with atomics (coalesced over Cout)
  int Cout, B, H, W;
  int tiles_x, tiles_y, ntiles;
  int zP, zp0, xP, xp0;           // plane counts / offsets of the dz and x views
};

template <int CIN>
struct WgCfg {
  static constexpr int GROUPS = CIN == 64 ? 2 : 1;             // tap groups (TMEM columns)
  static constexpr int TAPS0 = CIN == 64 ? 5 : 9;              // taps of group 0 (group 1: the rest)
  static constexpr int TMEM_COLS = 512;
  static constexpr int PW = WG_TW + 2, PH = WG_TH + 2;
  static constexpr int XPLANE = PH * PW * 16;                  // one 8-channel plane of the x halo patch
  static constexpr int ZPLANE = WG_TH * WG_TW * 16;            // one 8-channel plane of the dz tile (2 KB)
  static constexpr int ZBYTES = 16 * ZPLANE;                   // room for 128 dz channels
  static constexpr int XBYTES = (CIN / 8) * XPLANE;
  static constexpr int STAGE = ((ZBYTES + XBYTES + 127) / 128) * 128;
  static constexpr int NSTAGE = 3;
  static constexpr size_t SMEM = 1024 + NSTAGE * STAGE + 32768 /* slack: M = 128 reads past a short dz tile */ + 256;
};

// kind::f16 instruction descriptor with BOTH operands MN-major (bits 15 / 16), D = f32.
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// MMAs of one tile for the taps [TAP0, TAP0 + NTAP): accumulator t (TMEM columns t*CIN ..) += dz^T (16 pixels) . x shifted by
// the tap.  Descriptor fields (no swizzle): [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, version 1 at bit 46.
// MN-major: SBO steps along M / N (the next 8 channels = the next plane), LBO along K (the next 8 pixels) -- verified on
// B200 against fp32 autograd (tests/test_gpu_train16.py); the swapped reading gives garbage.
template <int CIN, int FMT, int TAP0, int NTAP>
__device__ __forceinline__ void issue_tile(uint32_t tmem_base, uint32_t zaddr, bool accumulate) {
  using Cfg = WgCfg<CIN>;
  constexpr uint32_t idesc = make_idesc_mn(128, CIN, Fmt16<FMT>::kUmmaFormat);
  constexpr uint32_t a_lbo = 128, a_sbo = Cfg::ZPLANE;
  constexpr uint32_t b_lbo = Cfg::PW * 16, b_sbo = Cfg::XPLANE;
  constexpr uint32_t a_hi = (a_sbo >> 4) | (1u << 14), b_hi = (b_sbo >> 4) | (1u << 14);
  const uint32_t a0 = (zaddr >> 4) | ((a_lbo >> 4) << 16);
  const uint32_t b0 = ((zaddr + Cfg::ZBYTES) >> 4) | ((b_lbo >> 4) << 16);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {                            // 16 pixels = tile rows 2ks, 2ks+1
#pragma unroll
    for (int t = 0; t < NTAP; ++t) {
      const int tap = TAP0 + t, dy = tap / 3, dx = tap % 3;   // compile-time after unrolling
      const uint32_t a_lo = a0 + (uint32_t)((ks * 256) >> 4);
      const uint32_t b_lo = b0 + (uint32_t)((((2 * ks + dy) * Cfg::PW + dx) * 16) >> 4);
      umma_f16(tmem_base + t * CIN, make_desc64(a_lo, a_hi), make_desc64(b_lo, b_hi), idesc, (accumulate || ks != 0) ? 1u : 0u);
    }
  }
}

template <int CIN, int FMT>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap zmap, const __grid_constant__ CUtensorMap xmap, const WgParams p) {
  using Cfg = WgCfg<CIN>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_s = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Cfg::NSTAGE * Cfg::STAGE + 32768);
  uint64_t* full = bars;               // [NSTAGE]
  uint64_t* empty = bars + 4;          // [NSTAGE]
  uint64_t* done = bars + 8;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    prefetch_tensormap(&zmap); prefetch_tensormap(&xmap);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tap group and tile walk of this CTA
  const int group = Cfg::GROUPS == 2 ? (blockIdx.x & 1) : 0;
  const int cta = Cfg::GROUPS == 2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int ncta = Cfg::GROUPS == 2 ? (gridDim.x >> 1) : gridDim.x;
  const int tap0 = group == 0 ? 0 : Cfg::TAPS0;
  const int ntap = group == 0 ? Cfg::TAPS0 : 9 - Cfg::TAPS0;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int zplanes = p.Cout / 8;

  if (warp == 0) {
    // ================================ TMA producer ================================
    int it = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
      const int s = it % Cfg::NSTAGE, ph = (it / Cfg::NSTAGE) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        const int b = tile / tiles_per_img, r = tile % tiles_per_img;
        const int y0 = (r / p.tiles_x) * WG_TH, x0 = (r % p.tiles_x) * WG_TW;
        unsigned char* zs = stage_s + s * Cfg::STAGE;
        mbar_arrive_expect_tx(&full[s], zplanes * Cfg::ZPLANE + Cfg::XBYTES);
        tma_load_3d(zs, &zmap, &full[s], x0 * 8, y0, b * p.zP + p.zp0);
        tma_load_3d(zs + Cfg::ZBYTES, &xmap, &full[s], (x0 - 1) * 8, y0 - 1, b * p.xP + p.xp0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // One elected thread issues 8 K-steps x ntap MMAs per tile.  An MMA takes ~40-49 tensor-core cycles, so the issue
    // sequence must stay well below that per MMA: the tap loop is unrolled at compile time for each tap group (all
    // descriptor offsets are immediates, one add per operand), as in conv_tc.cu.
    int it = 0;
    for (int tile = cta; tile < p.ntiles; tile += ncta, ++it) {
      const int s = it % Cfg::NSTAGE, ph = (it / Cfg::NSTAGE) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t zaddr = smem_u32(stage_s + s * Cfg::STAGE);
        if (group == 0) issue_tile<CIN, FMT, 0, Cfg::TAPS0>(tmem_base, zaddr, it != 0);
        else issue_tile<CIN, FMT, Cfg::TAPS0, 9 - Cfg::TAPS0>(tmem_base, zaddr, it != 0);
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);                       // every MMA of this CTA has retired
    __syncwarp();
  } else {
    // ================================ epilogue (warps 2..5) ========================
    const bool any = cta < p.ntiles;
    if (any) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int quad = warp & 3;
      const int co = quad * 32 + lane;                          // accumulator row = output channel
      for (int t = 0; t < ntap; ++t) {
        const int tap = tap0 + t;
#pragma unroll
        for (int c0 = 0; c0 < CIN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + t * CIN + c0, v);
          tmem_ld_wait();
          // lanes = consecutive output channels: every warp-wide atomic covers one 128-byte line of ws[tap][ci][:]
          if (co < p.Cout) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(p.ws + ((size_t)tap * CIN + c0 + j) * p.Cout + co, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ws [9][Cin][Cout] -> dw [Cout][Cin][3][3]
__global__ void wgrad_tc_finish_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cin, int Cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // index into dw
  if (i >= Cout * Cin * 9) return;
  const int tap = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
  dw[i] = ws[((size_t)tap * Cin + ci) * Cout + co];
}

int init_wgrad_tc() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess) return cuda_fail(e);
  if (q != cudaDriverEntryPointSuccess || fn == nullptr) return TGSR_ERR_UNSUPPORTED;
  g_encode_w = reinterpret_cast<EncodeTiledFnW>(fn);
  e = cudaFuncSetAttribute(wgrad_tc_kernel<64, TGSR_FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgCfg<64>::SMEM);
  if (e != cudaSuccess) return cuda_fail(e);
  e = cudaFuncSetAttribute(wgrad_tc_kernel<64, TGSR_FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgCfg<64>::SMEM);
  if (e != cudaSuccess) return cuda_fail(e);
  e = cudaFuncSetAttribute(wgrad_tc_kernel<32, TGSR_FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgCfg<32>::SMEM);
  if (e != cudaSuccess) return cuda_fail(e);
  e = cudaFuncSetAttribute(wgrad_tc_kernel<32, TGSR_FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WgCfg<32>::SMEM);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

static int make_box_w(CUtensorMap* m, const void* x, int B, int P, int H, int W, int npl, int fmt, int box_w, int box_h) {
  cuuint64_t gdim[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B * P};
  cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[3] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, (cuuint32_t)npl};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_w(m, fmt == TGSR_FMT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                          const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return TGSR_ERR_CUDA; }
  return TGSR_OK;
}

}  // namespace tgsr

extern "C" int tgsr_conv3x3_wgrad_tc(const void* x, int xP, int xp0, const void* dz, int zP, int zp0, float* dw, float* ws,
                                     int B, int Cin, int H, int W, int Cout, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(dz); TGSR_REQUIRE_PTR(dw); TGSR_REQUIRE_PTR(ws);
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(g_sm_count > 0 && g_encode_w != nullptr, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((Cin == 32 || Cin == 64) && (Cout == 32 || Cout == 64 || Cout == 128) && H % WG_TH == 0 && W % WG_TW == 0,
               TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(xp0 >= 0 && xp0 + Cin / 8 <= xP && zp0 >= 0 && zp0 + Cout / 8 <= zP, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * xP < (1LL << 31) && (long long)B * zP < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap zmap, xmap;
  int s = make_box_w(&zmap, dz, B, zP, H, W, Cout / 8, fmt, WG_TW, WG_TH);
  if (s != TGSR_OK) return s;
  s = make_box_w(&xmap, x, B, xP, H, W, Cin / 8, fmt, WG_TW + 2, WG_TH + 2);
  if (s != TGSR_OK) return s;
  TGSR_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)Cout * Cin * 9, st));
  WgParams p;
  p.ws = ws; p.Cout = Cout; p.B = B; p.H = H; p.W = W;
  p.tiles_x = W / WG_TW; p.tiles_y = H / WG_TH; p.ntiles = B * p.tiles_x * p.tiles_y;
  p.zP = zP; p.zp0 = zp0; p.xP = xP; p.xp0 = xp0;
  const int groups = Cin == 64 ? 2 : 1;
  // every CTA ends with 9 * Cin * Cout / groups atomics whatever its share of the tiles: small layers get fewer CTAs (at
  // least 4 tiles each) so that this fixed cost does not exceed the MMA work
  int grid = g_sm_count / groups * groups;
  const int by_tiles = ((p.ntiles + 3) / 4) * groups;
  if (by_tiles < grid) grid = by_tiles;
#define TGSR_WG_LAUNCH(C, F) wgrad_tc_kernel<C, F><<<grid, WG_THREADS, WgCfg<C>::SMEM, st>>>(zmap, xmap, p)
  if (Cin == 64) { if (fmt == TGSR_FMT_F16) TGSR_WG_LAUNCH(64, TGSR_FMT_F16); else TGSR_WG_LAUNCH(64, TGSR_FMT_BF16); }
  else { if (fmt == TGSR_FMT_F16) TGSR_WG_LAUNCH(32, TGSR_FMT_F16); else TGSR_WG_LAUNCH(32, TGSR_FMT_BF16); }
#undef TGSR_WG_LAUNCH
  wgrad_tc_finish_kernel<<<ceil_div(Cout * Cin * 9, 256), 256, 0, st>>>(ws, dw, Cin, Cout);
  return launch_status(2);
}
