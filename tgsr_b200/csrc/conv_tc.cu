// tcgen05 implicit-GEMM 3x3 convolution on bf16 "blocked" activations, with the TGSR block
// epilogues fused: folded BatchNorm, GLU, residual add, and the nearest-x2 upsample of upBlock
// computed algebraically as four 2x2 sub-pixel phases (2.25x fewer MACs than convolving the
// upsampled image).
//
// reference: util.py:62-65 (conv3x3), nn.BatchNorm2d (eval), util.py:45-53 (GLU),
//            util.py:110-130 (ResBlock), util.py:74-80 (upBlock).
//
// Data layout in HBM ("blocked planes"): bf16 [B][P][H][W][8] -- channel c lives in plane c/8,
// element c%8.  A pixel row of one plane is contiguous (W*16 B), so
//   * one 4-D TMA box [8ch, TW+2, TH+2, planes] brings a whole halo patch into shared memory
//     already in the UMMA no-swizzle K-major core-matrix layout (8 consecutive pixels x 16 B =
//     one 128-B core matrix); out-of-bounds rows/columns are zero-filled by TMA = conv padding;
//   * each of the 9 filter taps is just a different START ADDRESS into the same patch
//     (descriptor start + (dy*PW+dx)*16 B), so the activations are read from L2/HBM once per
//     tile (1.4x halo overhead) instead of 9x as with im2col;
//   * the epilogue's per-thread 16-B stores (8 channels of one pixel) coalesce to full 128-B
//     lines across the 8 pixels of a tile row.
//
// GEMM view per CTA tile: M = 128 pixels (16 rows x 8 cols), N = Cout (pre-GLU), K = taps*Cin.
// Weights (B operand) are packed once on the host side of the ABI (tgsr_pack_*) into the exact
// shared-memory image and stay resident in smem for the CTA's lifetime (persistent kernel).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> BN/GLU/residual -> bf16 -> global).
// Pipelines: NSTAGE patch buffers (full/empty mbarriers) and 2 TMEM accumulator stages
// (tmem_full/tmem_empty), so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "fmt16.cuh"
#include <cstdlib>

namespace tgsr {

using namespace ptx;
typedef uint16_t h16;   // storage element of either 16-bit format

constexpr int TC_TH = 16, TC_TW = 8;            // tile = 16 rows x 8 cols = 128 pixels (UMMA M)
// kernel modes: what the op table / epilogue compute
constexpr int MODE_CONV = 0;    // conv3x3 (+BN, +GLU, +residual) -> blocked bf16
constexpr int MODE_UP = 1;      // nearest-x2 + conv3x3 + BN + GLU as four sub-pixel phases -> blocked bf16 at 2H x 2W
constexpr int MODE_HEAD3 = 2;   // conv3x3 32->3 image head -> fp32 NCHW (+tanh, one*y + a*addend)
constexpr int MODE_HEAD5 = 3;   // conv5x5 32->3 image head -> fp32 NCHW (+tanh, one*y + a*addend)
constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static int g_tc_debug = 0;

// ------------------------------------------------------------------------------------------
// MMA op tables.  An op = one filter tap (or one tap x phase group of the 4-phase upsample):
// A start offset inside the patch, N, accumulator column offset; B blocks are stored in op order.
struct TcOp { int tap_r, tap_c, n_mul, dcol_mul; };   // N = n_mul*NPRE, dcol = dcol_mul*NPRE

template <int MODE> struct OpTable;
template <> struct OpTable<MODE_CONV> {
  static constexpr int NOPS = 9;
  __host__ __device__ static constexpr TcOp op(int i) { return TcOp{i / 3, i % 3, 1, 0}; }
};
// Image heads (Cout = 3): the kx taps are folded into N (column n = o*K + kx of a row-tap ky), so a KxK
// head needs K MMA ops instead of K*K; the epilogue sums the K shifted partials per output.
template <> struct OpTable<MODE_HEAD3> {
  static constexpr int NOPS = 3;
  __host__ __device__ static constexpr TcOp op(int i) { return TcOp{i, 0, 1, 0}; }
};
template <> struct OpTable<MODE_HEAD5> {
  static constexpr int NOPS = 5;
  __host__ __device__ static constexpr TcOp op(int i) { return TcOp{i, 0, 1, 0}; }
};
// 4-phase nearest-x2 + conv3x3: low-res tap (r,c) in {0,1,2}^2 (offset -1..+1) feeds the output
// parities ph = py*2+px listed below (see pack kernel for the pre-summed weights).
template <> struct OpTable<MODE_UP> {
  static constexpr int NOPS = 11;
  __host__ __device__ static constexpr TcOp op(int i) {
    constexpr TcOp t[11] = {
        {1, 1, 4, 0},                                          // centre: all four phases
        {0, 1, 2, 0}, {2, 1, 2, 2},                            // up: ph0,ph1   down: ph2,ph3
        {1, 0, 1, 0}, {1, 0, 1, 2}, {1, 2, 1, 1}, {1, 2, 1, 3},  // left: ph0, ph2   right: ph1, ph3
        {0, 0, 1, 0}, {0, 2, 1, 1}, {2, 0, 1, 2}, {2, 2, 1, 3}};  // corners
    return t[i];
  }
};

template <int CIN, int NPRE, int MODE>
struct TcCfg {
  static constexpr bool UP = MODE == MODE_UP;
  static constexpr bool HEAD = MODE == MODE_HEAD3 || MODE == MODE_HEAD5;
  static constexpr int R = MODE == MODE_HEAD5 ? 2 : 1;         // halo radius
  static constexpr int KS = 2 * R + 1;                         // filter size
  // Conv/up tiles: 16 rows x 8 cols; accumulator row m = (y, x), 8-row groups are PW*16 B apart.
  // Head tiles: the patch is 16 pixels wide and M walks it LINEARLY (m = y*16 + x', 8 rows of 16, 8-row groups
  // 128 B apart).  Columns x' < TWO = 16-2R are outputs, the rest is halo, so every output's kx neighbours
  // (x'+kx <= 15) sit in the same patch row = the same warp (a warp owns 2 rows): the epilogue combines the
  // kx-folded partial sums with warp shuffles, no shared-memory exchange.
  static constexpr int PW = HEAD ? 16 : TC_TW + 2 * R;
  static constexpr int TWO = HEAD ? 16 - 2 * R : TC_TW;        // output columns per tile (12 / 14 for 5x5 / 3x3 heads)
  static constexpr int MB = HEAD ? 4 : 1;                      // M=128 blocks per tile (heads: 4 x (8 rows x 16 px))
  static constexpr int TH = HEAD ? 8 * MB : TC_TH;             // output rows per tile
  static constexpr int PH = TH + 2 * R;
  static constexpr int SBO_A = HEAD ? 128 : PW * 16;
  static constexpr int NACC = 2;                               // TMEM accumulator stages
  static constexpr int PLANE_BYTES = PH * PW * 16;             // one 8-channel plane of the halo patch
  static constexpr int NPL = CIN / 8;                          // input planes
  static constexpr int KSTEPS = CIN / 16;                      // UMMA K = 16
  static constexpr int PATCH_BYTES = NPL * PLANE_BYTES;
  static constexpr int ACC_COLS = UP ? 4 * NPRE : (HEAD ? 32 * (HEAD ? 4 : 1) : NPRE);   // fp32 accumulator columns per stage
  static constexpr int TMEM_COLS = NACC * ACC_COLS;            // power of two >= 32
  static constexpr int NTAPB = UP ? 16 : OpTable<MODE>::NOPS;  // number of [NPRE x CIN] weight blocks
  static constexpr int WBYTES = NTAPB * CIN * NPRE * 2;
  static constexpr int NSCALE = NPRE;
  // Patch-ring depth: as many stages as fit next to the resident weights (TMA latency ~1-2 us must be
  // covered by stages in flight: small-channel layers and heads spend < 0.5 us of MMA time per tile).
  static constexpr int WB_PAD_ = ((WBYTES + 1023) / 1024) * 1024;
  static constexpr int STAGE_ROOM = (232448 - 1024 - WB_PAD_ - 2 * NSCALE * 4 - 512) / PATCH_BYTES;
  static constexpr int NSTAGE = STAGE_ROOM > 12 ? 12 : STAGE_ROOM;
  static_assert(NSTAGE >= 2, "not enough shared memory for a double-buffered patch ring");
  static constexpr int PBUF_BYTES = 0;
  static constexpr size_t SMEM = 1024 /*align slack*/ + ((WBYTES + 1023) / 1024) * 1024 + NSTAGE * PATCH_BYTES +
                                 2 * NSCALE * 4 + 512 + PBUF_BYTES;
};

struct TcParams {
  const void* wpacked;
  const float* scale;      // [NPRE] or null
  const float* shift;
  const h16* residual;   // blocked, or null
  h16* out;              // blocked
  int rP, rp0, yP, yp0;            // planes per image / first plane of the residual and output views
  int xP, xp0;                     // input view (for the TMA plane coordinate)
  int B, H, W;                     // INPUT spatial size
  int tiles_x, tiles_y, ntiles;
  // image-head epilogue (MODE_HEAD*): y = out_scale * act(conv) + addend_scale * addend, fp32 NCHW [B,3,H,W]
  float* out_f32;
  const float* addend;
  float out_scale, addend_scale;
  int act_tanh;
  int debug;   // profiling aid (TGSR_TC_DEBUG env): 1 skip MMAs, 2 skip epilogue work, 4 skip TMA patch loads
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

template <int CIN, int NPRE, int MODE, bool GLU, int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p) {
  using Cfg = TcCfg<CIN, NPRE, MODE>;
  using Ops = OpTable<MODE>;
  constexpr bool UP = Cfg::UP;
  constexpr bool HEAD = Cfg::HEAD;
  constexpr int TC_PW = Cfg::PW;
  constexpr int TC_PLANE_BYTES = Cfg::PLANE_BYTES;
  constexpr int WB_PAD = ((Cfg::WBYTES + 1023) / 1024) * 1024;
  constexpr int TC_NSTAGE = Cfg::NSTAGE;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* w_s = base;
  unsigned char* patch_s = w_s + WB_PAD;
  float* sc_s = reinterpret_cast<float*>(patch_s + TC_NSTAGE * Cfg::PATCH_BYTES);
  float* sh_s = sc_s + Cfg::NSCALE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh_s + Cfg::NSCALE);
  uint64_t* full = bars;                       // [NSTAGE]
  uint64_t* empty = bars + TC_NSTAGE;          // [NSTAGE]
  uint64_t* wbar = bars + 2 * TC_NSTAGE;       // [1]
  uint64_t* tfull = wbar + 1;                  // [NACC <= 4]
  uint64_t* tempty = tfull + 4;                // [NACC <= 4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // Programmatic dependent launch: let the next kernel of the stream start its own prologue (barrier init, TMEM
  // allocation, weight TMA) on every SM this grid has left, instead of after the whole grid has drained.  Everything
  // below that touches activations produced by earlier kernels sits behind griddepcontrol.wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(wbar, 1);
    for (int a = 0; a < Cfg::NACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
    fence_barrier_init();
    prefetch_tensormap(&tmap);
  }
  for (int i = threadIdx.x; i < Cfg::NSCALE; i += TC_THREADS) sh_s[i] = p.shift ? p.shift[i] : 0.0f;
  (void)sc_s;   // the BN scale is folded into the packed weights (tgsr_pack_conv3x3)
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp runs the loop (uniform control flow); one elected lane issues.
    if (elect_one()) {
      mbar_arrive_expect_tx(wbar, Cfg::WBYTES);
      constexpr int CH = 32768;
      for (int off = 0; off < Cfg::WBYTES; off += CH) {
        const int n = Cfg::WBYTES - off < CH ? Cfg::WBYTES - off : CH;
        bulk_g2s(w_s + off, reinterpret_cast<const unsigned char*>(p.wpacked) + off, n, wbar);
      }
    }
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");      // weights are not produced by the predecessor; patches are
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      const int b = tile / tiles_per_img, r = tile % tiles_per_img;
      const int y0 = (r / p.tiles_x) * Cfg::TH, x0 = (r % p.tiles_x) * Cfg::TWO;
      if (p.debug & 4) {
        if (elect_one()) mbar_arrive(&full[s]);
      } else if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], Cfg::PATCH_BYTES);
        tma_load_3d(patch_s + s * Cfg::PATCH_BYTES, &tmap, &full[s], (x0 - Cfg::R) * 8, y0 - Cfg::R, b * p.xP + p.xp0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    mbar_wait(wbar, 0);
    tc_fence_after();
    const uint32_t w_addr = smem_u32(w_s);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1;
      const int acc = it % Cfg::NACC, aph = (it / Cfg::NACC) & 1;
      mbar_wait(&tempty[acc], aph ^ 1);            // epilogue has drained this accumulator stage
      mbar_wait(&full[s], ph);                     // patch landed
      tc_fence_after();
      if (p.debug & 1) {
        if (elect_one()) { mbar_arrive(&empty[s]); mbar_arrive(&tfull[acc]); }
      } else if (elect_one()) {
        // Descriptors are "per-tile base + compile-time constant": only the 14-bit start-address field of
        // the low word changes between MMAs, so each one costs a single 32-bit add on the issuing thread.
        const uint32_t a_lo0 = (smem_u32(patch_s + s * Cfg::PATCH_BYTES) >> 4) | ((uint32_t)(TC_PLANE_BYTES >> 4) << 16);
        constexpr uint32_t a_hi = (uint32_t)(Cfg::SBO_A >> 4) | (1u << 14);
        constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t b_lo0 = w_addr >> 4;
        const uint32_t d_base = tmem_base + acc * Cfg::ACC_COLS;
#pragma unroll
        for (int j = 0; j < Cfg::MB; ++j) {         // heads: 4 M-blocks of 8 patch rows each per tile
          uint32_t b_off = 0;                       // compile-time after unrolling
#pragma unroll
          for (int o = 0; o < Ops::NOPS; ++o) {
            const TcOp op = Ops::op(o);
            const int N = op.n_mul * NPRE;
            const uint32_t idesc = make_idesc_f16(128, N, Fmt16<FMT>::kUmmaFormat);
#pragma unroll
            for (int k = 0; k < Cfg::KSTEPS; ++k) {
              const uint32_t a_lo = a_lo0 + (uint32_t)((((j * 8 + op.tap_r) * TC_PW + op.tap_c) * 16 + k * 2 * TC_PLANE_BYTES) >> 4);
              const uint32_t b_lo = b_lo0 + (uint32_t)((b_off + k * (N * 32)) >> 4) + ((uint32_t)((N * 16) >> 4) << 16);
              umma_f16(d_base + op.dcol_mul * NPRE + j * 32, make_desc64(a_lo, a_hi), make_desc64(b_lo, b_hi), idesc,
                       (o | k) != 0 ? 1u : 0u);
            }
            b_off += CIN * N * 2;
          }
        }
        umma_commit(&empty[s]);                      // patch slot reusable once these MMAs retire
        umma_commit(&tfull[acc]);                    // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // ================================ epilogue (warps 2..9) ========================
    // Two warps per TMEM lane quadrant: `half` selects which half of the output channels (or which
    // sub-pixel phases / RGB outputs) this warp produces, so two warps per SM sub-partition work on
    // every tile and the per-tile epilogue latency halves.
    asm volatile("griddepcontrol.wait;" ::: "memory");      // skip tensor / addend reads and all output writes follow
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                // 0: warps 2..5, 1: warps 6..9
    const int m = quad * 32 + lane;                  // accumulator row = pixel within the tile
    const int ty = m >> 3, tx = m & 7;
    const float4* sh4 = reinterpret_cast<const float4*>(sh_s);
    uint4 rnext[4], rnext2[4];                       // skip-tensor planes of the next two tiles (conv mode with residual)
    float adnext[2][3];                              // image-head addend of the next tile
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int acc = it % Cfg::NACC, aph = (it / Cfg::NACC) & 1;
      const int b = tile / tiles_per_img, r = tile % tiles_per_img;
      const int gy = (r / p.tiles_x) * TC_TH + ty, gx = (r % p.tiles_x) * TC_TW + tx;
      const uint32_t t_base = tmem_base + acc * Cfg::ACC_COLS + ((uint32_t)(quad * 32) << 16);
      if (p.debug & 2) {
        mbar_wait(&tfull[acc], aph);
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        continue;
      }

      if constexpr (HEAD) {
        // Image head.  A tile is MB=4 accumulator blocks (columns j*32..) of 8 patch rows each; accumulator row
        // m = (y, x') = (m/16, m%16) of the 16-wide patch; column o*K + kx holds the partial sum over (ky, c):
        //   out[y][x'][o] = sum_kx P[(y, x'+kx)][o*K + kx]  -> lane shuffles by kx inside the warp.
        // Warp group `half` produces blocks 2*half and 2*half+1.
        constexpr int K = Cfg::KS;
        const int hy = m >> 4, hx = m & 15;
        const int ox = (r % p.tiles_x) * Cfg::TWO + hx;
        const size_t hw = (size_t)p.H * p.W;
        bool valid[2]; size_t oi[2]; float ad[2][3];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int oy = (r / p.tiles_x) * Cfg::TH + (half * 2 + q) * 8 + hy;
          valid[q] = hx < Cfg::TWO && ox < p.W && oy < p.H;
          oi[q] = (size_t)b * 3 * hw + (size_t)oy * p.W + ox;
        }
        // the addend (SRb, fp32) of this tile was requested one tile ago: no DRAM round trip in the per-tile critical path
        auto load_addend = [&](int tl, float (&dst)[2][3]) {
          const int bb = tl / tiles_per_img, rr = tl % tiles_per_img;
          const int xx = (rr % p.tiles_x) * Cfg::TWO + hx;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int yy = (rr / p.tiles_x) * Cfg::TH + (half * 2 + q) * 8 + hy;
            const bool ok = hx < Cfg::TWO && xx < p.W && yy < p.H;
            const size_t o0 = (size_t)bb * 3 * hw + (size_t)yy * p.W + xx;
#pragma unroll
            for (int o = 0; o < 3; ++o) dst[q][o] = ok ? __ldg(p.addend + o0 + o * hw) : 0.f;
          }
        };
        if (p.addend) {
          if (it == 0) load_addend(tile, adnext);
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int o = 0; o < 3; ++o) ad[q][o] = adnext[q][o];
          if (tile + (int)gridDim.x < p.ntiles) load_addend(tile + (int)gridDim.x, adnext);
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q) ad[q][0] = ad[q][1] = ad[q][2] = 0.f;
        }
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        uint32_t v[2][16];
        tmem_ld16(t_base + (half * 2) * 32, v[0]);
        tmem_ld16(t_base + (half * 2 + 1) * 32, v[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);               // TMEM stage is free again
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float z[3];
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            z[o] = __uint_as_float(v[q][o * K]);
#pragma unroll
            for (int kx = 1; kx < K; ++kx) z[o] += __shfl_down_sync(0xffffffffu, __uint_as_float(v[q][o * K + kx]), kx);
          }
          if (valid[q]) {
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              float t = z[o];
              if (p.act_tanh) t = tanhf(t);
              __stcs(p.out_f32 + oi[q] + o * hw, fmaf(p.addend_scale, ad[q][o], t * p.out_scale));
            }
          }
        }
        continue;                                               // (tempty already signalled)
      } else if constexpr (!UP) {
        constexpr int COUT = GLU ? NPRE / 2 : NPRE;            // channels written
        constexpr int CW = COUT / 2;                           // channels per epilogue warp (32 or 16)
        const int c0 = half * CW;
        const size_t pix = ((size_t)gy * p.W + gx) * 8;
        const size_t plane_elems = (size_t)p.H * p.W * 8;
        h16* obase = p.out + ((size_t)b * p.yP + p.yp0 + c0 / 8) * plane_elems + pix;
        // The skip tensor of this tile was requested two tiles ago (rnext2 -> rnext -> rres): in the HBM-bound ResBlock conv the epilogue
        // would otherwise serialise on one DRAM round trip per tile.
        uint4 rres[CW / 8];
        if (p.residual) {
          auto load_res = [&](int tl, uint4* dst) {
            const int bb = tl / tiles_per_img, rr = tl % tiles_per_img;
            const int yy = (rr / p.tiles_x) * TC_TH + ty, xx = (rr % p.tiles_x) * TC_TW + tx;
            const h16* rbase = p.residual + ((size_t)bb * p.rP + p.rp0 + c0 / 8) * plane_elems + ((size_t)yy * p.W + xx) * 8;
#pragma unroll
            for (int j = 0; j < CW / 8; ++j) dst[j] = __ldg(reinterpret_cast<const uint4*>(rbase + (size_t)j * plane_elems));
          };
          const int g2 = (int)gridDim.x;
          if (it == 0) {
            load_res(tile, rnext);
            if (tile + g2 < p.ntiles) load_res(tile + g2, rnext2);
          }
#pragma unroll
          for (int j = 0; j < CW / 8; ++j) { rres[j] = rnext[j]; rnext[j] = rnext2[j]; }
          if (tile + 2 * g2 < p.ntiles) load_res(tile + 2 * g2, rnext2);      // two tiles ahead: ~32 KB of reads in flight per SM
        }
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        uint32_t va[CW], vg[CW];
        if constexpr (CW == 32) {
          tmem_ld32(t_base + c0, va);
          if constexpr (GLU) tmem_ld32(t_base + COUT + c0, vg);
        } else {
          tmem_ld16(t_base + c0, va);
          if constexpr (GLU) tmem_ld16(t_base + COUT + c0, vg);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {                      // planes of 8 channels
          float o[8], res[8];
          if (p.residual) unpack8_16<FMT>(rres[j], res);
          const float4 sa0 = sh4[(c0 + j * 8) / 4], sa1 = sh4[(c0 + j * 8) / 4 + 1];
          const float sa[8] = {sa0.x, sa0.y, sa0.z, sa0.w, sa1.x, sa1.y, sa1.z, sa1.w};
          float sg[8];
          if constexpr (GLU) {
            const float4 g0 = sh4[(COUT + c0 + j * 8) / 4], g1 = sh4[(COUT + c0 + j * 8) / 4 + 1];
            sg[0] = g0.x; sg[1] = g0.y; sg[2] = g0.z; sg[3] = g0.w; sg[4] = g1.x; sg[5] = g1.y; sg[6] = g1.z; sg[7] = g1.w;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float z = __uint_as_float(va[j * 8 + e]) + sa[e];
            if constexpr (GLU) z *= fast_sigmoid(__uint_as_float(vg[j * 8 + e]) + sg[e]);
            if (p.residual) z += res[e];
            o[e] = z;
          }
          *reinterpret_cast<uint4*>(obase + (size_t)j * plane_elems) = pack8_16<FMT>(o);
        }
      } else {
        // 4-phase upsample: accumulator columns [ph*NPRE, (ph+1)*NPRE) hold output pixel
        // (2*gy + ph/2, 2*gx + ph%2); GLU halves inside each phase block.  half h -> phases 2h, 2h+1
        // (= output row 2*gy + h, both columns: one 32-byte run per plane).
        static_assert(GLU && NPRE == 64, "upsample path is conv(C -> 64) + GLU -> 32 channels");
        constexpr int COUT = NPRE / 2;                          // 32
        const int Ho = 2 * p.H, Wo = 2 * p.W;
        const size_t plane_elems = (size_t)Ho * Wo * 8;
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        // both sub-pixel columns (phases 2*half, 2*half+1) of a plane are produced together: output pixels (oy, 2gx) and
        // (oy, 2gx+1) are 32 contiguous bytes -> one 256-bit store per plane, full 32-byte sectors on every request
        const int oy = 2 * gy + half, ox = 2 * gx;
        h16* obase = p.out + ((size_t)b * p.yP + p.yp0) * plane_elems + ((size_t)oy * Wo + ox) * 8;
        const uint32_t t0 = t_base + (half * 2) * NPRE, t1 = t0 + NPRE;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t a0[8], g0[8], a1[8], g1[8];
          tmem_ld8(t0 + j * 8, a0); tmem_ld8(t0 + COUT + j * 8, g0);
          tmem_ld8(t1 + j * 8, a1); tmem_ld8(t1 + COUT + j * 8, g1);
          tmem_ld_wait();
          const float4 sa0 = sh4[j * 2], sa1 = sh4[j * 2 + 1], sg0 = sh4[(COUT + j * 8) / 4], sg1 = sh4[(COUT + j * 8) / 4 + 1];
          const float sa[8] = {sa0.x, sa0.y, sa0.z, sa0.w, sa1.x, sa1.y, sa1.z, sa1.w};
          const float sg[8] = {sg0.x, sg0.y, sg0.z, sg0.w, sg1.x, sg1.y, sg1.z, sg1.w};
          float o0[8], o1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            o0[e] = (__uint_as_float(a0[e]) + sa[e]) * fast_sigmoid(__uint_as_float(g0[e]) + sg[e]);
            o1[e] = (__uint_as_float(a1[e]) + sa[e]) * fast_sigmoid(__uint_as_float(g1[e]) + sg[e]);
          }
          stg256(obase + (size_t)j * plane_elems, pack8_16<FMT>(o0), pack8_16<FMT>(o1));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------
// Weight packing (runs once per weight version).  Packed image = for each op, for each K chunk
// kc (8 input channels), for each group of 8 output rows: a 128-B core matrix [8 rows][8 k] bf16.
//   byte offset(op, n, c) = op_base + (c/8)*(N_op*16) + (n/8)*128 + (n%8)*16 + (c%8)*2
template <int FMT>
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, const float* __restrict__ scale, h16* __restrict__ out,
                                    int Cout, int Cin) {
  const int total = 9 * Cout * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % Cin, n = (i / Cin) % Cout, t = i / (Cin * Cout);
    const float v = w[((size_t)n * Cin + c) * 9 + t] * (scale ? scale[n] : 1.0f);   // BN scale folded, fp32
    const size_t off = (size_t)t * Cin * Cout + (size_t)(c / 8) * (Cout * 8) + (size_t)(n / 8) * 64 + (n % 8) * 8 + (c % 8);
    out[off] = Fmt16<FMT>::cvt(v);
  }
}

// Image heads: w [3][Cin][K][K] -> per ROW tap ky a [16 x Cin] block; row n = o*K + kx (n < 3K), rest zero.
template <int FMT>
__global__ void pack_head_kernel(const float* __restrict__ w, h16* __restrict__ out, int Cin, int K) {
  const int total = K * 16 * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % Cin, n = (i / Cin) % 16, ky = i / (Cin * 16);
    float v = 0.f;
    if (n < 3 * K) {
      const int o = n / K, kx = n % K;
      v = w[(((size_t)o * Cin + c) * K + ky) * K + kx];
    }
    const size_t off = (size_t)ky * Cin * 16 + (size_t)(c / 8) * (16 * 8) + (size_t)(n / 8) * 64 + (n % 8) * 8 + (c % 8);
    out[off] = Fmt16<FMT>::cvt(v);
  }
}

// Row sets of the 3x3 kernel that collapse onto low-res offset r (0,1,2 = -1,0,+1) for output
// parity py:  py=0: r=0 <- {ky 0}, r=1 <- {ky 1,2};  py=1: r=1 <- {ky 0,1}, r=2 <- {ky 2}.
__device__ __forceinline__ int up_taps(int parity, int r, int (&ks)[2]) {
  if (parity == 0) { if (r == 0) { ks[0] = 0; return 1; } if (r == 1) { ks[0] = 1; ks[1] = 2; return 2; } return 0; }
  if (r == 1) { ks[0] = 0; ks[1] = 1; return 2; }
  if (r == 2) { ks[0] = 2; return 1; }
  return 0;
}

template <int FMT>
__global__ void pack_up2x_kernel(const float* __restrict__ w, const float* __restrict__ scale, h16* __restrict__ out,
                                 int Cout, int Cin) {
  // 11 ops; op o covers phases {ph = dcol_mul .. dcol_mul + n_mul - 1}; rows of op = [phase][Cout]
  using Ops = OpTable<MODE_UP>;
  int op_base = 0;
  for (int o = 0; o < Ops::NOPS; ++o) {
    const TcOp op = Ops::op(o);
    const int N = op.n_mul * Cout;
    const int total = N * Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const int c = i % Cin, nn = i / Cin;
      const int phase = op.dcol_mul + nn / Cout, n = nn % Cout;
      const int py = phase >> 1, px = phase & 1;
      int kys[2], kxs[2];
      const int ny = up_taps(py, op.tap_r, kys), nx = up_taps(px, op.tap_c, kxs);
      float v = 0.f;                                        // pre-summed in fp32, rounded once
      for (int a = 0; a < ny; ++a)
        for (int bq = 0; bq < nx; ++bq) v += w[((size_t)n * Cin + c) * 9 + kys[a] * 3 + kxs[bq]];
      if (scale) v *= scale[n];
      const size_t off = (size_t)op_base + (size_t)(c / 8) * (N * 8) + (size_t)(nn / 8) * 64 + (nn % 8) * 8 + (c % 8);
      out[off] = Fmt16<FMT>::cvt(v);
    }
    op_base += N * Cin;
  }
}

// ------------------------------------------------------------------------------------------
// Layout converters between the reference's NCHW fp32 tensors and the blocked bf16 layout.
template <int FMT>
__global__ void nchw_to_blocked_kernel(const float* __restrict__ x, h16* __restrict__ y, int B, int C, int HW,
                                       int P, int p0) {
  // one thread per (b, plane, pixel): gathers 8 channels (coalesced across pixels), writes 16 B
  const size_t total = (size_t)B * (C / 8) * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int pl = (int)((i / HW) % (C / 8));
    const int b = (int)(i / ((size_t)HW * (C / 8)));
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __ldg(x + ((size_t)b * C + pl * 8 + e) * HW + pix);
    *reinterpret_cast<uint4*>(y + (((size_t)b * P + p0 + pl) * HW + pix) * 8) = pack8_16<FMT>(v);
  }
}
template <int FMT>
__global__ void blocked_to_nchw_kernel(const h16* __restrict__ x, float* __restrict__ y, int B, int C, int HW,
                                       int P, int p0) {
  const size_t total = (size_t)B * (C / 8) * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int pl = (int)((i / HW) % (C / 8));
    const int b = (int)(i / ((size_t)HW * (C / 8)));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * P + p0 + pl) * HW + pix) * 8));
    float v[8];
    unpack8_16<FMT>(u, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) y[((size_t)b * C + pl * 8 + e) * HW + pix] = v[e];
  }
}

// ------------------------------------------------------------------------------------------
template <int CIN, int NPRE, int MODE, bool GLU>
static int set_tc_attr() {
  cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<CIN, NPRE, MODE, GLU, TGSR_FMT_BF16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<CIN, NPRE, MODE>::SMEM);
  if (e != cudaSuccess) return cuda_fail(e);
  e = cudaFuncSetAttribute(conv3x3_tc_kernel<CIN, NPRE, MODE, GLU, TGSR_FMT_F16>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<CIN, NPRE, MODE>::SMEM);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

int init_conv_tc() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess) return cuda_fail(e);
  if (q != cudaDriverEntryPointSuccess || fn == nullptr) return TGSR_ERR_UNSUPPORTED;
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  if (const char* d = getenv("TGSR_TC_DEBUG")) g_tc_debug = atoi(d);
  int s;
  if ((s = set_tc_attr<64, 128, MODE_CONV, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<64, 64, MODE_CONV, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 64, MODE_CONV, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 32, MODE_CONV, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<64, 64, MODE_UP, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 64, MODE_UP, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 16, MODE_HEAD3, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 16, MODE_HEAD5, false>()) != TGSR_OK) return s;
  return TGSR_OK;
}

static int make_patch_tmap(CUtensorMap* m, const void* x, int B, int P, int H, int W, int npl, int fmt, int R = 1,
                           int head = 0) {
  const int TC_PW = head ? 16 : TC_TW + 2 * R;
  const int TC_PH = (head ? 32 : TC_TH) + 2 * R;
  // A pixel row of one plane is W*8 contiguous 16-bit elements, so (x, c%8) is described as ONE inner
  // dimension: the box's inner run is PW*16 bytes (160/192 B) instead of 16 B, which is what the TMA
  // engine's per-row cost wants; the bytes landing in shared memory are identical to a [8,PW,PH,planes] box.
  cuuint64_t gdim[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B * P};
  cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[3] = {(cuuint32_t)TC_PW * 8, (cuuint32_t)TC_PH, (cuuint32_t)npl};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_tiled(m, fmt == TGSR_FMT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                              const_cast<void*>(x), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return TGSR_ERR_CUDA; }
  return TGSR_OK;
}

template <int CIN, int NPRE, int MODE, bool GLU>
static int launch_tc(const CUtensorMap& tmap, const TcParams& p, int fmt, cudaStream_t st) {
  static const int use_pdl = [] { const char* e = getenv("TGSR_PDL"); return e ? atoi(e) : 1; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.ntiles < g_sm_count ? p.ntiles : g_sm_count);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcCfg<CIN, NPRE, MODE>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
  cudaError_t e;
  if (fmt == TGSR_FMT_F16) e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<CIN, NPRE, MODE, GLU, TGSR_FMT_F16>, tmap, p);
  else e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<CIN, NPRE, MODE, GLU, TGSR_FMT_BF16>, tmap, p);
  if (e != cudaSuccess) return cuda_fail(e);
  return launch_status();
}

}  // namespace tgsr

#define TGSR_CHECK_FMT(fmt) TGSR_REQUIRE((fmt) == TGSR_FMT_BF16 || (fmt) == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED)

namespace tgsr {
int img_head_mma(const void* x, int xP, int xp0, const void* wpacked, const float* addend, float out_scale, float addend_scale,
                 float* y, int B, int H, int W, int ksize, int act_tanh, int fmt, cudaStream_t st);   // head_mma.cu
}

extern "C" int tgsr_conv3x3_tc(const void* x, int xP, int xp0, const void* wpacked, const float* shift,
                               const void* residual, int rP, int rp0, void* y, int yP, int yp0, int B, int Cin,
                               int H, int W, int Cout, int up2x, int act, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(wpacked); TGSR_REQUIRE_PTR(y);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(g_sm_count > 0 && g_encode_tiled != nullptr, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(H % TC_TH == 0 && W % TC_TW == 0, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_GLU, TGSR_ERR_UNSUPPORTED);
  const int npl = Cin / 8;
  const int cout_planes = (act == TGSR_ACT_GLU ? Cout / 2 : Cout) / 8;
  TGSR_REQUIRE(xp0 >= 0 && xp0 + npl <= xP && yp0 >= 0 && yp0 + cout_planes <= yP, TGSR_ERR_BAD_SHAPE);
  if (residual) TGSR_REQUIRE(rp0 >= 0 && rp0 + cout_planes <= rP && !up2x && act == TGSR_ACT_NONE, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(wpacked) & 15) == 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * xP < (1LL << 31), TGSR_ERR_UNSUPPORTED);

  CUtensorMap tmap;
  int s = make_patch_tmap(&tmap, x, B, xP, H, W, npl, fmt);
  if (s != TGSR_OK) return s;
  TcParams p;
  p.wpacked = wpacked; p.scale = nullptr; p.shift = shift;
  p.residual = reinterpret_cast<const h16*>(residual);
  p.out = reinterpret_cast<h16*>(y);
  p.rP = rP; p.rp0 = rp0; p.yP = yP; p.yp0 = yp0; p.xP = xP; p.xp0 = xp0;
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = W / TC_TW; p.tiles_y = H / TC_TH; p.ntiles = B * p.tiles_x * p.tiles_y;
  p.out_f32 = nullptr; p.addend = nullptr; p.out_scale = 1.f; p.addend_scale = 0.f; p.act_tanh = 0;
  p.debug = g_tc_debug;
  cudaStream_t st = (cudaStream_t)stream;
  const bool glu = act == TGSR_ACT_GLU;
  if (!up2x) {
    if (Cin == 64 && Cout == 128 && glu) return launch_tc<64, 128, MODE_CONV, true>(tmap, p, fmt, st);
    if (Cin == 64 && Cout == 64 && !glu) return launch_tc<64, 64, MODE_CONV, false>(tmap, p, fmt, st);
    if (Cin == 32 && Cout == 64 && glu) return launch_tc<32, 64, MODE_CONV, true>(tmap, p, fmt, st);
    if (Cin == 32 && Cout == 32 && !glu) return launch_tc<32, 32, MODE_CONV, false>(tmap, p, fmt, st);
  } else {
    if (Cin == 64 && Cout == 64 && glu) return launch_tc<64, 64, MODE_UP, true>(tmap, p, fmt, st);
    if (Cin == 32 && Cout == 64 && glu) return launch_tc<32, 64, MODE_UP, true>(tmap, p, fmt, st);
  }
  return TGSR_ERR_UNSUPPORTED;
}

extern "C" int tgsr_pack_conv3x3(const float* w, const float* scale, void* packed, int Cout, int Cin, int up2x,
                                 int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(packed);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(Cout > 0 && Cin > 0 && Cout % 8 == 0 && Cin % 8 == 0, TGSR_ERR_BAD_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  h16* o = reinterpret_cast<h16*>(packed);
  if (fmt == TGSR_FMT_F16) {
    if (up2x) pack_up2x_kernel<TGSR_FMT_F16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin);
    else pack_conv3x3_kernel<TGSR_FMT_F16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin);
  } else {
    if (up2x) pack_up2x_kernel<TGSR_FMT_BF16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin);
    else pack_conv3x3_kernel<TGSR_FMT_BF16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin);
  }
  return launch_status();
}

extern "C" int tgsr_pack_head(const float* w, void* packed, int Cin, int ksize, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(packed);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(Cin > 0 && Cin % 8 == 0 && (ksize == 3 || ksize == 5), TGSR_ERR_BAD_SHAPE);
  if (fmt == TGSR_FMT_F16)
    pack_head_kernel<TGSR_FMT_F16><<<32, 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<h16*>(packed), Cin, ksize);
  else
    pack_head_kernel<TGSR_FMT_BF16><<<32, 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<h16*>(packed), Cin, ksize);
  return launch_status();
}

extern "C" int tgsr_img_head_tc(const void* x, int xP, int xp0, const void* wpacked, const float* addend,
                                float out_scale, float addend_scale, float* y, int B, int Cin, int H, int W,
                                int ksize, int act, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(wpacked); TGSR_REQUIRE_PTR(y);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(g_sm_count > 0 && g_encode_tiled != nullptr, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(Cin == 32 && (ksize == 3 || ksize == 5) && W % 8 == 0, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_TANH, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(xp0 >= 0 && xp0 + Cin / 8 <= xP, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * xP < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  // TGSR_HEAD_IMPL=mma selects the warp-level tensor-core head (head_mma.cu): parity-tested, but measured slower than the
  // tcgen05 head below (72.6 vs 48.9 us for the 3x3 head at 256x256, batch 32), so it is not the default.
  static const bool use_mma = [] { const char* e = getenv("TGSR_HEAD_IMPL"); return e != nullptr && e[0] == 'm'; }();
  if (use_mma)
    return img_head_mma(x, xP, xp0, wpacked, addend, out_scale, addend_scale, y, B, H, W, ksize, act == TGSR_ACT_TANH, fmt,
                        (cudaStream_t)stream);
  CUtensorMap tmap;
  int s = make_patch_tmap(&tmap, x, B, xP, H, W, Cin / 8, fmt, ksize / 2, 1);
  if (s != TGSR_OK) return s;
  TcParams p;
  p.wpacked = wpacked; p.scale = nullptr; p.shift = nullptr; p.residual = nullptr; p.out = nullptr;
  p.rP = p.rp0 = p.yP = p.yp0 = 0; p.xP = xP; p.xp0 = xp0;
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = W / TC_TW; p.tiles_y = H / TC_TH; p.ntiles = B * p.tiles_x * p.tiles_y;
  const int two = 16 - 2 * (ksize / 2);                      // output columns per head tile (32 rows x two cols)
  p.tiles_x = (W + two - 1) / two; p.tiles_y = (H + 31) / 32; p.ntiles = B * p.tiles_x * p.tiles_y;
  p.out_f32 = y; p.addend = addend; p.out_scale = out_scale; p.addend_scale = addend_scale;
  p.act_tanh = act == TGSR_ACT_TANH;
  p.debug = g_tc_debug;
  cudaStream_t st = (cudaStream_t)stream;
  if (ksize == 3) return launch_tc<32, 16, MODE_HEAD3, false>(tmap, p, fmt, st);
  return launch_tc<32, 16, MODE_HEAD5, false>(tmap, p, fmt, st);
}

extern "C" int tgsr_nchw_to_blocked(const float* x, void* y, int B, int C, int H, int W, int yP, int yp0, int fmt,
                                    tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(y);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0 && yp0 >= 0 && yp0 + C / 8 <= yP, TGSR_ERR_BAD_SHAPE);
  const size_t total = (size_t)B * (C / 8) * H * W;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (fmt == TGSR_FMT_F16)
    nchw_to_blocked_kernel<TGSR_FMT_F16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<h16*>(y), B, C, H * W, yP, yp0);
  else
    nchw_to_blocked_kernel<TGSR_FMT_BF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<h16*>(y), B, C, H * W, yP, yp0);
  return launch_status();
}

extern "C" int tgsr_blocked_to_nchw(const void* x, float* y, int B, int C, int H, int W, int xP, int xp0, int fmt,
                                    tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(y);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(B > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0 && xp0 >= 0 && xp0 + C / 8 <= xP, TGSR_ERR_BAD_SHAPE);
  const size_t total = (size_t)B * (C / 8) * H * W;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (fmt == TGSR_FMT_F16)
    blocked_to_nchw_kernel<TGSR_FMT_F16><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const h16*>(x), y, B, C, H * W, xP, xp0);
  else
    blocked_to_nchw_kernel<TGSR_FMT_BF16><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const h16*>(x), y, B, C, H * W, xP, xp0);
  return launch_status();
}
