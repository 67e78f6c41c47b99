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
#include <type_traits>

namespace tgsr {

using namespace ptx;
typedef uint16_t h16;   // storage element of either 16-bit format

constexpr int TC_TH = 16, TC_TW = 8;            // tile = 16 rows x 8 cols = 128 pixels (UMMA M)
// kernel modes: what the op table / epilogue compute
constexpr int MODE_CONV = 0;    // conv3x3 (+BN, +GLU, +residual) -> blocked bf16
constexpr int MODE_UP = 1;      // nearest-x2 + conv3x3 + BN + GLU as four sub-pixel phases -> blocked bf16 at 2H x 2W
constexpr int MODE_HEAD3 = 2;   // conv3x3 32->3 image head -> fp32 NCHW (+tanh, one*y + a*addend)
constexpr int MODE_HEAD5 = 3;   // conv5x5 32->3 image head -> fp32 NCHW (+tanh, one*y + a*addend)
constexpr int MODE_HEAD3W = 4;  // the same heads on 32-pixel-wide patches (images >= 128 wide): half the TMA rows per output
constexpr int MODE_HEAD5W = 5;
constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

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
template <> struct OpTable<MODE_HEAD3W> : OpTable<MODE_HEAD3> {};
template <> struct OpTable<MODE_HEAD5W> : OpTable<MODE_HEAD5> {};
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

template <int CIN, int NPRE, int MODE, bool GLU>
struct TcCfg {
  static constexpr bool UP = MODE == MODE_UP;
  static constexpr bool HEAD = MODE >= MODE_HEAD3;
  static constexpr bool WIDE = MODE == MODE_HEAD3W || MODE == MODE_HEAD5W;
  static constexpr int R = (MODE == MODE_HEAD5 || MODE == MODE_HEAD5W) ? 2 : 1;         // halo radius
  static constexpr int KS = 2 * R + 1;                         // filter size
  // Conv/up tiles: 16 rows x 8 cols; accumulator row m = (y, x), 8-row groups are PW*16 B apart.
  // Head tiles: the patch is HPW = 16 (or 32) pixels wide and M walks it LINEARLY (m = y*HPW + x', 8-pixel groups
  // 128 B apart).  Columns x' < TWO = HPW-2R are outputs, the rest is halo, so every output's kx neighbours
  // (x'+kx < HPW) sit in the same patch row = the same warp (a warp owns 2 rows of 16 or 1 row of 32): the epilogue
  // combines the kx-folded partial sums with warp shuffles, no shared-memory exchange.  The TMA engine spends ~10
  // cycles per box ROW whatever its length, and a head is bound by exactly that (136 rows per tile): the 32-wide
  // variant brings 2.1x the outputs per row and is used whenever the image is at least 128 wide.
  static constexpr int HPW = WIDE ? 32 : 16;
  static constexpr int PW = HEAD ? HPW : TC_TW + 2 * R;
  static constexpr int TWO = HEAD ? HPW - 2 * R : TC_TW;       // output columns per tile
  static constexpr int RB = HEAD ? 128 / HPW : TC_TH;          // patch rows per M=128 block (heads: 8 or 4)
  static constexpr int MB = HEAD ? 32 / RB : 1;                // M=128 blocks per tile (heads: 32 output rows)
  static constexpr int TH = HEAD ? RB * MB : TC_TH;            // output rows per tile
  static constexpr int PH = TH + 2 * R;
  static constexpr int SBO_A = HEAD ? 128 : PW * 16;
  static constexpr int PLANE_BYTES = PH * PW * 16;             // one 8-channel plane of the halo patch
  static constexpr int NPL = CIN / 8;                          // input planes
  static constexpr int KSTEPS = CIN / 16;                      // UMMA K = 16
  static constexpr int PATCH_BYTES = NPL * PLANE_BYTES;
  static constexpr int ACC_COLS = UP ? 4 * NPRE : (HEAD ? 32 * MB : NPRE);   // fp32 accumulator columns per stage
  // TMEM accumulator stages: the epilogue frees a stage as soon as its values are in registers, so two stages already
  // decouple it from the MMA issuer; the rest of the 512 columns is used when the tile is narrow.
  static constexpr int NACC = 512 / ACC_COLS > 4 ? 4 : 512 / ACC_COLS;
  static constexpr int TMEM_COLS = NACC * ACC_COLS;            // power of two >= 32
  static constexpr int NTAPB = UP ? 16 : OpTable<MODE>::NOPS;  // number of [NPRE x CIN] weight blocks
  static constexpr int WBYTES_TAPS = NTAPB * CIN * NPRE * 2;
  // Folded-BatchNorm shift: applied by ONE extra MMA per tile, D = ones[128 x 16] * shift[N x 16]^T, issued first with
  // accumulate = 0, so the epilogue starts from shifted accumulators (no per-element add, no shift table in shared
  // memory).  The shift is split into three 16-bit terms (k = 0,1,2 against three ones), exact to 2^-24.
  static constexpr bool BIAS = !HEAD;
  static constexpr int NBIAS = UP ? 4 * NPRE : NPRE;           // bias-block rows = accumulator columns
  static constexpr int BIAS_BYTES = BIAS ? NBIAS * 32 : 0;     // [N x 16] 16-bit, K-major core matrices
  static constexpr int ONES_BYTES = BIAS ? 4096 : 0;           // [128 x 16] 16-bit
  static constexpr int WBYTES = WBYTES_TAPS + BIAS_BYTES + ONES_BYTES;
  static constexpr int WB_PAD = ((WBYTES + 1023) / 1024) * 1024;
  // Skip-tensor ring (ResBlock second conv): the producer warp fetches the skip tile with TMA RSTAGE tiles ahead, so
  // the epilogue never waits for a DRAM round trip of its own.
  static constexpr bool RES = MODE == MODE_CONV && !GLU && NPRE <= 64;   // (the 128-channel act-free conv is training-only: no skip)
  static constexpr int RSTAGE = RES ? 3 : 0;
  static constexpr int RES_BYTES = RES ? TC_TH * TC_TW * NPRE * 2 : 0;
  // Patch-ring depth: as many stages as fit next to the resident weights (TMA latency ~1-2 us must be
  // covered by stages in flight: small-channel layers and heads spend < 0.5 us of MMA time per tile).
  // Kernels with small weights (Cin = 32, heads) stay below half of an SM's shared memory and TMEM: the CTA of the next
  // kernel in the stream (programmatic dependent launch) then becomes resident while this one still runs and has its
  // barriers, TMEM and weights ready when the grid dependency resolves -- these layers are launch-latency bound.
  static constexpr int HALF_SM = 113 * 1024;
  static constexpr bool SMALL = WB_PAD + RSTAGE * RES_BYTES + 4 * PATCH_BYTES + 1536 <= HALF_SM;
  static constexpr int BUDGET = SMALL ? HALF_SM : 232448;
  static_assert(!SMALL || TMEM_COLS <= 256, "a co-resident pair of small kernels must fit the 512 TMEM columns");
  static constexpr int STAGE_ROOM = (BUDGET - 1024 - WB_PAD - RSTAGE * RES_BYTES - 512) / PATCH_BYTES;
  static constexpr int NSTAGE = STAGE_ROOM > 12 ? 12 : STAGE_ROOM;
  static_assert(NSTAGE >= 2, "not enough shared memory for a double-buffered patch ring");
  static constexpr size_t SMEM = 1024 /*align slack*/ + WB_PAD + NSTAGE * PATCH_BYTES + RSTAGE * RES_BYTES + 512;
};

struct TcParams {
  const void* wpacked;   // [tap weight blocks | shift block | ones block] (tgsr_pack_conv3x3 / tgsr_pack_head)
  h16* out;              // blocked
  int has_res;                     // conv mode: add the skip tile fetched through `rmap`
  int rP, rp0, yP, yp0;            // planes per image / first plane of the residual and output views
  int xP, xp0;                     // input view (for the TMA plane coordinate)
  int B, H, W;                     // INPUT spatial size
  int tiles_x, tiles_y, ntiles;
  int step_x, step_y, step_b;      // gridDim.x decomposed in (tile x, tile y, image) units: the per-CTA tile stride
  int reverse;                     // walk the tiles from the last one backwards (see tgsr_conv3x3_tc flags)
  int l2_in_first, l2_out_last;    // L2 eviction hints: inputs are dead after this kernel / the output is read next (flags)
  // image-head epilogue (MODE_HEAD*): y = out_scale * act(conv) + addend_scale * addend, fp32 NCHW [B,3,H,W]
  float* out_f32;
  const float* addend;
  float out_scale, addend_scale;
  int act_tanh;
  int* nonfinite;        // optional device flag: set to 1 when a head accumulator is inf / NaN (16-bit overflow upstream)
};

// Tile walk of a persistent CTA: tile index t = (b * tiles_y + ty) * tiles_x + tx, advanced by gridDim.x with carries
// instead of two integer divisions per tile and role.
struct TileIter {
  int b, ty, tx;
  // k-th tile of this CTA = blockIdx.x + k * gridDim.x, counted from the last tile backwards when p.reverse is set.
  __device__ __forceinline__ void init(int first, const TcParams& p) {
    const int tile = p.reverse ? p.ntiles - 1 - first : first;
    tx = tile % p.tiles_x;
    const int q = tile / p.tiles_x;
    ty = q % p.tiles_y;
    b = q / p.tiles_y;
  }
  __device__ __forceinline__ void next(const TcParams& p) {
    if (!p.reverse) {
      tx += p.step_x;
      if (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
      ty += p.step_y;
      if (ty >= p.tiles_y) { ty -= p.tiles_y; ++b; }
      b += p.step_b;
    } else {
      tx -= p.step_x;
      if (tx < 0) { tx += p.tiles_x; --ty; }
      ty -= p.step_y;
      if (ty < 0) { ty += p.tiles_y; --b; }
      b -= p.step_b;
    }
  }
};

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GLU of one (value, gate) accumulator pair.  The packed weights and shift of BOTH halves carry a factor 1/2 (exact),
// so z * sigmoid(g) = z/2 + z/2 * tanh(g/2): one MUFU + one FMA per output instead of exp + reciprocal (the 4-phase
// up-block kernels evaluate 16384 gates per tile and were bound by the MUFU pipe).  tanh.approx has a relative error of
// 2^-11: below the bf16 rounding of the result and at the fp16 rounding of it.  -DTGSR_GLU_EXACT keeps exp + reciprocal.
#ifdef TGSR_GLU_EXACT
constexpr bool kGluTanh = false;
#else
constexpr bool kGluTanh = true;
#endif
template <int FMT>
__device__ __forceinline__ float glu_pair(uint32_t a, uint32_t g) {
  if constexpr (kGluTanh) {
    const float zh = __uint_as_float(a);
    return fmaf(zh, tanh_approx(__uint_as_float(g)), zh);
  } else {
    return __uint_as_float(a) * fast_sigmoid(__uint_as_float(g));
  }
}
__host__ __device__ constexpr float glu_prescale(bool glu, int fmt) { return (glu && kGluTanh) ? 0.5f : 1.0f; }

__device__ __forceinline__ uint4 lds128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}

template <int CIN, int NPRE, int MODE, bool GLU, int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap rmap, const TcParams p) {
  using Cfg = TcCfg<CIN, NPRE, MODE, GLU>;
  using Ops = OpTable<MODE>;
  constexpr bool UP = Cfg::UP;
  constexpr bool HEAD = Cfg::HEAD;
  constexpr int TC_PW = Cfg::PW;
  constexpr int TC_PLANE_BYTES = Cfg::PLANE_BYTES;
  constexpr int TC_NSTAGE = Cfg::NSTAGE;
  constexpr int RSTAGE = Cfg::RSTAGE;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* w_s = base;
  unsigned char* patch_s = w_s + Cfg::WB_PAD;
  unsigned char* res_s = patch_s + TC_NSTAGE * Cfg::PATCH_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(res_s + RSTAGE * Cfg::RES_BYTES);
  uint64_t* full = bars;                       // [NSTAGE <= 12]
  uint64_t* empty = bars + 12;                 // [NSTAGE <= 12]
  uint64_t* wbar = bars + 42;                  // [NOPS + 1 <= 12]: weight block of op o; [NOPS]: shift + ones blocks
  uint64_t* tfull = bars + 25;                 // [NACC <= 4]
  uint64_t* tempty = bars + 29;                // [NACC <= 4]
  uint64_t* rfull = bars + 33;                 // [RSTAGE <= 4]
  uint64_t* rempty = bars + 37;                // [RSTAGE <= 4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 41);
  static_assert(Ops::NOPS + 1 <= 12, "weight barriers");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // Programmatic dependent launch: let the next kernel of the stream start its own prologue (barrier init, TMEM
  // allocation, weight TMA) on every SM this grid has left, instead of after the whole grid has drained.  Everything
  // below that touches activations produced by earlier kernels sits behind griddepcontrol.wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int o = 0; o <= Ops::NOPS; ++o) mbar_init(&wbar[o], 1);
    for (int a = 0; a < Cfg::NACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
    for (int s = 0; s < RSTAGE; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rempty[s], 8); }
    fence_barrier_init();
    prefetch_tensormap(&tmap);
    if constexpr (Cfg::RES) prefetch_tensormap(&rmap);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // The whole warp runs the loop (uniform control flow); one elected lane issues.
    // Weights arrive block by block (shift + ones first, then the taps in issue order), each on its own barrier, so the
    // first tile's MMAs start after ~1/10 of the packed image has landed instead of all of it.
    if (elect_one()) {
      auto fetch = [&](uint64_t* bar, int off, int bytes) {
        mbar_arrive_expect_tx(bar, bytes);
        constexpr int CH = 32768;
        for (int o2 = 0; o2 < bytes; o2 += CH)
          bulk_g2s(w_s + off + o2, reinterpret_cast<const unsigned char*>(p.wpacked) + off + o2, bytes - o2 < CH ? bytes - o2 : CH, bar);
      };
      if constexpr (Cfg::BIAS) fetch(&wbar[Ops::NOPS], Cfg::WBYTES_TAPS, Cfg::BIAS_BYTES + Cfg::ONES_BYTES);
      int off = 0;
#pragma unroll
      for (int o = 0; o < Ops::NOPS; ++o) {
        const int bytes = CIN * Ops::op(o).n_mul * NPRE * 2;
        fetch(&wbar[o], off, bytes);
        off += bytes;
      }
    }
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");      // weights are not produced by the predecessor; patches are
    TileIter t; t.init(blockIdx.x, p);
    TileIter tr = t;                                         // skip tiles trail the patches by RLAG tiles
    const uint64_t pol_first = l2_policy_evict_first();
    constexpr int RLAG = 1;
    auto fetch_res = [&](int k) {
      const int rs = k % (RSTAGE > 0 ? RSTAGE : 1), rph = (k / (RSTAGE > 0 ? RSTAGE : 1)) & 1;
      mbar_wait(&rempty[rs], rph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&rfull[rs], Cfg::RES_BYTES);
        if (p.l2_in_first)
          tma_load_3d_hint(res_s + rs * Cfg::RES_BYTES, &rmap, &rfull[rs], tr.tx * Cfg::TWO * 8, tr.ty * Cfg::TH, tr.b * p.rP + p.rp0, pol_first);
        else
          tma_load_3d(res_s + rs * Cfg::RES_BYTES, &rmap, &rfull[rs], tr.tx * Cfg::TWO * 8, tr.ty * Cfg::TH, tr.b * p.rP + p.rp0);
      }
      __syncwarp();
      tr.next(p);
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it, t.next(p)) {
      const int s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1;
      const int y0 = t.ty * Cfg::TH, x0 = t.tx * Cfg::TWO;
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[s], Cfg::PATCH_BYTES);
        if (p.l2_in_first)
          tma_load_3d_hint(patch_s + s * Cfg::PATCH_BYTES, &tmap, &full[s], (x0 - Cfg::R) * 8, y0 - Cfg::R, t.b * p.xP + p.xp0, pol_first);
        else
          tma_load_3d(patch_s + s * Cfg::PATCH_BYTES, &tmap, &full[s], (x0 - Cfg::R) * 8, y0 - Cfg::R, t.b * p.xP + p.xp0);
      }
      __syncwarp();
      // The skip tile of tile k is only needed by the epilogue, after the MMAs of tile k: it is requested RLAG patches
      // later, so a full skip ring never holds back the patch stream that feeds the tensor core.
      if constexpr (Cfg::RES) { if (p.has_res && it >= RLAG) fetch_res(it - RLAG); }
    }
    if constexpr (Cfg::RES) {
      if (p.has_res) for (int k = it > RLAG ? it - RLAG : 0; k < it; ++k) fetch_res(k);
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    const uint32_t w_addr = smem_u32(w_s);
    // One tile's MMAs.  Descriptors are "per-tile base + compile-time constant": only the 14-bit start-address field of
    // the low word changes between MMAs.  FIRST (the CTA's first tile) also waits for the weight blocks, which land in
    // issue order; it is a separate instantiation so that the steady-state sequence is ONE basic block (with the waits
    // inline every op re-materialised its operands through R2UR and the issue thread spent ~58 cycles per MMA).
    auto issue_tile = [&](auto first_c, const uint32_t a_lo0, const uint32_t d_base) {
      constexpr bool FIRST = decltype(first_c)::value;
      constexpr uint32_t a_hi = (uint32_t)(Cfg::SBO_A >> 4) | (1u << 14);
      constexpr uint32_t b_hi = (uint32_t)(128 >> 4) | (1u << 14);
      const uint32_t b_lo0 = w_addr >> 4;
      if constexpr (Cfg::BIAS) {
        // accumulators start at the folded-BN shift: ones[128 x 16] (8-row groups 128 B apart, K chunks 2 KB apart)
        // times the shift block [NBIAS x 16]
        constexpr int NB = Cfg::NBIAS;
        if constexpr (FIRST) { mbar_wait(&wbar[Ops::NOPS], 0); tc_fence_after(); }
        const uint32_t one_lo = b_lo0 + (uint32_t)((Cfg::WBYTES_TAPS + Cfg::BIAS_BYTES) >> 4) + ((uint32_t)(2048 >> 4) << 16);
        const uint32_t bias_lo = b_lo0 + (uint32_t)(Cfg::WBYTES_TAPS >> 4) + ((uint32_t)((NB * 16) >> 4) << 16);
        umma_f16(d_base, make_desc64(one_lo, b_hi), make_desc64(bias_lo, b_hi),
                 make_idesc_f16(128, NB, Fmt16<FMT>::kUmmaFormat), 0u);
      }
#pragma unroll
      for (int j = 0; j < Cfg::MB; ++j) {         // heads: MB M-blocks of RB patch rows each per tile
        uint32_t b_off = 0;                       // compile-time after unrolling
#pragma unroll
        for (int o = 0; o < Ops::NOPS; ++o) {
          const TcOp op = Ops::op(o);
          const int N = op.n_mul * NPRE;
          const uint32_t idesc = make_idesc_f16(128, N, Fmt16<FMT>::kUmmaFormat);
          if constexpr (FIRST) { if (j == 0) { mbar_wait(&wbar[o], 0); tc_fence_after(); } }
#pragma unroll
          for (int k = 0; k < Cfg::KSTEPS; ++k) {
            const uint32_t a_lo = a_lo0 + (uint32_t)((((j * Cfg::RB + op.tap_r) * TC_PW + op.tap_c) * 16 + k * 2 * TC_PLANE_BYTES) >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)((b_off + k * (N * 32)) >> 4) + ((uint32_t)((N * 16) >> 4) << 16);
            umma_f16(d_base + op.dcol_mul * NPRE + j * 32, make_desc64(a_lo, a_hi), make_desc64(b_lo, b_hi), idesc,
                     (Cfg::BIAS || (o | k) != 0) ? 1u : 0u);
          }
          b_off += CIN * N * 2;
        }
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int s = it % TC_NSTAGE, ph = (it / TC_NSTAGE) & 1;
      const int acc = it % Cfg::NACC, aph = (it / Cfg::NACC) & 1;
      mbar_wait(&tempty[acc], aph ^ 1);            // epilogue has drained this accumulator stage
      mbar_wait(&full[s], ph);                     // patch landed
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo0 = (smem_u32(patch_s + s * Cfg::PATCH_BYTES) >> 4) | ((uint32_t)(TC_PLANE_BYTES >> 4) << 16);
        const uint32_t d_base = tmem_base + acc * Cfg::ACC_COLS;
        if (it == 0) issue_tile(std::true_type{}, a_lo0, d_base);
        else issue_tile(std::false_type{}, a_lo0, d_base);
        umma_commit(&empty[s]);                      // patch slot reusable once these MMAs retire
        umma_commit(&tfull[acc]);                    // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // ================================ epilogue (warps 2..9) ========================
    // Two warps per TMEM lane quadrant: `half` selects which half of the output channels (or which
    // sub-pixel phases / RGB outputs) this warp produces, so two warps per SM sub-partition work on
    // every tile.  A TMEM stage is handed back to the MMA issuer as soon as its values sit in registers.
    asm volatile("griddepcontrol.wait;" ::: "memory");      // skip tensor / addend reads and all output writes follow
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                // 0: warps 2..5, 1: warps 6..9
    const int m = quad * 32 + lane;                  // accumulator row = pixel within the tile
    const int ty = m >> 3, tx = m & 7;
    float adnext[Cfg::MB > 1 ? Cfg::MB / 2 : 1][3];  // image-head addend of the next tile
    const uint64_t pol_last = l2_policy_evict_last();
    TileIter t; t.init(blockIdx.x, p);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it, t.next(p)) {
      const int acc = it % Cfg::NACC, aph = (it / Cfg::NACC) & 1;
      const int b = t.b;
      const int gy = t.ty * TC_TH + ty, gx = t.tx * TC_TW + tx;
      const uint32_t t_base = tmem_base + acc * Cfg::ACC_COLS + ((uint32_t)(quad * 32) << 16);

      if constexpr (HEAD) {
        // Image head.  A tile is MB accumulator blocks (columns j*32..) of RB patch rows each; accumulator row
        // m = (y, x') = (m / HPW, m % HPW) of the HPW-wide patch; column o*K + kx holds the partial sum over (ky, c):
        //   out[y][x'][o] = sum_kx P[(y, x'+kx)][o*K + kx]  -> lane shuffles by kx inside the warp.
        // Warp group `half` produces blocks half*NQ .. half*NQ + NQ-1.
        constexpr int K = Cfg::KS;
        constexpr int NQ = Cfg::MB / 2;
        const int hy = m / Cfg::HPW, hx = m % Cfg::HPW;
        const int ox = t.tx * Cfg::TWO + hx;
        const size_t hw = (size_t)p.H * p.W;
        bool valid[NQ]; size_t oi[NQ]; float ad[NQ][3];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int oy = t.ty * Cfg::TH + (half * NQ + q) * Cfg::RB + hy;
          valid[q] = hx < Cfg::TWO && ox < p.W && oy < p.H;
          oi[q] = (size_t)b * 3 * hw + (size_t)oy * p.W + ox;
        }
        // the addend (SRb, fp32) of this tile was requested one tile ago: no DRAM round trip in the per-tile critical path
        auto load_addend = [&](const TileIter& tt, float (&dst)[NQ][3]) {
          const int xx = tt.tx * Cfg::TWO + hx;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int yy = tt.ty * Cfg::TH + (half * NQ + q) * Cfg::RB + hy;
            const bool ok = hx < Cfg::TWO && xx < p.W && yy < p.H;
            const size_t o0 = (size_t)tt.b * 3 * hw + (size_t)yy * p.W + xx;
#pragma unroll
            for (int o = 0; o < 3; ++o) dst[q][o] = ok ? __ldg(p.addend + o0 + o * hw) : 0.f;
          }
        };
        if (p.addend) {
          if (it == 0) load_addend(t, adnext);
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int o = 0; o < 3; ++o) ad[q][o] = adnext[q][o];
          if (tile + (int)gridDim.x < p.ntiles) { TileIter tn = t; tn.next(p); load_addend(tn, adnext); }
        } else {
#pragma unroll
          for (int q = 0; q < NQ; ++q) ad[q][0] = ad[q][1] = ad[q][2] = 0.f;
        }
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        uint32_t v[NQ][16];
#pragma unroll
        for (int q = 0; q < NQ; ++q) tmem_ld16(t_base + (half * NQ + q) * 32, v[q]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);               // TMEM stage is free again
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          float z[3];
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            z[o] = __uint_as_float(v[q][o * K]);
#pragma unroll
            for (int kx = 1; kx < K; ++kx) z[o] += __shfl_down_sync(0xffffffffu, __uint_as_float(v[q][o * K + kx]), kx);
          }
          if (valid[q]) {
            // inf / NaN anywhere upstream (an fp16 activation past 65504) reaches these sums; tanh would hide it
            if (p.nonfinite != nullptr && !(fabsf(z[0]) <= 3.0e38f && fabsf(z[1]) <= 3.0e38f && fabsf(z[2]) <= 3.0e38f))
              atomicOr(p.nonfinite, 1);
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              float tv = z[o];
              if (p.act_tanh) tv = tanhf(tv);
              __stcs(p.out_f32 + oi[q] + o * hw, fmaf(p.addend_scale, ad[q][o], tv * p.out_scale));
            }
          }
        }
      } else if constexpr (!UP) {
        constexpr int COUT = GLU ? NPRE / 2 : NPRE;            // channels written
        constexpr int CW = COUT / 2;                           // channels per epilogue warp (32 or 16)
        const int c0 = half * CW;
        const size_t pix = ((size_t)gy * p.W + gx) * 8;
        const size_t plane_elems = (size_t)p.H * p.W * 8;
        h16* obase = p.out + ((size_t)b * p.yP + p.yp0 + c0 / 8) * plane_elems + pix;
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        uint32_t va[CW], vg[GLU ? CW : 1];
        if constexpr (CW == 64) {
          tmem_ld32(t_base + c0, reinterpret_cast<uint32_t(&)[32]>(va[0]));
          tmem_ld32(t_base + c0 + 32, reinterpret_cast<uint32_t(&)[32]>(va[32]));
        } else if constexpr (CW == 32) {
          tmem_ld32(t_base + c0, va);
          if constexpr (GLU) tmem_ld32(t_base + COUT + c0, vg);
        } else {
          tmem_ld16(t_base + c0, va);
          if constexpr (GLU) tmem_ld16(t_base + COUT + c0, vg);
        }
        uint4 rres[CW / 8];
        bool has_res = false;
        if constexpr (Cfg::RES) {
          has_res = p.has_res != 0;
          if (has_res) {                                        // skip tile: [plane][16 x 8 pixels][8 ch], landed by TMA
            const int rs = it % RSTAGE, rph = (it / RSTAGE) & 1;
            mbar_wait(&rfull[rs], rph);
            const unsigned char* rp = res_s + rs * Cfg::RES_BYTES + (c0 / 8) * (TC_TH * TC_TW * 16) + m * 16;
#pragma unroll
            for (int j = 0; j < CW / 8; ++j) rres[j] = lds128(rp + j * (TC_TH * TC_TW * 16));
          }
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&tempty[acc]);                            // accumulators are in registers: stage free
          if constexpr (Cfg::RES) { if (has_res) mbar_arrive(&rempty[it % RSTAGE]); }
        }
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {                      // planes of 8 channels
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if constexpr (GLU) o[e] = glu_pair<FMT>(va[j * 8 + e], vg[j * 8 + e]);
            else o[e] = __uint_as_float(va[j * 8 + e]);
          }
          if constexpr (Cfg::RES) {
            if (has_res) {
              float res[8];
              unpack8_16<FMT>(rres[j], res);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] += res[e];
            }
          }
          if (p.l2_out_last) stg128_hint(obase + (size_t)j * plane_elems, pack8_16<FMT>(o), pol_last);
          else *reinterpret_cast<uint4*>(obase + (size_t)j * plane_elems) = pack8_16<FMT>(o);
        }
      } else {
        // 4-phase upsample: accumulator columns [ph*NPRE, (ph+1)*NPRE) hold output pixel
        // (2*gy + ph/2, 2*gx + ph%2); GLU halves inside each phase block.  half h -> phases 2h, 2h+1
        // (= output row 2*gy + h, both columns: one 32-byte run per plane).
        static_assert(!UP || NPRE == 64, "upsample path is conv(C -> 64) [+ GLU -> 32 channels]");
        const int Ho = 2 * p.H, Wo = 2 * p.W;
        const size_t plane_elems = (size_t)Ho * Wo * 8;
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        // both sub-pixel columns (phases 2*half, 2*half+1) of a plane are produced together: output pixels (oy, 2gx) and
        // (oy, 2gx+1) are 32 contiguous bytes -> one 256-bit store per plane, full 32-byte sectors on every request
        const int oy = 2 * gy + half, ox = 2 * gx;
        h16* obase = p.out + ((size_t)b * p.yP + p.yp0) * plane_elems + ((size_t)oy * Wo + ox) * 8;
        const uint32_t t0 = t_base + (half * 2) * NPRE, t1 = t0 + NPRE;
        if constexpr (GLU) {
          constexpr int COUT = NPRE / 2;                          // 32
          // software pipeline over the four output planes: the TMEM loads of plane j+1 are in flight while plane j is
          // evaluated and stored
          uint32_t a0[2][8], g0[2][8], a1[2][8], g1[2][8];
          tmem_ld8(t0, a0[0]); tmem_ld8(t0 + COUT, g0[0]); tmem_ld8(t1, a1[0]); tmem_ld8(t1 + COUT, g1[0]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cur = j & 1, nxt = cur ^ 1;
            tmem_ld_wait();
            if (j < 3) {
              tmem_ld8(t0 + (j + 1) * 8, a0[nxt]); tmem_ld8(t0 + COUT + (j + 1) * 8, g0[nxt]);
              tmem_ld8(t1 + (j + 1) * 8, a1[nxt]); tmem_ld8(t1 + COUT + (j + 1) * 8, g1[nxt]);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);           // last plane is in registers: stage free
            }
            float o0[8], o1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              o0[e] = glu_pair<FMT>(a0[cur][e], g0[cur][e]);
              o1[e] = glu_pair<FMT>(a1[cur][e], g1[cur][e]);
            }
            stg256(obase + (size_t)j * plane_elems, pack8_16<FMT>(o0), pack8_16<FMT>(o1));
          }
        } else {
          // act-free (training forward: the pre-BatchNorm conv output of an up-block): all NPRE channels of both phases
          uint32_t a0[2][8], a1[2][8];
          tmem_ld8(t0, a0[0]); tmem_ld8(t1, a1[0]);
#pragma unroll
          for (int j = 0; j < NPRE / 8; ++j) {
            const int cur = j & 1, nxt = cur ^ 1;
            tmem_ld_wait();
            if (j < NPRE / 8 - 1) {
              tmem_ld8(t0 + (j + 1) * 8, a0[nxt]); tmem_ld8(t1 + (j + 1) * 8, a1[nxt]);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);
            }
            float o0[8], o1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { o0[e] = __uint_as_float(a0[cur][e]); o1[e] = __uint_as_float(a1[cur][e]); }
            stg256(obase + (size_t)j * plane_elems, pack8_16<FMT>(o0), pack8_16<FMT>(o1));
          }
        }
      }
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
                                    int Cout, int Cin, float pre) {
  const int total = 9 * Cout * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % Cin, n = (i / Cin) % Cout, t = i / (Cin * Cout);
    const float v = w[((size_t)n * Cin + c) * 9 + t] * (scale ? scale[n] : 1.0f) * pre;   // BN scale folded, fp32
    const size_t off = (size_t)t * Cin * Cout + (size_t)(c / 8) * (Cout * 8) + (size_t)(n / 8) * 64 + (n % 8) * 8 + (c % 8);
    out[off] = Fmt16<FMT>::cvt(v);
  }
}

// Shift block [NB x 16] (row nn -> shift[nn % Cout], split into three 16-bit terms at k = 0,1,2) followed by the
// ones block [128 x 16] (k = 0,1,2 set): the operands of the per-tile bias MMA (see TcCfg::BIAS).
template <int FMT>
__global__ void pack_bias_kernel(const float* __restrict__ shift, h16* __restrict__ out, int Cout, int NB, float pre) {
  const int total = NB * 16 + 128 * 16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < NB * 16) {
      const int c = i % 16, nn = i / 16;
      float v = 0.f;
      if (c < 3 && shift != nullptr) {
        const float sft = shift[nn % Cout] * pre;
        const float hi = Fmt16<FMT>::cvtf(sft);
        const float mid = Fmt16<FMT>::cvtf(sft - hi);
        v = c == 0 ? hi : (c == 1 ? mid : Fmt16<FMT>::cvtf(sft - hi - mid));
      }
      out[(size_t)(c / 8) * (NB * 8) + (size_t)(nn / 8) * 64 + (nn % 8) * 8 + (c % 8)] = Fmt16<FMT>::cvt(v);
    } else {
      const int j = i - NB * 16, c = j % 16, r = j / 16;
      out[(size_t)NB * 16 + (size_t)(c / 8) * 1024 + (size_t)(r / 8) * 64 + (r % 8) * 8 + (c % 8)] =
          Fmt16<FMT>::cvt(c < 3 ? 1.0f : 0.0f);
    }
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
                                 int Cout, int Cin, float pre) {
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
      v *= pre;
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
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<CIN, NPRE, MODE, GLU>::SMEM);
  if (e != cudaSuccess) return cuda_fail(e);
  e = cudaFuncSetAttribute(conv3x3_tc_kernel<CIN, NPRE, MODE, GLU, TGSR_FMT_F16>,
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<CIN, NPRE, MODE, GLU>::SMEM);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

int init_conv_tc() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess) return cuda_fail(e);
  if (q != cudaDriverEntryPointSuccess || fn == nullptr) return TGSR_ERR_UNSUPPORTED;
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  int s;
  if ((s = set_tc_attr<64, 128, MODE_CONV, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<64, 64, MODE_CONV, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<64, 128, MODE_CONV, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 64, MODE_CONV, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 64, MODE_CONV, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 32, MODE_CONV, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<64, 64, MODE_UP, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 64, MODE_UP, true>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<64, 64, MODE_UP, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 64, MODE_UP, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 16, MODE_HEAD3, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 16, MODE_HEAD5, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 16, MODE_HEAD3W, false>()) != TGSR_OK) return s;
  if ((s = set_tc_attr<32, 16, MODE_HEAD5W, false>()) != TGSR_OK) return s;
  return TGSR_OK;
}

// Box [box_w pixels x 8 ch, box_h rows, npl planes] of a blocked view.  A pixel row of one plane is W*8 contiguous 16-bit
// elements, so (x, c%8) is described as ONE inner dimension: the box's inner run is box_w*16 bytes instead of 16 B, which is
// what the TMA engine's per-row cost wants; the bytes landing in shared memory are identical to a [8,box_w,box_h,planes] box.
static int make_box_tmap(CUtensorMap* m, const void* x, int B, int P, int H, int W, int npl, int fmt, int box_w, int box_h) {
  cuuint64_t gdim[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B * P};
  cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[3] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, (cuuint32_t)npl};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_tiled(m, fmt == TGSR_FMT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                              const_cast<void*>(x), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return TGSR_ERR_CUDA; }
  return TGSR_OK;
}

static int make_patch_tmap(CUtensorMap* m, const void* x, int B, int P, int H, int W, int npl, int fmt, int R = 1,
                           int head_pw = 0) {
  return make_box_tmap(m, x, B, P, H, W, npl, fmt, head_pw ? head_pw : TC_TW + 2 * R, (head_pw ? 32 : TC_TH) + 2 * R);
}

template <int CIN, int NPRE, int MODE, bool GLU>
static int launch_tc(const CUtensorMap& tmap, const CUtensorMap& rmap, TcParams& p, int fmt, cudaStream_t st) {
  static const int use_pdl = [] { const char* e = getenv("TGSR_PDL"); return e ? atoi(e) : 1; }();
  cudaLaunchConfig_t cfg = {};
  const int grid = p.ntiles < g_sm_count ? p.ntiles : g_sm_count;
  p.step_x = grid % p.tiles_x;
  p.step_y = (grid / p.tiles_x) % p.tiles_y;
  p.step_b = grid / (p.tiles_x * p.tiles_y);
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcCfg<CIN, NPRE, MODE, GLU>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
  cudaError_t e;
  if (fmt == TGSR_FMT_F16) e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<CIN, NPRE, MODE, GLU, TGSR_FMT_F16>, tmap, rmap, p);
  else e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<CIN, NPRE, MODE, GLU, TGSR_FMT_BF16>, tmap, rmap, p);
  if (e != cudaSuccess) return cuda_fail(e);
  return launch_status();
}

template <int CIN, int NPRE, int MODE, bool GLU>
static constexpr size_t packed_elems() { return TcCfg<CIN, NPRE, MODE, GLU>::WBYTES / 2; }

}  // namespace tgsr

#define TGSR_CHECK_FMT(fmt) TGSR_REQUIRE((fmt) == TGSR_FMT_BF16 || (fmt) == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED)

extern "C" size_t tgsr_packed_conv3x3_elems(int Cout, int Cin, int up2x) {
  if (Cout <= 0 || Cin <= 0) return 0;
  const size_t nb = (size_t)(up2x ? 4 : 1) * Cout;
  return (size_t)(up2x ? 16 : 9) * Cout * Cin + nb * 16 + 128 * 16;
}

extern "C" int tgsr_conv3x3_tc(const void* x, int xP, int xp0, const void* wpacked, const void* residual, int rP, int rp0,
                               void* y, int yP, int yp0, int B, int Cin, int H, int W, int Cout, int up2x, int act, int fmt,
                               int flags, tgsr_stream_t stream) {
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
               (reinterpret_cast<uintptr_t>(wpacked) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
               TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * xP < (1LL << 31) && (long long)B * (residual ? rP : 1) < (1LL << 31), TGSR_ERR_UNSUPPORTED);

  CUtensorMap tmap, rmap;
  int s = make_patch_tmap(&tmap, x, B, xP, H, W, npl, fmt);
  if (s != TGSR_OK) return s;
  if (residual) {
    s = make_box_tmap(&rmap, residual, B, rP, H, W, cout_planes, fmt, TC_TW, TC_TH);
    if (s != TGSR_OK) return s;
  } else {
    rmap = tmap;
  }
  TcParams p = {};
  p.wpacked = wpacked;
  p.out = reinterpret_cast<h16*>(y);
  p.has_res = residual != nullptr;
  p.reverse = (flags & TGSR_CONV_REVERSE) != 0;
  p.l2_in_first = (flags & TGSR_CONV_L2_IN_FIRST) != 0;
  p.l2_out_last = (flags & TGSR_CONV_L2_OUT_LAST) != 0;
  p.rP = rP; p.rp0 = rp0; p.yP = yP; p.yp0 = yp0; p.xP = xP; p.xp0 = xp0;
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = W / TC_TW; p.tiles_y = H / TC_TH; p.ntiles = B * p.tiles_x * p.tiles_y;
  p.out_f32 = nullptr; p.addend = nullptr; p.out_scale = 1.f; p.addend_scale = 0.f; p.act_tanh = 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool glu = act == TGSR_ACT_GLU;
  if (!up2x) {
    if (Cin == 64 && Cout == 128 && glu) return launch_tc<64, 128, MODE_CONV, true>(tmap, rmap, p, fmt, st);
    if (Cin == 64 && Cout == 64 && !glu) return launch_tc<64, 64, MODE_CONV, false>(tmap, rmap, p, fmt, st);
    if (Cin == 64 && Cout == 128 && !glu && !residual) return launch_tc<64, 128, MODE_CONV, false>(tmap, rmap, p, fmt, st);
    if (Cin == 32 && Cout == 64 && !glu && !residual) return launch_tc<32, 64, MODE_CONV, false>(tmap, rmap, p, fmt, st);
    if (Cin == 32 && Cout == 64 && glu) return launch_tc<32, 64, MODE_CONV, true>(tmap, rmap, p, fmt, st);
    if (Cin == 32 && Cout == 32 && !glu) return launch_tc<32, 32, MODE_CONV, false>(tmap, rmap, p, fmt, st);
  } else {
    if (Cin == 64 && Cout == 64 && glu) return launch_tc<64, 64, MODE_UP, true>(tmap, rmap, p, fmt, st);
    if (Cin == 32 && Cout == 64 && glu) return launch_tc<32, 64, MODE_UP, true>(tmap, rmap, p, fmt, st);
    if (Cin == 64 && Cout == 64 && !glu && !residual) return launch_tc<64, 64, MODE_UP, false>(tmap, rmap, p, fmt, st);
    if (Cin == 32 && Cout == 64 && !glu && !residual) return launch_tc<32, 64, MODE_UP, false>(tmap, rmap, p, fmt, st);
  }
  return TGSR_ERR_UNSUPPORTED;
}

extern "C" int tgsr_pack_conv3x3(const float* w, const float* scale, const float* shift, void* packed, int Cout, int Cin,
                                 int up2x, int act, int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(packed);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(Cout > 0 && Cin > 0 && Cout % 8 == 0 && Cin % 8 == 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_GLU, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  h16* o = reinterpret_cast<h16*>(packed);
  const float pre = glu_prescale(act == TGSR_ACT_GLU, fmt);
  const int nb = (up2x ? 4 : 1) * Cout;
  h16* ob = o + (size_t)(up2x ? 16 : 9) * Cout * Cin;
  if (fmt == TGSR_FMT_F16) {
    if (up2x) pack_up2x_kernel<TGSR_FMT_F16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin, pre);
    else pack_conv3x3_kernel<TGSR_FMT_F16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin, pre);
    pack_bias_kernel<TGSR_FMT_F16><<<16, 256, 0, st>>>(shift, ob, Cout, nb, pre);
  } else {
    if (up2x) pack_up2x_kernel<TGSR_FMT_BF16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin, pre);
    else pack_conv3x3_kernel<TGSR_FMT_BF16><<<64, 256, 0, st>>>(w, scale, o, Cout, Cin, pre);
    pack_bias_kernel<TGSR_FMT_BF16><<<16, 256, 0, st>>>(shift, ob, Cout, nb, pre);
  }
  return launch_status(2);
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
                                int ksize, int act, int fmt, int* nonfinite_flag, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(wpacked); TGSR_REQUIRE_PTR(y);
  TGSR_CHECK_FMT(fmt);
  TGSR_REQUIRE(g_sm_count > 0 && g_encode_tiled != nullptr, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(Cin == 32 && (ksize == 3 || ksize == 5) && W % 8 == 0, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(act == TGSR_ACT_NONE || act == TGSR_ACT_TANH, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(xp0 >= 0 && xp0 + Cin / 8 <= xP, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * xP < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  const int hpw = W >= 128 ? 32 : 16;                        // patch width: 32-pixel rows once the image is wide enough
  CUtensorMap tmap;
  int s = make_patch_tmap(&tmap, x, B, xP, H, W, Cin / 8, fmt, ksize / 2, hpw);
  if (s != TGSR_OK) return s;
  TcParams p = {};
  p.wpacked = wpacked; p.out = nullptr; p.has_res = 0;
  p.rP = p.rp0 = p.yP = p.yp0 = 0; p.xP = xP; p.xp0 = xp0;
  p.B = B; p.H = H; p.W = W;
  const int two = hpw - 2 * (ksize / 2);                     // output columns per head tile (32 rows x two cols)
  p.tiles_x = (W + two - 1) / two; p.tiles_y = (H + 31) / 32; p.ntiles = B * p.tiles_x * p.tiles_y;
  p.out_f32 = y; p.addend = addend; p.out_scale = out_scale; p.addend_scale = addend_scale;
  p.act_tanh = act == TGSR_ACT_TANH;
  p.nonfinite = nonfinite_flag;
  cudaStream_t st = (cudaStream_t)stream;
  if (hpw == 32) {
    if (ksize == 3) return launch_tc<32, 16, MODE_HEAD3W, false>(tmap, tmap, p, fmt, st);
    return launch_tc<32, 16, MODE_HEAD5W, false>(tmap, tmap, p, fmt, st);
  }
  if (ksize == 3) return launch_tc<32, 16, MODE_HEAD3, false>(tmap, tmap, p, fmt, st);
  return launch_tc<32, 16, MODE_HEAD5, false>(tmap, tmap, p, fmt, st);
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
