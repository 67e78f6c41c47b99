// Image heads on the warp-level tensor-core path: GET_IMAGE_G[_noAct].img (conv3x3 32->3, util.py:894-919) and
// conv_output (conv5x5 32->3 + tanh, then one*y + a*SRb, model.py:223-226,280) from blocked 16-bit activations to the
// fp32 NCHW image.
//
// EXPERIMENTAL, opt-in (TGSR_HEAD_IMPL=mma); the default head is the tcgen05 kernel in conv_tc.cu.
// With three output channels the op is bandwidth work: 64 B of activations are read per pixel for 12 B written.  The
// tcgen05 head (N = 16) is bound by shared-memory operand fetch -- every K x K tap re-reads its 128-pixel A tile for 16
// accumulator columns -- and runs at ~55 % of the HBM roofline.  Here each A fragment is fetched ONCE per (input row,
// horizontal tap, 16-channel step) with ldmatrix and reused by the K vertical taps, each accumulating into a different
// output row's registers.  Measured (B200, batch 32, 256x256, L2-cold): 72.6 us (3x3) / 112 us (5x5) against 48.9 / ~62 us
// for the tcgen05 head, in three variants (2 CTAs per SM, persistent double-buffered with 8 and with 16 warps) that all
// land at ~2.5 TB/s: ncu shows 34 % issue utilisation with fixed-latency "wait" stalls leading and the epilogue's
// scattered 4-byte stores plus the copy-issue loop taking 2/3 of the samples (profiles/r1_head_mma_experiment.md).
//
// Persistent CTAs (one per SM) walk tiles of 16 x 64 output pixels; the (16+2R) x (64+2R) x 32-channel halo patch and the
// fp32 addend tile arrive with zero-filling 16-byte cp.async copies into a double buffer, so the copies of tile i+1 are in
// flight while tile i is computed and stored; 16 warps = 4 column strips of 16 pixels x 4 groups of 4 rows; per warp 4 rows x
// 4 accumulator registers.  Weight fragments (mma.sync B operand, 3 real
// columns of 8) are rebuilt per CTA from the tcgen05 head's packed weights, so both kernels share one prepared buffer.
#include "common.cuh"
#include "fmt16.cuh"

namespace tgsr {
namespace {

constexpr int HM_TH = 16, HM_TW = 64, HM_THREADS = 512;
constexpr int HM_ROWS = 4;                       // output rows per warp: 16 warps = 4 column strips x 4 row groups

template <int K>
struct HmCfg {
  static constexpr int R = K / 2;
  static constexpr int PH = HM_TH + 2 * R, PW = HM_TW + 2 * R;
  static constexpr int PLANE = PH * PW * 16;              // bytes of one 8-channel plane of the patch
  static constexpr int PATCH = 4 * PLANE;                 // 32 channels
  static constexpr int WFRAG = K * K * 2 * 32 * 8;        // [dy][dx][k-step][lane] x 2 registers
  static constexpr int ADDEND = 3 * HM_TH * HM_TW * 4;    // fp32 addend tile [3][16][64]
  static constexpr int SMEM = 2 * PATCH + WFRAG + 2 * ADDEND;   // double-buffered patch and addend tiles
};

struct HmParams {
  const uint16_t* x;        // blocked activations [B][xP][H][W][8]
  const uint16_t* wpacked;  // tgsr_pack_head layout
  const float* addend;      // [B,3,H,W] or null
  float* y;                 // [B,3,H,W]
  float out_scale, addend_scale;
  int xP, xp0, B, H, W, tiles_x, tiles_y, act_tanh;
  int debug;                // TGSR_TC_DEBUG stage skipping (1 MMA loop, 2 epilogue, 4 copies): timing experiments only
};

template <int FMT>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FMT == TGSR_FMT_F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Issue the cp.async copies of one tile: the halo patch (one patch row per warp per step, no divisions in the inner
// loop; out-of-image chunks are zero-filled) and, if present, the fp32 addend tile.
template <int K>
__device__ __forceinline__ void hm_issue_tile(const HmParams& p, int tile, unsigned char* patch, float* adt) {
  using Cfg = HmCfg<K>;
  constexpr int R = Cfg::R, PH = Cfg::PH, PW = Cfg::PW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int b = tile / tiles_per_img, rt = tile - b * tiles_per_img;
  const int ty = rt / p.tiles_x, tx = rt - ty * p.tiles_x;
  const int y0 = ty * HM_TH, x0 = tx * HM_TW, H = p.H, W = p.W;
  const uint32_t patch_a = (uint32_t)__cvta_generic_to_shared(patch);
  const uint16_t* xb = p.x + ((size_t)b * p.xP + p.xp0) * H * W * 8;
  const bool interior = y0 - R >= 0 && y0 + HM_TH + R <= H && x0 - R >= 0 && x0 + HM_TW + R <= W;   // CTA-uniform
  if (interior) {
    const uint16_t* x00 = xb + ((size_t)(y0 - R) * W + (x0 - R)) * 8;
    for (int rr = warp; rr < 4 * PH; rr += HM_THREADS / 32) {
      const int plane = rr / PH, row = rr - plane * PH;
      const uint16_t* srow = x00 + ((size_t)plane * H + row) * W * 8;
      const uint32_t drow = patch_a + (uint32_t)(rr * PW) * 16u;
#pragma unroll
      for (int c = lane; c < PW; c += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(drow + (uint32_t)c * 16u), "l"(srow + c * 8) : "memory");
    }
  } else {
    for (int rr = warp; rr < 4 * PH; rr += HM_THREADS / 32) {
      const int plane = rr / PH, row = rr - plane * PH;
      const int gy = y0 - R + row;
      const bool okrow = gy >= 0 && gy < H;
      const uint16_t* srow = xb + ((size_t)plane * H + (okrow ? gy : 0)) * W * 8;
      const uint32_t drow = patch_a + (uint32_t)(rr * PW) * 16u;
#pragma unroll
      for (int c = lane; c < PW; c += 32) {
        const int gx = x0 - R + c;
        const bool ok = okrow && gx >= 0 && gx < W;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + (uint32_t)c * 16u), "l"(srow + (ok ? gx : 0) * 8),
                     "r"(ok ? 16 : 0) : "memory");
      }
    }
  }
  if (p.addend != nullptr) {
    const uint32_t adt_a = (uint32_t)__cvta_generic_to_shared(adt);
    for (int i = tid; i < 3 * 256; i += HM_THREADS) {              // channel k: 16 rows x 16 chunks of 4 pixels = 256 chunks
      const int k = i >> 8, row = (i >> 4) & 15, c4 = i & 15;
      const int gy = y0 + row, gx = x0 + 4 * c4;
      const bool ok = gy < H && gx < W;
      const float* src = ok ? p.addend + (((size_t)b * 3 + k) * H + gy) * W + gx : p.addend;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(adt_a + (uint32_t)i * 16u), "l"(src), "r"(ok ? 16 : 0)
                   : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Persistent: one CTA per SM walks the tiles; the copies of tile i+1 are in flight while tile i is computed and stored.
template <int K, int FMT>
__global__ void __launch_bounds__(HM_THREADS, 1)
head_mma_kernel(const HmParams p) {
  using Cfg = HmCfg<K>;
  constexpr int R = Cfg::R, PH = Cfg::PH, PW = Cfg::PW;
  extern __shared__ __align__(128) unsigned char hm_smem[];
  uint2* wfrag = reinterpret_cast<uint2*>(hm_smem + 2 * Cfg::PATCH);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // B fragments: column n = lane / 4 is output channel n (n < 3), rows k = 2*(lane%4) + {0,1} (+8) are input channels
  for (int i = tid; i < K * K * 2 * 32; i += HM_THREADS) {
    const int ln = i & 31, ks = (i >> 5) & 1, tap = i >> 6, dy = tap / K, dx = tap - dy * K;
    const int n = ln >> 2, c0 = ks * 16 + 2 * (ln & 3);
    uint32_t r[2] = {0u, 0u};
    if (n < 3) {
      const int col = n * K + dx;                                   // column of the packed [16 x Cin] block of row tap dy
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + 8 * h;
        const size_t off = (size_t)dy * 32 * 16 + (size_t)(c >> 3) * 128 + (size_t)(col >> 3) * 64 + (col & 7) * 8 + (c & 7);
        r[h] = (uint32_t)p.wpacked[off] | ((uint32_t)p.wpacked[off + 1] << 16);   // channels c, c+1 are adjacent
      }
    }
    wfrag[i] = make_uint2(r[0], r[1]);
  }
  const int ntiles = p.B * p.tiles_x * p.tiles_y;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int H = p.H, W = p.W;
  const size_t hw = (size_t)H * W;
  const int strip = warp & 3, rgrp = warp >> 2;
  const int g = lane >> 2, q = lane & 3;
  // ldmatrix row address of this lane: matrices (px 0-7 | 8-15) x (plane 2ks | 2ks+1)
  const uint32_t lane_off = (uint32_t)(((lane >> 4) * PH * PW + (rgrp * HM_ROWS) * PW + strip * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * 16);

  asm volatile("griddepcontrol.wait;" ::: "memory");          // the activations (and addend) come from predecessors
  int tile = blockIdx.x;
  if (tile < ntiles && !(p.debug & 4)) hm_issue_tile<K>(p, tile, hm_smem, reinterpret_cast<float*>(hm_smem + 2 * Cfg::PATCH + Cfg::WFRAG));
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    unsigned char* patch = hm_smem + buf * Cfg::PATCH;
    const float* adt = reinterpret_cast<const float*>(hm_smem + 2 * Cfg::PATCH + Cfg::WFRAG + buf * Cfg::ADDEND);
    const int next = tile + gridDim.x;
    if (next < ntiles && !(p.debug & 4)) {
      hm_issue_tile<K>(p, next, hm_smem + (buf ^ 1) * Cfg::PATCH,
                       reinterpret_cast<float*>(hm_smem + 2 * Cfg::PATCH + Cfg::WFRAG + (buf ^ 1) * Cfg::ADDEND));
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    float acc[HM_ROWS][4];
#pragma unroll
    for (int o = 0; o < HM_ROWS; ++o) { acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f; }
    const uint32_t patch_a = (uint32_t)__cvta_generic_to_shared(patch) + lane_off;
#pragma unroll 1
    for (int t2 = 0; t2 < ((p.debug & 1) ? 0 : K * 2); ++t2) {
      const int dx = t2 >> 1, ks = t2 & 1;
      uint2 bf[K];
#pragma unroll
      for (int dy = 0; dy < K; ++dy) bf[dy] = wfrag[((dy * K + dx) * 2 + ks) * 32 + lane];
      const uint32_t base = patch_a + (uint32_t)((ks * 2 * PH * PW + dx) * 16);
#pragma unroll
      for (int ir = 0; ir < HM_ROWS + 2 * R; ++ir) {
        uint32_t a[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base + (uint32_t)(ir * PW * 16)));
#pragma unroll
        for (int dy = 0; dy < K; ++dy) {
          const int orow = ir - dy;
          if (orow >= 0 && orow < HM_ROWS) mma16816<FMT>(acc[orow], a, bf[dy].x, bf[dy].y);
        }
      }
    }

    // epilogue: accumulator registers c0,c1 = pixel g, channels 2q,2q+1; c2,c3 = pixel g+8
    if (q < 2 && !(p.debug & 2)) {
      const int b = tile / tiles_per_img, rt = tile - b * tiles_per_img;
      const int ty = rt / p.tiles_x, tx = rt - ty * p.tiles_x;
      const int y0 = ty * HM_TH, x0 = tx * HM_TW;
      const int nch = q == 0 ? 2 : 1;                               // q=0: channels 0,1; q=1: channel 2
      const bool has_add = p.addend != nullptr;
      const int lx = strip * 16 + g, ly = rgrp * HM_ROWS;
      float* yb = p.y + ((size_t)b * 3 + 2 * q) * hw + (size_t)(y0 + ly) * W + (x0 + lx);
      const float* ab = adt + ((2 * q) * HM_TH + ly) * HM_TW + lx;
      const bool full = y0 + HM_TH <= H && x0 + HM_TW <= W;         // CTA-uniform: no per-store bounds checks inside
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j >= nch) continue;
        float* yj = yb + (size_t)j * hw;
#pragma unroll
        for (int o = 0; o < HM_ROWS; ++o) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (!full && (y0 + ly + o >= H || x0 + lx + 8 * h >= W)) continue;
            float t = acc[o][2 * h + j];
            if (p.act_tanh) t = tanhf(t);
            const float ad = has_add ? ab[(j * HM_TH + o) * HM_TW + 8 * h] : 0.f;
            __stcs(yj + (o * W + 8 * h), fmaf(p.addend_scale, ad, t * p.out_scale));
          }
        }
      }
    }
    __syncthreads();                                            // this buffer is refilled by the next iteration's copies
  }
}

template <int K, int FMT>
int launch_head_mma(const HmParams& p, cudaStream_t st) {
  static const int use_pdl = [] { const char* e = getenv("TGSR_PDL"); return e ? atoi(e) : 1; }();
  cudaLaunchConfig_t cfg = {};
  const int ntiles = p.B * p.tiles_x * p.tiles_y;
  cfg.gridDim = dim3((unsigned)(ntiles < g_sm_count ? ntiles : g_sm_count));
  cfg.blockDim = dim3(HM_THREADS);
  cfg.dynamicSmemBytes = HmCfg<K>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, head_mma_kernel<K, FMT>, p);
  if (e != cudaSuccess) return cuda_fail(e);
  return launch_status();
}

}  // namespace

int init_head_mma(void) {
  TGSR_CUDA(cudaFuncSetAttribute(head_mma_kernel<3, TGSR_FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HmCfg<3>::SMEM));
  TGSR_CUDA(cudaFuncSetAttribute(head_mma_kernel<3, TGSR_FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HmCfg<3>::SMEM));
  TGSR_CUDA(cudaFuncSetAttribute(head_mma_kernel<5, TGSR_FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HmCfg<5>::SMEM));
  TGSR_CUDA(cudaFuncSetAttribute(head_mma_kernel<5, TGSR_FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HmCfg<5>::SMEM));
  return TGSR_OK;
}

// Called by tgsr_img_head_tc (conv_tc.cu) after argument validation.
int img_head_mma(const void* x, int xP, int xp0, const void* wpacked, const float* addend, float out_scale, float addend_scale,
                 float* y, int B, int H, int W, int ksize, int act_tanh, int fmt, cudaStream_t st) {
  HmParams p;
  p.x = reinterpret_cast<const uint16_t*>(x); p.wpacked = reinterpret_cast<const uint16_t*>(wpacked);
  p.addend = addend; p.y = y; p.out_scale = out_scale; p.addend_scale = addend_scale;
  p.xP = xP; p.xp0 = xp0; p.B = B; p.H = H; p.W = W;
  p.tiles_x = ceil_div(W, HM_TW); p.tiles_y = ceil_div(H, HM_TH); p.act_tanh = act_tanh;
  static const int dbg = [] { const char* e = getenv("TGSR_TC_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = dbg;
  if ((long long)B * p.tiles_x * p.tiles_y >= (1LL << 31)) return TGSR_ERR_UNSUPPORTED;
  if (ksize == 3) return fmt == TGSR_FMT_F16 ? launch_head_mma<3, TGSR_FMT_F16>(p, st) : launch_head_mma<3, TGSR_FMT_BF16>(p, st);
  return fmt == TGSR_FMT_F16 ? launch_head_mma<5, TGSR_FMT_F16>(p, st) : launch_head_mma<5, TGSR_FMT_BF16>(p, st);
}

}  // namespace tgsr
