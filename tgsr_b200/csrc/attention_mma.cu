// Fused word-region attention on blocked 16-bit activations (GlobalAttention.py:87-130).
//
// Two kernels:
//   attn_project_kernel   S[b] = W . ctx[b]  (GlobalAttention.py:99-101), once per caption (it does not depend on
//                         the pixels), written as a per-lane tensor-core operand image ("sfrag").
//   attn_blocked_kernel   per pixel: logits L[q,t] = sum_i x[i,q] S[i,t], caption mask (with the reference's
//                         row-order quirk), softmax over words, weightedContext out[i,q] = sum_t S[i,t] P[q,t].
//
// attn_blocked_kernel is HBM-bound by design: x is read once (64 B/pixel), c_code (64 B/pixel) and the fp32
// attention map (4*T B/pixel) are written once.  To keep the arithmetic below the memory time the two small
// contractions run on the warp-level tensor-core path (mma.sync m16n8k16 / m16n8k8, fp32 accumulate); CUDA-core
// FMAs (2*32*T per pixel) would take about twice the memory time.
//
// Data movement: persistent CTAs (one per SM, 16 warps); every warp owns a TMA ring of 32-pixel tiles
// (cp.async.bulk, 4 planes x 512 B per tile, mbarrier completion).  In blocked planes [P][Q][8] a pixel's 8
// channels of one plane are 16 contiguous bytes, i.e. 8 consecutive pixels of a plane are exactly one 8x8
// 16-bit matrix with 16-byte rows: ldmatrix.x4 on (plane 2s, plane 2s+1) x (pixels 0-7, 8-15) returns the A
// fragment of K-step s (channels 16s..16s+15) in MMA register order, conflict-free and without any shuffling.
// The attention map leaves through a warp-private shared-memory tile so that every global store is a full
// 128-byte row segment.
//
// Accuracy: S is computed in fp32 and split into three 16-bit pieces (exact to 2^-24), x is exact in its
// storage format, products are exact and accumulate in fp32: logits and the fp32 attention maps keep fp32
// accuracy.  The second contraction runs in fp16 whatever the activation format (P lies in [0,1]; S rounded to
// 11 bits): 8x finer than the 16-bit rounding of its result.
#include "common.cuh"
#include "fmt16.cuh"
#include "ptx_sm100.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace tgsr {

typedef uint16_t h16;

constexpr int AM_IDF = 32;
// NT = word tiles of 8: 3 (T <= 24, the production captions), 4 (T <= 32) or 6 (T <= 48).  The operand words of a caption
// live in registers (16 * NT per lane), so the long-caption variants run 8 warps per CTA with a deeper ring instead of 16.
template <int NT>
struct AmCfg {
  static constexpr int TP = NT * 8;
  static constexpr int NL = 12 * NT;          // logits operand words per lane: 2 K-steps x NT x 3 pieces x (b0, b1)
  static constexpr int NFRAG = 16 * NT;       // + weighted-sum operand: 4 planes x NT words
  static constexpr int THREADS = NT == 3 ? 512 : 256;
  static constexpr int WARPS = THREADS / 32;
  static constexpr int DEPTH = NT == 3 ? 3 : 6;   // TMA ring depth per warp (pixel tiles in flight)
};
template <int NT, bool WIDE = (NT * 8 > 32)> struct AmMask { typedef uint32_t type; };
template <int NT> struct AmMask<NT, true> { typedef unsigned long long type; };
__host__ __device__ constexpr int am_nt_for(int T) { return T <= 24 ? 3 : (T <= 32 ? 4 : 6); }
constexpr int AM_PLANE_B = 512;       // 32 pixels * 16 B
constexpr int AM_TILE_B = 4 * AM_PLANE_B;
constexpr int AM_LDP = 36;            // padded row (32 pixels) of the per-warp attention-map staging tile
constexpr float kLog2e = 1.4426950408889634f;

template <int FMT> struct MmaOp;
template <> struct MmaOp<TGSR_FMT_BF16> {
  __device__ __forceinline__ static void k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ __forceinline__ static void k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(b0));
  }
};
template <> struct MmaOp<TGSR_FMT_F16> {
  __device__ __forceinline__ static void k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ __forceinline__ static void k8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(b0));
  }
};

// (v0, v1) -> N packed 16-bit pairs whose sum reproduces the inputs (piece k holds the residual of k-1).
template <int FMT, int N>
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t (&out)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const uint32_t p = Fmt16<FMT>::pack2(v0, v1);
    out[k] = p;
    if (k + 1 < N) {
      const float2 f = Fmt16<FMT>::unpack2(p);
      v0 -= f.x; v1 -= f.y;
    }
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// volatile: the operand words must stay in registers, not be re-loaded from global memory every tile
__device__ __forceinline__ uint4 ldg128_keep(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// sfrag word f of lane l of caption b
__host__ __device__ inline size_t sfrag_index(int b, int f, int lane, int nfrag) {
  return (((size_t)b * (nfrag / 4) + (f >> 2)) * 32 + lane) * 4 + (f & 3);
}

// =====================================================================================================
// Projection: grid (B, 4); CTA (b, r) computes rows 8r..8r+7 of S[b] (= plane r of the 32 channels): the cdf
// loop is split 16 ways over the 256 threads, operands come straight from global memory in one or two rounds
// of independent loads, partial sums meet in shared memory.  The CTA then writes every operand word that
// depends on its rows: register r%2 of the logits fragments of K-step r/2 and the weighted-sum fragments of
// plane r.
constexpr int AP_THREADS = 256;
constexpr int AP_ROWS = 8;
constexpr int AP_KS = 16;

template <int FMT, int NT>
__global__ void __launch_bounds__(AP_THREADS)
attn_project_kernel(const float* __restrict__ context, const float* __restrict__ w_ctx0, const float* __restrict__ w_ctx1,
                    const float* __restrict__ w_ctx2, uint32_t* __restrict__ sfrag0, uint32_t* __restrict__ sfrag1,
                    uint32_t* __restrict__ sfrag2, int cdf, int T) {
  constexpr int AM_NT = NT, AM_TP = AmCfg<NT>::TP, NL = AmCfg<NT>::NL, NFRAG = AmCfg<NT>::NFRAG;
  __shared__ float part[AP_KS][AP_ROWS][AM_TP];
  __shared__ float Srow[AP_ROWS][AM_TP + 1];
  const int tid = threadIdx.x, b = blockIdx.x, r = blockIdx.y;
  // blockIdx.z: which projection weight (the three generator stages share the captions: one launch for all of them)
  const float* __restrict__ w_ctx = blockIdx.z == 0 ? w_ctx0 : (blockIdx.z == 1 ? w_ctx1 : w_ctx2);
  uint32_t* __restrict__ sfrag = blockIdx.z == 0 ? sfrag0 : (blockIdx.z == 1 ? sfrag1 : sfrag2);
  {
    const int tq = tid & 7, ig = (tid >> 3) & 1, ks = tid >> 4;
    const int nquad = cdf >> 2;
    const int c_begin = 4 * ((ks * nquad) / AP_KS), c_end = 4 * (((ks + 1) * nquad) / AP_KS);
    float acc[4][AM_NT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < AM_NT; ++j) acc[i][j] = 0.f;
    int tt[AM_NT];
#pragma unroll
    for (int j = 0; j < AM_NT; ++j) tt[j] = min(tq + 8 * j, T - 1);      // clamped; padded words are zeroed below
    const float* wr = w_ctx + (size_t)(r * AP_ROWS + ig * 4) * cdf;
    const float* cb = context + (size_t)b * cdf * T;
#pragma unroll 2
    for (int c = c_begin; c < c_end; c += 4) {
      float4 w4[4];
      float cv[4][AM_NT];
#pragma unroll
      for (int i = 0; i < 4; ++i) w4[i] = __ldg(reinterpret_cast<const float4*>(wr + (size_t)i * cdf + c));
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int j = 0; j < AM_NT; ++j) cv[cc][j] = __ldg(cb + (c + cc) * T + tt[j]);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float w = cc == 0 ? w4[i].x : (cc == 1 ? w4[i].y : (cc == 2 ? w4[i].z : w4[i].w));
#pragma unroll
          for (int j = 0; j < AM_NT; ++j) acc[i][j] = fmaf(w, cv[cc][j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < AM_NT; ++j) part[ks][ig * 4 + i][tq + 8 * j] = acc[i][j];
  }
  __syncthreads();
  for (int idx = tid; idx < AP_ROWS * AM_TP; idx += AP_THREADS) {
    const int i = idx / AM_TP, t = idx - i * AM_TP;
    float v = 0.f;
    if (t < T) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
      for (int k = 0; k < AP_KS; k += 4) {
        v0 += part[k][i][t]; v1 += part[k + 1][i][t]; v2 += part[k + 2][i][t]; v3 += part[k + 3][i][t];
      }
      v = (v0 + v1) + (v2 + v3);
    }
    Srow[i][t] = v;
  }
  __syncthreads();
  // logits operand:  B[k][n] = log2(e) * S[16s + k][t = nt*8 + g]  (the log2(e) factor turns the softmax exponential
  // into a bare ex2).  Plane r = K-step s = r/2, register h = r%2 of the (b0,b1) pair; lane (g,tig) holds channels
  // 8r + 2tig + {0,1}; word f = (((s*NT + nt)*3 + piece)*2 + h).
  for (int u = tid; u < 32 * AM_NT; u += AP_THREADS) {
    const int lane = u & 31, nt = u >> 5, g = lane >> 2, tig = lane & 3, sk = r >> 1, h = r & 1;
    uint32_t p[3];
    split_pair<FMT, 3>(kLog2e * Srow[2 * tig][nt * 8 + g], kLog2e * Srow[2 * tig + 1][nt * 8 + g], p);
#pragma unroll
    for (int k = 0; k < 3; ++k) sfrag[sfrag_index(b, ((sk * AM_NT + nt) * 3 + k) * 2 + h, lane, NFRAG)] = p[k];
  }
  for (int u = tid; u < 32 * AM_NT; u += AP_THREADS) {
    // weighted-sum operand:  B[k = t][n = ch] = S[pl*8 + g][t]; word rr holds words 8rr + 2tig, +1 (k16 steps take the
    // words (2j, 2j+1) as (b0, b1); an odd last block is a k8 step)
    const int lane = u & 31, rr = u >> 5, g = lane >> 2, tig = lane & 3;
    // always fp16 (11-bit mantissa; |S| < 65504 assumed): P lies in [0,1], so this contraction is 8x more accurate
    // than the 16-bit rounding of its result whatever the activation format
    sfrag[sfrag_index(b, NL + r * AM_NT + rr, lane, NFRAG)] =
        Fmt16<TGSR_FMT_F16>::pack2(Srow[g][rr * 8 + 2 * tig], Srow[g][rr * 8 + 2 * tig + 1]);
  }
}

// =====================================================================================================
// smem: ptile [16 warps][AM_TP][AM_LDP] f32 | ring [16 warps][AM_DEPTH][AM_TILE_B] | bars [16][AM_DEPTH] u64 |
//       maskw [B] u32
template <int FMT, int NT>
__global__ void __launch_bounds__(AmCfg<NT>::THREADS, 1)
attn_blocked_kernel(const h16* __restrict__ xin, int xP, int xp0, const uint32_t* __restrict__ sfrag,
                    const uint8_t* __restrict__ mask, h16* __restrict__ yout, int yP, int yp0,
                    float* __restrict__ attn, int B, int Q, int T, int tiles_img, int n_tiles) {
  constexpr int AM_NT = NT, AM_TP = AmCfg<NT>::TP, NL = AmCfg<NT>::NL, AM_NFRAG = AmCfg<NT>::NFRAG;
  constexpr int AM_THREADS = AmCfg<NT>::THREADS, AM_WARPS = AmCfg<NT>::WARPS, AM_DEPTH = AmCfg<NT>::DEPTH;
  typedef typename AmMask<NT>::type MaskT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ptile = reinterpret_cast<float*>(smem_raw);
  unsigned char* ring = smem_raw + AM_WARPS * AM_TP * AM_LDP * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + AM_WARPS * AM_DEPTH * AM_TILE_B);
  MaskT* maskw = reinterpret_cast<MaskT*>(bars + AM_WARPS * AM_DEPTH);

  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler (uniform datapath)
  // this CTA's contiguous run of 32-pixel tiles (global index = b * tiles_img + tile-in-image)
  const int t_begin = (int)(((long long)blockIdx.x * n_tiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);

  // Programmatic dependent launch: the next kernel of the stream may start its prologue on SMs this grid has left.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- TMA ring: one elected lane requests the four 512-byte plane rows of a tile.  Each warp owns its barriers.
  uint64_t* my_bars = bars + warp * AM_DEPTH;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < AM_DEPTH; ++d) ptx::mbar_init(my_bars + d, 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();
  unsigned char* my_ring = ring + warp * (AM_DEPTH * AM_TILE_B);
  const size_t plane = (size_t)Q * 8;
  // (image, tile-in-image) of the next tile to request and of the next tile to consume, advanced without divisions
  int tile = t_begin + warp;
  int rb = min(tile, n_tiles - 1) / tiles_img, rt = tile - rb * tiles_img, rtile = tile;
  int b = rb, ti = rt;
  auto request = [&](int stage) {
    if (rtile < t_end && ptx::elect_one()) {
      const h16* src = xin + ((size_t)rb * xP + xp0) * plane + (size_t)rt * 256;
      ptx::mbar_arrive_expect_tx(my_bars + stage, 4 * 512);
#pragma unroll
      for (int pl = 0; pl < 4; ++pl)
        ptx::bulk_g2s(my_ring + stage * AM_TILE_B + pl * AM_PLANE_B, src + pl * plane, 512, my_bars + stage);
    }
    rtile += AM_WARPS; rt += AM_WARPS;
    while (rt >= tiles_img) { rt -= tiles_img; ++rb; }
  };

  // operand words and caption mask words come from kernels that finished long before the predecessor: they are read
  // while the predecessor (the conv that produces x) may still be draining on other SMs
  uint32_t fr[AM_NFRAG];                                         // operand words of the current caption
  auto load_operands = [&](int bb) {
    const uint4* sf = reinterpret_cast<const uint4*>(sfrag) + (size_t)bb * (AM_NFRAG / 4) * 32 + lane;
#pragma unroll
    for (int k = 0; k < AM_NFRAG / 4; ++k) {
      const uint4 v = ldg128_keep(sf + k * 32);
      fr[4 * k] = v.x; fr[4 * k + 1] = v.y; fr[4 * k + 2] = v.z; fr[4 * k + 3] = v.w;
    }
  };
  int cur_b = b;
  load_operands(cur_b);

  const MaskT pad_bits = T >= (int)(8 * sizeof(MaskT)) ? (MaskT)0 : (MaskT) ~(((MaskT)1 << T) - (MaskT)1);   // words t >= T are always masked
  if (mask) {
    for (int c = tid; c < B; c += AM_THREADS) {
      MaskT bits = pad_bits;
#pragma unroll 6
      for (int t = 0; t < T; ++t) bits |= (MaskT)(__ldg(mask + c * T + t) != 0) << t;
      maskw[c] = bits;
    }
    __syncthreads();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");             // x (and the output planes) belong to the predecessor until here
#pragma unroll
  for (int d = 0; d < AM_DEPTH; ++d) request(d);

  // ldmatrix row address of this lane: matrices 0..3 = (plane 2s, px 0-7), (plane 2s, px 8-15), (plane 2s+1, px 0-7),
  // (plane 2s+1, px 8-15); lanes 8j..8j+7 address the rows of matrix j
  const uint32_t frag_addr = ptx::smem_u32(my_ring) + (uint32_t)((lane >> 4) * AM_PLANE_B + (lane & 15) * 16);
  uint32_t cap = 0, cap_step = 0;                                // caption of pixel row g of the tile: (b*Q + q) % B
  if (mask) {
    cap = (uint32_t)(((unsigned long long)tile * 32u + (unsigned)g) % (unsigned)B);
    cap_step = (uint32_t)((AM_WARPS * 32) % B);
  }
  const MaskT my_pad = pad_bits >> (2 * tig);
  float* const pt_w = ptile + warp * (AM_TP * AM_LDP) + (2 * tig) * AM_LDP + g;        // + (nt*8+j)*AM_LDP + row
  const float* const pt_r = ptile + warp * (AM_TP * AM_LDP) + (lane >> 3) * AM_LDP + (lane & 7) * 4;
  const uint32_t row_stride4 = 4u * (uint32_t)Q;                 // elements between rows t and t + 4
  const uint32_t plane32 = (uint32_t)plane;
  const bool has_mask = mask != nullptr;

  int stage = 0;
  uint32_t phase = 0;
  for (; tile < t_end; tile += AM_WARPS) {
    const int q0 = ti * 32;
    if (b != cur_b) {                                            // warp-uniform; at most a few times per warp
      cur_b = b;
      load_operands(b);
    }
    ptx::mbar_wait(my_bars + stage, phase);
    uint32_t a[2][2][4];                                         // [K-step][16-row tile][MMA A registers]
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) ldmatrix_x4(a[s][mt], frag_addr + stage * AM_TILE_B + s * (2 * AM_PLANE_B) + mt * 256);
    __syncwarp();
    request(stage);                                              // refill the slot just drained
    if (++stage == AM_DEPTH) { stage = 0; phase ^= 1u; }

    // ---- logits  (GlobalAttention.py:107): two 16-row tiles, six independent accumulator chains
    float acc[2][AM_NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < AM_NT; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int sp = 0; sp < 3; ++sp)
#pragma unroll
        for (int nt = 0; nt < AM_NT; ++nt)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
            MmaOp<FMT>::k16(acc[mt][nt], a[s][mt], fr[((s * AM_NT + nt) * 3 + sp) * 2], fr[((s * AM_NT + nt) * 3 + sp) * 2 + 1]);
    // ---- caption mask (row r = b*Q+q uses caption r % B: GlobalAttention.py:109-116) + softmax over words
    MaskT w[4] = {my_pad, my_pad, my_pad, my_pad};               // rows g, g+8, g+16, g+24
    if (has_mask) {
      const uint32_t inc8 = 8u % (unsigned)B;
      uint32_t c = cap;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        w[r] = maskw[c] >> (2 * tig);
        c += inc8; c = c >= (unsigned)B ? c - (unsigned)B : c;
      }
      cap += cap_step; cap = cap >= (unsigned)B ? cap - (unsigned)B : cap;
    }
    float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < AM_NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (w[2 * mt] & ((MaskT)1 << (nt * 8 + j))) acc[mt][nt][j] = -CUDART_INF_F;
          if (w[2 * mt + 1] & ((MaskT)1 << (nt * 8 + j))) acc[mt][nt][2 + j] = -CUDART_INF_F;
          mx[2 * mt] = fmaxf(mx[2 * mt], acc[mt][nt][j]);
          mx[2 * mt + 1] = fmaxf(mx[2 * mt + 1], acc[mt][nt][2 + j]);
        }
#pragma unroll
    for (int r = 0; r < 4; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
#pragma unroll
    for (int r = 0; r < 4; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    float sm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < AM_NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          acc[mt][nt][j] = ex2_approx(acc[mt][nt][j] - mx[2 * mt]);            // ex2(-inf) = 0: masked words are exactly 0;
          acc[mt][nt][2 + j] = ex2_approx(acc[mt][nt][2 + j] - mx[2 * mt + 1]);  // an all-masked row gives NaN, as torch does
          sm[2 * mt] += acc[mt][nt][j]; sm[2 * mt + 1] += acc[mt][nt][2 + j];
        }
#pragma unroll
    for (int r = 0; r < 4; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], 1);
#pragma unroll
    for (int r = 0; r < 4; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], 2);
#pragma unroll
    for (int r = 0; r < 4; ++r) sm[r] = rcp_approx(sm[r]);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < AM_NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          acc[mt][nt][j] *= sm[2 * mt]; acc[mt][nt][2 + j] *= sm[2 * mt + 1];
          pt_w[(nt * 8 + j) * AM_LDP + mt * 16] = acc[mt][nt][j];
          pt_w[(nt * 8 + j) * AM_LDP + mt * 16 + 8] = acc[mt][nt][2 + j];
        }
    // ---- weightedContext = S . P^T  (GlobalAttention.py:126)
    float o[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) { o[mt][pl][0] = o[mt][pl][1] = o[mt][pl][2] = o[mt][pl][3] = 0.f; }
    {
      using H = Fmt16<TGSR_FMT_F16>;                             // P in [0,1] and S as fp16 (see attn_project_kernel)
#pragma unroll
      for (int j = 0; j < AM_NT / 2; ++j) {                      // k16 steps: words 16j .. 16j+15
        uint32_t pa[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          pa[mt][0] = H::pack2(acc[mt][2 * j][0], acc[mt][2 * j][1]); pa[mt][1] = H::pack2(acc[mt][2 * j][2], acc[mt][2 * j][3]);
          pa[mt][2] = H::pack2(acc[mt][2 * j + 1][0], acc[mt][2 * j + 1][1]); pa[mt][3] = H::pack2(acc[mt][2 * j + 1][2], acc[mt][2 * j + 1][3]);
        }
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
            MmaOp<TGSR_FMT_F16>::k16(o[mt][pl], pa[mt], fr[NL + pl * AM_NT + 2 * j], fr[NL + pl * AM_NT + 2 * j + 1]);
      }
      if (AM_NT & 1) {                                           // odd last block of 8 words: one k8 step
        uint32_t pb[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          pb[mt][0] = H::pack2(acc[mt][AM_NT - 1][0], acc[mt][AM_NT - 1][1]); pb[mt][1] = H::pack2(acc[mt][AM_NT - 1][2], acc[mt][AM_NT - 1][3]);
        }
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) MmaOp<TGSR_FMT_F16>::k8(o[mt][pl], pb[mt][0], pb[mt][1], fr[NL + pl * AM_NT + AM_NT - 1]);
      }
    }
    {
      h16* const yp_base = yout + ((size_t)b * yP + yp0) * plane + (size_t)(q0 + g) * 8 + 2 * tig;
      uint32_t yoff = 0;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) {
        uint32_t* yp = reinterpret_cast<uint32_t*>(yp_base + yoff);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          yp[mt * 64] = Fmt16<FMT>::pack2(o[mt][pl][0], o[mt][pl][1]);            // pixel rows g (+16): 16 B per pixel
          yp[mt * 64 + 32] = Fmt16<FMT>::pack2(o[mt][pl][2], o[mt][pl][3]);       // pixel rows g + 8 (+16)
        }
        yoff += plane32;
      }
    }
    // ---- attn[b,t,q]  (GlobalAttention.py:118-122): coalesced 128-byte row segments out of the staging tile
    __syncwarp();
    {
      float* const ap = attn + (size_t)b * T * Q + (size_t)(lane >> 3) * Q + q0 + (lane & 7) * 4;
      uint32_t aoff = 0;
#pragma unroll
      for (int k = 0; k < AM_TP / 4; ++k) {
        if ((lane >> 3) + 4 * k < T)
          __stcs(reinterpret_cast<float4*>(ap + aoff), *reinterpret_cast<const float4*>(pt_r + 4 * k * AM_LDP));
        aoff += row_stride4;
      }
    }
    __syncwarp();
    ti += AM_WARPS;
    while (ti >= tiles_img) { ti -= tiles_img; ++b; }
  }
}

template <int NT>
static size_t attn_smem_bytes(int B, bool has_mask) {
  using Cfg = AmCfg<NT>;
  return (size_t)Cfg::WARPS * Cfg::TP * AM_LDP * 4 + (size_t)Cfg::WARPS * Cfg::DEPTH * AM_TILE_B + (size_t)Cfg::WARPS * Cfg::DEPTH * 8 +
         (has_mask ? (size_t)B * sizeof(typename AmMask<NT>::type) : 0);
}

template <int NT>
static cudaError_t set_attn_attr() {
  cudaError_t e = cudaFuncSetAttribute(attn_blocked_kernel<TGSR_FMT_BF16, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attn_blocked_kernel<TGSR_FMT_F16, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

int init_attention_mma() {
  cudaError_t e = set_attn_attr<3>();
  if (e == cudaSuccess) e = set_attn_attr<4>();
  if (e == cudaSuccess) e = set_attn_attr<6>();
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

}  // namespace tgsr

extern "C" size_t tgsr_attn_sfrag_bytes(int B) { return (size_t)B * tgsr::AmCfg<3>::NFRAG * 32 * 4; }
/* operand image of B captions of T words (T <= 48) */
extern "C" size_t tgsr_attn_sfrag_bytes_t(int B, int T) { return (size_t)B * 16 * tgsr::am_nt_for(T) * 32 * 4; }

static int attn_project_launch(const float* context, const float* const w[3], void* const sf[3], int n, int B, int idf, int cdf,
                               int T, int fmt, cudaStream_t st) {
  using namespace tgsr;
  dim3 grid(B, AM_IDF / AP_ROWS, n);
  auto s0 = reinterpret_cast<uint32_t*>(sf[0]); auto s1 = reinterpret_cast<uint32_t*>(sf[1]); auto s2 = reinterpret_cast<uint32_t*>(sf[2]);
#define TGSR_AP_LAUNCH(F, N) attn_project_kernel<F, N><<<grid, AP_THREADS, 0, st>>>(context, w[0], w[1], w[2], s0, s1, s2, cdf, T)
#define TGSR_AP_FMT(N) do { if (fmt == TGSR_FMT_F16) TGSR_AP_LAUNCH(TGSR_FMT_F16, N); else TGSR_AP_LAUNCH(TGSR_FMT_BF16, N); } while (0)
  switch (am_nt_for(T)) {
    case 3: TGSR_AP_FMT(3); break;
    case 4: TGSR_AP_FMT(4); break;
    default: TGSR_AP_FMT(6); break;
  }
#undef TGSR_AP_FMT
#undef TGSR_AP_LAUNCH
  return launch_status();
}

extern "C" int tgsr_attn_project(const float* context, const float* w_ctx, void* sfrag, int B, int idf, int cdf, int T,
                                 int fmt, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE_PTR(context); TGSR_REQUIRE_PTR(w_ctx); TGSR_REQUIRE_PTR(sfrag);
  TGSR_REQUIRE(B > 0 && idf > 0 && cdf > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(idf == AM_IDF && T <= 48 && cdf % 4 == 0 && B <= 65535, TGSR_ERR_UNSUPPORTED);
  const float* w[3] = {w_ctx, w_ctx, w_ctx};
  void* sf[3] = {sfrag, sfrag, sfrag};
  return attn_project_launch(context, w, sf, 1, B, idf, cdf, T, fmt, (cudaStream_t)stream);
}

/* Up to three projections of the same captions (the generator stages' conv_context weights) in one launch. */
extern "C" int tgsr_attn_project_multi(const float* context, const float* w0, const float* w1, const float* w2, void* sfrag0,
                                       void* sfrag1, void* sfrag2, int n, int B, int idf, int cdf, int T, int fmt,
                                       tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(n >= 1 && n <= 3, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE_PTR(context); TGSR_REQUIRE_PTR(w0); TGSR_REQUIRE_PTR(sfrag0);
  if (n > 1) { TGSR_REQUIRE_PTR(w1); TGSR_REQUIRE_PTR(sfrag1); }
  if (n > 2) { TGSR_REQUIRE_PTR(w2); TGSR_REQUIRE_PTR(sfrag2); }
  TGSR_REQUIRE(B > 0 && idf > 0 && cdf > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(idf == AM_IDF && T <= 48 && cdf % 4 == 0 && B <= 65535, TGSR_ERR_UNSUPPORTED);
  const float* w[3] = {w0, n > 1 ? w1 : w0, n > 2 ? w2 : w0};
  void* sf[3] = {sfrag0, n > 1 ? sfrag1 : sfrag0, n > 2 ? sfrag2 : sfrag0};
  return attn_project_launch(context, w, sf, n, B, idf, cdf, T, fmt, (cudaStream_t)stream);
}

extern "C" int tgsr_attn_fwd_blocked(const void* x, int xP, int xp0, const void* sfrag, const uint8_t* mask, void* y,
                                     int yP, int yp0, float* attn, int B, int idf, int Q, int T, int fmt,
                                     tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE(fmt == TGSR_FMT_BF16 || fmt == TGSR_FMT_F16, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(sfrag); TGSR_REQUIRE_PTR(y); TGSR_REQUIRE_PTR(attn);
  TGSR_REQUIRE(B > 0 && idf > 0 && Q > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(g_sm_count > 0, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE(idf == AM_IDF && T <= 48 && B <= 4096 && Q % 32 == 0, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(xp0 >= 0 && xp0 + idf / 8 <= xP && yp0 >= 0 && yp0 + idf / 8 <= yP, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE((long long)B * Q < (1LL << 31) && (long long)48 * Q < (1LL << 30), TGSR_ERR_UNSUPPORTED);
  const int nt = am_nt_for(T);
  const size_t smem = nt == 3 ? attn_smem_bytes<3>(B, mask != nullptr) : (nt == 4 ? attn_smem_bytes<4>(B, mask != nullptr) : attn_smem_bytes<6>(B, mask != nullptr));
  TGSR_REQUIRE(smem <= 200 * 1024, TGSR_ERR_UNSUPPORTED);
  const int warps = nt == 3 ? AmCfg<3>::WARPS : AmCfg<4>::WARPS;
  const int tiles_img = Q / 32;
  const int n_tiles = B * tiles_img;
  // persistent CTAs, one per SM; every CTA owns an equal contiguous run of tiles (+-1)
  const int grid = n_tiles < warps * g_sm_count ? ceil_div(n_tiles, warps) : g_sm_count;
  cudaStream_t st = (cudaStream_t)stream;
  auto xi = reinterpret_cast<const h16*>(x);
  auto yo = reinterpret_cast<h16*>(y);
  auto sf = reinterpret_cast<const uint32_t*>(sfrag);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(warps * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  static const int use_pdl = [] { const char* e = getenv("TGSR_PDL"); return e ? atoi(e) : 1; }();
  cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
  cudaError_t e;
#define TGSR_AB_LAUNCH(F, N) e = cudaLaunchKernelEx(&cfg, attn_blocked_kernel<F, N>, xi, xP, xp0, sf, mask, yo, yP, yp0, attn, B, Q, T, tiles_img, n_tiles)
#define TGSR_AB_FMT(N) do { if (fmt == TGSR_FMT_F16) TGSR_AB_LAUNCH(TGSR_FMT_F16, N); else TGSR_AB_LAUNCH(TGSR_FMT_BF16, N); } while (0)
  switch (nt) {
    case 3: TGSR_AB_FMT(3); break;
    case 4: TGSR_AB_FMT(4); break;
    default: TGSR_AB_FMT(6); break;
  }
#undef TGSR_AB_FMT
#undef TGSR_AB_LAUNCH
  if (e != cudaSuccess) return cuda_fail(e);
  return launch_status();
}
