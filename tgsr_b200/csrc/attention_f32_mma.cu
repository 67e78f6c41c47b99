// Fused word-region attention with the reference's fp32 NCHW I/O (GlobalAttention.py:87-130) on the warp-level
// tensor-core path, at fp32 accuracy.
//
// Same split as attention_mma.cu: attn_project_f32_kernel computes S[b] = W . ctx[b] once per caption (fp32);
// attn_f32_mma_kernel streams the pixels: logits, caption mask, softmax over words, weighted sum.
//
// fp32 accuracy on tensor cores: every operand is split into a TF32 "hi" part and a TF32 "lo" residual and each
// contraction issues hi*hi + lo*hi + hi*lo (m16n8k8, fp32 accumulate): products carry ~2^-21 relative error,
// inside the 1e-5 bar on the attention maps.  2*idf*T FMAs per pixel on CUDA cores cost several times the memory
// time of this kernel (measured: 5-30 % of HBM peak); on the tensor pipe they stay below it.
//
// Access pattern: lane (g = lane/4, tig = lane%4) owns the four consecutive pixels 4g..4g+3 of a 32-pixel tile.
// MMA rows are a free permutation of pixels and MMA K-slots a free permutation of channels / words, chosen so
// that every global access is a float4 and every warp instruction touches full 128-byte lines:
//   rows (mt, g), (mt, g+8)  <->  pixels 4g + 2mt, 4g + 2mt + 1
//   logits K-slots tig, tig+4 of step ks  <->  channels 8ks + tig, 8ks + tig + 4   (two float4 loads)
//   weighted-sum K-slots tig, tig+4 of step kt  <->  words 8kt + 2tig, 8kt + 2tig + 1  (= the logits C fragment)
// so x[ch][4g..4g+3], attn[t][4g..4g+3] and wctx[ch][4g..4g+3] all move as float4 straight from/to registers.
// x streams through a per-lane cp.async ring (each lane copies exactly the words it consumes itself).
#include "common.cuh"
#include "ptx_sm100.cuh"
#include <math_constants.h>

namespace tgsr {

constexpr int AF_THREADS = 512;
constexpr int AF_WARPS = AF_THREADS / 32;
constexpr int AF_DEPTH = 6;            // cp.async ring: K-steps (1 KB per warp each) in flight
constexpr float kLog2e_f = 1.4426950408889634f;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ float ex2_approx_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// =====================================================================================================
// S[b][i][t] = sum_c W[i][c] ctx[b][c][t]   ->  S fp32 [B][idf][TP], TP = 8*NT, zero for t >= T.
// grid (B, idf/8); the cdf loop is split 16 ways over 256 threads, partial sums meet in shared memory.
template <int NT>
__global__ void __launch_bounds__(256)
attn_project_f32_kernel(const float* __restrict__ context, const float* __restrict__ w_ctx, float* __restrict__ S,
                        int idf, int cdf, int T) {
  constexpr int TP = NT * 8, KS = 16;
  __shared__ float part[KS][8][TP];
  const int tid = threadIdx.x, b = blockIdx.x, r = blockIdx.y;
  const int tq = tid & 7, ig = (tid >> 3) & 1, ks = tid >> 4;
  const int nquad = cdf >> 2;
  const int c_begin = 4 * ((ks * nquad) / KS), c_end = 4 * (((ks + 1) * nquad) / KS);
  float acc[4][NT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  int tt[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) tt[j] = min(tq + 8 * j, T - 1);          // clamped; padded words are zeroed below
  const float* wr = w_ctx + (size_t)(r * 8 + ig * 4) * cdf;
  const float* cb = context + (size_t)b * cdf * T;
#pragma unroll 2
  for (int c = c_begin; c < c_end; c += 4) {
    float4 w4[4];
    float cv[4][NT];
#pragma unroll
    for (int i = 0; i < 4; ++i) w4[i] = __ldg(reinterpret_cast<const float4*>(wr + (size_t)i * cdf + c));
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int j = 0; j < NT; ++j) cv[cc][j] = __ldg(cb + (c + cc) * T + tt[j]);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float w = cc == 0 ? w4[i].x : (cc == 1 ? w4[i].y : (cc == 2 ? w4[i].z : w4[i].w));
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(w, cv[cc][j], acc[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) part[ks][ig * 4 + i][tq + 8 * j] = acc[i][j];
  __syncthreads();
  for (int idx = tid; idx < 8 * TP; idx += 256) {
    const int i = idx / TP, t = idx - i * TP;
    float v = 0.f;
    if (t < T) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
      for (int k = 0; k < KS; k += 4) {
        v0 += part[k][i][t]; v1 += part[k + 1][i][t]; v2 += part[k + 2][i][t]; v3 += part[k + 3][i][t];
      }
      v = (v0 + v1) + (v2 + v3);
    }
    S[((size_t)b * idf + r * 8 + i) * TP + t] = v;
  }
}

// =====================================================================================================
// smem: Sl [idf/8][NT][32 lanes] uint4 (logits B operand: b0 hi, b1 hi, b0 lo, b1 lo)
//     | So [NT][idf/8][32 lanes] uint4 (weighted-sum B operand, same packing)
//     | ring [16 warps][AF_DEPTH][32 lanes][2 x 16 B] | maskbits [B] u64
template <int NT>
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_f32_mma_kernel(const float* __restrict__ x, const float* __restrict__ S, const uint8_t* __restrict__ mask,
                    float* __restrict__ wctx, float* __restrict__ attn, int B, int idf, int Q, int T, int tiles_img,
                    int n_tiles) {
  constexpr int TP = NT * 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K8 = idf >> 3;
  uint4* Sl = reinterpret_cast<uint4*>(smem_raw);
  uint4* So = Sl + K8 * NT * 32;
  unsigned char* ring = reinterpret_cast<unsigned char*>(So + NT * K8 * 32);
  unsigned long long* maskbits = reinterpret_cast<unsigned long long*>(ring + AF_WARPS * AF_DEPTH * 1024);

  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int t_begin = (int)(((long long)blockIdx.x * n_tiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);

  const unsigned long long pad_bits = ~((1ull << T) - 1ull);      // T <= 48
  if (mask) {
    for (int c = tid; c < B; c += AF_THREADS) {
      unsigned long long bits = pad_bits;
      for (int t = 0; t < T; ++t) bits |= (unsigned long long)(__ldg(mask + c * T + t) != 0) << t;
      maskbits[c] = bits;
    }
  }
  const uint32_t ring_w = ptx::smem_u32(ring) + (uint32_t)(warp * AF_DEPTH * 1024 + lane * 32);
  const uint32_t sl_addr = ptx::smem_u32(Sl) + lane * 16, so_addr = ptx::smem_u32(So) + lane * 16;
  const size_t chQ = (size_t)Q;

  // the CTA walks its run of tiles image by image: the operand images in shared memory belong to one caption
  for (int seg_begin = t_begin; seg_begin < t_end;) {
    const int b = seg_begin / tiles_img;
    const int seg_end = min(t_end, (b + 1) * tiles_img);
    __syncthreads();                                              // previous segment's operands no longer in use
    // ---- operand images of caption b (TF32 hi/lo), built straight from S in global memory (L2-resident)
    const float* Sb = S + (size_t)b * idf * TP;
    for (int e = tid; e < K8 * NT * 32; e += AF_THREADS) {
      const int l = e & 31, nt = (e >> 5) % NT, ks = e / (32 * NT), gg = l >> 2, tg = l & 3;
      uint4 v;
      tf32_split(kLog2e_f * __ldg(Sb + (size_t)(8 * ks + tg) * TP + nt * 8 + gg), v.x, v.z);        // b0: k = tig
      tf32_split(kLog2e_f * __ldg(Sb + (size_t)(8 * ks + tg + 4) * TP + nt * 8 + gg), v.y, v.w);    // b1: k = tig + 4
      Sl[e] = v;
    }
    for (int e = tid; e < NT * K8 * 32; e += AF_THREADS) {
      const int l = e & 31, cn = (e >> 5) % K8, kt = e / (32 * K8), gg = l >> 2, tg = l & 3;
      uint4 v;
      tf32_split(__ldg(Sb + (size_t)(8 * cn + gg) * TP + kt * 8 + 2 * tg), v.x, v.z);               // slot tig   <-> word 2tig
      tf32_split(__ldg(Sb + (size_t)(8 * cn + gg) * TP + kt * 8 + 2 * tg + 1), v.y, v.w);           // slot tig+4 <-> word 2tig+1
      So[e] = v;
    }
    __syncthreads();

    const float* xb = x + (size_t)b * idf * chQ + (size_t)tig * chQ + 4 * g;       // + (8ks)*Q + q0 ; second word + 4Q
    float* const ob = wctx + (size_t)b * idf * chQ + (size_t)(2 * tig) * chQ + 4 * g;
    float* const ab = attn + (size_t)b * T * chQ + (size_t)(2 * tig) * chQ + 4 * g;
    const int seg0 = seg_begin - b * tiles_img + warp;              // this warp's first tile inside the image
    const int segN = seg_end - b * tiles_img;
    // linear stream of (tile, K-step) requests of this warp
    int rt = seg0, rk = 0, slot_w = 0;
    auto request = [&]() {
      if (rt < segN) {
        const float* src = xb + (size_t)(8 * rk) * chQ + (size_t)rt * 32;
        cp_async16(ring_w + slot_w * 1024, src);
        cp_async16(ring_w + slot_w * 1024 + 16, src + 4 * chQ);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (++rk == K8) { rk = 0; rt += AF_WARPS; }
      if (++slot_w == AF_DEPTH) slot_w = 0;
    };
#pragma unroll
    for (int d = 0; d < AF_DEPTH - 1; ++d) request();
    int slot_r = 0;

    for (int ti = seg0; ti < segN; ti += AF_WARPS) {
      const int q0 = ti * 32;
      float acc[2][NT][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }
      // ---- logits  (GlobalAttention.py:107)
      for (int ks = 0; ks < K8; ++ks) {
        request();
        asm volatile("cp.async.wait_group %0;" ::"n"(AF_DEPTH - 1) : "memory");
        const float4 v = lds128f(ring_w + slot_r * 1024), u = lds128f(ring_w + slot_r * 1024 + 16);
        if (++slot_r == AF_DEPTH) slot_r = 0;
        uint32_t ah[2][4], al[2][4];
        tf32_split(v.x, ah[0][0], al[0][0]); tf32_split(v.y, ah[0][1], al[0][1]);
        tf32_split(u.x, ah[0][2], al[0][2]); tf32_split(u.y, ah[0][3], al[0][3]);
        tf32_split(v.z, ah[1][0], al[1][0]); tf32_split(v.w, ah[1][1], al[1][1]);
        tf32_split(u.z, ah[1][2], al[1][2]); tf32_split(u.w, ah[1][3], al[1][3]);
        uint4 bw[NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) bw[nt] = lds128u(sl_addr + (ks * NT + nt) * 512);
        // hi*hi, lo*hi, hi*lo -- product outermost, so consecutive MMAs hit different accumulators (no dependent chain)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][nt], ah[mt], bw[nt].x, bw[nt].y);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][nt], al[mt], bw[nt].x, bw[nt].y);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_tf32(acc[mt][nt], ah[mt], bw[nt].z, bw[nt].w);
      }
      // ---- caption mask (row r = b*Q+q uses caption r % B: GlobalAttention.py:109-116) + softmax over words
      // rows: [0] = px 4g (mt 0, row g), [1] = px 4g+1 (mt 0, row g+8), [2] = px 4g+2 (mt 1, g), [3] = px 4g+3
      unsigned long long w[4] = {pad_bits, pad_bits, pad_bits, pad_bits};
      if (mask) {
        uint32_t c = ((unsigned)b * (unsigned)Q + (unsigned)(q0 + 4 * g)) % (unsigned)B;       // B*Q < 2^31
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          w[r] = maskbits[c];
          c = (c + 1u == (unsigned)B) ? 0u : c + 1u;
        }
      }
      uint32_t wl[4], wh[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const unsigned long long s = w[r] >> (2 * tig);
        wl[r] = (uint32_t)s; wh[r] = (uint32_t)(s >> 32);
      }
      float mx[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int bit = nt * 8 + j;
            const uint32_t m0 = bit < 32 ? (wl[2 * mt] >> bit) : (wh[2 * mt] >> (bit - 32));
            const uint32_t m1 = bit < 32 ? (wl[2 * mt + 1] >> bit) : (wh[2 * mt + 1] >> (bit - 32));
            if (m0 & 1u) acc[mt][nt][j] = -CUDART_INF_F;
            if (m1 & 1u) acc[mt][nt][2 + j] = -CUDART_INF_F;
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
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            acc[mt][nt][j] = ex2_approx_f(acc[mt][nt][j] - mx[2 * mt]);             // ex2(-inf) = 0: masked words exactly 0
            acc[mt][nt][2 + j] = ex2_approx_f(acc[mt][nt][2 + j] - mx[2 * mt + 1]);   // all-masked row: NaN, as torch
            sm[2 * mt] += acc[mt][nt][j]; sm[2 * mt + 1] += acc[mt][nt][2 + j];
          }
#pragma unroll
      for (int r = 0; r < 4; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], 1);
#pragma unroll
      for (int r = 0; r < 4; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], 2);
#pragma unroll
      for (int r = 0; r < 4; ++r) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sm[r]) : "f"(sm[r]));
      // ---- attn[b,t,q] (GlobalAttention.py:118-122): four consecutive pixels of word 8nt + 2tig + j per store
      {
        float* ap = ab + q0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            acc[0][nt][j] *= sm[0]; acc[0][nt][2 + j] *= sm[1]; acc[1][nt][j] *= sm[2]; acc[1][nt][2 + j] *= sm[3];
            if (nt * 8 + 2 * tig + j < T)
              __stcs(reinterpret_cast<float4*>(ap + (size_t)(nt * 8 + j) * chQ),
                     make_float4(acc[0][nt][j], acc[0][nt][2 + j], acc[1][nt][j], acc[1][nt][2 + j]));
          }
      }
      // ---- weightedContext = S . P^T (GlobalAttention.py:126), 32 channels at a time
      for (int c4 = 0; c4 < K8; c4 += 4) {
        float o[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int cn = 0; cn < 4; ++cn) { o[mt][cn][0] = o[mt][cn][1] = o[mt][cn][2] = o[mt][cn][3] = 0.f; }
#pragma unroll
        for (int kt = 0; kt < NT; ++kt) {
          uint32_t ph[2][4], pl[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {                       // A = {row g slot tig, row g+8 slot tig, row g slot tig+4, row g+8 slot tig+4}
            tf32_split(acc[mt][kt][0], ph[mt][0], pl[mt][0]); tf32_split(acc[mt][kt][2], ph[mt][1], pl[mt][1]);
            tf32_split(acc[mt][kt][1], ph[mt][2], pl[mt][2]); tf32_split(acc[mt][kt][3], ph[mt][3], pl[mt][3]);
          }
          uint4 bo4[4];
#pragma unroll
          for (int cn = 0; cn < 4; ++cn) bo4[cn] = (c4 + cn < K8) ? lds128u(so_addr + (kt * K8 + c4 + cn) * 512) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int cn = 0; cn < 4; ++cn)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_tf32(o[mt][cn], ph[mt], bo4[cn].x, bo4[cn].y);
#pragma unroll
          for (int cn = 0; cn < 4; ++cn)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_tf32(o[mt][cn], pl[mt], bo4[cn].x, bo4[cn].y);
#pragma unroll
          for (int cn = 0; cn < 4; ++cn)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_tf32(o[mt][cn], ph[mt], bo4[cn].z, bo4[cn].w);
        }
        float* op = ob + (size_t)(8 * c4) * chQ + q0;
#pragma unroll
        for (int cn = 0; cn < 4; ++cn)
          if (c4 + cn < K8) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              *reinterpret_cast<float4*>(op + (size_t)(8 * cn + j) * chQ) =
                  make_float4(o[0][cn][j], o[0][cn][2 + j], o[1][cn][j], o[1][cn][2 + j]);
          }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    seg_begin = seg_end;
  }
}

static size_t af_smem_bytes(int idf, int NT, int B, bool has_mask) {
  return (size_t)2 * (idf / 8) * NT * 512 + (size_t)AF_WARPS * AF_DEPTH * 1024 + (has_mask ? (size_t)B * 8 : 0);
}

template <int NT> static int af_set_attr() {
  cudaError_t e = cudaFuncSetAttribute(attn_f32_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

int init_attention_f32_mma() {
  int s;
  if ((s = af_set_attr<3>()) != TGSR_OK) return s;
  if ((s = af_set_attr<4>()) != TGSR_OK) return s;
  return af_set_attr<6>();
}

template <int NT>
static int af_launch(const float* x, const float* context, const float* w_ctx, const uint8_t* mask, float* S, float* wctx,
                     float* attn, int B, int idf, int cdf, int Q, int T, cudaStream_t st) {
  attn_project_f32_kernel<NT><<<dim3(B, idf / 8), 256, 0, st>>>(context, w_ctx, S, idf, cdf, T);
  int rc = launch_status();
  if (rc != TGSR_OK) return rc;
  const int tiles_img = Q / 32, n_tiles = B * tiles_img;
  const int grid = n_tiles < AF_WARPS * g_sm_count ? ceil_div(n_tiles, AF_WARPS) : g_sm_count;
  attn_f32_mma_kernel<NT><<<grid, AF_THREADS, af_smem_bytes(idf, NT, B, mask != nullptr), st>>>(
      x, S, mask, wctx, attn, B, idf, Q, T, tiles_img, n_tiles);
  return launch_status();
}

// envelope of the tensor-core path; outside it tgsr_attn_fwd_f32 keeps the CUDA-core kernel (attention.cu)
bool attn_f32_mma_supported(int B, int idf, int cdf, int Q, int T, bool has_mask) {
  if (idf % 8 != 0 || idf < 8 || idf > 256 || T > 48 || cdf % 4 != 0 || Q % 32 != 0 || B > 65535) return false;
  const int NT = T <= 24 ? 3 : (T <= 32 ? 4 : 6);
  return af_smem_bytes(idf, NT, B, has_mask) <= 227 * 1024;
}

size_t attn_f32_ws_floats(int B, int idf, int T) {
  const int NT = T <= 24 ? 3 : (T <= 32 ? 4 : 6);
  return (size_t)B * idf * NT * 8;
}

int attn_project_f32_launch(const float* context, const float* w_ctx, float* S, int B, int idf, int cdf, int T, cudaStream_t st) {
  dim3 grid(B, idf / 8);
  if (T <= 24) attn_project_f32_kernel<3><<<grid, 256, 0, st>>>(context, w_ctx, S, idf, cdf, T);
  else if (T <= 32) attn_project_f32_kernel<4><<<grid, 256, 0, st>>>(context, w_ctx, S, idf, cdf, T);
  else attn_project_f32_kernel<6><<<grid, 256, 0, st>>>(context, w_ctx, S, idf, cdf, T);
  return launch_status();
}

int attn_f32_mma_launch(const float* x, const float* context, const float* w_ctx, const uint8_t* mask, float* S, float* wctx,
                        float* attn, int B, int idf, int cdf, int Q, int T, cudaStream_t st) {
  if (T <= 24) return af_launch<3>(x, context, w_ctx, mask, S, wctx, attn, B, idf, cdf, Q, T, st);
  if (T <= 32) return af_launch<4>(x, context, w_ctx, mask, S, wctx, attn, B, idf, cdf, Q, T, st);
  return af_launch<6>(x, context, w_ctx, mask, S, wctx, attn, B, idf, cdf, Q, T, st);
}

}  // namespace tgsr

extern "C" size_t tgsr_attn_f32_ws_floats(int B, int idf, int T) { return tgsr::attn_f32_ws_floats(B, idf, T); }

extern "C" int tgsr_attn_f32_tc_supported(int B, int idf, int cdf, int Q, int T, int has_mask) {
  return tgsr::attn_f32_mma_supported(B, idf, cdf, Q, T, has_mask != 0) ? 1 : 0;
}

extern "C" int tgsr_attn_fwd_f32_tc(const float* input, const float* context, const float* w_ctx, const uint8_t* mask,
                                    float* ws, float* wctx, float* attn, int B, int idf, int cdf, int Q, int T,
                                    tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(input); TGSR_REQUIRE_PTR(context); TGSR_REQUIRE_PTR(w_ctx); TGSR_REQUIRE_PTR(ws);
  TGSR_REQUIRE_PTR(wctx); TGSR_REQUIRE_PTR(attn);
  TGSR_REQUIRE(B > 0 && idf > 0 && cdf > 0 && Q > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(g_sm_count > 0, TGSR_ERR_NOT_INIT);
  TGSR_REQUIRE((long long)B * Q < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(attn_f32_mma_supported(B, idf, cdf, Q, T, mask != nullptr), TGSR_ERR_UNSUPPORTED);
  return attn_f32_mma_launch(input, context, w_ctx, mask, ws, wctx, attn, B, idf, cdf, Q, T, (cudaStream_t)stream);
}
