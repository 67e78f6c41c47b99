// TF32 tensor-core convolutions for the training step (fp32 NCHW tensors, fp32 accumulate).
//
// PyTorch's default for fp32 convolutions on Ampere+ is TF32 (torch.backends.cudnn.allow_tf32 = True), so the
// reference's GPU training already computes its convs at this precision; the CUDA-core fp32 kernels
// (conv_direct.cu, train_ops.cu) stay as the exact-precision path and for channel counts outside the envelope.
//
//   conv_tf32_kernel   y = conv_kxk(x [nearest-upsampled x2 if up2x], w), stride 1, same padding: implicit GEMM with
//                      M = 16x16 output pixels of one image, N = NB output channels, K = 8 input channels x taps.
//                      Used for the training forward and (with flipped/transposed weights) the data gradient.
//   wgrad_tf32_kernel  dw[co][ci][tap] = sum_px dy[co][px] x[ci][px+tap]: M = COB output channels, N = 32 input
//                      channels, K = the pixels of 16x16 tiles; the dy fragment of a K-step is reused by all taps.
// Both use mma.sync m16n8k8 (measured on B200: one HMMA per 2.1 cycles per SM = 275 TFLOP/s TF32).  Weights are rounded
// to TF32 once per step (cvt.rna); activations and gradients are staged (cp.async double buffer in the conv kernel) as raw fp32 bits -- the tensor core reads only the
// TF32 part of an operand (truncation), and cvt.rna is a ~6-instruction integer emulation on sm_100 that would double the
// instruction count of the staging loops.  Shared-memory strides make every fragment load bank-conflict free.
#include "common.cuh"

namespace tgsr {

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------- weight preparation
// wk[(ci*KK + t)*Cout + co] = tf32(w[co][ci][t])                       (forward)
// wk[(co*KK + t)*Cin + ci]  = tf32(w[co][ci][KK-1-t])                   (data gradient: channels swap, taps flip)
__global__ void conv_prep_tf32_kernel(const float* __restrict__ w, uint32_t* __restrict__ wk, int Cout, int Cin, int KK,
                                      int dgrad) {
  const int n = Cout * Cin * KK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i % KK, ci = (i / KK) % Cin, co = i / (KK * Cin);
    const uint32_t v = to_tf32(w[i]);
    if (dgrad) wk[((size_t)co * KK + (KK - 1 - t)) * Cin + ci] = v;
    else wk[((size_t)ci * KK + t) * Cout + co] = v;
  }
}

// ---------------------------------------------------------------- implicit-GEMM conv
// grid (tiles_x*tiles_y*B, N/NB); 256 threads: warp = (mw = warp&3: output rows 4mw..4mw+3, nw = warp>>2: NB/2 channels).
// smem: xs [8 ch][TH][RS] tf32 (PLANE % 32 == 8) | ws [KK][8][NB+8] tf32.
template <int K, int NB>
__global__ void __launch_bounds__(256, 2)
conv_tf32_kernel(const float* __restrict__ x, const uint32_t* __restrict__ wk, float* __restrict__ y, int Cin, int H, int W,
                 int N, int up2x) {
  constexpr int R = K / 2, KK = K * K, TH = 16 + K - 1, RS = (K == 3) ? 20 : 22, PLANE = TH * RS, LDW = NB + 8;
  constexpr int NT = NB / 16;                                   // n-tiles (8 channels) per warp
  static_assert(PLANE % 32 == 8 || PLANE % 32 == 24, "conflict-free A fragments");
  extern __shared__ __align__(16) uint32_t csm[];               // two buffers of { xs [8][PLANE] | ws [KK][8][LDW] }
  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const int tiles_x = ceil_div(Wo, 16), tiles_y = ceil_div(Ho, 16);
  const int b = blockIdx.x / (tiles_x * tiles_y), t = blockIdx.x % (tiles_x * tiles_y);
  const int ty0 = (t / tiles_x) * 16, tx0 = (t % tiles_x) * 16;
  const int n0 = blockIdx.y * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int mw = warp & 3, nw = warp >> 2;

  float acc[4][NT][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }

  // halo-tile positions owned by this thread (the same for every channel and every chunk): source offset inside a
  // channel plane, or -1 outside the image (zero padding)
  constexpr int TW = 16 + K - 1, NPOS = (TH * TW + 255) / 256;
  int src_off[NPOS], dst_off[NPOS];
#pragma unroll
  for (int k = 0; k < NPOS; ++k) {
    const int r = tid + 256 * k, yy = r / TW, xx = r - yy * TW;
    const int gy = ty0 + yy - R, gx = tx0 + xx - R;
    const bool in = r < TH * TW && gy >= 0 && gy < Ho && gx >= 0 && gx < Wo;
    src_off[k] = in ? (up2x ? (gy >> 1) * W + (gx >> 1) : gy * W + gx) : -1;
    dst_off[k] = r < TH * TW ? yy * RS + xx : -1;
  }
  const size_t chan = (size_t)H * W;
  const float* xb = x + (size_t)b * Cin * chan;

  // cp.async double buffer over 8-channel chunks: chunk c+1 streams global -> shared memory (no registers: 4-byte copies
  // with zero fill for the halo / padding, 16-byte copies for the weight rows) while the tensor cores work on chunk c.
  constexpr int XS_WORDS = 8 * PLANE, WS_WORDS = KK * 8 * LDW, BUF_WORDS = XS_WORDS + WS_WORDS;
  constexpr int NW16 = (KK * 8 * (NB / 4) + 255) / 256;          // 16-byte weight copies per thread
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(csm);
  auto issue_chunk = [&](int c0, int bufi) {
    const uint32_t xs_a = smem0 + (uint32_t)(bufi * BUF_WORDS) * 4u, ws_a = xs_a + XS_WORDS * 4u;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const bool cv = c0 + c < Cin;
      const float* xc = xb + (size_t)(cv ? c0 + c : 0) * chan;
#pragma unroll
      for (int k = 0; k < NPOS; ++k) {
        if (dst_off[k] >= 0) {
          const bool ok = cv && src_off[k] >= 0;
          const float* src = xc + (ok ? src_off[k] : 0);
          const int nbytes = ok ? 4 : 0;                           // 0 source bytes: the 4 destination bytes are zero-filled
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(xs_a + (uint32_t)(c * PLANE + dst_off[k]) * 4u), "l"(src),
                       "r"(nbytes) : "memory");
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NW16; ++k) {
      const int idx = tid + 256 * k;                              // (tap, channel, group of 4 output channels)
      if (idx < KK * 8 * (NB / 4)) {
        const int n4 = idx % (NB / 4), c = (idx / (NB / 4)) % 8, tp = idx / ((NB / 4) * 8);
        const bool ok = c0 + c < Cin && n0 + 4 * n4 < N;          // N % 32 == 0: a group is valid as a whole
        const uint32_t* src = wk + (ok ? ((size_t)(c0 + c) * KK + tp) * N + n0 + 4 * n4 : 0);
        const int nbytes = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ws_a + (uint32_t)((tp * 8 + c) * LDW + 4 * n4) * 4u), "l"(src),
                     "r"(nbytes) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_chunk(0, 0);
  int buf = 0;
  for (int c0 = 0; c0 < Cin; c0 += 8) {
    const bool more = c0 + 8 < Cin;
    if (more) {
      issue_chunk(c0 + 8, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                                              // chunk c0 has landed for every thread
    const uint32_t* xs = csm + buf * BUF_WORDS;
    const uint32_t* ws = xs + XS_WORDS;
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        uint32_t bf[NT][2];
        const uint32_t* wp = ws + ((ky * K + kx) * 8 + tig) * LDW + nw * (NB / 2) + g;
#pragma unroll
        for (int n = 0; n < NT; ++n) { bf[n][0] = wp[n * 8]; bf[n][1] = wp[4 * LDW + n * 8]; }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const uint32_t* xp = xs + tig * PLANE + (mw * 4 + m + ky) * RS + g + kx;
          const uint32_t a0 = xp[0], a1 = xp[8], a2 = xp[4 * PLANE], a3 = xp[4 * PLANE + 8];
#pragma unroll
          for (int n = 0; n < NT; ++n) mma_tf32_16x8x8(acc[m][n], a0, a1, a2, a3, bf[n][0], bf[n][1]);
        }
      }
    __syncthreads();                                              // all warps done with this buffer before it is refilled
    buf ^= 1;
  }
  // epilogue: D (row g -> pixel x = g / g+8, cols 2tig, 2tig+1 -> channel)
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int oy = ty0 + mw * 4 + m;
    if (oy >= Ho) continue;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int co = n0 + nw * (NB / 2) + n * 8 + 2 * tig;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ox = tx0 + g + 8 * h;
        if (ox >= Wo) continue;
        if (co < N) y[(((size_t)b * N + co) * Ho + oy) * Wo + ox] = acc[m][n][2 * h];
        if (co + 1 < N) y[(((size_t)b * N + co + 1) * Ho + oy) * Wo + ox] = acc[m][n][2 * h + 1];
      }
    }
  }
}

// ---------------------------------------------------------------- weight gradient
// grid (Cout/COB, Cin/32, splits); 256 threads: warp = (cw = warp&1: COB/2 output channels, nw = warp>>1: 8 input channels).
// smem: ds [COB][260] tf32 (16x16 tile of dy, plane stride 260 = 4 mod 32) | xs [32][XP] tf32 (halo tile, plane stride 4 mod 32).
template <int K, int COB>
__global__ void __launch_bounds__(256)
wgrad_tf32_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int B, int Cin, int H,
                  int W, int Cout, int up2x) {
  constexpr int R = K / 2, KK = K * K, TH = 16 + K - 1, XP = (K == 3) ? 18 * 18 : 20 * 20 + 4, DP = 260, MT = COB / 32;
  static_assert(K != 3 || (XP % 32 == 4 && DP % 32 == 4), "conflict-free fragments");
  extern __shared__ __align__(16) uint32_t csm[];
  uint32_t* ds = csm;
  uint32_t* xs = ds + COB * DP;
  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const int tiles_x = ceil_div(Wo, 16), tiles_y = ceil_div(Ho, 16);
  const int n_work = B * tiles_x * tiles_y;
  const int co0 = blockIdx.x * COB, ci0 = blockIdx.y * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int cw = warp & 1, nw = warp >> 1;
  float acc[MT][KK][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int k = 0; k < KK; ++k) { acc[m][k][0] = acc[m][k][1] = acc[m][k][2] = acc[m][k][3] = 0.f; }

  for (int wk = blockIdx.z; wk < n_work; wk += gridDim.z) {
    const int b = wk / (tiles_x * tiles_y), t = wk % (tiles_x * tiles_y);
    const int ty0 = (t / tiles_x) * 16, tx0 = (t % tiles_x) * 16;
    __syncthreads();
    // staged in batches of 16 independent loads followed by 16 stores (one global-load latency per batch)
    {
      const int yy = tid >> 4, xx = tid & 15, gy = ty0 + yy, gx = tx0 + xx;
      const bool in = gy < Ho && gx < Wo;
      const float* dp = dy + (((size_t)b * Cout + co0) * Ho + (in ? gy : 0)) * Wo + (in ? gx : 0);
      const size_t cs = (size_t)Ho * Wo;
#pragma unroll
      for (int c0 = 0; c0 < COB; c0 += 16) {
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __ldg(dp + (size_t)(c0 + c) * cs);
#pragma unroll
        for (int c = 0; c < 16; ++c) ds[(c0 + c) * DP + tid] = in ? __float_as_uint(v[c]) : 0u;
      }
    }
    {
      constexpr int NPX = (TH * TH + 255) / 256;
      int src[NPX], dst[NPX];
#pragma unroll
      for (int k = 0; k < NPX; ++k) {
        const int r = tid + 256 * k, yy = r / TH, xx = r - yy * TH;
        const int gy = ty0 + yy - R, gx = tx0 + xx - R;
        const bool in = r < TH * TH && gy >= 0 && gy < Ho && gx >= 0 && gx < Wo;
        src[k] = in ? (up2x ? (gy >> 1) * W + (gx >> 1) : gy * W + gx) : -1;
        dst[k] = r < TH * TH ? r : -1;
      }
      const float* xb = x + ((size_t)b * Cin + ci0) * ((size_t)H * W);
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 8) {
        float v[8][NPX];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int k = 0; k < NPX; ++k) v[c][k] = __ldg(xb + (size_t)(c0 + c) * H * W + (src[k] >= 0 ? src[k] : 0));
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int k = 0; k < NPX; ++k)
            if (dst[k] >= 0) xs[(c0 + c) * XP + dst[k]] = src[k] >= 0 ? __float_as_uint(v[c][k]) : 0u;
      }
    }
    __syncthreads();
    // K-steps: 8 consecutive pixels of one tile row
    for (int yy = 0; yy < 16; ++yy)
#pragma unroll
      for (int xh = 0; xh < 2; ++xh) {
        uint32_t a[MT][4];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint32_t* dp = ds + (cw * (COB / 2) + m * 16 + g) * DP + yy * 16 + xh * 8 + tig;
          a[m][0] = dp[0]; a[m][1] = dp[8 * DP]; a[m][2] = dp[4]; a[m][3] = dp[8 * DP + 4];
        }
        const uint32_t* xp = xs + (nw * 8 + g) * XP + yy * TH + xh * 8 + tig;
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const uint32_t b0 = xp[ky * TH + kx], b1 = xp[ky * TH + kx + 4];
#pragma unroll
            for (int m = 0; m < MT; ++m) mma_tf32_16x8x8(acc[m][ky * K + kx], a[m][0], a[m][1], a[m][2], a[m][3], b0, b1);
          }
      }
  }
  // D: rows g / g+8 -> output channel, cols 2tig, 2tig+1 -> input channel
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int k = 0; k < KK; ++k)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int co = co0 + cw * (COB / 2) + m * 16 + g + 8 * h;
        const int ci = ci0 + nw * 8 + 2 * tig;
        atomicAdd(dw + ((size_t)co * Cin + ci) * KK + k, acc[m][k][2 * h]);
        atomicAdd(dw + ((size_t)co * Cin + ci + 1) * KK + k, acc[m][k][2 * h + 1]);
      }
}

template <int K, int NB> static size_t conv_tf32_smem() {
  constexpr int TH = 16 + K - 1, RS = (K == 3) ? 20 : 22;
  return (size_t)2 * (8 * TH * RS + K * K * 8 * (NB + 8)) * 4;       // double-buffered
}
template <int K, int COB> static size_t wgrad_tf32_smem() {
  constexpr int XP = (K == 3) ? 18 * 18 : 20 * 20 + 4;
  return (size_t)(COB * 260 + 32 * XP) * 4;
}

int init_conv_tf32() {
  cudaError_t e;
#define TGSR_ATTR(fn, bytes) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); if (e != cudaSuccess) return cuda_fail(e)
  TGSR_ATTR((conv_tf32_kernel<3, 64>), (conv_tf32_smem<3, 64>()));
  TGSR_ATTR((conv_tf32_kernel<3, 32>), (conv_tf32_smem<3, 32>()));
  TGSR_ATTR((conv_tf32_kernel<5, 32>), (conv_tf32_smem<5, 32>()));
  TGSR_ATTR((wgrad_tf32_kernel<3, 64>), (wgrad_tf32_smem<3, 64>()));
  TGSR_ATTR((wgrad_tf32_kernel<3, 32>), (wgrad_tf32_smem<3, 32>()));
#undef TGSR_ATTR
  return TGSR_OK;
}

}  // namespace tgsr

using namespace tgsr;

/* 1 when tgsr_conv2d_tf32 / tgsr_conv2d_wgrad_tf32 cover the shape (else use the fp32 CUDA-core entry points). */
extern "C" int tgsr_conv_tf32_supported(int Cin, int Cout, int ksize, int wgrad) {
  if (wgrad) return ksize == 3 && Cout % 32 == 0 && Cin % 32 == 0;
  if (ksize == 3) return Cout % 32 == 0;
  return ksize == 5 && Cout % 32 == 0;
}

extern "C" int tgsr_conv_prep_tf32(const float* w, void* wk, int Cout, int Cin, int ksize, int dgrad, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(w); TGSR_REQUIRE_PTR(wk);
  TGSR_REQUIRE(Cout > 0 && Cin > 0 && (ksize == 3 || ksize == 5), TGSR_ERR_BAD_SHAPE);
  const int n = Cout * Cin * ksize * ksize;
  conv_prep_tf32_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<uint32_t*>(wk), Cout, Cin,
                                                                           ksize * ksize, dgrad);
  return launch_status();
}

/* y [B,N,Ho,Wo] = conv(x [B,Cin,H,W] (nearest x2 if up2x), wk) with wk from tgsr_conv_prep_tf32 (N = its output
 * channel count: Cout forward, the original Cin for the data gradient). */
extern "C" int tgsr_conv2d_tf32(const float* x, const void* wk, float* y, int B, int Cin, int H, int W, int N, int ksize,
                                int up2x, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(wk); TGSR_REQUIRE_PTR(y);
  TGSR_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && N > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(tgsr_conv_tf32_supported(Cin, N, ksize, 0), TGSR_ERR_UNSUPPORTED);
  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const long long tiles = (long long)B * ceil_div(Wo, 16) * ceil_div(Ho, 16);
  TGSR_REQUIRE(tiles < (1LL << 31), TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  auto w = reinterpret_cast<const uint32_t*>(wk);
  if (ksize == 3 && N % 64 == 0)
    conv_tf32_kernel<3, 64><<<dim3((unsigned)tiles, N / 64), 256, conv_tf32_smem<3, 64>(), st>>>(x, w, y, Cin, H, W, N, up2x);
  else if (ksize == 3)
    conv_tf32_kernel<3, 32><<<dim3((unsigned)tiles, N / 32), 256, conv_tf32_smem<3, 32>(), st>>>(x, w, y, Cin, H, W, N, up2x);
  else
    conv_tf32_kernel<5, 32><<<dim3((unsigned)tiles, N / 32), 256, conv_tf32_smem<5, 32>(), st>>>(x, w, y, Cin, H, W, N, up2x);
  return launch_status();
}

extern "C" int tgsr_conv2d_wgrad_tf32(const float* x, const float* dy, float* dw, int B, int Cin, int H, int W, int Cout,
                                      int ksize, int up2x, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(x); TGSR_REQUIRE_PTR(dy); TGSR_REQUIRE_PTR(dw);
  TGSR_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(tgsr_conv_tf32_supported(Cin, Cout, ksize, 1), TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cout * Cin * 9 * sizeof(float), st));
  const int Ho = up2x ? 2 * H : H, Wo = up2x ? 2 * W : W;
  const int n_work = B * ceil_div(Wo, 16) * ceil_div(Ho, 16);
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  const int COB = Cout % 64 == 0 ? 64 : 32;
  const int blocks_xy = (Cout / COB) * (Cin / 32);
  int splits = ceil_div(2 * sms, blocks_xy);
  splits = splits < 1 ? 1 : (splits > n_work ? n_work : splits);
  if (splits > 65535) splits = 65535;
  dim3 grid(Cout / COB, Cin / 32, splits);
  if (COB == 64) wgrad_tf32_kernel<3, 64><<<grid, 256, wgrad_tf32_smem<3, 64>(), st>>>(x, dy, dw, B, Cin, H, W, Cout, up2x);
  else wgrad_tf32_kernel<3, 32><<<grid, 256, wgrad_tf32_smem<3, 32>(), st>>>(x, dy, dw, B, Cin, H, W, Cout, up2x);
  return launch_status();
}
