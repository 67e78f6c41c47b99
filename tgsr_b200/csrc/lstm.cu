// RNN_ENCODER forward: Embedding -> 1-layer bidirectional LSTM (eval), fp32.
// reference: util.py:233-260 (forward), util.py:222-231 (zero initial state).
//
// Packed-sequence semantics are handled in-kernel from device-resident lengths (the reference
// syncs the host with cap_lens.tolist(), util.py:239): the forward direction runs t=0..len-1,
// the reverse direction t=len-1..0, outputs beyond len are zero, sent_emb is the final hidden
// state of each direction.
//
// Three launches: (1) transpose W_hh so the recurrence reads it coalesced, (2) the input
// projection for every token (one GEMM-like pass, W_ih read once per 8 tokens), (3) the
// recurrence: one CTA per (direction, group of LSTM_BT captions), thread j owns gate row j.
#include "common.cuh"

namespace tgsr {

constexpr int LSTM_TOK = 8;  // tokens per CTA in the input projection
constexpr int LSTM_BT = 8;   // captions per CTA in the recurrence

__global__ void lstm_transpose_whh(const float* __restrict__ w_hh, float* __restrict__ w_hhT, int H) {
  // w_hh [2][4H][H] -> w_hhT [2][H][4H]
  const int n = 2 * 4 * H * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int d = i / (4 * H * H), r = i % (4 * H * H), j = r / H, k = r % H;
    w_hhT[((size_t)d * H + k) * 4 * H + j] = w_hh[i];
  }
}

// xproj[tok][d][j] = emb[cap[tok]] . w_ih[d][j] + b_ih[d][j] + b_hh[d][j]
__global__ void __launch_bounds__(256)
lstm_xproj_kernel(const int64_t* __restrict__ captions, const float* __restrict__ emb,
                  const float* __restrict__ w_ih, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                  float* __restrict__ xproj, int ntok, int ntoken, int ninput, int H) {
  extern __shared__ float xs[];  // [LSTM_TOK][ninput]
  const int tok0 = blockIdx.x * LSTM_TOK;
  for (int i = threadIdx.x; i < LSTM_TOK * ninput; i += blockDim.x) {
    const int tk = i / ninput, k = i % ninput;
    float v = 0.f;
    if (tok0 + tk < ntok) {
      long long id = captions ? captions[tok0 + tk] : (long long)(tok0 + tk);
      if (id < 0) id = 0;
      if (id >= ntoken) id = ntoken - 1;   // nn.Embedding would raise; clamp keeps the kernel memory-safe
      v = __ldg(emb + (size_t)id * ninput + k);
    }
    xs[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarp = blockDim.x / 32;
  const int rows = 2 * 4 * H;
  for (int r = warp; r < rows; r += nwarp) {
    const float* wr = w_ih + (size_t)r * ninput;
    float acc[LSTM_TOK];
#pragma unroll
    for (int tk = 0; tk < LSTM_TOK; ++tk) acc[tk] = 0.f;
    for (int k = lane; k < ninput; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int tk = 0; tk < LSTM_TOK; ++tk) acc[tk] = fmaf(wv, xs[tk * ninput + k], acc[tk]);
    }
#pragma unroll
    for (int tk = 0; tk < LSTM_TOK; ++tk)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[tk] += __shfl_xor_sync(0xffffffffu, acc[tk], o);
    if (lane == 0) {
      const float bsum = b_ih[r] + b_hh[r];
#pragma unroll
      for (int tk = 0; tk < LSTM_TOK; ++tk)
        if (tok0 + tk < ntok) xproj[(size_t)(tok0 + tk) * rows + r] = acc[tk] + bsum;
    }
  }
}

// grid (2, ceil(B/LSTM_BT)); block 4H threads.
__global__ void __launch_bounds__(512)
lstm_recurrence_kernel(const int64_t* __restrict__ cap_lens, const float* __restrict__ xproj,
                       const float* __restrict__ w_hhT, float* __restrict__ words, float* __restrict__ sent,
                       int B, int Tcap, int Tmax, int H) {
  extern __shared__ float sm[];
  float* h_s = sm;                       // [LSTM_BT][H]
  float* g_s = sm + LSTM_BT * H;         // [LSTM_BT][4H]
  __shared__ int len_s[LSTM_BT];
  const int d = blockIdx.x, b0 = blockIdx.y * LSTM_BT;
  const int j = threadIdx.x, G = 4 * H;
  if (j < LSTM_BT) {
    int L = 0;
    if (b0 + j < B) {
      long long v = cap_lens[b0 + j];
      L = (int)(v < 0 ? 0 : (v > Tcap ? Tcap : v));
    }
    len_s[j] = L;
  }
  for (int i = j; i < LSTM_BT * H; i += blockDim.x) h_s[i] = 0.f;
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int bb = 0; bb < LSTM_BT; ++bb) maxlen = max(maxlen, len_s[bb]);

  // cell state: thread (bb, u) for tid < LSTM_BT*H keeps c in a register; with 4H threads and
  // LSTM_BT*H cells each thread owns LSTM_BT/4 cells.
  constexpr int CPT = LSTM_BT / 4;
  float c_reg[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) c_reg[i] = 0.f;

  const float* wT = w_hhT + (size_t)d * H * G;
  for (int s = 0; s < maxlen; ++s) {
    float acc[LSTM_BT];
#pragma unroll
    for (int bb = 0; bb < LSTM_BT; ++bb) {
      const int L = len_s[bb];
      const int t = d == 0 ? s : L - 1 - s;
      acc[bb] = (s < L) ? __ldg(xproj + ((size_t)(b0 + bb) * Tcap + t) * 2 * G + (size_t)d * G + j) : 0.f;
    }
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      const float wv = __ldg(wT + (size_t)k * G + j);
#pragma unroll
      for (int bb = 0; bb < LSTM_BT; ++bb) acc[bb] = fmaf(wv, h_s[bb * H + k], acc[bb]);
    }
#pragma unroll
    for (int bb = 0; bb < LSTM_BT; ++bb) g_s[bb * G + j] = acc[bb];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int cell = j + i * G;            // 0 .. LSTM_BT*H-1
      const int bb = cell / H, u = cell % H;
      const int L = len_s[bb];
      if (s < L) {
        const float* g = g_s + bb * G;
        const float ig = 1.0f / (1.0f + expf(-g[u]));
        const float fg = 1.0f / (1.0f + expf(-g[H + u]));
        const float gg = tanhf(g[2 * H + u]);
        const float og = 1.0f / (1.0f + expf(-g[3 * H + u]));
        const float c = fmaf(fg, c_reg[i], ig * gg);
        c_reg[i] = c;
        const float h = og * tanhf(c);
        h_s[bb * H + u] = h;
        const int t = d == 0 ? s : L - 1 - s;
        if (t < Tmax) words[((size_t)(b0 + bb) * 2 * H + d * H + u) * Tmax + t] = h;
        if (s == L - 1) sent[(size_t)(b0 + bb) * 2 * H + d * H + u] = h;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Persistent cluster recurrence (H = 128).  A cluster of 2 CTAs owns LC_BT captions of one
// direction for all time steps.  CTA `rank` owns hidden units [64*rank, 64*rank+64): thread r holds
// gate row (g = r/64, u = 64*rank + r%64) of W_hh ENTIRELY IN REGISTERS (128 floats), so the
// recurrent GEMV reads no weights from memory after the prologue.  The input projection is a
// gather from the precomputed table xtab[token][dir][4H] = E.W_ih^T + b_ih + b_hh; the values of
// step s+1 are fetched into registers while step s computes, so the serial loop never waits for
// global memory.  Every thread applies its own gate's nonlinearity (the five transcendental
// evaluations of a cell run in parallel on four threads instead of in sequence on one).  Each step
// the new h of a CTA's 64 units is written to both CTAs' shared memory (DSMEM) and one cluster
// barrier publishes it.  Two captions per cluster: the step time is latency (barrier, DSMEM,
// transcendental chains), so more, smaller clusters finish sooner than fewer, wider ones.
constexpr int LC_BT = 2;
constexpr int LC_H = 128;

__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void st_remote_f32(const float* local_smem_ptr, uint32_t target_rank, float v) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local_smem_ptr);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(target_rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
lstm_cluster_kernel(const int64_t* __restrict__ captions, const int64_t* __restrict__ cap_lens,
                    const float* __restrict__ xtab, const float* __restrict__ w_hh, float* __restrict__ words,
                    float* __restrict__ sent, int B, int Tcap, int Tmax, int ntoken) {
  constexpr int H = LC_H, G = 4 * H;
  static_assert(LC_BT == 2, "the hidden state is read as float2 (two captions per unit)");
  __shared__ __align__(16) float h_s[2][H][LC_BT];      // double-buffered hidden state, [unit][caption]
  __shared__ float g_s[4][64][LC_BT];                   // this CTA's activated gates
  __shared__ int len_s[LC_BT];
  __shared__ int tok_s[LC_BT * 64];                     // token ids of this cluster's captions (Tcap <= 64)
  const uint32_t rank = cluster_rank();
  const int d = blockIdx.y, b0 = blockIdx.z * LC_BT;
  const int r = threadIdx.x, g = r >> 6, ul = r & 63, u = (int)rank * 64 + ul, row = g * H + u;

  if (r < LC_BT) {
    int L = 0;
    if (b0 + r < B) {
      const long long v = cap_lens[b0 + r];
      L = (int)(v < 0 ? 0 : (v > Tcap ? Tcap : v));
    }
    len_s[r] = L;
  }
  for (int i = r; i < H * LC_BT; i += 256) (&h_s[0][0][0])[i] = 0.f;
  for (int i = r; i < LC_BT * Tcap; i += 256) {
    const int bb = i / Tcap, t = i % Tcap;
    long long id = (b0 + bb < B) ? captions[(size_t)(b0 + bb) * Tcap + t] : 0;
    tok_s[i] = (int)(id < 0 ? 0 : (id >= ntoken ? ntoken - 1 : id));
  }
  float w[H];                                           // W_hh[d][row][:] in registers
  {
    const float4* wr = reinterpret_cast<const float4*>(w_hh + ((size_t)d * G + row) * H);
#pragma unroll
    for (int k4 = 0; k4 < H / 4; ++k4) {
      const float4 v = __ldg(wr + k4);
      w[4 * k4] = v.x; w[4 * k4 + 1] = v.y; w[4 * k4 + 2] = v.z; w[4 * k4 + 3] = v.w;
    }
  }
  __syncthreads();
  int len[LC_BT];
  int maxlen = 0;
#pragma unroll
  for (int bb = 0; bb < LC_BT; ++bb) { len[bb] = len_s[bb]; maxlen = max(maxlen, len[bb]); }
  // input projection of this thread's gate row at step s, one value per caption (zero once a caption has ended)
  auto fetch_x = [&](int s, float (&x)[LC_BT]) {
#pragma unroll
    for (int bb = 0; bb < LC_BT; ++bb) {
      x[bb] = 0.f;
      if (s < len[bb]) {
        const int t = d == 0 ? s : len[bb] - 1 - s;
        x[bb] = __ldg(xtab + ((size_t)tok_s[bb * Tcap + t] * 2 + d) * G + row);
      }
    }
  };
  float xc[LC_BT], xn[LC_BT];
  fetch_x(0, xc);
  cluster_sync_all();                                   // both CTAs initialised before any remote store

  float c_reg = 0.f;                                    // cell (unit cu, caption cb) lives in thread cb*64 + cu
  const int cb = r >> 6, cu = r & 63, cunit = (int)rank * 64 + cu;
  const bool cell_thread = r < 64 * LC_BT;
  const int cL = cell_thread ? len_s[cb] : 0;
  for (int s = 0; s < maxlen; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    fetch_x(s + 1, xn);                                 // lands while the recurrent product below runs
    float acc[LC_BT];
#pragma unroll
    for (int bb = 0; bb < LC_BT; ++bb) acc[bb] = xc[bb];
    const float2* h2 = reinterpret_cast<const float2*>(&h_s[cur][0][0]);
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const float2 hv = h2[k];                          // broadcast: h[k] of the 2 captions
      acc[0] = fmaf(w[k], hv.x, acc[0]); acc[1] = fmaf(w[k], hv.y, acc[1]);
    }
#pragma unroll
    for (int bb = 0; bb < LC_BT; ++bb)                  // gate order i, f, g, o: g is tanh, the rest sigmoid
      g_s[g][ul][bb] = g == 2 ? tanhf(acc[bb]) : 1.0f / (1.0f + expf(-acc[bb]));
    __syncthreads();
    if (cell_thread && s < cL) {
      c_reg = fmaf(g_s[1][cu][cb], c_reg, g_s[0][cu][cb] * g_s[2][cu][cb]);
      const float h = g_s[3][cu][cb] * tanhf(c_reg);
      h_s[nxt][cunit][cb] = h;
      st_remote_f32(&h_s[nxt][cunit][cb], rank ^ 1u, h);
    }
    cluster_arrive();
    if (cell_thread && s < cL) {                        // outputs leave while the barrier completes
      const float h = h_s[nxt][cunit][cb];
      const int t = d == 0 ? s : cL - 1 - s;
      if (t < Tmax) words[((size_t)(b0 + cb) * 2 * H + d * H + cunit) * Tmax + t] = h;
      if (s == cL - 1) sent[(size_t)(b0 + cb) * 2 * H + d * H + cunit] = h;
    }
#pragma unroll
    for (int bb = 0; bb < LC_BT; ++bb) xc[bb] = xn[bb];
    cluster_wait();
  }
}

int init_lstm() {
  return TGSR_OK;
}

}  // namespace tgsr

extern "C" int tgsr_lstm_prepare_f32(const float* emb, const float* w_ih, const float* b_ih, const float* b_hh,
                                     float* xtab, int ntoken, int ninput, int H, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(emb); TGSR_REQUIRE_PTR(w_ih); TGSR_REQUIRE_PTR(b_ih); TGSR_REQUIRE_PTR(b_hh); TGSR_REQUIRE_PTR(xtab);
  TGSR_REQUIRE(ntoken > 0 && ninput > 0 && H > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(ninput <= 1024, TGSR_ERR_UNSUPPORTED);
  lstm_xproj_kernel<<<ceil_div(ntoken, LSTM_TOK), 256, LSTM_TOK * ninput * sizeof(float), (cudaStream_t)stream>>>(
      nullptr, emb, w_ih, b_ih, b_hh, xtab, ntoken, ntoken, ninput, H);
  return launch_status();
}

extern "C" int tgsr_lstm_bidir_fwd_tab_f32(const int64_t* captions, const int64_t* cap_lens, const float* xtab,
                                           const float* w_hh, float* words, float* sent, int B, int Tcap, int Tmax,
                                           int ntoken, int H, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(captions); TGSR_REQUIRE_PTR(cap_lens); TGSR_REQUIRE_PTR(xtab); TGSR_REQUIRE_PTR(w_hh);
  TGSR_REQUIRE_PTR(words); TGSR_REQUIRE_PTR(sent);
  TGSR_REQUIRE(B > 0 && Tcap > 0 && Tmax > 0 && Tmax <= Tcap && ntoken > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(H == LC_H && B <= 65535 * LC_BT && Tcap <= 48, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  TGSR_CUDA(cudaMemsetAsync(words, 0, (size_t)B * 2 * H * Tmax * sizeof(float), st));
  TGSR_CUDA(cudaMemsetAsync(sent, 0, (size_t)B * 2 * H * sizeof(float), st));
  lstm_cluster_kernel<<<dim3(2, 2, ceil_div(B, LC_BT)), 256, 0, st>>>(captions, cap_lens, xtab, w_hh, words, sent,
                                                                             B, Tcap, Tmax, ntoken);
  return launch_status();
}

extern "C" int tgsr_lstm_bidir_fwd_f32(const int64_t* captions, const int64_t* cap_lens, const float* emb,
                                       const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                       float* ws, float* words, float* sent, int B, int Tcap, int Tmax, int ntoken,
                                       int ninput, int H, tgsr_stream_t stream) {
  using namespace tgsr;
  TGSR_REQUIRE_PTR(captions); TGSR_REQUIRE_PTR(cap_lens); TGSR_REQUIRE_PTR(emb); TGSR_REQUIRE_PTR(w_ih);
  TGSR_REQUIRE_PTR(w_hh); TGSR_REQUIRE_PTR(b_ih); TGSR_REQUIRE_PTR(b_hh); TGSR_REQUIRE_PTR(ws);
  TGSR_REQUIRE_PTR(words); TGSR_REQUIRE_PTR(sent);
  TGSR_REQUIRE(B > 0 && Tcap > 0 && Tmax > 0 && Tmax <= Tcap && ntoken > 0 && ninput > 0 && H > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(H <= 128 && H % 32 == 0 && ninput <= 1024 && B <= 65535 * LSTM_BT, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  const int G = 4 * H;
  float* xproj = ws;
  float* w_hhT = ws + (size_t)B * Tcap * 2 * G;
  TGSR_CUDA(cudaMemsetAsync(words, 0, (size_t)B * 2 * H * Tmax * sizeof(float), st));
  TGSR_CUDA(cudaMemsetAsync(sent, 0, (size_t)B * 2 * H * sizeof(float), st));
  lstm_transpose_whh<<<64, 256, 0, st>>>(w_hh, w_hhT, H);
  const int ntok = B * Tcap;
  lstm_xproj_kernel<<<ceil_div(ntok, LSTM_TOK), 256, LSTM_TOK * ninput * sizeof(float), st>>>(
      captions, emb, w_ih, b_ih, b_hh, xproj, ntok, ntoken, ninput, H);
  const size_t smem = (size_t)(LSTM_BT * H + LSTM_BT * G) * sizeof(float);
  lstm_recurrence_kernel<<<dim3(2, ceil_div(B, LSTM_BT)), G, smem, st>>>(cap_lens, xproj, w_hhT, words, sent, B,
                                                                        Tcap, Tmax, H);
  return launch_status(3);
}
