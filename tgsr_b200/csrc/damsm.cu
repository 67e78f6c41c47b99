// DAMSM word / sentence matching losses with gradients to the image side (miscc/losses.py:21-136, :499-522;
// func_attention GlobalAttention.py:33-74), fp32.
//
// The reference loops over the captions in Python: for caption i it repeats the words B times, runs func_attention
// against all B images (two softmaxes, three transposed copies, two bmm), a cosine similarity and a log-sum-exp,
// and concatenates B columns -- O(B) launches of ~15 kernels each, O(B^2) work.  Here every (image j, caption i) pair
// is one CTA of ONE kernel:
//   raw[s,t] = sum_d F_j[d,s] W_i[d,t]            p = softmax_t(raw)            a[t,:] = softmax_s(gamma1 * p[:,t])
//   c[d,t]   = sum_s F_j[d,s] a[t,s]              r_t = cos(W_i[:,t], c[:,t])   sim[j,i] = log sum_t exp(gamma2 r_t)
// followed by one small kernel for the two cross entropies (and d sim), one per-pair backward kernel and one
// per-image kernel that gathers the gradient of the region features without atomics:
//   dF_j = sum_i ( dc_ji . a_ji  +  W_i . draw_ji^T )
// Gradients reach the image side only (region features, cnn_code): the text encoder is frozen in the training step.
#include "common.cuh"
#include <math_constants.h>

namespace tgsr {

constexpr int DM_THREADS = 256;
constexpr int DM_NDF = 256;          // cfg.TEXT.EMBEDDING_DIM of every TGSR config (util.py:185)
constexpr int DM_TP = 24;            // words per caption, padded (WORDS_NUM = 18)
constexpr int DM_SC = 32;            // regions per staged chunk
constexpr int DM_FLD = 36;           // row stride of the [d][32 regions] chunk
constexpr int DM_FLD2 = DM_NDF + 4;  // row stride of the transposed [region][d] chunk

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// out[s][t] (row stride DM_TP+1) = sum_d F[d][s] * X[d][t]   for s < S, t < DM_TP;  X in shared memory [256][DM_TP]
// F: global [256][S].  Fc: staging [256][DM_FLD], part: [2][32][DM_TP+1].
__device__ void gemm_FtX(const float* __restrict__ F, int S, const float* X, float* out, float* Fc, float* part) {
  const int tid = threadIdx.x, s4 = (tid & 7) * 4, tp = (tid >> 3) & 15, kh = tid >> 7;
  for (int s0 = 0; s0 < S; s0 += DM_SC) {
    __syncthreads();
    for (int idx = tid; idx < DM_NDF * DM_SC; idx += DM_THREADS) {
      const int d = idx >> 5, s = idx & 31;
      Fc[d * DM_FLD + s] = (s0 + s < S) ? __ldg(F + (size_t)d * S + s0 + s) : 0.f;
    }
    __syncthreads();
    if (tp < DM_TP / 2) {
      float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
      const int d0 = kh * (DM_NDF / 2);
#pragma unroll 4
      for (int d = d0; d < d0 + DM_NDF / 2; ++d) {
        const float4 f = *reinterpret_cast<const float4*>(Fc + d * DM_FLD + s4);
        const float2 x = *reinterpret_cast<const float2*>(X + d * DM_TP + 2 * tp);
        acc[0][0] = fmaf(f.x, x.x, acc[0][0]); acc[0][1] = fmaf(f.x, x.y, acc[0][1]);
        acc[1][0] = fmaf(f.y, x.x, acc[1][0]); acc[1][1] = fmaf(f.y, x.y, acc[1][1]);
        acc[2][0] = fmaf(f.z, x.x, acc[2][0]); acc[2][1] = fmaf(f.z, x.y, acc[2][1]);
        acc[3][0] = fmaf(f.w, x.x, acc[3][0]); acc[3][1] = fmaf(f.w, x.y, acc[3][1]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        part[(kh * 32 + s4 + k) * (DM_TP + 1) + 2 * tp] = acc[k][0];
        part[(kh * 32 + s4 + k) * (DM_TP + 1) + 2 * tp + 1] = acc[k][1];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < DM_SC * DM_TP; idx += DM_THREADS) {
      const int s = idx / DM_TP, t = idx - s * DM_TP;
      if (s0 + s < S) out[(s0 + s) * (DM_TP + 1) + t] = part[s * (DM_TP + 1) + t] + part[(32 + s) * (DM_TP + 1) + t];
    }
  }
  __syncthreads();
}

// acc[k][m] = sum_s F[d4+k][s] * A[s][tq + 4m]   (d4 = 4*(tid&63), tq = tid>>6, m < 6: t < 24)
// A: shared [S][DM_TP+1].  Fc2: staging [32][DM_FLD2].
__device__ void gemm_FA(const float* __restrict__ F, int S, const float* A, float* Fc2, float (&acc)[4][6]) {
  const int tid = threadIdx.x, d4 = (tid & 63) * 4, tq = tid >> 6;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int m = 0; m < 6; ++m) acc[k][m] = 0.f;
  for (int s0 = 0; s0 < S; s0 += DM_SC) {
    __syncthreads();
    for (int idx = tid; idx < DM_NDF * DM_SC; idx += DM_THREADS) {
      const int d = idx >> 5, s = idx & 31;
      Fc2[s * DM_FLD2 + d] = (s0 + s < S) ? __ldg(F + (size_t)d * S + s0 + s) : 0.f;
    }
    __syncthreads();
    const int ns = min(DM_SC, S - s0);
    for (int s = 0; s < ns; ++s) {
      const float4 f = *reinterpret_cast<const float4*>(Fc2 + s * DM_FLD2 + d4);
      const float* ar = A + (s0 + s) * (DM_TP + 1) + tq;
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        const float a = ar[4 * m];
        acc[0][m] = fmaf(f.x, a, acc[0][m]); acc[1][m] = fmaf(f.y, a, acc[1][m]);
        acc[2][m] = fmaf(f.z, a, acc[2][m]); acc[3][m] = fmaf(f.w, a, acc[3][m]);
      }
    }
  }
  __syncthreads();
}

// shared-memory layout helpers (floats)
constexpr int DM_OFF_W = 0;                                   // Wq   [256][24]
constexpr int DM_OFF_A = DM_OFF_W + DM_NDF * DM_TP;           // raw / a [S_pad][25]
__host__ __device__ inline int dm_spad(int S) { return (S + 31) / 32 * 32; }

// ---------------------------------------------------------------- forward, one CTA per (image j, caption i)
// grid (n_cap, n_img).  diag != 0: func_attention mode, grid (B, 1), pair (b, b), every word counts (cap_lens NULL).
// outputs (any may be NULL): sim [n_img][n_cap]; attn_out [pair][T][S] (a); p_out [pair][S][DM_TP];
// c_out [pair][256][T] (weightedContext).
__global__ void __launch_bounds__(DM_THREADS)
damsm_pair_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ words, const long long* __restrict__ cap_lens,
                      float gamma1, float gamma2, float eps, float* __restrict__ sim, float* __restrict__ attn_out,
                      float* __restrict__ p_out, float* __restrict__ c_out, int S, int T, int diag) {
  extern __shared__ __align__(16) float dsm[];
  const int SP = dm_spad(S);
  float* Wq = dsm + DM_OFF_W;
  float* A = dsm + DM_OFF_A;                                  // [SP][25]
  float* Fc = A + SP * (DM_TP + 1);                           // max(256*36, 32*260)
  float* part = Fc + DM_NDF * DM_FLD;                         // [2][32][25]
  float* red = part + 2 * 32 * (DM_TP + 1);                   // [3][DM_TP] + scratch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x, j = diag ? blockIdx.x : blockIdx.y, n_cap = gridDim.x;
  const size_t pair = diag ? (size_t)i : (size_t)j * n_cap + i;
  const int len = cap_lens ? (int)min((long long)T, cap_lens[i]) : T;
  const float* F = feats + (size_t)j * DM_NDF * S;
  const float* W = words + (size_t)i * DM_NDF * T;
  for (int idx = tid; idx < DM_NDF * DM_TP; idx += DM_THREADS) {
    const int d = idx / DM_TP, t = idx - d * DM_TP;
    Wq[idx] = t < len ? __ldg(W + (size_t)d * T + t) : 0.f;     // words beyond the caption's length do not exist (:79-81)
  }
  gemm_FtX(F, S, Wq, A, Fc, part);                            // raw[s][t]   (GlobalAttention.py:51)
  // softmax over the caption's words, per region (:54)
  for (int s = tid; s < S; s += DM_THREADS) {
    float* r = A + s * (DM_TP + 1);
    float m = -CUDART_INF_F;
    for (int t = 0; t < len; ++t) m = fmaxf(m, r[t]);
    float sum = 0.f;
    for (int t = 0; t < len; ++t) { const float e = expf(r[t] - m); r[t] = e; sum += e; }
    const float inv = 1.0f / sum;
    for (int t = 0; t < len; ++t) r[t] *= inv;
    for (int t = len; t < DM_TP; ++t) r[t] = 0.f;
    if (p_out)
      for (int t = 0; t < DM_TP; ++t) p_out[(pair * S + s) * DM_TP + t] = r[t];
  }
  __syncthreads();
  // softmax over regions of gamma1 * p, per word (:59-63): warp per word
  for (int t = warp; t < DM_TP; t += DM_THREADS / 32) {
    if (t < len) {
      float m = -CUDART_INF_F;
      for (int s = lane; s < S; s += 32) m = fmaxf(m, gamma1 * A[s * (DM_TP + 1) + t]);
      m = warp_max_f(m);
      float sum = 0.f;
      for (int s = lane; s < S; s += 32) { const float e = expf(gamma1 * A[s * (DM_TP + 1) + t] - m); A[s * (DM_TP + 1) + t] = e; sum += e; }
      sum = warp_sum_f(sum);
      const float inv = 1.0f / sum;
      for (int s = lane; s < S; s += 32) {
        const float a = A[s * (DM_TP + 1) + t] * inv;
        A[s * (DM_TP + 1) + t] = a;
        if (attn_out) attn_out[(pair * T + t) * S + s] = a;
      }
    } else if (attn_out && t < T) {
      for (int s = lane; s < S; s += 32) attn_out[(pair * T + t) * S + s] = 0.f;
    }
  }
  __syncthreads();
  float acc[4][6];
  gemm_FA(F, S, A, Fc, acc);                                  // c[d][t]  (:70)
  const int d4 = (tid & 63) * 4, tq = tid >> 6;
  if (c_out) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int m = 0; m < 6; ++m)
        if (tq + 4 * m < T) c_out[(pair * DM_NDF + d4 + k) * T + tq + 4 * m] = acc[k][m];
  }
  if (!sim) return;
  // cosine similarity per word (losses.py:12-18,102) and log-sum-exp over the caption's words (:107-109)
  if (tid < 3 * DM_TP) red[tid] = 0.f;
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    const int t = tq + 4 * m;
    float wc = 0.f, cc = 0.f, ww = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float w = Wq[(d4 + k) * DM_TP + t], c = acc[k][m];
      wc = fmaf(w, c, wc); cc = fmaf(c, c, cc); ww = fmaf(w, w, ww);
    }
    wc = warp_sum_f(wc); cc = warp_sum_f(cc); ww = warp_sum_f(ww);
    if (lane == 0) { atomicAdd(red + t, wc); atomicAdd(red + DM_TP + t, cc); atomicAdd(red + 2 * DM_TP + t, ww); }
  }
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;
    for (int t = 0; t < len; ++t) {
      const float r = red[t] / fmaxf(sqrtf(red[DM_TP + t]) * sqrtf(red[2 * DM_TP + t]), eps);
      sum += expf(gamma2 * r);
    }
    sim[pair] = logf(sum);
  }
}

// ---------------------------------------------------------------- cross entropies over the B x B score matrix
// scores = gamma3 * sim (+ -inf where class mask); loss0 = CE over rows (image -> caption), loss1 = CE over columns
// (losses.py:124-133 / :47-59).  One CTA.  d0/d1 = d loss0 / d sim and d loss1 / d sim (gamma3 included).
__global__ void __launch_bounds__(256)
damsm_ce_kernel(const float* __restrict__ sim, const unsigned char* __restrict__ cmask, const long long* __restrict__ labels,
                float gamma3, float* __restrict__ loss, float* __restrict__ d0, float* __restrict__ d1, int B) {
  __shared__ float lsum[2];
  if (threadIdx.x < 2) lsum[threadIdx.x] = 0.f;
  __syncthreads();
  for (int r = threadIdx.x; r < 2 * B; r += 256) {
    const int dir = r / B, k = r - dir * B;                  // dir 0: row k; dir 1: column k
    float m = -CUDART_INF_F;
    for (int n = 0; n < B; ++n) {
      const int jj = dir ? n : k, ii = dir ? k : n;
      float v = gamma3 * sim[jj * B + ii];
      if (cmask && cmask[jj * B + ii]) v = -CUDART_INF_F;
      m = fmaxf(m, v);
    }
    float sum = 0.f;
    for (int n = 0; n < B; ++n) {
      const int jj = dir ? n : k, ii = dir ? k : n;
      float v = gamma3 * sim[jj * B + ii];
      if (cmask && cmask[jj * B + ii]) v = -CUDART_INF_F;
      sum += expf(v - m);
    }
    const float lse = m + logf(sum);
    const int lab = (int)labels[k];
    {
      const int jj = dir ? lab : k, ii = dir ? k : lab;
      atomicAdd(lsum + dir, (lse - gamma3 * sim[jj * B + ii]) / B);
    }
    float* dd = dir ? d1 : d0;
    for (int n = 0; n < B; ++n) {
      const int jj = dir ? n : k, ii = dir ? k : n;
      float v = gamma3 * sim[jj * B + ii];
      const bool masked = cmask && cmask[jj * B + ii];
      const float sm = masked ? 0.f : expf(v - lse);
      dd[jj * B + ii] = masked ? 0.f : gamma3 * (sm - (n == lab ? 1.f : 0.f)) / B;   // .data.masked_fill_: no gradient
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) loss[threadIdx.x] = lsum[threadIdx.x];
}

// ---------------------------------------------------------------- backward per pair
// g = g0*d0[j,i] + g1*d1[j,i];  writes dcT [pair][DM_TP][256] (d sim / d c, transposed) and draw [pair][DM_TP][S]
// (d sim / d raw, word-major), both scaled by g.
__global__ void __launch_bounds__(DM_THREADS)
damsm_pair_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ words, const long long* __restrict__ cap_lens,
                      const float* __restrict__ attn_in, const float* __restrict__ p_in, const float* __restrict__ c_in,
                      const float* __restrict__ d0, const float* __restrict__ d1, const float* __restrict__ g0,
                      const float* __restrict__ g1, float gamma1, float gamma2, float eps, float* __restrict__ dcT,
                      float* __restrict__ draw, int S, int T) {
  extern __shared__ __align__(16) float dsm[];
  const int SP = dm_spad(S);
  float* X = dsm;                                             // dc [256][24]
  float* A = X + DM_NDF * DM_TP;                              // da -> dz -> draw [SP][25]
  float* Fc = A + SP * (DM_TP + 1);
  float* part = Fc + DM_NDF * DM_FLD;
  float* red = part + 2 * 32 * (DM_TP + 1);                   // [4][DM_TP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x, j = blockIdx.y, n_cap = gridDim.x;
  const size_t pair = (size_t)j * n_cap + i;
  const int len = (int)min((long long)T, cap_lens[i]);
  const float g = g0[0] * d0[pair] + g1[0] * d1[pair];
  const float* F = feats + (size_t)j * DM_NDF * S;
  const float* W = words + (size_t)i * DM_NDF * T;
  // cosine terms from the saved c
  if (tid < 4 * DM_TP) red[tid] = 0.f;
  __syncthreads();
  for (int t = warp; t < len; t += DM_THREADS / 32) {
    float wc = 0.f, cc = 0.f, ww = 0.f;
    for (int d = lane; d < DM_NDF; d += 32) {
      const float w = __ldg(W + (size_t)d * T + t), c = __ldg(c_in + (pair * DM_NDF + d) * T + t);
      wc = fmaf(w, c, wc); cc = fmaf(c, c, cc); ww = fmaf(w, w, ww);
    }
    wc = warp_sum_f(wc); cc = warp_sum_f(cc); ww = warp_sum_f(ww);
    if (lane == 0) { red[t] = wc; red[DM_TP + t] = cc; red[2 * DM_TP + t] = ww; }
  }
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;
    for (int t = 0; t < len; ++t) {
      const float r = red[t] / fmaxf(sqrtf(red[DM_TP + t]) * sqrtf(red[2 * DM_TP + t]), eps);
      const float e = expf(gamma2 * r);
      red[3 * DM_TP + t] = e; sum += e;
    }
    for (int t = 0; t < len; ++t) red[3 * DM_TP + t] = g * gamma2 * red[3 * DM_TP + t] / sum;   // d sim / d r_t (times g)
  }
  __syncthreads();
  // dc[d][t] = dr_t * ( w/(|w||c|) - r c/|c|^2 )   (zero where the clamp is active or t >= len)
  for (int idx = tid; idx < DM_NDF * DM_TP; idx += DM_THREADS) {
    const int d = idx / DM_TP, t = idx - d * DM_TP;
    float v = 0.f;
    if (t < len) {
      const float nc = sqrtf(red[DM_TP + t]), nw = sqrtf(red[2 * DM_TP + t]), den = nc * nw;
      if (den > eps) {
        const float w = __ldg(W + (size_t)d * T + t), c = __ldg(c_in + (pair * DM_NDF + d) * T + t);
        v = red[3 * DM_TP + t] * (w / den - red[t] * c / (den * red[DM_TP + t]));
      }
    }
    X[idx] = v;
    dcT[(pair * DM_TP + t) * DM_NDF + d] = v;
  }
  gemm_FtX(F, S, X, A, Fc, part);                             // da[s][t] = sum_d F[d][s] dc[d][t]
  // word-wise softmax backward over regions: dz[t][s] = a (da - sum_s a da); dp[s][t] = gamma1 dz
  for (int t = warp; t < DM_TP; t += DM_THREADS / 32) {
    if (t < len) {
      float dot = 0.f;
      for (int s = lane; s < S; s += 32) dot = fmaf(__ldg(attn_in + (pair * T + t) * S + s), A[s * (DM_TP + 1) + t], dot);
      dot = warp_sum_f(dot);
      for (int s = lane; s < S; s += 32)
        A[s * (DM_TP + 1) + t] = gamma1 * __ldg(attn_in + (pair * T + t) * S + s) * (A[s * (DM_TP + 1) + t] - dot);
    } else {
      for (int s = lane; s < S; s += 32) A[s * (DM_TP + 1) + t] = 0.f;
    }
  }
  __syncthreads();
  // region-wise softmax backward over words: draw[s][t] = p (dp - sum_t p dp)
  for (int s = tid; s < S; s += DM_THREADS) {
    const float* p = p_in + (pair * S + s) * DM_TP;
    float* r = A + s * (DM_TP + 1);
    float dot = 0.f;
    for (int t = 0; t < len; ++t) dot = fmaf(__ldg(p + t), r[t], dot);
    for (int t = 0; t < DM_TP; ++t) r[t] = t < len ? __ldg(p + t) * (r[t] - dot) : 0.f;
  }
  __syncthreads();
  for (int idx = tid; idx < DM_TP * S; idx += DM_THREADS) {
    const int t = idx / S, s = idx - t * S;
    draw[(pair * DM_TP + t) * S + s] = A[s * (DM_TP + 1) + t];
  }
}

// ---------------------------------------------------------------- gradient of the region features, per image
// dF[j][d][s] = sum_i sum_t ( dcT[j,i][t][d] * a[j,i][t][s] + wT[i][t][d] * draw[j,i][t][s] )
// grid (ceil(S/32), n_img); thread (d4 = 4*(tid&63), sg = tid>>6): 4 channels x 8 regions.
__global__ void __launch_bounds__(DM_THREADS)
damsm_dfeat_kernel(const float* __restrict__ dcT, const float* __restrict__ attn_in, const float* __restrict__ wT,
                   const float* __restrict__ draw, float* __restrict__ dF, int n_cap, int S, int T) {
  const int tid = threadIdx.x, d4 = (tid & 63) * 4, sg = tid >> 6;
  const int j = blockIdx.y, s0 = blockIdx.x * 32 + sg * 8;
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[k][m] = 0.f;
  for (int i = 0; i < n_cap; ++i) {
    const size_t pair = (size_t)j * n_cap + i;
    for (int t = 0; t < T; ++t) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(dcT + (pair * DM_TP + t) * DM_NDF + d4));
      const float4 w = __ldg(reinterpret_cast<const float4*>(wT + ((size_t)i * T + t) * DM_NDF + d4));
      const float* ar = attn_in + (pair * T + t) * S;
      const float* dr = draw + (pair * DM_TP + t) * S;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int s = s0 + m;
        const float a = s < S ? __ldg(ar + s) : 0.f, r = s < S ? __ldg(dr + s) : 0.f;
        acc[0][m] = fmaf(x.x, a, fmaf(w.x, r, acc[0][m])); acc[1][m] = fmaf(x.y, a, fmaf(w.y, r, acc[1][m]));
        acc[2][m] = fmaf(x.z, a, fmaf(w.z, r, acc[2][m])); acc[3][m] = fmaf(x.w, a, fmaf(w.w, r, acc[3][m]));
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (s0 + m < S) dF[((size_t)j * DM_NDF + d4 + k) * S + s0 + m] = acc[k][m];
}

// ---------------------------------------------------------------- sentence similarity (losses.py:41-46) and its gradient
// sim[j][i] = <C_j, E_i> / max(|C_j||E_i|, eps)
__global__ void __launch_bounds__(256)
sent_sim_kernel(const float* __restrict__ cnn, const float* __restrict__ rnn, float eps, float* __restrict__ sim, int B, int D) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * B) return;
  const int j = warp / B, i = warp - j * B;
  float dot = 0.f, cc = 0.f, ee = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float c = __ldg(cnn + (size_t)j * D + d), e = __ldg(rnn + (size_t)i * D + d);
    dot = fmaf(c, e, dot); cc = fmaf(c, c, cc); ee = fmaf(e, e, ee);
  }
  dot = warp_sum_f(dot); cc = warp_sum_f(cc); ee = warp_sum_f(ee);
  if (lane == 0) sim[warp] = dot / fmaxf(sqrtf(cc) * sqrtf(ee), eps);
}
// dC[j][d] = sum_i (g0 d0 + g1 d1)[j,i] * ( E_i[d]/(|C_j||E_i|) - sim[j,i] C_j[d]/|C_j|^2 )
__global__ void __launch_bounds__(256)
sent_bwd_kernel(const float* __restrict__ cnn, const float* __restrict__ rnn, const float* __restrict__ sim,
                const float* __restrict__ d0, const float* __restrict__ d1, const float* __restrict__ g0,
                const float* __restrict__ g1, float eps, float* __restrict__ dC, int B, int D) {
  extern __shared__ float en[];                                // |E_i|^2, then the pair weights
  __shared__ float nrm[2];
  const int j = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float cc = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) { const float c = cnn[(size_t)j * D + d]; cc = fmaf(c, c, cc); }
  cc = warp_sum_f(cc);
  if (threadIdx.x == 0) nrm[0] = 0.f;
  __syncthreads();
  if (lane == 0) atomicAdd(nrm, cc);
  for (int i = warp; i < B; i += 8) {
    float ee = 0.f;
    for (int k = lane; k < D; k += 32) { const float e = __ldg(rnn + (size_t)i * D + k); ee = fmaf(e, e, ee); }
    ee = warp_sum_f(ee);
    if (lane == 0) en[i] = ee;
  }
  __syncthreads();
  const float nc2 = nrm[0], nc = sqrtf(nc2);
  for (int d = threadIdx.x; d < D; d += 256) {
    const float c = cnn[(size_t)j * D + d];
    float acc = 0.f;
    for (int i = 0; i < B; ++i) {
      const float g = g0[0] * d0[j * B + i] + g1[0] * d1[j * B + i];
      const float den = nc * sqrtf(en[i]);
      if (den > eps) acc += g * (__ldg(rnn + (size_t)i * D + d) / den - sim[j * B + i] * c / nc2);
    }
    dC[(size_t)j * D + d] = acc;
  }
}

// ---------------------------------------------------------------- MSE (losses.py:779-784: nn.MSELoss, mean reduction)
__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ loss, float* __restrict__ grad,
           size_t n, float inv_n) {
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float d = __ldg(a + i) - __ldg(b + i);
    s = fmaf(d, d, s);
    if (grad) grad[i] = 2.0f * d * inv_n;
  }
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, s * inv_n);
}

__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g, float* __restrict__ grad,
               size_t n, float inv_n) {
  const float s = 2.0f * inv_n * g[0];
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) grad[i] = s * (__ldg(a + i) - __ldg(b + i));
}

static size_t dm_fwd_smem(int S) {
  return (size_t)(DM_NDF * DM_TP + dm_spad(S) * (DM_TP + 1) + DM_NDF * DM_FLD + 2 * 32 * (DM_TP + 1) + 4 * DM_TP) * sizeof(float);
}

int init_damsm() {
  cudaError_t e = cudaFuncSetAttribute(damsm_pair_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return cuda_fail(e);
  e = cudaFuncSetAttribute(damsm_pair_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  return e == cudaSuccess ? TGSR_OK : cuda_fail(e);
}

}  // namespace tgsr

using namespace tgsr;

extern "C" int tgsr_damsm_pairs_fwd_f32(const float* feats, const float* words, const int64_t* cap_lens, float gamma1,
                                        float gamma2, float eps, float* sim, float* attn, float* p, float* c, int n_img,
                                        int n_cap, int ndf, int S, int T, int diag, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(feats); TGSR_REQUIRE_PTR(words);
  TGSR_REQUIRE(n_img > 0 && n_cap > 0 && S > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(ndf == DM_NDF && T <= DM_TP && S <= 1024 && n_img <= 65535, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE(!diag || n_img == n_cap, TGSR_ERR_BAD_SHAPE);
  static_assert(DM_NDF * DM_FLD >= 32 * DM_FLD2, "staging buffer holds both chunk layouts");
  const size_t smem = dm_fwd_smem(S);
  TGSR_REQUIRE(smem <= 200 * 1024, TGSR_ERR_UNSUPPORTED);
  dim3 grid(n_cap, diag ? 1 : n_img);
  damsm_pair_fwd_kernel<<<grid, DM_THREADS, smem, (cudaStream_t)stream>>>(
      feats, words, reinterpret_cast<const long long*>(cap_lens), gamma1, gamma2, eps, sim, attn, p, c, S, T, diag);
  return launch_status();
}

extern "C" int tgsr_damsm_ce_f32(const float* sim, const uint8_t* class_mask, const int64_t* labels, float gamma3,
                                 float* loss2, float* d0, float* d1, int B, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(sim); TGSR_REQUIRE_PTR(labels); TGSR_REQUIRE_PTR(loss2); TGSR_REQUIRE_PTR(d0); TGSR_REQUIRE_PTR(d1);
  TGSR_REQUIRE(B > 0, TGSR_ERR_BAD_SHAPE);
  damsm_ce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sim, class_mask, reinterpret_cast<const long long*>(labels), gamma3,
                                                      loss2, d0, d1, B);
  return launch_status();
}

extern "C" int tgsr_damsm_pairs_bwd_f32(const float* feats, const float* words, const float* words_t, const int64_t* cap_lens,
                                        const float* attn, const float* p, const float* c, const float* d0, const float* d1,
                                        const float* g0, const float* g1, float gamma1, float gamma2, float eps, float* dcT,
                                        float* draw, float* d_feats, int n_img, int n_cap, int ndf, int S, int T,
                                        tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(feats); TGSR_REQUIRE_PTR(words); TGSR_REQUIRE_PTR(words_t); TGSR_REQUIRE_PTR(cap_lens); TGSR_REQUIRE_PTR(attn);
  TGSR_REQUIRE_PTR(p); TGSR_REQUIRE_PTR(c); TGSR_REQUIRE_PTR(d0); TGSR_REQUIRE_PTR(d1); TGSR_REQUIRE_PTR(g0); TGSR_REQUIRE_PTR(g1);
  TGSR_REQUIRE_PTR(dcT); TGSR_REQUIRE_PTR(draw); TGSR_REQUIRE_PTR(d_feats);
  TGSR_REQUIRE(n_img > 0 && n_cap > 0 && S > 0 && T > 0, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(ndf == DM_NDF && T <= DM_TP && S <= 1024 && n_img <= 65535, TGSR_ERR_UNSUPPORTED);
  const size_t smem = dm_fwd_smem(S);
  TGSR_REQUIRE(smem <= 200 * 1024, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  damsm_pair_bwd_kernel<<<dim3(n_cap, n_img), DM_THREADS, smem, st>>>(
      feats, words, reinterpret_cast<const long long*>(cap_lens), attn, p, c, d0, d1, g0, g1, gamma1, gamma2, eps, dcT, draw, S, T);
  int rc = launch_status();
  if (rc != TGSR_OK) return rc;
  damsm_dfeat_kernel<<<dim3(ceil_div(S, 32), n_img), DM_THREADS, 0, st>>>(dcT, attn, words_t, draw, d_feats, n_cap, S, T);
  return launch_status();
}

extern "C" int tgsr_sent_sim_f32(const float* cnn_code, const float* rnn_code, float eps, float* sim, int B, int D,
                                 tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(cnn_code); TGSR_REQUIRE_PTR(rnn_code); TGSR_REQUIRE_PTR(sim);
  TGSR_REQUIRE(B > 0 && D > 0, TGSR_ERR_BAD_SHAPE);
  sent_sim_kernel<<<ceil_div(B * B * 32, 256), 256, 0, (cudaStream_t)stream>>>(cnn_code, rnn_code, eps, sim, B, D);
  return launch_status();
}

extern "C" int tgsr_sent_bwd_f32(const float* cnn_code, const float* rnn_code, const float* sim, const float* d0,
                                 const float* d1, const float* g0, const float* g1, float eps, float* d_cnn, int B, int D,
                                 tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(cnn_code); TGSR_REQUIRE_PTR(rnn_code); TGSR_REQUIRE_PTR(sim); TGSR_REQUIRE_PTR(d0); TGSR_REQUIRE_PTR(d1);
  TGSR_REQUIRE_PTR(g0); TGSR_REQUIRE_PTR(g1); TGSR_REQUIRE_PTR(d_cnn);
  TGSR_REQUIRE(B > 0 && D > 0, TGSR_ERR_BAD_SHAPE);
  sent_bwd_kernel<<<B, 256, (size_t)B * sizeof(float), (cudaStream_t)stream>>>(cnn_code, rnn_code, sim, d0, d1, g0, g1, eps, d_cnn, B, D);
  return launch_status();
}

/* loss[0] += mean((a-b)^2) (loss must be zeroed by the caller, so several terms can accumulate: losses.py:779-784);
 * grad (or NULL) = d mean / d a. */
extern "C" int tgsr_mse_f32(const float* a, const float* b, float* loss, float* grad, size_t n, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(a); TGSR_REQUIRE_PTR(b); TGSR_REQUIRE_PTR(loss);
  TGSR_REQUIRE(n > 0, TGSR_ERR_BAD_SHAPE);
  const size_t sms = g_sm_count > 0 ? g_sm_count : 148;
  size_t blocks = (n + 255) / 256;
  if (blocks > sms * 8) blocks = sms * 8;
  mse_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, loss, grad, n, 1.0f / (float)n);
  return launch_status();
}

/* grad = g[0] * d mean((a-b)^2) / d a  (g: device scalar, the upstream gradient). */
extern "C" int tgsr_mse_bwd_f32(const float* a, const float* b, const float* g, float* grad, size_t n, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(a); TGSR_REQUIRE_PTR(b); TGSR_REQUIRE_PTR(g); TGSR_REQUIRE_PTR(grad);
  TGSR_REQUIRE(n > 0, TGSR_ERR_BAD_SHAPE);
  const size_t sms = g_sm_count > 0 ? g_sm_count : 148;
  size_t blocks = (n + 255) / 256;
  if (blocks > sms * 8) blocks = sms * 8;
  mse_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, g, grad, n, 1.0f / (float)n);
  return launch_status();
}
