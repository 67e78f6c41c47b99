// Input pyramid of the SR flows on the device: the reference builds its LR / real / upsampled ("bic") / blurred inputs on
// the host with PIL, per image, inside the DataLoader workers (datasets.py:151-196 get_imgs_blur, :112-148 get_imgs):
//   transforms.Resize(s)(img)                      -> PIL Image.resize(BILINEAR): separable triangle filter with
//                                                     antialiasing, 8-bit fixed point (22 fraction bits), horizontal pass
//                                                     rounded to uint8, then vertical pass rounded to uint8;
//   img.filter(ImageFilter.GaussianBlur(radius=2)) -> three extended box blurs along x, then three along y, 8.24 fixed point,
//                                                     edge pixels replicated;
//   ToTensor() + Normalize(.5,.5) (datasets.py:1563) -> ((x / 255) - 0.5) / 0.5 in fp32, HWC -> CHW.
// All of it is byte/integer arithmetic, so the results here are bit-identical to Pillow's (tests/test_gpu_pyramid.py).
// Filter coefficients are computed in the kernels in fp64 with explicitly rounded operations (no FMA contraction), the
// same operation order as the host library uses, so no host tables are needed.
//
// Two sets of kernels:
//   * generic one-thread-per-output kernels for any size / channel count (resize, box-blur pass, normalise);
//   * pyramid_plane_kernel: one CTA per (image, channel, level, row band) that keeps its rows in shared memory from the
//     first resample tap to the last blur pass: every output byte is written once, nothing intermediate reaches HBM.
#include "common.cuh"
#include <math.h>
#include <string.h>

namespace tgsr {
namespace {

constexpr int PREC_BITS = 22;              // 32 - 8 - 2: fraction bits of the 8-bit resample path
constexpr int PYR_THREADS = 256;
constexpr int PYR_MAX_JOBS = 4;
constexpr int PYR_BANDS_PER_JOB = 4;
constexpr int PYR_MAX_BANDS = PYR_MAX_JOBS * PYR_BANDS_PER_JOB;
constexpr int PYR_BAND_ROWS = 64;
constexpr int PYR_HALO = 6;                // three vertical box passes, each reaching 2 rows (radius 1 + the fractional tap)
constexpr int PYR_MAX_SMEM = 200 * 1024;

__host__ __device__ inline int resample_ksize(int in_size, int out_size) {
  double fs = (double)in_size / (double)out_size;
  if (fs < 1.0) fs = 1.0;
  return (int)ceil(fs) * 2 + 1;
}

// Taps of output sample xx of one axis: kk[0..ksize) fixed-point weights, bounds = (first input sample, tap count).
__device__ void resample_taps(int in_size, int out_size, int ksize, int xx, int* __restrict__ kk, int* __restrict__ bounds) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fs = scale < 1.0 ? 1.0 : scale;                 // filter scale == support (triangle filter support is 1)
  const double ss = __ddiv_rn(1.0, fs);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, fs), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, fs), 0.5);
  if (xmax > in_size) xmax = in_size;
  const int n = xmax - xmin;
  auto tap = [&](int x) -> double {
    const double t = fabs(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
    return t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
  };
  double tot = 0.0;
  for (int x = 0; x < n; ++x) tot = __dadd_rn(tot, tap(x));
  for (int x = 0; x < ksize; ++x) {
    double w = x < n ? tap(x) : 0.0;
    if (x < n && tot != 0.0) w = __ddiv_rn(w, tot);
    kk[x] = (int)__dadd_rn(0.5, __dmul_rn(w, (double)(1 << PREC_BITS)));
  }
  bounds[0] = xmin;
  bounds[1] = n;
}

__device__ __forceinline__ uint8_t clip8(int acc) {
  const int v = acc >> PREC_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ float norm_u8(unsigned int v) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), 0.5f), 0.5f);
}

// ----------------------------------------------------------------------------------------------- generic kernels

__global__ void resample_table_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx < out_size) resample_taps(in_size, out_size, ksize, xx, kk + (size_t)xx * ksize, bounds + 2 * xx);
}

// [rows, iw, C] -> [rows, ow, C]
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t total, int iw, int ow, int C, int ks,
                const int* __restrict__ bounds, const int* __restrict__ kk) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = (int)(idx % C);
    const int x = (int)((idx / C) % ow);
    const size_t row = idx / ((size_t)C * ow);
    const int x0 = __ldg(bounds + 2 * x), n = __ldg(bounds + 2 * x + 1);
    const uint8_t* p = src + (row * iw + x0) * C + c;
    const int* k = kk + (size_t)x * ks;
    int acc = 1 << (PREC_BITS - 1);
    for (int i = 0; i < n; ++i) acc += (int)__ldg(p + (size_t)i * C) * __ldg(k + i);
    dst[idx] = clip8(acc);
  }
}

// [B, ih, WC] -> [B, oh, WC]
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t total, int ih, int oh, int WC, int ks,
                const int* __restrict__ bounds, const int* __restrict__ kk) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int xc = (int)(idx % WC);
    const int y = (int)((idx / WC) % oh);
    const size_t b = idx / ((size_t)WC * oh);
    const int y0 = __ldg(bounds + 2 * y), n = __ldg(bounds + 2 * y + 1);
    const uint8_t* p = src + (b * ih + y0) * WC + xc;
    const int* k = kk + (size_t)y * ks;
    int acc = 1 << (PREC_BITS - 1);
    for (int i = 0; i < n; ++i) acc += (int)__ldg(p + (size_t)i * WC) * __ldg(k + i);
    dst[idx] = clip8(acc);
  }
}

// One extended box blur along a strided axis: element e of line `line` lives at base(line) + e * estride.
// out[x] = (ww * sum_{|i|<=r} in[x+i] + fw * (in[x-r-1] + in[x+r+1]) + 2^23) >> 24, indices clamped to [0, len).
__global__ void __launch_bounds__(256)
box_blur_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t total, int len, int inner, int r,
                uint32_t ww, uint32_t fw) {
  // element idx = (outer * len + x) * inner + in;  the blurred axis has stride `inner`.
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int in = (int)(idx % inner);
    const int x = (int)((idx / inner) % len);
    const size_t outer = idx / ((size_t)inner * len);
    const uint8_t* p = src + outer * len * inner + in;
    uint32_t acc = 0;
    for (int i = -r; i <= r; ++i) {
      int xi = x + i; xi = xi < 0 ? 0 : (xi >= len ? len - 1 : xi);
      acc += __ldg(p + (size_t)xi * inner);
    }
    int xl = x - r - 1; xl = xl < 0 ? 0 : xl;
    int xr = x + r + 1; xr = xr >= len ? len - 1 : xr;
    const uint32_t far = (uint32_t)__ldg(p + (size_t)xl * inner) + (uint32_t)__ldg(p + (size_t)xr * inner);
    dst[idx] = (uint8_t)((acc * ww + far * fw + (1u << 23)) >> 24);
  }
}

// [B, H, W, C] uint8 -> [B, C, H, W] fp32
__global__ void __launch_bounds__(256)
normalize_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t total, int HW, int C) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int q = (int)(idx % HW);
    const int c = (int)((idx / HW) % C);
    const size_t b = idx / ((size_t)HW * C);
    dst[idx] = norm_u8(__ldg(src + (b * HW + q) * C + c));
  }
}

// ----------------------------------------------------------------------------------------------- fused plane kernel
//
// One CTA per (image, channel, level, row band).  A band is up to a quarter of the level's rows; with the blur it carries a
// 6-row halo on each interior side (three vertical passes, each reaching 2 rows), so the rows it finally stores are exact
// while every band stays small enough for several CTAs per SM.  At the real image edges the halo is absent and the passes
// replicate the edge row, which is the filter's own boundary rule.

struct PyrJob {
  const uint8_t* src;      // [B, ih, iw, C]
  float* out;              // [B, C, oh, ow] normalised, nullable
  float* out_blur;         // [B, C, oh, ow] normalised blurred, nullable
  uint8_t* out_u8;         // [B, oh, ow, C] resized bytes, nullable (source of a later level)
  int ih, iw, oh, ow, ksx, ksy;
};
struct PyrJobs {
  PyrJob j[PYR_MAX_JOBS];
  int C;
  uint32_t ww, fw;
  uint8_t band_job[PYR_MAX_BANDS];
  short band_r0[PYR_MAX_BANDS], band_r1[PYR_MAX_BANDS];
};

__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return (w >> (8 * i)) & 0xffu; }

// Division by a loop-invariant width: a shift when it is a power of two (every production size), a real division otherwise.
struct FastDiv { int d, shift; };
__device__ __forceinline__ FastDiv fast_div(int d) { return FastDiv{d, (d & (d - 1)) == 0 ? __ffs(d) - 1 : -1}; }
__device__ __forceinline__ void divmod(int i, const FastDiv& f, int& q, int& r) {
  if (f.shift >= 0) { q = i >> f.shift; r = i & (f.d - 1); }
  else { q = i / f.d; r = i - q * f.d; }
}

// One extended box blur step (integer radius 1) on four byte lanes at once.  f0,a,b,c,f1 hold, per byte lane, the far-left,
// left, centre, right and far-right sample: out = (ww*(a+b+c) + fw*(f0+f1) + 2^23) >> 24.  The sums run two pixels per
// register in 16-bit lanes (max 765, no carry between lanes); the products are per pixel; the four top bytes are gathered
// with byte permutes.
__device__ __forceinline__ uint32_t blur_word(uint32_t f0, uint32_t a, uint32_t b, uint32_t c, uint32_t f1, uint32_t ww,
                                              uint32_t fw) {
  const uint32_t se = __byte_perm(a, 0, 0x4240) + __byte_perm(b, 0, 0x4240) + __byte_perm(c, 0, 0x4240);   // pixels 0, 2
  const uint32_t so = __byte_perm(a, 0, 0x4341) + __byte_perm(b, 0, 0x4341) + __byte_perm(c, 0, 0x4341);   // pixels 1, 3
  const uint32_t fe = __byte_perm(f0, 0, 0x4240) + __byte_perm(f1, 0, 0x4240);
  const uint32_t fo = __byte_perm(f0, 0, 0x4341) + __byte_perm(f1, 0, 0x4341);
  const uint32_t r0 = (se & 0xffffu) * ww + ((fe & 0xffffu) * fw + (1u << 23));
  const uint32_t r1 = (so & 0xffffu) * ww + ((fo & 0xffffu) * fw + (1u << 23));
  const uint32_t r2 = (se >> 16) * ww + ((fe >> 16) * fw + (1u << 23));
  const uint32_t r3 = (so >> 16) * ww + ((fo >> 16) * fw + (1u << 23));
  return __byte_perm(__byte_perm(r0, r1, 0x0073), __byte_perm(r2, r3, 0x0073), 0x5410);
}

// Box blur along x of an [h][w] byte plane, four pixels per thread: the shifted neighbourhoods are funnel shifts of the
// three words around the output word; the row ends replicate the edge pixel.
__device__ __forceinline__ void plane_blur_x(const uint8_t* __restrict__ s, uint8_t* __restrict__ d, int h, int w, uint32_t ww,
                                             uint32_t fw) {
  const int nw = w >> 2;
  const FastDiv dn = fast_div(nw);
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
  for (int i = threadIdx.x; i < h * nw; i += PYR_THREADS) {
    int y, j;
    divmod(i, dn, y, j);
    const uint32_t w1 = s32[i];
    const uint32_t w0 = j > 0 ? s32[i - 1] : (w1 & 0xffu) * 0x01010101u;
    const uint32_t w2 = j < nw - 1 ? s32[i + 1] : (w1 >> 24) * 0x01010101u;
    d32[i] = blur_word(__funnelshift_r(w0, w1, 16), __funnelshift_r(w0, w1, 24), w1, __funnelshift_r(w1, w2, 8),
                       __funnelshift_r(w1, w2, 16), ww, fw);
  }
}

// Same along y (rows clamp to the band's buffer: exact at image edges, absorbed by the halo elsewhere).
__device__ __forceinline__ void plane_blur_y(const uint8_t* __restrict__ s, uint8_t* __restrict__ d, int h, int w, uint32_t ww,
                                             uint32_t fw) {
  const int nw = w >> 2;
  const FastDiv dn = fast_div(nw);
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
  for (int i = threadIdx.x; i < h * nw; i += PYR_THREADS) {
    int y, j;
    divmod(i, dn, y, j);
    const int up1 = y > 0 ? nw : 0, up2 = y > 1 ? nw : 0;             // row steps, zero where the edge row replicates
    const int dn1 = y < h - 1 ? nw : 0, dn2 = y < h - 2 ? nw : 0;
    d32[i] = blur_word(s32[i - up1 - up2], s32[i - up1], s32[i], s32[i + dn1], s32[i + dn1 + dn2], ww, fw);
  }
}

// Horizontal resample of rows [0, ns) of one channel with the KS taps of a column held in registers: a thread owns one
// output column (ow a power of two <= PYR_THREADS) and walks down the rows, so the per-tap work is one byte load and one
// multiply-add.  Columns whose tap window would run past the row end are shifted left with their weights.
template <int KS, int CC>
__device__ __forceinline__ void hpass_columns(const uint8_t* __restrict__ sp, uint8_t* __restrict__ T, int ns, int iw, int ow,
                                              int ow_shift, const int* __restrict__ txb, const int* __restrict__ txk) {
  const int x = threadIdx.x & (ow - 1);
  const int ystep = PYR_THREADS >> ow_shift;
  int x0 = txb[2 * x];
  const int shift = max(x0 - (iw - KS), 0);
  x0 -= shift;
  int kreg[KS];
#pragma unroll
  for (int t = 0; t < KS; ++t) kreg[t] = (t >= shift) ? txk[x * KS + t - shift] : 0;
  const uint8_t* p = sp + (size_t)x0 * CC;
  for (int y = threadIdx.x >> ow_shift; y < ns; y += ystep) {
    const uint8_t* q = p + (size_t)y * iw * CC;
    int acc = 1 << (PREC_BITS - 1);
#pragma unroll
    for (int t = 0; t < KS; ++t) acc += (int)__ldg(q + t * CC) * kreg[t];
    T[y * ow + x] = clip8(acc);
  }
}

__device__ __forceinline__ void plane_store_norm(const uint8_t* __restrict__ s, float* __restrict__ g, int n4,
                                                 const float* __restrict__ lut) {
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
  float4* g4 = reinterpret_cast<float4*>(g);
  for (int i = threadIdx.x; i < n4; i += PYR_THREADS) {
    const uint32_t w = s32[i];
    g4[i] = make_float4(lut[byte_of(w, 0)], lut[byte_of(w, 1)], lut[byte_of(w, 2)], lut[byte_of(w, 3)]);
  }
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

// Rows [e0, e1) of the level that a band holding output rows [r0, r1) keeps in shared memory.
__host__ __device__ inline void band_extent(int r0, int r1, int oh, bool blur, int* e0, int* e1) {
  const int halo = blur ? PYR_HALO : 0;
  *e0 = imax(r0 - halo, 0);
  *e1 = imin(r1 + halo, oh);
}

// Shared-memory layout of one band (host and device agree through this function).  ns is an upper bound of the source rows
// the band's vertical taps touch.
struct PyrLayout { size_t tx, ty, T, A, Bp, total; };
__host__ __device__ inline PyrLayout pyr_layout(const PyrJob& J, bool blur, int ne) {
  PyrLayout L;
  const bool vpass = J.ih != J.oh;
  const int ns = vpass ? imin(J.ih, (ne * J.ih + J.oh - 1) / J.oh + J.ksy + 2) : ne;
  size_t off = 256 * sizeof(float);                                  // normalisation table
  L.tx = off; off += align16(J.iw != J.ow ? (size_t)J.ow * (J.ksx + 2) * 4 : 0);
  L.ty = off; off += align16(vpass ? (size_t)ne * (J.ksy + 2) * 4 : 0);
  L.T = off; off += align16((size_t)ns * J.ow);
  if (vpass) { L.A = off; off += align16((size_t)ne * J.ow); } else L.A = L.T;
  L.Bp = off; off += align16(blur ? (size_t)ne * J.ow : 0);
  L.total = off;
  return L;
}

__global__ void __launch_bounds__(PYR_THREADS)
pyramid_plane_kernel(const __grid_constant__ PyrJobs jobs) {
  extern __shared__ __align__(16) uint8_t smem[];
  const PyrJob& J = jobs.j[jobs.band_job[blockIdx.y]];
  const int r0 = jobs.band_r0[blockIdx.y], r1 = jobs.band_r1[blockIdx.y];
  const int C = jobs.C;
  const int b = blockIdx.x / C, c = blockIdx.x - b * C;
  const int ih = J.ih, iw = J.iw, oh = J.oh, ow = J.ow;
  const bool hpass = iw != ow, vpass = ih != oh, blur = J.out_blur != nullptr;
  int e0, e1;
  band_extent(r0, r1, oh, blur, &e0, &e1);
  const int ne = e1 - e0;
  const PyrLayout L = pyr_layout(J, blur, ne);
  float* lut = reinterpret_cast<float*>(smem);
  int* txb = reinterpret_cast<int*>(smem + L.tx);       // [ow][2] bounds, then [ow][ksx] weights
  int* txk = txb + 2 * ow;
  int* tyb = reinterpret_cast<int*>(smem + L.ty);       // [ne][2] bounds, then [ne][ksy] weights (rows e0..e1-1)
  int* tyk = tyb + 2 * ne;
  uint8_t* T = smem + L.T;
  uint8_t* A = smem + L.A;
  uint8_t* Bp = smem + L.Bp;

  if (threadIdx.x < 256) lut[threadIdx.x] = norm_u8(threadIdx.x);
  if (hpass) for (int x = threadIdx.x; x < ow; x += PYR_THREADS) resample_taps(iw, ow, J.ksx, x, txk + x * J.ksx, txb + 2 * x);
  if (vpass) for (int y = threadIdx.x; y < ne; y += PYR_THREADS) resample_taps(ih, oh, J.ksy, e0 + y, tyk + y * J.ksy, tyb + 2 * y);
  __syncthreads();
  // source rows [s0, s1) feed this band
  const int s0 = vpass ? tyb[0] : e0;
  const int s1 = vpass ? tyb[2 * (ne - 1)] + tyb[2 * (ne - 1) + 1] : e1;
  const int ns = s1 - s0;

  // stage 1: T[ns][ow] = horizontally resampled (or copied) rows of the channel plane
  const uint8_t* sp = J.src + ((size_t)b * ih + s0) * iw * C + c;
  const FastDiv dow = fast_div(ow);
  const bool columns = hpass && C == 3 && dow.shift >= 0 && ow <= PYR_THREADS && iw >= J.ksx &&
                       (J.ksx == 3 || J.ksx == 5 || J.ksx == 9 || J.ksx == 17);
  if (columns) {
    switch (J.ksx) {
      case 3: hpass_columns<3, 3>(sp, T, ns, iw, ow, dow.shift, txb, txk); break;
      case 5: hpass_columns<5, 3>(sp, T, ns, iw, ow, dow.shift, txb, txk); break;
      case 9: hpass_columns<9, 3>(sp, T, ns, iw, ow, dow.shift, txb, txk); break;
      default: hpass_columns<17, 3>(sp, T, ns, iw, ow, dow.shift, txb, txk); break;
    }
  } else if (hpass) {
    for (int i = threadIdx.x; i < ns * ow; i += PYR_THREADS) {
      int y, x;
      divmod(i, dow, y, x);
      const int x0 = txb[2 * x], n = txb[2 * x + 1];
      const uint8_t* p = sp + ((size_t)y * iw + x0) * C;
      const int* k = txk + x * J.ksx;
      int acc = 1 << (PREC_BITS - 1);
      for (int t = 0; t < n; ++t) acc += (int)__ldg(p + t * C) * k[t];
      T[i] = clip8(acc);
    }
  } else {
    for (int i = threadIdx.x; i < ns * ow; i += PYR_THREADS) T[i] = __ldg(sp + (size_t)i * C);
  }
  __syncthreads();
  // stage 2: A[ne][ow] = vertically resampled rows e0..e1-1 (A aliases T when the height is unchanged)
  if (vpass) {
    // four adjacent columns per thread: one word load and one weight load per tap
    const int nw = ow >> 2;
    const FastDiv dnw = fast_div(nw);
    const uint32_t* T32 = reinterpret_cast<const uint32_t*>(T);
    uint32_t* A32 = reinterpret_cast<uint32_t*>(A);
    for (int i = threadIdx.x; i < ne * nw; i += PYR_THREADS) {
      int y, j;
      divmod(i, dnw, y, j);
      const int y0 = tyb[2 * y] - s0, n = tyb[2 * y + 1];
      const uint32_t* p = T32 + y0 * nw + j;
      const int* k = tyk + y * J.ksy;
      int a0 = 1 << (PREC_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
      for (int t = 0; t < n; ++t) {
        const uint32_t w = p[t * nw];
        const int kt = k[t];
        a0 += (int)byte_of(w, 0) * kt; a1 += (int)byte_of(w, 1) * kt;
        a2 += (int)byte_of(w, 2) * kt; a3 += (int)byte_of(w, 3) * kt;
      }
      A32[i] = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
    }
    __syncthreads();
  }
  const size_t plane = ((size_t)b * C + c) * oh * ow + (size_t)r0 * ow;
  const uint8_t* Aout = A + (r0 - e0) * ow;
  const int nout = (r1 - r0) * ow;
  if (J.out != nullptr) plane_store_norm(Aout, J.out + plane, nout >> 2, lut);
  if (J.out_u8 != nullptr) {
    uint8_t* g = J.out_u8 + ((size_t)b * oh + r0) * ow * C + c;
    for (int i = threadIdx.x; i < nout; i += PYR_THREADS) g[(size_t)i * C] = Aout[i];
  }
  if (!blur) return;
  plane_blur_x(A, Bp, ne, ow, jobs.ww, jobs.fw); __syncthreads();
  plane_blur_x(Bp, A, ne, ow, jobs.ww, jobs.fw); __syncthreads();
  plane_blur_x(A, Bp, ne, ow, jobs.ww, jobs.fw); __syncthreads();
  plane_blur_y(Bp, A, ne, ow, jobs.ww, jobs.fw); __syncthreads();
  plane_blur_y(A, Bp, ne, ow, jobs.ww, jobs.fw); __syncthreads();
  plane_blur_y(Bp, A, ne, ow, jobs.ww, jobs.fw); __syncthreads();
  plane_store_norm(Aout, J.out_blur + plane, nout >> 2, lut);
}

// BoxBlur.c _gaussian_blur_radius + the 8.24 weights of ImagingHorizontalBoxBlur, float arithmetic as the host library.
void gaussian_box_params(float radius, int passes, int* r, uint32_t* ww, uint32_t* fw) {
  volatile float sigma2 = radius * radius / (float)passes;
  volatile float L = (float)sqrt(12.0 * sigma2 + 1.0);
  volatile float l = (float)floor((L - 1.0) / 2.0);
  volatile float a = (float)((2 * l + 1) * (l * (l + 1) - 3.0 * sigma2));
  a = a / (6 * (sigma2 - (l + 1) * (l + 1)));
  volatile float fr = l + a;
  *r = (int)fr;
  *ww = (uint32_t)((float)(1 << 24) / (fr * 2 + 1));
  *fw = ((1u << 24) - (uint32_t)(*r * 2 + 1) * *ww) / 2;
}

inline unsigned grid_for(size_t total) {
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace

int init_pyramid(void) {
  TGSR_CUDA(cudaFuncSetAttribute(pyramid_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PYR_MAX_SMEM));
  return TGSR_OK;
}

}  // namespace tgsr

using namespace tgsr;

extern "C" size_t tgsr_resize_u8_ws_bytes(int B, int ih, int iw, int oh, int ow, int C) {
  if (B <= 0 || ih <= 0 || iw <= 0 || oh <= 0 || ow <= 0 || C <= 0) return 0;
  const size_t tables = align16(((size_t)ow * (resample_ksize(iw, ow) + 2) + (size_t)oh * (resample_ksize(ih, oh) + 2)) * 4);
  return tables + align16((size_t)B * ih * ow * C);
}

extern "C" int tgsr_resize_u8(const uint8_t* src, uint8_t* dst, void* ws, int B, int ih, int iw, int oh, int ow, int C,
                              tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(src); TGSR_REQUIRE_PTR(dst);
  TGSR_REQUIRE(B > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0 && C > 0, TGSR_ERR_BAD_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  const bool hpass = iw != ow, vpass = ih != oh;
  if (!hpass && !vpass) {
    TGSR_CUDA(cudaMemcpyAsync(dst, src, (size_t)B * ih * iw * C, cudaMemcpyDeviceToDevice, st));
    return TGSR_OK;
  }
  TGSR_REQUIRE_PTR(ws);
  const int ksx = resample_ksize(iw, ow), ksy = resample_ksize(ih, oh);
  int* hb = reinterpret_cast<int*>(ws);
  int* hk = hb + 2 * ow;
  int* vb = hk + (size_t)ow * ksx;
  int* vk = vb + 2 * oh;
  uint8_t* tmp = reinterpret_cast<uint8_t*>(ws) + align16(((size_t)ow * (ksx + 2) + (size_t)oh * (ksy + 2)) * 4);
  int launches = 0;
  if (hpass) { resample_table_kernel<<<ceil_div(ow, 128), 128, 0, st>>>(iw, ow, ksx, hb, hk); ++launches; }
  if (vpass) { resample_table_kernel<<<ceil_div(oh, 128), 128, 0, st>>>(ih, oh, ksy, vb, vk); ++launches; }
  const uint8_t* vsrc = src;
  if (hpass) {
    uint8_t* hdst = vpass ? tmp : dst;
    const size_t total = (size_t)B * ih * ow * C;
    resize_h_kernel<<<grid_for(total), 256, 0, st>>>(src, hdst, total, iw, ow, C, ksx, hb, hk); ++launches;
    vsrc = hdst;
  }
  if (vpass) {
    const size_t total = (size_t)B * oh * ow * C;
    resize_v_kernel<<<grid_for(total), 256, 0, st>>>(vsrc, dst, total, ih, oh, ow * C, ksy, vb, vk); ++launches;
  }
  return launch_status(launches);
}

extern "C" int tgsr_gaussian_blur_u8(const uint8_t* src, uint8_t* dst, uint8_t* tmp, int B, int H, int W, int C, float radius,
                                     int passes, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(src); TGSR_REQUIRE_PTR(dst); TGSR_REQUIRE_PTR(tmp);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && passes > 0 && passes <= 8 && radius > 0.0f, TGSR_ERR_BAD_SHAPE);
  int r; uint32_t ww, fw;
  gaussian_box_params(radius, passes, &r, &ww, &fw);
  TGSR_REQUIRE(r >= 0 && r <= 4096, TGSR_ERR_UNSUPPORTED);
  TGSR_REQUIRE((uint64_t)ww * (uint64_t)(2 * r + 1) * 255 + (uint64_t)fw * 510 + (1u << 23) <= 0xffffffffull, TGSR_ERR_UNSUPPORTED);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * H * W * C;
  const unsigned grid = grid_for(total);
  // 2 * passes launches, ping-ponging tmp/dst so that the last one lands in dst
  const uint8_t* in = src;
  for (int p = 0; p < 2 * passes; ++p) {
    uint8_t* out = (p & 1) ? dst : tmp;
    if (p < passes) box_blur_kernel<<<grid, 256, 0, st>>>(in, out, total, W, C, r, ww, fw);          // along x: stride C
    else box_blur_kernel<<<grid, 256, 0, st>>>(in, out, total, H, W * C, r, ww, fw);                  // along y: stride W*C
    in = out;
  }
  return launch_status(2 * passes);
}

extern "C" int tgsr_normalize_u8_f32(const uint8_t* src, float* dst, int B, int H, int W, int C, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(src); TGSR_REQUIRE_PTR(dst);
  TGSR_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, TGSR_ERR_BAD_SHAPE);
  const size_t total = (size_t)B * H * W * C;
  normalize_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(src, dst, total, H * W, C);
  return launch_status();
}

// Row bands of one level: at most PYR_BANDS_PER_JOB, about PYR_BAND_ROWS rows each.
static int pyr_bands(int oh, int* rows) {
  int nb = ceil_div(oh, PYR_BAND_ROWS);
  if (nb > PYR_BANDS_PER_JOB) nb = PYR_BANDS_PER_JOB;
  *rows = ceil_div(oh, nb);
  return ceil_div(oh, *rows);
}

static size_t pyr_job_smem(const PyrJob& J, bool blur) {
  int rows;
  const int nb = pyr_bands(J.oh, &rows);
  size_t smem = 0;
  for (int k = 0; k < nb; ++k) {
    int e0, e1;
    band_extent(k * rows, imin((k + 1) * rows, J.oh), J.oh, blur, &e0, &e1);
    const size_t s = pyr_layout(J, blur, e1 - e0).total;
    smem = s > smem ? s : smem;
  }
  return smem;
}

static bool pyr_job_ok(const PyrJob& J, bool blur) {
  if (J.ow % 4 != 0 || J.oh < 1 || J.ow < 4 || J.oh > 32767) return false;
  return pyr_job_smem(J, blur) <= (size_t)PYR_MAX_SMEM;
}

extern "C" int tgsr_pyramid_supported(int ih, int iw, int oh, int ow, int blur) {
  if (ih <= 0 || iw <= 0 || oh <= 0 || ow <= 0) return 0;
  PyrJob J{};
  J.ih = ih; J.iw = iw; J.oh = oh; J.ow = ow; J.ksx = resample_ksize(iw, ow); J.ksy = resample_ksize(ih, oh);
  return pyr_job_ok(J, blur != 0) ? 1 : 0;
}

extern "C" int tgsr_pyramid_levels_u8(const uint8_t* src, int B, int ih, int iw, int C, int nlevels, const int* oh,
                                      const int* ow, float* const* out, float* const* out_blur, uint8_t* const* out_u8,
                                      float blur_radius, tgsr_stream_t stream) {
  TGSR_REQUIRE_PTR(src); TGSR_REQUIRE_PTR(oh); TGSR_REQUIRE_PTR(ow);
  TGSR_REQUIRE(B > 0 && ih > 0 && iw > 0 && C > 0 && nlevels > 0 && nlevels <= PYR_MAX_JOBS, TGSR_ERR_BAD_SHAPE);
  TGSR_REQUIRE(g_sm_count > 0, TGSR_ERR_NOT_INIT);
  PyrJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  jobs.C = C;
  int r = 1;
  bool any_blur = false;
  for (int i = 0; i < nlevels; ++i) any_blur |= out_blur != nullptr && out_blur[i] != nullptr;
  if (any_blur) {
    gaussian_box_params(blur_radius, 3, &r, &jobs.ww, &jobs.fw);
    TGSR_REQUIRE(r == 1, TGSR_ERR_UNSUPPORTED);          // GaussianBlur(radius=2), the one the reference uses
  }
  size_t smem = 0;
  int nbands = 0;
  for (int i = 0; i < nlevels; ++i) {
    PyrJob& J = jobs.j[i];
    J.src = src; J.ih = ih; J.iw = iw; J.oh = oh[i]; J.ow = ow[i];
    TGSR_REQUIRE(J.oh > 0 && J.ow > 0, TGSR_ERR_BAD_SHAPE);
    J.ksx = resample_ksize(iw, J.ow); J.ksy = resample_ksize(ih, J.oh);
    J.out = out != nullptr ? out[i] : nullptr;
    J.out_blur = out_blur != nullptr ? out_blur[i] : nullptr;
    J.out_u8 = out_u8 != nullptr ? out_u8[i] : nullptr;
    TGSR_REQUIRE(J.out != nullptr || J.out_blur != nullptr || J.out_u8 != nullptr, TGSR_ERR_NULL);
    const bool blur = J.out_blur != nullptr;
    TGSR_REQUIRE(pyr_job_ok(J, blur), TGSR_ERR_UNSUPPORTED);
    TGSR_REQUIRE((reinterpret_cast<uintptr_t>(J.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(J.out_blur) & 15) == 0,
                 TGSR_ERR_BAD_SHAPE);
    const size_t s = pyr_job_smem(J, blur);
    smem = s > smem ? s : smem;
    int rows;
    const int nb = pyr_bands(J.oh, &rows);
    for (int k = 0; k < nb; ++k) {
      jobs.band_job[nbands] = (uint8_t)i;
      jobs.band_r0[nbands] = (short)(k * rows);
      jobs.band_r1[nbands] = (short)imin((k + 1) * rows, J.oh);
      ++nbands;
    }
  }
  pyramid_plane_kernel<<<dim3((unsigned)(B * C), (unsigned)nbands), PYR_THREADS, smem, (cudaStream_t)stream>>>(jobs);
  return launch_status();
}
