/*
 * tgsr.h -- C ABI of libtgsr_sm100.so, the B200 (sm_100a) implementation of the TGSR
 * text-conditioned SR generator forward path.
 *
 * The reference (cxm12/TGSR) has no FFI: the path sits behind plain torch.nn.Module classes
 * (SURVEY.md section 8b).  Each entry point below replaces the arithmetic of the reference
 * Python code cited next to it; the host-side mirror of the module API (tgsr_b200/*.py) binds
 * these symbols with ctypes.  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer borrowed from the caller (torch tensors); the library
 *     never allocates, frees or retains caller memory; workspaces are passed in;
 *   - tensors are dense, row-major in the index order written in the comment;
 *   - `stream` is a cudaStream_t (CUstream); work is enqueued, never synchronised;
 *   - returns 0 (TGSR_OK) or a negative tgsr_status; never throws; launch errors are
 *     captured with cudaGetLastError and returned as TGSR_ERR_CUDA (tgsr_last_cuda_error()).
 *   - thread-safe with respect to distinct streams after tgsr_init(device).
 */
#ifndef TGSR_H_
#define TGSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tgsr_stream_t; /* == cudaStream_t */

typedef enum {
  TGSR_OK = 0,
  TGSR_ERR_BAD_SHAPE = -1,    /* a size is <= 0, misaligned or inconsistent                  */
  TGSR_ERR_UNSUPPORTED = -2,  /* valid in the reference but outside this kernel's envelope   */
  TGSR_ERR_CUDA = -3,         /* CUDA runtime/driver error, see tgsr_last_cuda_error()        */
  TGSR_ERR_NULL = -4,         /* a required pointer is NULL                                  */
  TGSR_ERR_NOT_INIT = -5      /* tgsr_init(device) has not been called for this device       */
} tgsr_status;

/* activation selectors for the conv entry points */
enum { TGSR_ACT_NONE = 0, TGSR_ACT_GLU = 1, TGSR_ACT_TANH = 2 };
/* 16-bit operand/storage format of the tensor-core path (both: tcgen05 kind::f16, fp32 accumulate) */
enum { TGSR_FMT_BF16 = 0, TGSR_FMT_F16 = 1 };

int tgsr_version(void);
const char* tgsr_error_string(int status);
int tgsr_last_cuda_error(void); /* cudaError_t of the last TGSR_ERR_CUDA on this thread */
/* One-time per device: kernel attributes (dynamic smem opt-in), TMA driver entry point, SM count. */
int tgsr_init(int device);
int tgsr_sm_count(void);
/* Number of kernels this library has enqueued since load (monotonic; not thread-exact). */
unsigned long long tgsr_launch_count(void);

/* ---- a1: GlobalAttentionGeneral.forward, GlobalAttention.py:87-130 (fp32 NCHW I/O) ---------
 * input [B,idf,Q], context [B,cdf,T], w_ctx [idf,cdf] (conv_context.weight), mask uint8 [B,T] or
 * NULL, wctx [B,idf,Q], attn [B,T,Q].  One fused kernel: word projection, logits, caption mask
 * (reference row-order quirk: pixel q of image b uses caption (b*Q+q) % B), softmax over T,
 * weighted sum.  idf <= 256, cdf <= 512 and cdf % 4 == 0, T <= 48. */
int tgsr_attn_fwd_f32(const float* input, const float* context, const float* w_ctx,
                      const uint8_t* mask, float* wctx, float* attn, int B, int idf, int cdf,
                      int Q, int T, tgsr_stream_t stream);

/* a1 at fp32 accuracy on the warp-level tensor-core path (3xTF32 split operands, fp32 accumulate): the word
 * projection runs once per caption into `ws` (tgsr_attn_f32_ws_floats(B, idf, T) floats, caller-owned), then one
 * persistent kernel streams the pixels with float4-only global accesses.  Same tensors and semantics as
 * tgsr_attn_fwd_f32.  Envelope (tgsr_attn_f32_tc_supported): idf % 8 == 0, idf <= 256, T <= 48, Q % 32 == 0 and the
 * operand images fit shared memory ((idf/8) * ceil8(T)/8 <= ~120). */
size_t tgsr_attn_f32_ws_floats(int B, int idf, int T);
int tgsr_attn_f32_tc_supported(int B, int idf, int cdf, int Q, int T, int has_mask);
int tgsr_attn_fwd_f32_tc(const float* input, const float* context, const float* w_ctx, const uint8_t* mask,
                         float* ws, float* wctx, float* attn, int B, int idf, int cdf, int Q, int T,
                         tgsr_stream_t stream);

/* ---- a2: func_attention, GlobalAttention.py:33-74 (DAMSM double softmax), forward ----------
 * query [B,ndf,Tq], context [B,ndf,S], wctx [B,ndf,Tq], attn [B,Tq,S].  Tq <= 48, S <= 1024. */
int tgsr_func_attn_fwd_f32(const float* query, const float* context, float gamma1, float* wctx,
                           float* attn, int B, int ndf, int Tq, int S, tgsr_stream_t stream);

/* ---- a3/a4/a5/a8: conv (3x3 / 5x5, stride 1, same padding, no bias) with fused epilogue ----
 * util.py:62-69 conv, nn.BatchNorm2d eval (folded to scale/shift), util.py:45-53 GLU,
 * util.py:126-130 residual add, util.py:76 nearest x2 upsample of the INPUT, tanh image heads
 * (util.py:899-901, model.py:224) and model.py:280 `one*y + a*SRb`.  fp32 NCHW.
 *   x [B,Cin,H,W]; if up2x the convolution runs on the nearest-upsampled 2H x 2W image.
 *   w [Cout,Cin,k,k]; scale/shift [Cout] or NULL (identity).
 *   z = conv*scale + shift; act GLU: o[c] = z[c]*sigmoid(z[c+Cout/2]) (Cout/2 channels out);
 *   act TANH: o = tanh(z).  Then o += residual (same shape as o, or NULL);
 *   then o = out_scale*o + addend_scale*addend (addend same shape as o, or NULL).
 *   y [B, Cout or Cout/2, Ho, Wo]. */
int tgsr_conv2d_f32(const float* x, const float* w, const float* scale, const float* shift,
                    const float* residual, const float* addend, float out_scale,
                    float addend_scale, float* y, int B, int Cin, int H, int W, int Cout,
                    int ksize, int up2x, int act, tgsr_stream_t stream);

/* BatchNorm eval fold: scale = gamma/sqrt(var+eps), shift = beta - mean*scale (fp32, [C]). */
int tgsr_bn_fold_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                     float eps, float* scale, float* shift, int C, tgsr_stream_t stream);

/* ---- a12/a13: Linear (+bias) (+BatchNorm1d fold) + GLU, util.py:379-386 and :417-420 -------
 * x [B,K], w [N,K], bias/scale/shift [N] or NULL; y [B,N/2] = GLU((x w^T + bias)*scale+shift). */
int tgsr_linear_glu_f32(const float* x, const float* w, const float* bias, const float* scale,
                        const float* shift, float* y, int B, int K, int N, tgsr_stream_t stream);
/* CA_NET.reparametrize util.py:388-395: c = eps*exp(0.5*logvar)+mu; mulogvar [B,2*c_dim]. */
int tgsr_ca_reparam_f32(const float* mulogvar, const float* eps, float* c_code, int B, int c_dim,
                        tgsr_stream_t stream);

/* ---- a11: RNN_ENCODER.forward, util.py:233-260 (Embedding -> 1-layer bi-LSTM, eval) --------
 * captions int64 [B,Tcap]; cap_lens int64 [B] (device; any order); emb [ntoken,ninput];
 * w_ih [2][4H,ninput], w_hh [2][4H,H], b_ih/b_hh [2][4H] (dir 0 = forward, 1 = reverse; gate
 * order i,f,g,o).  xproj_ws: workspace of B*Tcap*8H + 8H*H floats.  words [B,2H,Tmax] (zero beyond
 * each caption's length), sent [B,2H].  No host sync.  H <= 128. */
int tgsr_lstm_bidir_fwd_f32(const int64_t* captions, const int64_t* cap_lens, const float* emb,
                            const float* w_ih, const float* w_hh, const float* b_ih,
                            const float* b_hh, float* xproj_ws, float* words, float* sent, int B,
                            int Tcap, int Tmax, int ntoken, int ninput, int H,
                            tgsr_stream_t stream);

/* Fast path of a11 for H == 128 (the TGSR encoder): the input projection becomes a table lookup.
 * tgsr_lstm_prepare_f32 builds xtab [ntoken][2][4H] = emb.w_ih^T + b_ih + b_hh once per weight
 * version; tgsr_lstm_bidir_fwd_tab_f32 runs the recurrence as one persistent cluster kernel
 * (W_hh register-resident, h exchanged through distributed shared memory).  w_hh [2][4H,H]. */
int tgsr_lstm_prepare_f32(const float* emb, const float* w_ih, const float* b_ih, const float* b_hh,
                          float* xtab, int ntoken, int ninput, int H, tgsr_stream_t stream);
int tgsr_lstm_bidir_fwd_tab_f32(const int64_t* captions, const int64_t* cap_lens, const float* xtab,
                                const float* w_hh, float* words, float* sent, int B, int Tcap, int Tmax,
                                int ntoken, int H, tgsr_stream_t stream);

/* =============================== 16-bit tensor-core path ===================================
 * Every entry point below takes `fmt` (TGSR_FMT_BF16 or TGSR_FMT_F16): the storage format of the
 * blocked activations and packed weights it touches.
 * Activations live in HBM as "blocked planes": 16-bit [B][P][H][W][8]; channel c is element c%8 of
 * plane c/8.  A view of C channels inside a buffer with P planes per image is (ptr, P, p0): it
 * covers planes [p0, p0 + C/8).  Channel concatenation (util.py:773,817) is therefore free: the
 * producers write disjoint plane ranges of one buffer. */

/* NCHW fp32 [B,C,H,W] <-> blocked 16-bit view (edge conversions; the reference API is NCHW fp32). */
int tgsr_nchw_to_blocked(const float* x, void* y, int B, int C, int H, int W, int yP, int yp0, int fmt,
                         tgsr_stream_t stream);
int tgsr_blocked_to_nchw(const void* x, float* y, int B, int C, int H, int W, int xP, int xp0, int fmt,
                         tgsr_stream_t stream);

/* Weight prep: conv3x3 weight [Cout,Cin,3,3] fp32 + folded BatchNorm (scale, shift) -> packed 16-bit UMMA operand
 * image of tgsr_packed_conv3x3_elems(Cout,Cin,up2x) elements: [tap weight blocks | shift block | ones block].
 * up2x == 0: 9 taps.  up2x == 1: the nearest-x2 upsample of util.py:76 is folded into the weights as four 2x2
 * sub-pixel phases (pre-summed in fp32, rounded once), 16 blocks.  scale [Cout] (or NULL = 1) is multiplied into the
 * weights in fp32 before the single rounding; shift [Cout] (or NULL = 0) is applied by one extra MMA per tile
 * (ones x shift, the shift split into three 16-bit terms, exact to 2^-24), so the conv epilogue adds nothing.
 * act: TGSR_ACT_NONE or TGSR_ACT_GLU as in the conv call (bf16 GLU layers carry an exact factor 1/2 in weights and
 * shift: z*sigmoid(g) = z/2 + z/2*tanh(g/2)). */
size_t tgsr_packed_conv3x3_elems(int Cout, int Cin, int up2x);
int tgsr_pack_conv3x3(const float* w, const float* scale, const float* shift, void* packed, int Cout, int Cin,
                      int up2x, int act, int fmt, tgsr_stream_t stream);

/* a3/a4/a5: conv3x3 (+ folded BN, both parts inside wpacked) (+GLU) (+residual) on tcgen05,
 * 16-bit operands, fp32 accumulation in TMEM; util.py:62-65, :45-53, :110-130, :74-80.
 * x view (xP,xp0) of Cin channels at HxW; y view (yP,yp0) of Cout (or Cout/2 with GLU) channels at
 * HxW, or at 2Hx2W when up2x (wpacked must then come from tgsr_pack_conv3x3(up2x=1)).
 * residual view (rP,rp0) or NULL (only with act NONE, no up2x, Cout <= 64).  Supported (Cin,Cout,act,up2x):
 * (64,128,GLU,0) (64,64,NONE,0) (32,64,GLU,0) (32,32,NONE,0) (64,64,GLU,1) (32,64,GLU,1), and for the training step's
 * raw convolutions (64,128,NONE,0) (32,64,NONE,0);
 * H % 16 == 0, W % 8 == 0.
 * flags: TGSR_CONV_REVERSE walks the output tiles from the last one backwards.  Results are identical; a kernel that
 * consumes what its predecessor just wrote in forward order finds the most recently written (still L2-resident) part
 * of its input first, which matters when a layer's tensors exceed the 126 MB L2 (64 channels at 128x128, batch 32). */
#define TGSR_CONV_REVERSE 1
#define TGSR_CONV_L2_IN_FIRST 2   /* the inputs (patch and skip tiles) are dead after this call: load them with L2 evict_first */
#define TGSR_CONV_L2_OUT_LAST 4   /* the output is the next kernel's input: store it with L2 evict_last (plain conv mode) */
int tgsr_conv3x3_tc(const void* x, int xP, int xp0, const void* wpacked, const void* residual, int rP, int rp0,
                    void* y, int yP, int yp0, int B, int Cin, int H, int W, int Cout, int up2x, int act, int fmt,
                    int flags, tgsr_stream_t stream);

/* a8 image heads on tensor cores: GET_IMAGE_G / GET_IMAGE_G_noAct (conv3x3 32->3, util.py:894-919) and
 * conv_output (conv5x5 32->3 + Tanh, model.py:223-226) with the fused `one*y + a*SRb` of model.py:280.
 * x: blocked view of 32 channels; wpacked from tgsr_pack_head (w [3,Cin,k,k] fp32);
 * y fp32 NCHW [B,3,H,W] = out_scale * act(conv(x)) + addend_scale * addend (addend NCHW or NULL).
 * act NONE or TANH; W % 8 == 0 (any H).  The kx taps are folded into the MMA N dimension (N = 3k <= 16), so
 * a kxk head issues k row-tap MMAs per K-step; packed weights hold k*16*Cin elements.
 * nonfinite_flag: optional device int; the kernel ORs 1 into it when a pre-activation sum is inf or NaN -- an
 * overflow of the 16-bit activations anywhere upstream ends up there (tanh would otherwise hide it). */
int tgsr_pack_head(const float* w, void* packed, int Cin, int ksize, int fmt, tgsr_stream_t stream);
int tgsr_img_head_tc(const void* x, int xP, int xp0, const void* wpacked, const float* addend,
                     float out_scale, float addend_scale, float* y, int B, int Cin, int H, int W,
                     int ksize, int act, int fmt, int* nonfinite_flag, tgsr_stream_t stream);

/* im2f / convin: conv3x3(3 -> 2C) + folded BN + GLU from the NCHW fp32 LR image (minus x_sub if not
 * NULL: model.py:268-270 `LR - LRb`) into a blocked 16-bit view of C channels; util.py:741-744,
 * model.py:228.  w [2C,3,3,3], scale/shift [2C] or NULL.  C % 8 == 0, C <= 64. */
int tgsr_conv_in_blocked(const float* x, const float* x_sub, const float* w, const float* scale,
                         const float* shift, void* y, int yP, int yp0, int B, int H, int W, int C, int fmt,
                         tgsr_stream_t stream);

/* a1 on blocked activations (GlobalAttention.py:87-130), two entry points.
 * tgsr_attn_project: the word projection S[b] = w_ctx . context[b] (conv_context, GlobalAttention.py:99-101) in
 * fp32, written as the per-lane tensor-core operand image `sfrag` (tgsr_attn_sfrag_bytes_t(B, T) bytes, caller-owned
 * workspace; tgsr_attn_sfrag_bytes(B) is the T <= 24 size).  It depends on the caption only: one call per (captions,
 * conv_context weight), reused by every pixel tile.  context fp32 [B,cdf,T], w_ctx fp32 [idf,cdf]; idf == 32, T <= 48
 * (three kernel variants: word tiles for T <= 24, <= 32, <= 48), cdf % 4 == 0.
 * tgsr_attn_fwd_blocked: reads the idf=32 channel view (xP,xp0), writes weightedContext (16-bit) into the view
 * (yP,yp0) -- normally the next planes of the same concat buffer -- and the attention map attn fp32 [B,T,Q]; caption
 * mask uint8 [B,T] or NULL with the reference's row-order quirk (pixel q of image b uses caption (b*Q+q) % B).
 * Logits, mask, softmax and the weighted sum are one persistent kernel (TMA-fed, warp-level tensor cores, S split
 * into three 16-bit pieces so the fp32 maps keep fp32 accuracy).  Q % 32 == 0. */
size_t tgsr_attn_sfrag_bytes(int B);
size_t tgsr_attn_sfrag_bytes_t(int B, int T);
int tgsr_attn_project(const float* context, const float* w_ctx, void* sfrag, int B, int idf, int cdf, int T,
                      int fmt, tgsr_stream_t stream);
/* n (1..3) projections of the same captions with different conv_context weights in ONE launch (a forward computes the
 * operand images of all generator stages up front). */
int tgsr_attn_project_multi(const float* context, const float* w0, const float* w1, const float* w2, void* sfrag0,
                            void* sfrag1, void* sfrag2, int n, int B, int idf, int cdf, int T, int fmt,
                            tgsr_stream_t stream);
int tgsr_attn_fwd_blocked(const void* x, int xP, int xp0, const void* sfrag, const uint8_t* mask, void* y,
                          int yP, int yp0, float* attn, int B, int idf, int Q, int T, int fmt,
                          tgsr_stream_t stream);

/* ============================ training-mode generator (fp32 NCHW) ===============================
 * The reference trains through autograd + cuDNN; these are the explicit forward/backward kernels of the same
 * arithmetic (nn.BatchNorm2d batch statistics, GLU util.py:45-53, ResBlock skip util.py:126-130, nn.Upsample(2)
 * util.py:76, conv3x3/conv5x5 util.py:62-69, tanh head + `one*y + a*SRb` model.py:224,280, GlobalAttention.py:87-130). */

/* BatchNorm2d train-mode statistics of x [B,C,HW]: mean[c], invstd[c] = 1/sqrt(biased var + eps); updates
 * running_mean / running_var (unbiased var, momentum) in place when not NULL.  ws: 2*C doubles of scratch (the channel
 * sums of the multi-CTA reduction). */
int tgsr_bn_stats_f32(const float* x, float* mean, float* invstd, float* running_mean, float* running_var,
                      float momentum, float eps, double* ws, int B, int C, int HW, tgsr_stream_t stream);
/* y = act(gamma*(x-mean)*invstd + beta) (+ residual, shaped like y).  act NONE: y [B,C,HW]; act GLU: y [B,C/2,HW]. */
int tgsr_bn_act_f32(const float* x, const float* mean, const float* invstd, const float* gamma,
                    const float* beta, const float* residual, float* y, int B, int C, int HW, int act,
                    tgsr_stream_t stream);
/* Backward of tgsr_bn_act_f32 through the batch statistics: dx [B,C,HW], dgamma [C], dbeta [C] from dy (shape of y).
 * (The residual's gradient is dy itself.)  ws: 2*C doubles of scratch. */
int tgsr_bn_act_bwd_f32(const float* x, const float* mean, const float* invstd, const float* gamma,
                        const float* beta, const float* dy, float* dx, float* dgamma, float* dbeta, double* ws,
                        int B, int C, int HW, int act, tgsr_stream_t stream);
/* wt[ci][co][k-1-ky][k-1-kx] = w[co][ci][ky][kx]: tgsr_conv2d_f32(dy, wt) is the data gradient of conv(x, w). */
int tgsr_conv_weight_flip_f32(const float* w, float* wt, int Cout, int Cin, int ksize, tgsr_stream_t stream);
/* Backward of the nearest x2 upsample folded into a conv: y [BC,H,W] = 2x2 block sums of x [BC,2H,2W]. */
int tgsr_sumpool2x2_f32(const float* x, float* y, int BC, int H, int W, tgsr_stream_t stream);
/* dw [Cout,Cin,k,k] = weight gradient of conv(x [B,Cin,H,W] (nearest-upsampled x2 if up2x), w) given dy
 * [B,Cout,Ho,Wo]; dw is overwritten. */
int tgsr_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, int B, int Cin, int H, int W, int Cout,
                          int ksize, int up2x, tgsr_stream_t stream);
/* y = one*tanh(conv) + a*srb (model.py:224,280): dconv = dy*one*(1-t^2), t = (y - a*srb)/one; dsrb = a*dy if not NULL.
 * srb may be NULL (a plain tanh head). */
int tgsr_tanh_head_bwd_f32(const float* y, const float* srb, const float* dy, float one, float a, float* dconv,
                           float* dsrb, size_t n, tgsr_stream_t stream);
/* Backward of a1: d_input [B,idf,Q] and d_w_ctx [idf,cdf] from d_wctx [B,idf,Q] (and d_attn [B,T,Q] or NULL), with
 * the forward's attention map `attn`.  ws: 2 * tgsr_attn_f32_ws_floats(B, idf, T) floats.  No gradient to `context`. */
int tgsr_attn_bwd_f32(const float* input, const float* context, const float* w_ctx, const float* attn,
                      const float* d_wctx, const float* d_attn, float* ws, float* d_input, float* d_w_ctx,
                      int B, int idf, int cdf, int Q, int T, tgsr_stream_t stream);

/* TF32 tensor-core convolutions of the training step (fp32 tensors, TF32 operands, fp32 accumulate -- the precision
 * PyTorch's cuDNN convs use by default).  tgsr_conv_prep_tf32 converts w [Cout,Cin,k,k] once per step into the K-major
 * operand `wk` (Cout*Cin*k*k words): dgrad == 0 for the forward, dgrad == 1 for the data gradient (channels swapped, taps
 * flipped).  tgsr_conv2d_tf32: y [B,N,Ho,Wo] = conv(x [B,Cin,H,W] (nearest x2 if up2x), wk), N = wk's output channels.
 * tgsr_conv2d_wgrad_tf32: as tgsr_conv2d_wgrad_f32.  tgsr_conv_tf32_supported(Cin, Cout, k, wgrad) gives the envelope
 * (output channels % 32 == 0; weight gradient: 3x3, both channel counts % 32 == 0). */
int tgsr_conv_tf32_supported(int Cin, int Cout, int ksize, int wgrad);
int tgsr_conv_prep_tf32(const float* w, void* wk, int Cout, int Cin, int ksize, int dgrad, tgsr_stream_t stream);
int tgsr_conv2d_tf32(const float* x, const void* wk, float* y, int B, int Cin, int H, int W, int N, int ksize,
                     int up2x, tgsr_stream_t stream);
int tgsr_conv2d_wgrad_tf32(const float* x, const float* dy, float* dw, int B, int Cin, int H, int W, int Cout,
                           int ksize, int up2x, tgsr_stream_t stream);

/* ---- a2 / a14: DAMSM matching losses (miscc/losses.py:21-136; func_attention GlobalAttention.py:33-74), fp32 ----
 * One CTA per (image j, caption i) pair replaces the reference's Python loop over captions.
 * feats [n_img,256,S] (region features, S = 17*17), words [n_cap,256,T], cap_lens int64 [n_cap] or NULL (all T words).
 * Outputs, any may be NULL: sim [n_img,n_cap] = log sum_t exp(gamma2 * cos(w_t, c_t)) (losses.py:102-109);
 * attn [pair,T,S] (second softmax, rows >= cap_len are zero); p [pair,S,24] (first softmax, for the backward);
 * c [pair,256,T] (weightedContext).  pair = j*n_cap + i.  diag != 0: pairs (b,b) only, pair = b (func_attention).
 * ndf == 256, T <= 24, S <= 1024. */
int tgsr_damsm_pairs_fwd_f32(const float* feats, const float* words, const int64_t* cap_lens, float gamma1,
                             float gamma2, float eps, float* sim, float* attn, float* p, float* c, int n_img,
                             int n_cap, int ndf, int S, int T, int diag, tgsr_stream_t stream);
/* Both cross entropies over scores = gamma3 * sim [B,B] (class_mask uint8 [B,B] or NULL: masked -> -inf, no gradient):
 * loss2[0] = CE over rows (image -> caption), loss2[1] = CE over columns; d0/d1 [B,B] = d loss / d sim. */
int tgsr_damsm_ce_f32(const float* sim, const uint8_t* class_mask, const int64_t* labels, float gamma3,
                      float* loss2, float* d0, float* d1, int B, tgsr_stream_t stream);
/* d_feats [n_img,256,S] = d (g0*loss0 + g1*loss1) / d feats through the pair kernel.  words_t [n_cap,T,256] is words
 * transposed; g0, g1 device scalars; dcT [pairs,24,256] and draw [pairs,24,S] are caller-owned scratch. */
int tgsr_damsm_pairs_bwd_f32(const float* feats, const float* words, const float* words_t, const int64_t* cap_lens,
                             const float* attn, const float* p, const float* c, const float* d0, const float* d1,
                             const float* g0, const float* g1, float gamma1, float gamma2, float eps, float* dcT,
                             float* draw, float* d_feats, int n_img, int n_cap, int ndf, int S, int T,
                             tgsr_stream_t stream);
/* sent_loss (losses.py:21-62): sim [B,B] = cos(cnn_code[j], rnn_code[i]); its gradient to cnn_code given d0/d1 of
 * tgsr_damsm_ce_f32. */
int tgsr_sent_sim_f32(const float* cnn_code, const float* rnn_code, float eps, float* sim, int B, int D,
                      tgsr_stream_t stream);
int tgsr_sent_bwd_f32(const float* cnn_code, const float* rnn_code, const float* sim, const float* d0,
                      const float* d1, const float* g0, const float* g1, float eps, float* d_cnn, int B, int D,
                      tgsr_stream_t stream);
/* nn.MSELoss (losses.py:779-784): loss[0] += mean((a-b)^2) (caller zeroes loss so terms accumulate); grad or NULL. */
int tgsr_mse_f32(const float* a, const float* b, float* loss, float* grad, size_t n, tgsr_stream_t stream);
/* grad = g[0] * 2 (a - b) / n  (g: device scalar, the upstream gradient of the mean). */
int tgsr_mse_bwd_f32(const float* a, const float* b, const float* g, float* grad, size_t n, tgsr_stream_t stream);

/* Fused Adam over a flat bucket (torch.optim.Adam semantics, no weight decay): g = grad*grad_scale; m = b1 m + (1-b1) g;
 * v = b2 v + (1-b2) g^2; param -= (lr / bias_corr1) * m / (sqrt(v) / bias_corr2_sqrt + eps). */
int tgsr_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                  float beta1, float beta2, float eps, float bias_corr1, float bias_corr2_sqrt, float grad_scale,
                  tgsr_stream_t stream);

/* Same update with the 1-based step count read from device memory (CUDA-graph capturable training step). */
int tgsr_adam_devstep_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                          float beta1, float beta2, float eps, const int* step, float grad_scale,
                          tgsr_stream_t stream);

/* Eval-flow edge: uint8 image quantisation of trainer_objective.py:153-155 on the device,
 * y[i] = uint8(round_half_even(clamp((x[i] + 1.0f) * 127.5f, 0, 255))), fp32 arithmetic as numpy does it.
 * x fp32 (16-byte aligned), y uint8 (4-byte aligned), n elements. */
int tgsr_quantize_u8(const float* x, uint8_t* y, size_t n, tgsr_stream_t stream);

/* Input pyramid on the device (replaces the per-image PIL work of datasets.py:151-196 get_imgs_blur / :112-148 get_imgs
 * done in the DataLoader workers).  Byte-exact with Pillow: Image.resize(BILINEAR) (antialiased triangle filter, 8-bit fixed
 * point, horizontal then vertical pass), ImageFilter.GaussianBlur (3 extended box blurs per axis, edge replication) and
 * ToTensor()+Normalize(.5,.5) (datasets.py:1563).  Images are uint8 [B, H, W, C] (decoded, HWC as PIL holds them);
 * normalised outputs are fp32 [B, C, H, W].
 *   tgsr_resize_u8        : transforms.Resize([oh, ow])(img) for a batch; ws >= tgsr_resize_u8_ws_bytes(...).
 *   tgsr_gaussian_blur_u8 : img.filter(GaussianBlur(radius)) with `passes` box blurs per axis (PIL uses 3); tmp = B*H*W*C bytes.
 *   tgsr_normalize_u8_f32 : ToTensor()+Normalize.
 *   tgsr_pyramid_levels_u8: up to 4 levels of one source in ONE launch; per level i the source is resized to (oh[i], ow[i])
 *                           (or taken as is when the size matches) and any of out[i] (normalised), out_blur[i] (normalised
 *                           GaussianBlur(blur_radius) of the resized image) and out_u8[i] (the resized bytes, [B,oh,ow,C]) that
 *                           is non-NULL is written.  out / out_blur / out_u8 are HOST arrays of device pointers (or NULL).
 *                           Envelope (tgsr_pyramid_supported): ow % 4 == 0, plane fits shared memory (all sizes up to 256x256
 *                           from sources up to 256 rows), blur_radius == 2. */
size_t tgsr_resize_u8_ws_bytes(int B, int ih, int iw, int oh, int ow, int C);
int tgsr_resize_u8(const uint8_t* src, uint8_t* dst, void* ws, int B, int ih, int iw, int oh, int ow, int C,
                   tgsr_stream_t stream);
int tgsr_gaussian_blur_u8(const uint8_t* src, uint8_t* dst, uint8_t* tmp, int B, int H, int W, int C, float radius,
                          int passes, tgsr_stream_t stream);
int tgsr_normalize_u8_f32(const uint8_t* src, float* dst, int B, int H, int W, int C, tgsr_stream_t stream);
int tgsr_pyramid_supported(int ih, int iw, int oh, int ow, int blur);
int tgsr_pyramid_levels_u8(const uint8_t* src, int B, int ih, int iw, int C, int nlevels, const int* oh, const int* ow,
                           float* const* out, float* const* out_blur, uint8_t* const* out_u8, float blur_radius,
                           tgsr_stream_t stream);

/* Word-axis reductions of the attention maps (miscc/utils.py:122-125 attn.max(dim=1); :389-392 conf_score; the max-map also
 * weighs weight_MSE, miscc/losses.py:796).  attn fp32 [B, T, Q] (Q % 4 == 0, T <= 255) ->
 *   maxv   fp32  [B, Q]     max over the T words,
 *   argmax uint8 [B, Q]     first word index attaining it,
 *   conf   fp64  [B, T+1]   replicas * sum over pixels of v * (v > 2*thresh): column 0 for the max-map, column 1+t for
 *                           word t (the reference sums a 3-channel replica of each map: replicas = 3). */
int tgsr_attn_word_stats_f32(const float* attn, float* maxv, uint8_t* argmax, double* conf, int B, int T, int Q, float thresh,
                             int replicas, tgsr_stream_t stream);

/* Stand-alone losses of miscc/losses.py.
 *   KL_loss (losses.py:806-810): out[0] = -0.5 * mean(1 + logvar - mu^2 - exp(logvar)) over n elements; the backward takes
 *   the upstream scalar gradient g[0] (device) and writes d_mu, d_logvar.
 *   cosine_similarity (losses.py:12-18) along the middle axis of an [outer, D, inner] view:
 *   out[o,i] = sum_d x1*x2 / max(|x1|*|x2|, eps); stats (nullable; 3 floats per output: dot, |x1|, |x2|) feeds the backward,
 *   which writes d1, d2 (same shape as the inputs) from the upstream gradient g [outer, inner]. */
int tgsr_kl_loss_f32(const float* mu, const float* logvar, float* out, size_t n, tgsr_stream_t stream);
int tgsr_kl_loss_bwd_f32(const float* mu, const float* logvar, const float* g, float* d_mu, float* d_logvar, size_t n,
                         tgsr_stream_t stream);
int tgsr_cosine_similarity_f32(const float* x1, const float* x2, float* out, float* stats, size_t outer, int D, int inner,
                               float eps, tgsr_stream_t stream);
int tgsr_cosine_similarity_bwd_f32(const float* x1, const float* x2, const float* stats, const float* g, float* d1, float* d2,
                                   size_t outer, int D, int inner, float eps, tgsr_stream_t stream);


/* ---- training step on blocked 16-bit activations [B][C/8][H][W][8] (csrc/train16.cu); statistics and gradients in fp32 /
 * fp64, one 16-bit rounding per layer.  nn.BatchNorm2d train mode + GLU + skip (util.py:45-53,64,110-130), nn.Upsample(2)
 * (util.py:76) and their adjoints. */
/* mean / invstd [C] of z over (B, H*W); running stats updated like nn.BatchNorm2d(momentum) when given.  ws: 2*C doubles. */
int tgsr_bn16_stats(const void* z, float* mean, float* invstd, float* running_mean, float* running_var, float momentum,
                    float eps, double* ws, int B, int C, int HW, int fmt, tgsr_stream_t stream);
/* y = act(gamma * (z - mean) * invstd + beta) (+ residual): act NONE -> C channels out (residual allowed), GLU -> C/2. */
int tgsr_bn16_act(const void* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                  const void* residual, void* y, int B, int C, int HW, int act, int fmt, tgsr_stream_t stream);
/* Backward of tgsr_bn16_act through the batch statistics: dz [B,C,..], dgamma / dbeta [C] from dy.  ws: 2*C doubles. */
int tgsr_bn16_act_bwd(const void* z, const float* mean, const float* invstd, const float* gamma, const float* beta,
                      const void* dy, void* dz, float* dgamma, float* dbeta, double* ws, int B, int C, int HW, int act,
                      int fmt, tgsr_stream_t stream);
/* Weight gradient of conv3x3 (pad 1, no bias) on tcgen05: dw [Cout,Cin,3,3] fp32 (overwritten) from the blocked 16-bit
 * views x (xP,xp0; Cin channels) and dz (zP,zp0; Cout channels) at HxW.  Both operands are read from the blocked layout
 * as MN-major UMMA operands (K = pixels); fp32 accumulation in TMEM, per-CTA partial sums added with fp32 atomics.
 * Cin in {32,64}, Cout in {32,64,128}, H % 16 == 0, W % 8 == 0.  ws: caller's scratch of 9*Cin*Cout floats (the partial sums
 * are accumulated there as [tap][ci][co], one 128-byte line per warp-wide atomic, then transposed into dw). */
int tgsr_conv3x3_wgrad_tc(const void* x, int xP, int xp0, const void* dz, int zP, int zp0, float* dw, float* ws, int B, int Cin,
                          int H, int W, int Cout, int fmt, tgsr_stream_t stream);
/* Weight gradient of a 3-output-channel image head (GET_IMAGE_G[_noAct].img util.py:894-919, conv_output model.py:223-226;
 * autograd in the reference): dw [3][Cin][k][k] fp32 = sum over pixels of dz [B,3,H,W] fp32 (x) the blocked 16-bit features
 * (view xP,xp0), zero padding.  Warp-level tensor-core path, dz enters as a 16-bit value + 16-bit residual.  Cin == 32,
 * k in {3,5}; dw is overwritten. */
int tgsr_head_wgrad16(const void* x, int xP, int xp0, const float* dz, float* dw, int B, int Cin, int H, int W, int ksize,
                      int fmt, tgsr_stream_t stream);
/* F.max_pool2d(x, 3, stride 2) (no padding, floor mode) of the frozen CNN_ENCODER (util.py:262-368, Inception-v3 stem and
 * Mixed_6a / Mixed_7a) on channels-last 16-bit tensors [B,H,W,C], C % 8 == 0: y [B,Ho,Wo,C], idx uint8 [B,Ho,Wo,C] = window
 * position of the first maximum; backward gathers gy through idx into gx [B,H,W,C] (no atomics). */
int tgsr_maxpool3s2_nhwc16(const void* x, void* y, uint8_t* idx, int B, int H, int W, int C, int fmt, tgsr_stream_t stream);
int tgsr_maxpool3s2_nhwc16_bwd(const void* gy, const uint8_t* idx, void* gx, int B, int H, int W, int C, int fmt,
                               tgsr_stream_t stream);
/* Data gradient of the same heads into the blocked layout: dh (view hP,hp0; Cin = 32 channels, 16-bit) = transposed conv of
 * dz [B,3,H,W] fp32 with w [3][Cin][k][k] fp32, zero padding; warp-level tensor cores, K = 3*k*k.  k in {3,5}. */
int tgsr_head_dgrad16(const float* dz, const float* w, void* dh, int hP, int hp0, int B, int Cin, int H, int W, int ksize, int fmt,
                      tgsr_stream_t stream);
/* ReLU backward on channels-last 16-bit tensors: out[p][c] = y[p][c] > 0 ? gy[p][c] : 0 for npix pixels of C channels; gy's
 * pixels are gy_pixel_stride elements apart (a channel slice of a wider tensor: the gradient of a concatenation), y and out
 * are dense.  C % 8 == 0, stride % 8 == 0, 16-byte aligned pointers. */
int tgsr_relu_bwd_nhwc16(const void* gy, long long gy_pixel_stride, const void* y, void* out, long long npix, int C, int fmt,
                         tgsr_stream_t stream);
/* y [B,P,2H,2W,8] = nearest x2 of x [B,P,H,W,8];  y [B,P,H,W,8] = 2x2 block sums of x [B,P,2H,2W,8] (its adjoint). */
int tgsr_upsample2x_blocked(const void* x, void* y, int B, int P, int H, int W, tgsr_stream_t stream);
int tgsr_sumpool2x2_blocked(const void* x, void* y, int B, int P, int H, int W, int fmt, tgsr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TGSR_H_ */
