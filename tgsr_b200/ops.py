"""Tensor-level wrappers over the C ABI (include/tgsr.h).  PyTorch is plumbing here: it owns the
device memory and the stream; every arithmetic op below is a libtgsr_sm100 kernel."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GLU, ACT_NONE, ACT_TANH, call, ptr

Tensor = torch.Tensor


def _f32c(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError('tgsr_b200: %s must be a CUDA tensor (no CPU path)' % name)
    if t.dtype != torch.float32:
        raise TypeError('tgsr_b200: %s must be float32, got %s' % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------ a1
def attention_fwd(inp: Tensor, context: Tensor, w_ctx: Tensor, mask: Optional[Tensor],
                  tensor_cores: bool = True) -> Tuple[Tensor, Tensor]:
    """GlobalAttentionGeneral.forward (GlobalAttention.py:87-130) -> (weightedContext, attn).
    tensor_cores=False keeps the CUDA-core kernel even inside the tensor-core kernel's envelope (tests)."""
    inp = _f32c(inp, 'input'); context = _f32c(context, 'context'); w_ctx = _f32c(w_ctx, 'conv_context.weight')
    if inp.dim() != 4 or context.dim() != 3:
        raise ValueError('expected input [B,idf,ih,iw] and context [B,cdf,T]')
    B, idf, ih, iw = inp.shape
    cdf, T = context.shape[1], context.shape[2]
    if context.shape[0] != B or w_ctx.numel() != idf * cdf:
        raise ValueError('attention shape mismatch: input %s context %s weight %s'
                         % (tuple(inp.shape), tuple(context.shape), tuple(w_ctx.shape)))
    m8 = None
    if mask is not None:
        if tuple(mask.shape) != (B, T):
            raise ValueError('mask must be [B,T]=%s, got %s' % ((B, T), tuple(mask.shape)))
        m8 = mask.to(device=inp.device, dtype=torch.uint8).contiguous()
    wctx = torch.empty_like(inp)
    attn = torch.empty((B, T, ih, iw), device=inp.device, dtype=torch.float32)
    lib = _lib.load()
    if tensor_cores and lib.tgsr_attn_f32_tc_supported(B, idf, cdf, ih * iw, T, int(m8 is not None)):
        # tensor-core path (3xTF32, fp32 accuracy): per-caption projection workspace + persistent pixel kernel
        ws = torch.empty(lib.tgsr_attn_f32_ws_floats(B, idf, T), device=inp.device, dtype=torch.float32)
        call('tgsr_attn_fwd_f32_tc', inp.device, ptr(inp), ptr(context), ptr(w_ctx), ptr(m8), ptr(ws), ptr(wctx), ptr(attn),
             B, idf, cdf, ih * iw, T)
    else:
        call('tgsr_attn_fwd_f32', inp.device, ptr(inp), ptr(context), ptr(w_ctx), ptr(m8), ptr(wctx), ptr(attn),
             B, idf, cdf, ih * iw, T)
    return wctx, attn


# ------------------------------------------------------------------ a2
def func_attention_fwd(query: Tensor, context: Tensor, gamma1: float) -> Tuple[Tensor, Tensor]:
    """func_attention (GlobalAttention.py:33-74), forward only.  At the DAMSM shape (ndf 256, <= 24 words) it runs as
    the diagonal of the fused pair kernel of the matching losses (csrc/damsm.cu); other shapes keep the generic kernel."""
    query = _f32c(query, 'query'); context = _f32c(context, 'context')
    B, ndf, Tq = query.shape
    ih, iw = context.shape[2], context.shape[3]
    wctx = torch.empty_like(query)
    attn = torch.empty((B, Tq, ih, iw), device=query.device, dtype=torch.float32)
    if ndf == 256 and Tq <= 24 and ih * iw <= 1024 and B <= 65535:
        call('tgsr_damsm_pairs_fwd_f32', query.device, ptr(context), ptr(query), None, float(gamma1), 0.0, 1e-8, None, ptr(attn),
             None, ptr(wctx), B, B, ndf, ih * iw, Tq, 1)
    else:
        call('tgsr_func_attn_fwd_f32', query.device, ptr(query), ptr(context), float(gamma1), ptr(wctx), ptr(attn),
             B, ndf, Tq, ih * iw)
    return wctx, attn


# ------------------------------------------------------------------ a3/a4/a5/a8
def bn_fold(gamma: Tensor, beta: Tensor, mean: Tensor, var: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    scale = torch.empty_like(gamma); shift = torch.empty_like(gamma)
    call('tgsr_bn_fold_f32', gamma.device, ptr(gamma), ptr(beta), ptr(mean), ptr(var), float(eps), ptr(scale),
         ptr(shift), gamma.numel())
    return scale, shift


def conv2d_fused(x: Tensor, w: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], act: int,
                 residual: Optional[Tensor] = None, addend: Optional[Tensor] = None, out_scale: float = 1.0,
                 addend_scale: float = 0.0, up2x: bool = False) -> Tensor:
    x = _f32c(x, 'x'); w = _f32c(w, 'weight')
    B, Cin, H, W = x.shape
    Cout, Cin_w, k, k2 = w.shape
    if Cin_w != Cin or k != k2:
        raise ValueError('conv weight %s does not match input %s' % (tuple(w.shape), tuple(x.shape)))
    if act == ACT_GLU and Cout % 2 != 0:
        raise AssertionError('channels dont divide 2!')          # util.py:51
    Cf = Cout // 2 if act == ACT_GLU else Cout
    Ho, Wo = (2 * H, 2 * W) if up2x else (H, W)
    y = torch.empty((B, Cf, Ho, Wo), device=x.device, dtype=torch.float32)
    for nm, t in (('residual', residual), ('addend', addend)):
        if t is not None and tuple(t.shape) != tuple(y.shape):
            raise ValueError('%s shape %s != output shape %s' % (nm, tuple(t.shape), tuple(y.shape)))
    residual = None if residual is None else _f32c(residual, 'residual')
    addend = None if addend is None else _f32c(addend, 'addend')
    call('tgsr_conv2d_f32', x.device, ptr(x), ptr(w), ptr(scale), ptr(shift), ptr(residual), ptr(addend),
         float(out_scale), float(addend_scale), ptr(y), B, Cin, H, W, Cout, k, int(up2x), act)
    return y


# ------------------------------------------------------------------ a12/a13
def linear_glu(x: Tensor, w: Tensor, bias: Optional[Tensor], scale: Optional[Tensor] = None,
               shift: Optional[Tensor] = None) -> Tensor:
    x = _f32c(x, 'x'); w = _f32c(w, 'weight')
    B, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError('linear weight %s does not match input %s' % (tuple(w.shape), tuple(x.shape)))
    y = torch.empty((B, N // 2), device=x.device, dtype=torch.float32)
    call('tgsr_linear_glu_f32', x.device, ptr(x), ptr(w), ptr(bias), ptr(scale), ptr(shift), ptr(y), B, K, N)
    return y


def ca_reparam(mulogvar: Tensor, eps: Tensor) -> Tensor:
    B, two_c = mulogvar.shape
    c = torch.empty((B, two_c // 2), device=mulogvar.device, dtype=torch.float32)
    call('tgsr_ca_reparam_f32', mulogvar.device, ptr(mulogvar), ptr(_f32c(eps, 'eps')), ptr(c), B, two_c // 2)
    return c


# ------------------------------------------------------------------ a11
def lstm_bidir(captions: Tensor, cap_lens: Tensor, emb: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor,
               b_hh: Tensor, Tmax: int) -> Tuple[Tensor, Tensor]:
    """Embedding + bidirectional LSTM.  w_ih [2,4H,ninput], w_hh [2,4H,H], b_* [2,4H] (stacked fwd, reverse)."""
    dev = emb.device
    captions = captions.to(device=dev, dtype=torch.int64).contiguous()
    cap_lens = cap_lens.to(device=dev, dtype=torch.int64).contiguous()
    B, Tcap = captions.shape
    H = w_hh.shape[2]
    ntoken, ninput = emb.shape
    ws = torch.empty(B * Tcap * 8 * H + 8 * H * H, device=dev, dtype=torch.float32)
    words = torch.empty((B, 2 * H, Tmax), device=dev, dtype=torch.float32)
    sent = torch.empty((B, 2 * H), device=dev, dtype=torch.float32)
    call('tgsr_lstm_bidir_fwd_f32', dev, ptr(captions), ptr(cap_lens), ptr(_f32c(emb, 'emb')), ptr(_f32c(w_ih, 'w_ih')),
         ptr(_f32c(w_hh, 'w_hh')), ptr(_f32c(b_ih, 'b_ih')), ptr(_f32c(b_hh, 'b_hh')), ptr(ws), ptr(words), ptr(sent),
         B, Tcap, Tmax, ntoken, ninput, H)
    return words, sent


def lstm_prepare(emb: Tensor, w_ih: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """xtab [ntoken, 2, 4H] = emb @ w_ih^T + b_ih + b_hh (weight prep for lstm_bidir_tab)."""
    ntoken, ninput = emb.shape
    H = w_ih.shape[1] // 4
    xtab = torch.empty((ntoken, 2, 4 * H), device=emb.device, dtype=torch.float32)
    call('tgsr_lstm_prepare_f32', emb.device, ptr(_f32c(emb, 'emb')), ptr(_f32c(w_ih, 'w_ih')), ptr(_f32c(b_ih, 'b_ih')),
         ptr(_f32c(b_hh, 'b_hh')), ptr(xtab), ntoken, ninput, H)
    return xtab


def lstm_bidir_tab(captions: Tensor, cap_lens: Tensor, xtab: Tensor, w_hh: Tensor, Tmax: int) -> Tuple[Tensor, Tensor]:
    dev = xtab.device
    captions = captions.to(device=dev, dtype=torch.int64).contiguous()
    cap_lens = cap_lens.to(device=dev, dtype=torch.int64).contiguous()
    B, Tcap = captions.shape
    H = w_hh.shape[2]
    words = torch.empty((B, 2 * H, Tmax), device=dev, dtype=torch.float32)
    sent = torch.empty((B, 2 * H), device=dev, dtype=torch.float32)
    call('tgsr_lstm_bidir_fwd_tab_f32', dev, ptr(captions), ptr(cap_lens), ptr(xtab), ptr(_f32c(w_hh, 'w_hh')), ptr(words),
         ptr(sent), B, Tcap, Tmax, xtab.shape[0], H)
    return words, sent


# ------------------------------------------------------------------ eval-flow edge
def quantize_u8(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """uint8 image of trainer_objective.py:153-155: round(clip((x + 1) * 127.5, 0, 255)), same shape as x."""
    x = _f32c(x, 'x')
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    elif out.dtype != torch.uint8 or out.numel() != x.numel() or not out.is_contiguous() or out.device != x.device:
        raise ValueError('out must be a contiguous uint8 tensor with as many elements as x on the same device')
    import ctypes as C
    call('tgsr_quantize_u8', x.device, ptr(x), ptr(out), C.c_size_t(x.numel()))
    return out


def attention_word_stats(attn: Tensor, thresh: float = 0.0) -> Tuple[Tensor, Tensor, Tensor]:
    """Word-axis reductions of attention maps `attn` [B, T, ih, iw] in one pass on the device:
    (attn_max [B,1,ih,iw], attn_argmax int64 [B,1,ih,iw]) = attn.max(dim=1, keepdim=True) (miscc/utils.py:122-125; the same
    max-map weighs weight_MSE, miscc/losses.py:796) and conf [B, T+1] fp64 = the thresholded confidence sums of
    miscc/utils.py:389-392 (`np.sum(one_map * (one_map > 2*thresh))` over the 3-channel replica of the max-map, column 0,
    and of each word map, columns 1..T)."""
    attn = _f32c(attn, 'attn')
    if attn.dim() != 4:
        raise ValueError('expected attn [B,T,ih,iw]')
    B, T, ih, iw = attn.shape
    Q = ih * iw
    if Q % 4 != 0 or T > 255:
        raise ValueError('attention_word_stats needs ih*iw %% 4 == 0 and T <= 255, got %s' % (tuple(attn.shape),))
    maxv = torch.empty((B, 1, ih, iw), device=attn.device, dtype=torch.float32)
    arg = torch.empty((B, 1, ih, iw), device=attn.device, dtype=torch.uint8)
    conf = torch.empty((B, T + 1), device=attn.device, dtype=torch.float64)
    import ctypes as C
    call('tgsr_attn_word_stats_f32', attn.device, ptr(attn), ptr(maxv), ptr(arg), ptr(conf), B, T, Q, C.c_float(thresh), 3)
    return maxv, arg.long(), conf
