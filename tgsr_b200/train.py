"""Training-mode execution of the generator blocks (fp32 NCHW): autograd nodes whose forward AND backward are
libtgsr_sm100 kernels.  torch.autograd only schedules the backward; no ATen arithmetic runs on this path except
`torch.cat` of (h_code, c_code) (util.py:773,817) and gradient accumulation.

reference: the generator modules in train() mode (nn.BatchNorm2d batch statistics, util.py / model.py); the
reference's gradients come from autograd over cuDNN/cuBLAS -- these are the explicit kernels (csrc/train_ops.cu,
csrc/attention_bwd.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import ACT_GLU, ACT_NONE, ACT_TANH, call, ptr

Tensor = torch.Tensor


def _empty_like_shape(shape, ref: Tensor) -> Tensor:
    return torch.empty(shape, device=ref.device, dtype=torch.float32)


# ---------------------------------------------------------------- conv (raw, no BN / activation)
_CONV_PRECISION = 'fp32'


def set_conv_precision(mode: str) -> None:
    """'fp32': exact CUDA-core convolutions (gradient-parity tests).  'tf32': TF32 tensor-core convolutions wherever the
    kernels cover the shape (fp32 accumulate) -- the precision PyTorch's cuDNN convs use by default, i.e. what the
    reference's own GPU training computes."""
    global _CONV_PRECISION
    if mode not in ('fp32', 'tf32'):
        raise ValueError("conv precision must be 'fp32' or 'tf32'")
    _CONV_PRECISION = mode


def _tf32_ok(Cin: int, Cout: int, k: int, wgrad: bool) -> bool:
    return _CONV_PRECISION == 'tf32' and bool(_lib.load().tgsr_conv_tf32_supported(Cin, Cout, k, int(wgrad)))


def _conv_tf32(x: Tensor, w: Tensor, up2x: bool, dgrad: bool) -> Tensor:
    Cout, Cin, k, _ = w.shape
    wk = torch.empty(w.numel(), device=w.device, dtype=torch.int32)
    call('tgsr_conv_prep_tf32', w.device, ptr(w), ptr(wk), Cout, Cin, k, int(dgrad))
    B, Cx, H, W = x.shape
    N = Cin if dgrad else Cout
    Ho, Wo = (2 * H, 2 * W) if up2x else (H, W)
    y = torch.empty((B, N, Ho, Wo), device=x.device, dtype=torch.float32)
    call('tgsr_conv2d_tf32', x.device, ptr(x), ptr(wk), ptr(y), B, Cx, H, W, N, k, int(up2x))
    return y


def _conv_raw(x: Tensor, w: Tensor, up2x: bool) -> Tensor:
    if _tf32_ok(w.shape[1], w.shape[0], w.shape[2], False):
        return _conv_tf32(x, w, up2x, False)
    return ops.conv2d_fused(x, w, None, None, ACT_NONE, up2x=up2x)


def _conv_dgrad(dy: Tensor, w: Tensor, up2x: bool) -> Tensor:
    """Data gradient of conv(x [nearest x2 if up2x], w): the same conv kernel on dy with the transposed, flipped weights;
    the folded upsample becomes a 2x2 sum-pool."""
    Cout, Cin, k, _ = w.shape
    if _tf32_ok(Cout, Cin, k, False):
        dx = _conv_tf32(dy, w, False, True)
    else:
        wt = torch.empty((Cin, Cout, k, k), device=w.device, dtype=torch.float32)
        call('tgsr_conv_weight_flip_f32', w.device, ptr(w), ptr(wt), Cout, Cin, k)
        dx = ops.conv2d_fused(dy, wt, None, None, ACT_NONE)
    if up2x:
        B, _, Ho, Wo = dx.shape
        out = _empty_like_shape((B, Cin, Ho // 2, Wo // 2), dx)
        call('tgsr_sumpool2x2_f32', dx.device, ptr(dx), ptr(out), B * Cin, Ho // 2, Wo // 2)
        dx = out
    return dx


def _conv_wgrad(x: Tensor, dy: Tensor, w_shape, up2x: bool) -> Tensor:
    Cout, Cin, k, _ = w_shape
    B, _, H, W = x.shape
    dw = _empty_like_shape(tuple(w_shape), x)
    name = 'tgsr_conv2d_wgrad_tf32' if _tf32_ok(Cin, Cout, k, True) else 'tgsr_conv2d_wgrad_f32'
    call(name, x.device, ptr(x), ptr(dy), ptr(dw), B, Cin, H, W, Cout, k, int(up2x))
    return dw


class ConvFn(torch.autograd.Function):
    """y = conv_kxk(x [nearest-upsampled x2 if up2x], w), stride 1, same padding, no bias (util.py:62-69, :76)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, up2x: bool):
        x = x.contiguous(); w = w.contiguous()
        ctx.save_for_backward(x, w)
        ctx.up2x = up2x
        return _conv_raw(x, w, up2x)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = _conv_dgrad(dy, w, ctx.up2x) if ctx.needs_input_grad[0] else None
        dw = _conv_wgrad(x, dy, w.shape, ctx.up2x) if ctx.needs_input_grad[1] else None
        return dx, dw, None


# ---------------------------------------------------------------- BatchNorm (batch statistics) + GLU + residual
class BnActFn(torch.autograd.Function):
    """y = act(BN_train(x)) (+ residual).  Updates the module's running statistics like nn.BatchNorm2d.train()."""

    @staticmethod
    def forward(ctx, x: Tensor, gamma: Tensor, beta: Tensor, residual: Optional[Tensor], bn: nn.BatchNorm2d, act: int):
        x = x.contiguous()
        B, Cc, H, W = x.shape
        mean = _empty_like_shape((Cc,), x); invstd = _empty_like_shape((Cc,), x)
        track = bn.track_running_stats and bn.running_mean is not None
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        ws = torch.empty(2 * Cc, device=x.device, dtype=torch.float64)
        call('tgsr_bn_stats_f32', x.device, ptr(x), ptr(mean), ptr(invstd), ptr(bn.running_mean if track else None),
             ptr(bn.running_var if track else None), mom, float(bn.eps), ptr(ws), B, Cc, H * W)
        if track:
            _lib.bump_weights_epoch()          # running stats were written through raw pointers: folded-BN caches are stale
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        Cf = Cc // 2 if act == ACT_GLU else Cc
        y = _empty_like_shape((B, Cf, H, W), x)
        res = None if residual is None else residual.contiguous()
        call('tgsr_bn_act_f32', x.device, ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(res), ptr(y), B, Cc, H * W, act)
        ctx.save_for_backward(x, mean, invstd, gamma, beta)
        ctx.act = act
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, mean, invstd, gamma, beta = ctx.saved_tensors
        dy = dy.contiguous()
        B, Cc, H, W = x.shape
        dx = torch.empty_like(x); dg = torch.empty_like(gamma); db = torch.empty_like(beta)
        ws = torch.empty(2 * Cc, device=x.device, dtype=torch.float64)
        call('tgsr_bn_act_bwd_f32', x.device, ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(dy), ptr(dx), ptr(dg), ptr(db),
             ptr(ws), B, Cc, H * W, ctx.act)
        return dx, dg, db, (dy if ctx.has_res else None), None, None


# ---------------------------------------------------------------- image heads
class TanhHeadFn(torch.autograd.Function):
    """y = one * tanh(conv(x, w)) + a * srb   (conv_output model.py:223-226,280; GET_IMAGE_G util.py:899-901 with srb None)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, srb: Optional[Tensor], one: float, a: float):
        x = x.contiguous(); w = w.contiguous()
        srb_c = None if srb is None else srb.contiguous()
        y = ops.conv2d_fused(x, w, None, None, ACT_TANH, addend=srb_c, out_scale=one, addend_scale=a)
        ctx.save_for_backward(x, w, y, srb_c)
        ctx.one, ctx.a = float(one), float(a)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, w, y, srb = ctx.saved_tensors
        dy = dy.contiguous()
        dconv = torch.empty_like(y)
        dsrb = torch.empty_like(y) if (srb is not None and ctx.needs_input_grad[2]) else None
        call('tgsr_tanh_head_bwd_f32', y.device, ptr(y), ptr(srb), ptr(dy), ctx.one, ctx.a, ptr(dconv), ptr(dsrb),
             C.c_size_t(y.numel()))
        dx = _conv_dgrad(dconv, w, False) if ctx.needs_input_grad[0] else None
        dw = _conv_wgrad(x, dconv, w.shape, False) if ctx.needs_input_grad[1] else None
        return dx, dw, dsrb, None, None


# ---------------------------------------------------------------- word-region attention
class AttentionFn(torch.autograd.Function):
    """GlobalAttentionGeneral.forward (GlobalAttention.py:87-130) with gradients to the pixel features and the word
    projection weight.  `context` (the frozen text encoder's word embeddings) receives no gradient."""

    @staticmethod
    def forward(ctx, inp: Tensor, w_ctx: Tensor, context: Tensor, mask: Optional[Tensor]):
        inp = inp.contiguous(); context = context.contiguous()
        w2 = w_ctx.reshape(w_ctx.shape[0], -1).contiguous()
        wctx, attn = ops.attention_fwd(inp, context, w2, mask)
        ctx.save_for_backward(inp, w2, context, attn)
        ctx.w_shape = tuple(w_ctx.shape)
        ctx.set_materialize_grads(False)
        return wctx, attn

    @staticmethod
    def backward(ctx, d_wctx: Optional[Tensor], d_attn: Optional[Tensor]):
        inp, w2, context, attn = ctx.saved_tensors
        B, idf, ih, iw = inp.shape
        cdf, T = context.shape[1], context.shape[2]
        if d_wctx is None:
            d_wctx = torch.zeros_like(inp)
        d_wctx = d_wctx.contiguous()
        d_attn = None if d_attn is None else d_attn.contiguous()
        lib = _lib.load()
        ws = torch.empty(2 * lib.tgsr_attn_f32_ws_floats(B, idf, T), device=inp.device, dtype=torch.float32)
        d_inp = torch.empty_like(inp); d_w = torch.empty_like(w2)
        call('tgsr_attn_bwd_f32', inp.device, ptr(inp), ptr(context), ptr(w2), ptr(attn), ptr(d_wctx), ptr(d_attn), ptr(ws),
             ptr(d_inp), ptr(d_w), B, idf, cdf, ih * iw, T)
        return d_inp, d_w.view(ctx.w_shape), None, None


# ---------------------------------------------------------------- block-level entry used by blocks._conv
def wants_grad(x: Tensor, conv: nn.Conv2d, bn) -> bool:
    """True when the call must run on the training path: the layer is in train() mode and either has a BatchNorm in
    train() mode (batch statistics) or autograd is recording and the input or the weights need a gradient."""
    if not conv.training:
        return False
    if bn is not None:
        return bn.training
    return torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad)


def conv_bn_act(x: Tensor, conv: nn.Conv2d, bn, act: int, residual: Optional[Tensor] = None, addend: Optional[Tensor] = None,
                out_scale: float = 1.0, addend_scale: float = 0.0, up2x: bool = False) -> Tensor:
    from . import is_16bit
    if is_16bit():
        raise RuntimeError('tgsr_b200: the training path is fp32 only; call tgsr_b200.set_precision("fp32")')
    if bn is not None:
        if not bn.training:
            raise RuntimeError('tgsr_b200: gradients through an eval-mode BatchNorm are not implemented; call .train()')
        if addend is not None or out_scale != 1.0:
            raise NotImplementedError('addend/out_scale are image-head options (no BatchNorm)')
        z = ConvFn.apply(x, conv.weight, up2x)
        return BnActFn.apply(z, bn.weight, bn.bias, residual, bn, act)
    if act == ACT_TANH:
        if residual is not None or up2x:
            raise NotImplementedError('tanh heads have no residual / upsample')
        return TanhHeadFn.apply(x, conv.weight, addend, out_scale, addend_scale)
    if act == ACT_NONE and addend is None and out_scale == 1.0 and residual is None:
        return ConvFn.apply(x, conv.weight, up2x)
    raise NotImplementedError('training path: unsupported block form (act %d without BatchNorm)' % act)
