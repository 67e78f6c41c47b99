"""Training-mode execution of the generator on BLOCKED 16-bit activations: the bf16 / tcgen05 twin of train.py.

Activations between layers are bf16 tensors [B, C/8, H, W, 8] (the layout of csrc/conv_tc.cu).  Per layer:
  * conv3x3 forward and data gradient: the tcgen05 implicit-GEMM kernel (bf16 operands packed from the fp32 MASTER weights
    every step, fp32 accumulation in TMEM); the data gradient is the same kernel on dz with the transposed, flipped weights
    (input channels beyond 64 are fed in two accumulating passes through the kernel's skip input);
  * weight gradient: tcgen05 as well (csrc/wgrad_tc.cu): both operands read from the blocked tensors as MN-major UMMA
    operands with the pixels as K, fp32 accumulation in TMEM over all tiles of a CTA;
  * BatchNorm (batch statistics) + GLU + skip, forward and backward: csrc/train16.cu, arithmetic in fp32 / fp64;
  * nearest x2 upsample / 2x2 sum-pool adjoint: csrc/train16.cu.
The image heads (3 output channels), im2f / convin (3 input channels) and the word attention keep their fp32 kernels
(train.py); `to_blocked` / `to_nchw` convert at those edges and are differentiable.

reference: the generator modules in train() mode under autograd (util.py:45-130,726-823, model.py:34-78,212-298); cfg 5 of
BASELINE.json asks for this step in bf16.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib, train
from ._lib import ACT_GLU, ACT_NONE, ACT_TANH, FMT_BF16, call, ptr

Tensor = torch.Tensor
FMT = FMT_BF16
DT = torch.bfloat16


def _blk(B, C, H, W, dev) -> Tensor:
    return torch.empty((B, C // 8, H, W, 8), device=dev, dtype=DT)


# ---------------------------------------------------------------- layout edges (differentiable)
class ToBlockedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor):
        x = x.contiguous()
        B, C, H, W = x.shape
        y = _blk(B, C, H, W, x.device)
        call('tgsr_nchw_to_blocked', x.device, ptr(x), ptr(y), B, C, H, W, C // 8, 0, FMT)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        return ToNchwFn.forward(None, dy)


class ToNchwFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xb: Tensor):
        xb = xb.contiguous()
        B, P, H, W, _ = xb.shape
        y = torch.empty((B, P * 8, H, W), device=xb.device, dtype=torch.float32)
        call('tgsr_blocked_to_nchw', xb.device, ptr(xb), ptr(y), B, P * 8, H, W, P, 0, FMT)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        return ToBlockedFn.forward(None, dy)


def to_blocked(x: Tensor) -> Tensor:
    return ToBlockedFn.apply(x)


def to_nchw(xb: Tensor) -> Tensor:
    return ToNchwFn.apply(xb)


# ---------------------------------------------------------------- conv3x3 (raw: no BN, no activation)
def _pack(w: Tensor) -> Tensor:
    Cout, Cin = w.shape[0], w.shape[1]
    n = _lib.load().tgsr_packed_conv3x3_elems(Cout, Cin, 0)
    packed = torch.empty(n, device=w.device, dtype=DT)
    call('tgsr_pack_conv3x3', w.device, ptr(w), None, None, ptr(packed), Cout, Cin, 0, ACT_NONE, FMT)
    return packed


def _conv_tc(xb: Tensor, p0: int, Cin: int, w: Tensor, residual: Optional[Tensor] = None) -> Tensor:
    """tcgen05 conv of planes [p0, p0 + Cin/8) of xb with fp32 weights w [Cout, Cin, 3, 3] (+ residual) -> blocked bf16."""
    B, P, H, W, _ = xb.shape
    Cout = w.shape[0]
    y = _blk(B, Cout, H, W, xb.device)
    rP = 0 if residual is None else residual.shape[1]
    call('tgsr_conv3x3_tc', xb.device, ptr(xb), P, p0, ptr(_pack(w.contiguous())), ptr(residual), rP, 0, ptr(y), Cout // 8, 0,
         B, Cin, H, W, Cout, 0, ACT_NONE, FMT, 0)
    return y


_TC_SHAPES = {(64, 128), (64, 64), (32, 64), (32, 32)}
import os as _os
WGRAD_TC = _os.environ.get('TGSR_WGRAD_TC', '1') != '0'      # 0: weight gradients on the TF32 kernel (fp32 NCHW copies)


class Conv16Fn(torch.autograd.Function):
    """z = conv3x3(x, w), stride 1, pad 1, no bias (util.py:62-65) on blocked bf16 activations."""

    @staticmethod
    def forward(ctx, xb: Tensor, w: Tensor):
        xb = xb.contiguous()
        Cout, Cin = w.shape[0], w.shape[1]
        if (Cin, Cout) not in _TC_SHAPES or xb.shape[1] * 8 != Cin or xb.shape[2] % 16 or xb.shape[3] % 8:
            raise NotImplementedError('train16 conv: unsupported shape Cin %d Cout %d at %s' % (Cin, Cout, tuple(xb.shape)))
        ctx.save_for_backward(xb, w)
        return _conv_tc(xb, 0, Cin, w.detach())

    @staticmethod
    def backward(ctx, dz: Tensor):
        xb, w = ctx.saved_tensors
        dz = dz.contiguous()
        Cout, Cin = w.shape[0], w.shape[1]
        B, _, H, W, _ = xb.shape
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # data gradient = conv of dz with the transposed, flipped weights wt [Cin, Cout, 3, 3]; the kernel takes at most 64
            # input channels, so Cout = 128 goes in two passes, the second accumulating onto the first through the skip input
            wt = torch.empty((Cin, Cout, 3, 3), device=w.device, dtype=torch.float32)
            call('tgsr_conv_weight_flip_f32', w.device, ptr(w.detach().contiguous()), ptr(wt), Cout, Cin, 3)
            step = Cin if Cout == 2 * Cin else Cout               # input-channel slice per pass: an instantiated (Cin', Cout') pair
            if (step, Cin) not in _TC_SHAPES:
                step = Cout
            for c0 in range(0, Cout, step):
                dx = _conv_tc(dz, c0 // 8, step, wt[:, c0:c0 + step].contiguous(), residual=dx)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            if WGRAD_TC:                                            # tcgen05, straight from the blocked bf16 tensors
                call('tgsr_conv3x3_wgrad_tc', w.device, ptr(xb), Cin // 8, 0, ptr(dz), Cout // 8, 0, ptr(dw), ptr(torch.empty_like(dw)), B, Cin, H,
                     W, Cout, FMT)
            else:                                                   # TF32 kernel on fp32 NCHW copies
                x32 = ToNchwFn.forward(None, xb); dz32 = ToNchwFn.forward(None, dz)
                name = 'tgsr_conv2d_wgrad_tf32' if _lib.load().tgsr_conv_tf32_supported(Cin, Cout, 3, 1) else 'tgsr_conv2d_wgrad_f32'
                call(name, w.device, ptr(x32), ptr(dz32), ptr(dw), B, Cin, H, W, Cout, 3, 0)
        return dx, dw


# ---------------------------------------------------------------- BatchNorm (batch statistics) + GLU + residual
class BnAct16Fn(torch.autograd.Function):
    """y = act(BN_train(z)) (+ residual) on blocked bf16; updates the module's running statistics like nn.BatchNorm2d."""

    @staticmethod
    def forward(ctx, z: Tensor, gamma: Tensor, beta: Tensor, residual: Optional[Tensor], bn: nn.BatchNorm2d, act: int):
        z = z.contiguous()
        B, P, H, W, _ = z.shape
        Cc = P * 8
        dev = z.device
        mean = torch.empty(Cc, device=dev, dtype=torch.float32); invstd = torch.empty(Cc, device=dev, dtype=torch.float32)
        ws = torch.empty(2 * Cc, device=dev, dtype=torch.float64)
        track = bn.track_running_stats and bn.running_mean is not None
        mom = 0.1 if bn.momentum is None else float(bn.momentum)
        call('tgsr_bn16_stats', dev, ptr(z), ptr(mean), ptr(invstd), ptr(bn.running_mean if track else None),
             ptr(bn.running_var if track else None), mom, float(bn.eps), ptr(ws), B, Cc, H * W, FMT)
        if track:
            _lib.bump_weights_epoch()
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        Cf = Cc // 2 if act == ACT_GLU else Cc
        y = _blk(B, Cf, H, W, dev)
        res = None if residual is None else residual.contiguous()
        call('tgsr_bn16_act', dev, ptr(z), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(res), ptr(y), B, Cc, H * W, act, FMT)
        ctx.save_for_backward(z, mean, invstd, gamma, beta)
        ctx.act, ctx.has_res = act, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        z, mean, invstd, gamma, beta = ctx.saved_tensors
        dy = dy.contiguous()
        B, P, H, W, _ = z.shape
        Cc = P * 8
        dz = torch.empty_like(z); dg = torch.empty_like(gamma); db = torch.empty_like(beta)
        ws = torch.empty(2 * Cc, device=z.device, dtype=torch.float64)
        call('tgsr_bn16_act_bwd', z.device, ptr(z), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(dy), ptr(dz), ptr(dg), ptr(db),
             ptr(ws), B, Cc, H * W, ctx.act, FMT)
        return dz, dg, db, (dy if ctx.has_res else None), None, None


class Up2x16Fn(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, mode='nearest') (util.py:76) on blocked bf16; backward = 2x2 sum-pool."""

    @staticmethod
    def forward(ctx, xb: Tensor):
        xb = xb.contiguous()
        B, P, H, W, _ = xb.shape
        y = torch.empty((B, P, 2 * H, 2 * W, 8), device=xb.device, dtype=DT)
        call('tgsr_upsample2x_blocked', xb.device, ptr(xb), ptr(y), B, P, H, W)
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        dy = dy.contiguous()
        B, P, H2, W2, _ = dy.shape
        dx = torch.empty((B, P, H2 // 2, W2 // 2, 8), device=dy.device, dtype=DT)
        call('tgsr_sumpool2x2_blocked', dy.device, ptr(dy), ptr(dx), B, P, H2 // 2, W2 // 2, FMT)
        return dx


# ---------------------------------------------------------------- blocks (util.py:74-130)
def conv_bn_act(xb: Tensor, conv: nn.Conv2d, bn: nn.BatchNorm2d, act: int, residual: Optional[Tensor] = None) -> Tensor:
    if not bn.training:
        raise RuntimeError('tgsr_b200: gradients through an eval-mode BatchNorm are not implemented; call .train()')
    return BnAct16Fn.apply(Conv16Fn.apply(xb, conv.weight), bn.weight, bn.bias, residual, bn, act)


def res_block(xb: Tensor, rb) -> Tensor:
    blk = rb.block
    y = conv_bn_act(xb, blk[0], blk[1], ACT_GLU)
    return conv_bn_act(y, blk[3], blk[4], ACT_NONE, residual=xb)


def body_noskip(xb: Tensor, seq) -> Tensor:
    return conv_bn_act(conv_bn_act(xb, seq[0], seq[1], ACT_GLU), seq[3], seq[4], ACT_NONE)


class UpConv16Fn(torch.autograd.Function):
    """z = conv3x3(nearest_x2(x), w) (util.py:74-80) without materialising the upsampled tensor in the forward pass: the
    4-phase sub-pixel kernel of the eval path (conv_tc.cu MODE_UP, act-free instantiation; 16/36 of the MMAs).  Backward is
    the explicit formulation: the data gradient at the output resolution followed by the 2x2 sum-pool (the adjoint of the
    upsample), the weight gradient against the re-created upsampled input."""

    @staticmethod
    def forward(ctx, xb: Tensor, w: Tensor):
        xb = xb.contiguous()
        Cout, Cin = w.shape[0], w.shape[1]
        B, P, H, W, _ = xb.shape
        ctx.save_for_backward(xb, w)
        n = _lib.load().tgsr_packed_conv3x3_elems(Cout, Cin, 1)
        packed = torch.empty(n, device=w.device, dtype=DT)
        wc = w.detach().contiguous()
        call('tgsr_pack_conv3x3', w.device, ptr(wc), None, None, ptr(packed), Cout, Cin, 1, ACT_NONE, FMT)
        z = _blk(B, Cout, 2 * H, 2 * W, xb.device)
        call('tgsr_conv3x3_tc', xb.device, ptr(xb), P, 0, ptr(packed), None, 0, 0, ptr(z), Cout // 8, 0, B, Cin, H, W, Cout, 1,
             ACT_NONE, FMT, 0)
        return z

    @staticmethod
    def backward(ctx, dz: Tensor):
        xb, w = ctx.saved_tensors
        dx = dw = None
        with torch.no_grad():
            x_up = Up2x16Fn.forward(None, xb)
            need = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])

            class _C:                                              # reuse Conv16Fn.backward on the explicit tensors
                saved_tensors = (x_up, w)
                needs_input_grad = need
            dx_up, dw = Conv16Fn.backward(_C, dz)
            if dx_up is not None:
                dx = Up2x16Fn.backward(None, dx_up)
        return dx, dw


_UPCONV = _os.environ.get('TGSR_TRAIN_UPCONV', '1') != '0'       # 0: explicit upsample + conv in the forward pass too


def up_block(xb: Tensor, seq) -> Tensor:
    conv, bn = seq[1], seq[2]
    Cout, Cin = conv.weight.shape[0], conv.weight.shape[1]
    if _UPCONV and Cout == 64 and Cin in (32, 64) and xb.shape[1] * 8 == Cin and xb.shape[2] % 16 == 0 and xb.shape[3] % 8 == 0:
        if not bn.training:
            raise RuntimeError('tgsr_b200: gradients through an eval-mode BatchNorm are not implemented; call .train()')
        return BnAct16Fn.apply(UpConv16Fn.apply(xb, conv.weight), bn.weight, bn.bias, None, bn, ACT_GLU)
    return conv_bn_act(Up2x16Fn.apply(xb), conv, bn, ACT_GLU)


def conv_in(seq, x: Tensor, x_sub: Optional[Tensor] = None) -> Tensor:
    """[conv3x3(3, 2C), BN, GLU] from the fp32 image: fp32 training kernels (K = 27), then into the blocked layout."""
    if x_sub is not None:
        x = torch.sub(x, x_sub)
    z = train.ConvFn.apply(x, seq[0].weight, False)
    return to_blocked(train.BnActFn.apply(z, seq[1].weight, seq[1].bias, None, seq[1], ACT_GLU))


class Head16Fn(torch.autograd.Function):
    """y = one * act(conv_kxk(h, w)) + a * srb with 3 output channels (GET_IMAGE_G[_noAct].img util.py:894-919, conv_output +
    Tanh model.py:223-226,280): forward on the tcgen05 head kernel straight from the blocked features; the backward (K = 3
    input channels for the data gradient) uses the fp32 kernels of train.py on the NCHW copy of h."""

    @staticmethod
    def forward(ctx, hb: Tensor, w: Tensor, srb: Optional[Tensor], one: float, a: float, tanh: bool):
        hb = hb.contiguous()
        B, P, H, W, _ = hb.shape
        k = w.shape[2]
        packed = torch.empty(k * 16 * w.shape[1], device=w.device, dtype=DT)
        call('tgsr_pack_head', w.device, ptr(w.detach().contiguous()), ptr(packed), w.shape[1], k, FMT)
        y = torch.empty((B, 3, H, W), device=hb.device, dtype=torch.float32)
        srb_c = None if srb is None else srb.contiguous()
        from .fast import nonfinite_flag
        call('tgsr_img_head_tc', hb.device, ptr(hb), P, 0, ptr(packed), ptr(srb_c), float(one), float(a), ptr(y), B, P * 8, H, W, k,
             ACT_TANH if tanh else ACT_NONE, FMT, ptr(nonfinite_flag(hb.device)))
        ctx.save_for_backward(hb, w, y, srb_c)
        ctx.one, ctx.a, ctx.tanh = float(one), float(a), tanh
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        import ctypes as C
        hb, w, y, srb = ctx.saved_tensors
        dy = dy.contiguous()
        dsrb = None
        if ctx.tanh:
            dconv = torch.empty_like(y)
            dsrb = torch.empty_like(y) if (srb is not None and ctx.needs_input_grad[2]) else None
            call('tgsr_tanh_head_bwd_f32', y.device, ptr(y), ptr(srb), ptr(dy), ctx.one, ctx.a, ptr(dconv), ptr(dsrb),
                 C.c_size_t(y.numel()))
        else:
            dconv = dy if ctx.one == 1.0 else dy * ctx.one
            if srb is not None and ctx.needs_input_grad[2]:
                dsrb = dy * ctx.a
        dh = None
        if ctx.needs_input_grad[0]:
            if w.shape[1] == 32:                                   # tensor cores, straight into the blocked layout
                B, P, H, W, _ = hb.shape
                dh = torch.empty_like(hb)
                dconv, wc = dconv.contiguous(), w.detach().contiguous()      # raw pointers: keep both alive across the call
                call('tgsr_head_dgrad16', hb.device, ptr(dconv), ptr(wc), ptr(dh), P, 0, B, P * 8, H, W, w.shape[2], FMT)
            else:
                dh = ToBlockedFn.forward(None, train._conv_dgrad(dconv, w.detach(), False))
        dw = None
        if ctx.needs_input_grad[1]:
            if w.shape[1] == 32:
                B, P, H, W, _ = hb.shape
                dw = torch.empty(tuple(w.shape), device=w.device, dtype=torch.float32)
                call('tgsr_head_wgrad16', hb.device, ptr(hb), P, 0, ptr(dconv.contiguous()), ptr(dw), B, P * 8, H, W, w.shape[2], FMT)
            else:
                dw = train._conv_wgrad(ToNchwFn.forward(None, hb), dconv, w.shape, False)
        return dh, dw, dsrb, None, None, None


def image_head(hb: Tensor, head_seq, addend: Optional[Tensor] = None, a: float = 0.0, one: float = 1.0) -> Tensor:
    tanh = any(isinstance(m, nn.Tanh) for m in head_seq)
    return Head16Fn.apply(hb, head_seq[0].weight, addend, one, a, tanh)


def stage(st, hb: Optional[Tensor], LR: Optional[Tensor], words: Tensor, mask):
    """INIT_STAGE_GImgup (LR given) / NEXT_STAGE_G (hb given) in train mode; util.py:763-777 / :807-823."""
    if hb is None:
        hb = conv_in(st.im2f, LR)
    st.att.applyMask(mask)
    c, att = train.AttentionFn.apply(to_nchw(hb), st.att.conv_context.weight, words, mask)
    x = torch.cat((hb, to_blocked(c)), 1)                     # util.py:773,817: channel concat = plane concat
    for rb in st.residual:
        x = res_block(x, rb)
    return up_block(x, st.upsample), att


def g_sr_net_low(net, LR, sent, words, mask, stages, heads):
    """G_SR_NET_low.forward in train mode (model.py:48-78 / models16.py:16-39)."""
    imgs: List[Tensor] = []
    atts: List[Tensor] = []
    c_code, mu, logvar = net.ca_net(sent)
    h = None
    for i, (st, hd) in enumerate(zip(stages, heads)):
        h, att = stage(st, h, LR if i == 0 else None, words, mask)
        imgs.append(image_head(h, hd.img))
        atts.append(att)
    return imgs, atts, mu, logvar


def netg_high(net, x: Tensor, x_sub: Optional[Tensor], SRb, a: float, one: float, tail) -> List[Tensor]:
    """Trunk of NetG_high / NetG_highweight in train mode (model.py:264-298)."""
    h = conv_in(net.convin, x, x_sub)
    for rb in net.residual:
        h = res_block(h, rb)
    h = up_block(h, net.upscale2x)
    ims = [image_head(h, net.conv_output, SRb[0], a, one)]
    for k, (body, up) in enumerate(tail):
        h = up_block(body_noskip(h, body), up)
        ims.append(image_head(h, net.conv_output, SRb[k + 1], a, one))
    return ims
