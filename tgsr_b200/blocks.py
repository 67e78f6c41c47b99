"""Generator building blocks with the reference's constructors, forward signatures and
state-dict keys (util.py).  torch.nn leaf modules (Conv2d, BatchNorm2d, Linear, ...) are used
ONLY as parameter holders so that ``state_dict()`` / ``load_state_dict()`` match the shipped
checkpoints key for key; their own forward() is never called -- every block's forward launches
the fused sm_100a kernels through the C ABI.

Eval-mode BatchNorm is folded to (scale, shift) once per weight version and cached outside the
state dict.  In train() mode (or whenever autograd records and a gradient is needed) the fp32 blocks run the
training kernels of tgsr_b200/train.py: batch-statistics BatchNorm and explicit backward kernels.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .attention import GlobalAttentionGeneral as ATT_NET
from .attention import conv1x1  # noqa: F401  (re-export, util.py:56)
from .config import cfg
from ._lib import ACT_GLU, ACT_NONE, ACT_TANH, tensor_key


# ---------------------------------------------------------------- derived-parameter cache
def _folded_bn(bn: Optional[nn.Module]):
    """(scale, shift) for an eval-mode BatchNorm, cached on the module (not a buffer)."""
    if bn is None:
        return None, None
    if bn.training:
        raise RuntimeError('tgsr_b200: training-mode BatchNorm is not implemented on the CUDA path; call .eval()')
    src = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple(tensor_key(t) for t in src)
    cached = bn.__dict__.get('_tgsr_fold')
    if cached is None or cached[0] != key:
        scale, shift = ops.bn_fold(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)
        cached = (key, scale, shift)
        bn.__dict__['_tgsr_fold'] = cached
    return cached[1], cached[2]


def _conv(x, conv: nn.Conv2d, bn, act, **kw):
    from . import train
    if train.wants_grad(x, conv, bn):               # train() mode / autograd recording: explicit fwd+bwd kernels
        return train.conv_bn_act(x, conv, bn, act, **kw)
    scale, shift = _folded_bn(bn)
    return ops.conv2d_fused(x, conv.weight.detach(), scale, shift, act, **kw)


def _bf16():
    from . import is_16bit
    return is_16bit()


def _fmt():
    from . import fmt_for
    return fmt_for('low')


def _tc_ok(conv: nn.Conv2d, act, up2x, x) -> bool:
    """bf16 mode and the layer is inside the tcgen05 kernel's envelope (else: fp32 CUDA-core kernel)."""
    from . import tc
    return _bf16() and tc.tc_supported(conv.weight.shape[1], conv.weight.shape[0], act, up2x, x.shape[2], x.shape[3])


# ---------------------------------------------------------------- leaf helpers (util.py:45-69)
class GLU(nn.Module):
    """util.py:45-53.  Inside the blocks below GLU is fused into the conv epilogue; this
    standalone module exists for constructor/state-dict parity."""

    def forward(self, x):
        nc = x.size(1)
        assert nc % 2 == 0, 'channels dont divide 2!'
        raise NotImplementedError('tgsr_b200: GLU only exists fused into conv/linear kernels')


def conv3x3(in_planes, out_planes):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=1, padding=1, bias=False)


def conv5x5(in_planes, out_planes):
    return nn.Conv2d(in_planes, out_planes, kernel_size=5, stride=1, padding=2, bias=False)


class _ConvBnGlu(nn.Sequential):
    """[conv3x3, (BatchNorm2d), GLU] -> one kernel (Block3x3_relu, util.py:102-107)."""

    def forward(self, x):
        mods = list(self)
        bn = mods[1] if isinstance(mods[1], nn.BatchNorm2d) else None
        if _tc_ok(mods[0], ACT_GLU, False, x):
            from . import fast, tc
            return tc.to_nchw(fast.conv_bn_act(tc.to_blocked(x, fmt=_fmt()), mods[0], bn, ACT_GLU))
        return _conv(x, mods[0], bn, ACT_GLU)


class _UpBlock(nn.Sequential):
    """[Upsample(nearest,2), conv3x3, (BatchNorm2d), GLU] -> one kernel; the upsample is folded
    into the conv's input indexing (upBlock, util.py:74-80; upBlocknoBN :83-88)."""

    def forward(self, x):
        mods = list(self)
        bn = mods[2] if isinstance(mods[2], nn.BatchNorm2d) else None
        if _tc_ok(mods[1], ACT_GLU, True, x):
            from . import fast, tc
            return tc.to_nchw(fast.up_block(tc.to_blocked(x, fmt=_fmt()), self))
        return _conv(x, mods[1], bn, ACT_GLU, up2x=True)


class _ConvBnGluConvBn(nn.Sequential):
    """[conv3x3, BN, GLU, conv3x3, BN] without skip (residual24/48, model.py:229-232)."""

    def forward(self, x, residual=None):
        mods = list(self)
        if isinstance(mods[1], nn.BatchNorm2d) and _tc_ok(mods[0], ACT_GLU, False, x) \
                and _tc_ok(mods[3], ACT_NONE, False, x):
            from . import fast, tc
            xb = tc.to_blocked(x, fmt=_fmt())
            y = fast.conv_bn_act(xb, mods[0], mods[1], ACT_GLU)
            rb = xb if residual is x else (None if residual is None else tc.to_blocked(residual, fmt=_fmt()))
            return tc.to_nchw(fast.conv_bn_act(y, mods[3], mods[4], ACT_NONE, residual=rb))
        if isinstance(mods[1], nn.BatchNorm2d):
            y = _conv(x, mods[0], mods[1], ACT_GLU)
            return _conv(y, mods[3], mods[4], ACT_NONE, residual=residual)
        y = _conv(x, mods[0], None, ACT_GLU)
        return _conv(y, mods[2], None, ACT_NONE, residual=residual)


class _ImageHead(nn.Sequential):
    """[conv, (Tanh)] image head.  ``forward(h, addend, a, one)`` fuses ``one*y + a*addend``."""

    def forward(self, h, addend=None, a: float = 0.0, one: float = 1.0):
        act = ACT_TANH if any(isinstance(m, nn.Tanh) for m in self) else ACT_NONE
        if _bf16() and h.shape[1] == 32 and self[0].weight.shape[0] == 3 and h.shape[3] % 8 == 0:
            from . import fast, tc
            return fast.image_head(tc.to_blocked(h, fmt=_fmt()), self, addend, a, one)
        return _conv(h, self[0], None, act, addend=addend, out_scale=one, addend_scale=a)


def upBlock(in_planes, out_planes):
    return _UpBlock(nn.Upsample(scale_factor=2, mode='nearest'), conv3x3(in_planes, out_planes * 2),
                    nn.BatchNorm2d(out_planes * 2), GLU())


def upBlocknoBN(in_planes, out_planes):
    return _UpBlock(nn.Upsample(scale_factor=2, mode='nearest'), conv3x3(in_planes, out_planes * 2), GLU())


def Block3x3_relu(in_planes, out_planes):
    return _ConvBnGlu(conv3x3(in_planes, out_planes * 2), nn.BatchNorm2d(out_planes * 2), GLU())


class ResBlock(nn.Module):
    """util.py:110-130: x + BN(conv(GLU(BN(conv(x))))) as two fused kernels."""

    def __init__(self, channel_num, batchnorm=True):
        super().__init__()
        if batchnorm:
            self.block = _ConvBnGluConvBn(conv3x3(channel_num, channel_num * 2), nn.BatchNorm2d(channel_num * 2),
                                          GLU(), conv3x3(channel_num, channel_num), nn.BatchNorm2d(channel_num))
        else:
            self.block = _ConvBnGluConvBn(conv3x3(channel_num, channel_num * 2), GLU(),
                                          conv3x3(channel_num, channel_num))

    def forward(self, x):
        return self.block(x, residual=x)


# ---------------------------------------------------------------- CA_NET (util.py:372-400)
class CA_NET(nn.Module):
    def __init__(self):
        super().__init__()
        self.t_dim = cfg.TEXT.EMBEDDING_DIM
        self.c_dim = cfg.GAN.CONDITION_DIM
        self.fc = nn.Linear(self.t_dim, self.c_dim * 4, bias=True)
        self.relu = GLU()

    def _mulogvar(self, text_embedding):
        from ._lib import no_grad_path
        no_grad_path('CA_NET(text_embedding)', text_embedding)
        if self.training and torch.is_grad_enabled() and self.fc.weight.requires_grad and not CA_NET._warned:
            import warnings
            CA_NET._warned = True
            warnings.warn('tgsr_b200: CA_NET has no backward kernel -- mu / logvar carry no gradient, so KL_loss on them trains '
                          'nothing (c_code is unused by the SR nets, model.py:51-52); freeze ca_net or leave it out of the optimiser')
        return ops.linear_glu(text_embedding, self.fc.weight.detach(), self.fc.bias.detach())

    _warned = False

    def encode(self, text_embedding):
        x = self._mulogvar(text_embedding)
        return x[:, :self.c_dim], x[:, self.c_dim:]

    def reparametrize(self, mu, logvar):
        eps = torch.randn(mu.shape, device=mu.device, dtype=torch.float32)     # util.py:390-393 normal_()
        return ops.ca_reparam(torch.cat((mu, logvar), 1).contiguous(), eps)

    def forward(self, text_embedding):
        x = self._mulogvar(text_embedding)
        eps = torch.randn((x.shape[0], self.c_dim), device=x.device, dtype=torch.float32)
        c_code = ops.ca_reparam(x, eps)
        return c_code, x[:, :self.c_dim], x[:, self.c_dim:]


# ---------------------------------------------------------------- generator stages
class _FcBnGlu(nn.Sequential):
    """[Linear(no bias), BatchNorm1d, GLU] (INIT_STAGE_G.fc, util.py:417-420)."""

    def forward(self, x):
        scale, shift = _folded_bn(self[1])
        return ops.linear_glu(x, self[0].weight.detach(), None, scale, shift)


class INIT_STAGE_G(nn.Module):
    """AttnGAN noise stage, util.py:404-443 (not reached by the shipped SR flows)."""

    def __init__(self, ngf, ncf):
        super().__init__()
        self.gf_dim = ngf
        self.in_dim = cfg.GAN.Z_DIM + ncf
        self.define_module()

    def define_module(self):
        nz, ngf = self.in_dim, self.gf_dim
        self.fc = _FcBnGlu(nn.Linear(nz, ngf * 4 * 4 * 2, bias=False), nn.BatchNorm1d(ngf * 4 * 4 * 2), GLU())
        self.upsample1 = upBlock(ngf, ngf // 2)
        self.upsample2 = upBlock(ngf // 2, ngf // 4)
        self.upsample3 = upBlock(ngf // 4, ngf // 8)
        self.upsample4 = upBlock(ngf // 8, ngf // 16)

    def forward(self, z_code, c_code):
        c_z_code = torch.cat((c_code, z_code), 1)
        out_code = self.fc(c_z_code).view(-1, self.gf_dim, 4, 4)
        out_code = self.upsample1(out_code)
        out_code = self.upsample2(out_code)
        out_code32 = self.upsample3(out_code)
        return self.upsample4(out_code32)


def _run_stage_bf16(stage, h_code, LR, word_embs, mask):
    """Module-level call in bf16 mode: NCHW fp32 in/out around the blocked tensor-core stage."""
    from . import fast, tc
    stage.att.applyMask(mask)
    hb = None
    if h_code is not None:
        B, C, H, W = h_code.shape
        hb = tc.to_blocked(h_code, out=tc.Blocked.empty(B, C, H, W, h_code.device, planes=2 * C // 8, fmt=_fmt()))
    out, att, _ = fast.stage(stage, hb, LR, word_embs, mask, None, _fmt())
    return tc.to_nchw(out), att


def _run_stage(stage, h_code, word_embs, mask):
    """Shared body of INIT_STAGE_GImgup / NEXT_STAGE_G (util.py:769-776, :814-821)."""
    stage.att.applyMask(mask)
    c_code, att = stage.att(h_code, word_embs)
    h_c_code = torch.cat((h_code, c_code), 1)
    out_code = stage.residual(h_c_code)
    return stage.upsample(out_code), att


class INIT_STAGE_GImgup(nn.Module):
    """util.py:726-777: image -> features, word attention, residual blocks, x2 upsample."""

    def __init__(self, ngf, ncf, nef, batchnorm=True):
        super().__init__()
        self.gf_dim = ngf
        self.in_dim = cfg.GAN.Z_DIM + ncf
        self.ef_dim = nef
        self.batchnorm = batchnorm
        self.att = ATT_NET(self.gf_dim, self.ef_dim)
        if batchnorm:
            self.im2f = _ConvBnGlu(conv3x3(3, ngf * 2), nn.BatchNorm2d(ngf * 2), GLU())
            self.upsample = upBlock(ngf * 2, ngf)
        else:
            self.im2f = _ConvBnGlu(conv3x3(3, ngf * 2), GLU())
            self.upsample = upBlocknoBN(ngf * 2, ngf)
        self.residual = nn.Sequential(*[ResBlock(ngf * 2, batchnorm) for _ in range(cfg.GAN.R_NUM)])

    def forward(self, c_code0, LR, word_embs, mask):
        from . import fast
        if _bf16() and fast.stage_supported(self, LR.shape[2], LR.shape[3]):
            return _run_stage_bf16(self, None, LR, word_embs, mask)
        return _run_stage(self, self.im2f(LR), word_embs, mask)


class NEXT_STAGE_G(nn.Module):
    """util.py:781-823."""

    def __init__(self, ngf, nef, ncf, weightatten=False):
        super().__init__()
        if weightatten:
            raise NotImplementedError('GlobalAttentionGeneral_weight is out of scope (unused by the SR flows)')
        self.weightattn = weightatten
        self.gf_dim = ngf
        self.ef_dim = nef
        self.cf_dim = ncf
        self.num_residual = cfg.GAN.R_NUM
        self.att = ATT_NET(ngf, self.ef_dim)
        self.residual = nn.Sequential(*[ResBlock(ngf * 2) for _ in range(cfg.GAN.R_NUM)])
        self.upsample = upBlock(ngf * 2, ngf)

    def forward(self, h_code, c_code0, word_embs, mask):
        from . import fast
        if _bf16() and fast.stage_supported(self, h_code.shape[2], h_code.shape[3]):
            return _run_stage_bf16(self, h_code, None, word_embs, mask)
        return _run_stage(self, h_code, word_embs, mask)


class GET_IMAGE_G(nn.Module):
    """util.py:894-905: conv3x3(ngf,3) + tanh."""

    def __init__(self, ngf):
        super().__init__()
        self.gf_dim = ngf
        self.img = _ImageHead(conv3x3(ngf, 3), nn.Tanh())

    def forward(self, h_code):
        return self.img(h_code)


class GET_IMAGE_G_noAct(nn.Module):
    """util.py:909-919: conv3x3(ngf,3)."""

    def __init__(self, ngf):
        super().__init__()
        self.gf_dim = ngf
        self.img = _ImageHead(conv3x3(ngf, 3))

    def forward(self, h_code):
        return self.img(h_code)
