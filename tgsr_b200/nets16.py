"""x16 SR nets (models16.py): 16x16 -> 256x256 with WEIGHT-SHARED stages.  One NEXT_STAGE_G object
serves h_net2..4 and one GET_IMAGE_G (tanh) serves all four heads (models16.py:13-14), so the
state dict carries the shared tensors under every alias, exactly like the reference."""
from __future__ import annotations

import torch
import torch.nn as nn

from .blocks import (CA_NET, GET_IMAGE_G, GLU, INIT_STAGE_GImgup, NEXT_STAGE_G, ResBlock, _ConvBnGlu,
                     _ConvBnGluConvBn, _ImageHead, conv3x3, conv5x5, upBlock)
from .config import cfg


class G_SR_NET_low(nn.Module):
    """models16.py:5-39."""

    def __init__(self):
        super().__init__()
        ngf = cfg.GAN.GF_DIM
        nef = cfg.TEXT.EMBEDDING_DIM
        ncf = cfg.GAN.CONDITION_DIM
        self.ca_net = CA_NET()
        self.h_net1 = INIT_STAGE_GImgup(ngf, ncf, nef)
        self.h_net4 = self.h_net3 = self.h_net2 = NEXT_STAGE_G(ngf, nef, ncf)
        self.img_net4 = self.img_net3 = self.img_net2 = self.img_net1 = GET_IMAGE_G(ngf)

    def forward(self, LR, sent_emb, word_embs, mask):
        from . import fast, fmt_for, is_16bit
        stages = [self.h_net1, self.h_net2, self.h_net3, self.h_net4]
        heads = [self.img_net1, self.img_net2, self.img_net3, self.img_net4]
        if self.training and is_16bit():                        # bf16 training step (tgsr_b200/train16.py); the shared
            from . import train16                               # stage / head parameters accumulate their gradients in autograd
            return train16.g_sr_net_low(self, LR, sent_emb, word_embs, mask, stages, heads)
        if is_16bit() and fast.stage_supported(self.h_net1, LR.shape[2], LR.shape[3]):
            self.h_net1.att.applyMask(mask); self.h_net2.att.applyMask(mask)
            return fast.g_sr_net_low(self, LR, sent_emb, word_embs, mask,
                                     [self.h_net1, self.h_net2, self.h_net3, self.h_net4],
                                     [self.img_net1, self.img_net2, self.img_net3, self.img_net4], False, fmt_for('low'))
        fake_imgs, att_maps = [], []
        c_code, mu, logvar = self.ca_net(sent_emb)
        h_code, att = self.h_net1(None, LR, word_embs, mask)
        fake_imgs.append(self.img_net1(h_code)); att_maps.append(att)
        for h_net, img_net in ((self.h_net2, self.img_net2), (self.h_net3, self.img_net3),
                               (self.h_net4, self.img_net4)):
            h_code, att = h_net(h_code, None, word_embs, mask)
            fake_imgs.append(img_net(h_code)); att_maps.append(att)
        return fake_imgs, att_maps, mu, logvar


def _fast_ok(net, LR) -> bool:
    """16-bit mode and the trunk is inside the tcgen05 kernels' envelope: every layer then runs on blocked activations
    without layout conversions between them (tgsr_b200/fast.py)."""
    from . import is_16bit
    return (is_16bit() and net.convin[0].weight.shape[0] == 64 and LR.shape[2] % 16 == 0 and LR.shape[3] % 8 == 0
            and not net.training)


def _cached_float(mod, t, name) -> float:
    """Host copy of a one-element parameter, cached per weight version (a device->host read only when it changed)."""
    from ._lib import tensor_key
    key = tensor_key(t)
    c = mod.__dict__.get(name)
    if c is None or c[0] != key:
        c = (key, float(t.detach()))
        mod.__dict__[name] = c
    return c[1]


def _res_body(ngf):
    return _ConvBnGluConvBn(conv3x3(ngf, ngf * 2), nn.BatchNorm2d(ngf * 2), GLU(), conv3x3(ngf, ngf),
                            nn.BatchNorm2d(ngf))


class NetG_high(nn.Module):
    """models16.py:42-94: every upscale and every residual body is ONE shared module."""

    def __init__(self, cat=False):
        super().__init__()
        ngf = cfg.GAN.GF_DIM
        self.residual = nn.Sequential(*[ResBlock(channel_num=32) for _ in range(6)])
        self.upscale16x = self.upscale8x = self.upscale4x = self.upscale2x = upBlock(ngf, ngf)
        self.cat = cat
        self.conv_output = _ImageHead(conv5x5(ngf, 3), nn.Tanh())
        self.convin = _ConvBnGlu(conv3x3(3, ngf * 2), nn.BatchNorm2d(ngf * 2), GLU())
        self.residual816 = self.residual48 = self.residual24 = _res_body(ngf)

    def forward(self, LR, SRb, LRb, low=False):
        from . import fast, fmt_for, is_16bit
        if self.training and is_16bit():                        # bf16 training step (tgsr_b200/train16.py)
            from . import train16
            x, x_sub = (LRb, None) if low else (LR, LRb)
            return train16.netg_high(self, x, x_sub, SRb, 1.0, 1.0,
                                     [(self.residual24, self.upscale4x), (self.residual48, self.upscale8x),
                                      (self.residual816, self.upscale16x)])
        if _fast_ok(self, LR):
            x, x_sub = (LRb, None) if low else (LR, LRb)
            return fast.netg_high(self, x, x_sub, SRb, 1.0, 1.0,
                                  [(self.residual24, self.upscale4x), (self.residual48, self.upscale8x),
                                   (self.residual816, self.upscale16x)], None, fmt_for('high'))
        out = self.convin(LRb if low else torch.sub(LR, LRb))
        out = self.residual(out)
        out = self.upscale2x(out)
        ims = [self.conv_output(out, SRb[0], 1.0, 1.0)]
        for body, up, srb in ((self.residual24, self.upscale4x, SRb[1]), (self.residual48, self.upscale8x, SRb[2]),
                              (self.residual816, self.upscale16x, SRb[3])):
            out = up(body(out))
            ims.append(self.conv_output(out, srb, 1.0, 1.0))
        return ims


class NetG_highweight(nn.Module):
    """models16.py:97-179 with weightmap=False.  The reference reuses residual48/upscale8x for the
    last stage (models16.py:172-173) and then adds SRb8 (128^2) to a 256^2 tensor (:178), which is a
    shape error; here the last stage keeps the reuse and adds SRb[3] (documented deviation)."""

    def __init__(self, weightmap=False, low='lr-lrblur'):
        super().__init__()
        if weightmap:
            raise NotImplementedError('weightmap=True is out of scope')
        ngf = cfg.GAN.GF_DIM
        self.low = low
        self.weightmap = weightmap
        self.residual = nn.Sequential(*[ResBlock(channel_num=32) for _ in range(6)])
        self.upscale4x = upBlock(ngf, ngf)
        self.upscale2x = upBlock(ngf, ngf)
        self.upscale8x = upBlock(ngf, ngf)
        self.upscale16x = upBlock(ngf, ngf)
        self.conv_output = _ImageHead(conv5x5(ngf, 3), nn.Tanh())
        self.convin = _ConvBnGlu(conv3x3(3, ngf * 2), nn.BatchNorm2d(ngf * 2), GLU())
        self.residual24 = _res_body(ngf)
        self.residual48 = _res_body(ngf)
        self.residual816 = _res_body(ngf)
        self.a = nn.Parameter(torch.FloatTensor([0.5]))          # registered here (models16.py:131)
        self.one = torch.ones([1])

    def forward(self, LR, SRb, LRb):
        if self.low == 'lrblur':
            x = LRb
        elif self.low == 'lr-lrblur':
            x = torch.sub(LR, LRb)
        elif self.low == 'lr':
            x = LR
        else:
            raise ValueError(self.low)
        a, one = _cached_float(self, self.a, '_tgsr_a'), 1.0      # `a` is trainable here: read once per weight version
        from . import fast, fmt_for, is_16bit
        if self.training and is_16bit():                        # bf16 training step; `a` enters as a constant, as in the x8 net
            from . import train16
            x, x_sub = (LR, LRb) if self.low == 'lr-lrblur' else (x, None)
            ims = train16.netg_high(self, x, x_sub, SRb, a, one,
                                    [(self.residual24, self.upscale4x), (self.residual48, self.upscale8x),
                                     (self.residual48, self.upscale8x)])
            return ims, self.a, self.one.to(ims[0].device)
        if _fast_ok(self, LR):
            x, x_sub = (LR, LRb) if self.low == 'lr-lrblur' else (x, None)
            ims = fast.netg_high(self, x, x_sub, SRb, a, one,
                                 [(self.residual24, self.upscale4x), (self.residual48, self.upscale8x),
                                  (self.residual48, self.upscale8x)], None, fmt_for('high'))
            return ims, self.a, self.one.to(ims[0].device)
        out = self.convin(x)
        out = self.residual(out)
        out = self.upscale2x(out)
        ims = [self.conv_output(out, SRb[0], a, one)]
        for body, up, srb in ((self.residual24, self.upscale4x, SRb[1]), (self.residual48, self.upscale8x, SRb[2]),
                              (self.residual48, self.upscale8x, SRb[3])):
            out = up(body(out))
            ims.append(self.conv_output(out, srb, a, one))
        return ims, self.a, self.one.to(ims[0].device)
