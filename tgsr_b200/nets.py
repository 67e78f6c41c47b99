"""x8 SR nets (model.py): G_SR_NET_low, NetG_high, NetG_highweight -- same constructors, forward
signatures, return structures and state-dict keys as the reference; the shipped
Checkpoint/face_S8 state dicts load with strict=True."""
from __future__ import annotations

import torch
import torch.nn as nn

from .blocks import (CA_NET, GET_IMAGE_G, GET_IMAGE_G_noAct, GLU, INIT_STAGE_GImgup, NEXT_STAGE_G, ResBlock,
                     _ConvBnGlu, _ConvBnGluConvBn, _ImageHead, conv3x3, conv5x5, upBlock)
from .config import cfg


class G_SR_NET_low(nn.Module):
    """model.py:34-78: text-attention low-frequency net, LR 32x32 -> 64/128/256."""

    def __init__(self):
        super().__init__()
        ngf = cfg.GAN.GF_DIM
        nef = cfg.TEXT.EMBEDDING_DIM
        ncf = cfg.GAN.CONDITION_DIM
        self.ca_net = CA_NET()
        self.h_net1 = INIT_STAGE_GImgup(ngf, ncf, nef)
        self.h_net2 = NEXT_STAGE_G(ngf, nef, ncf)
        self.h_net3 = NEXT_STAGE_G(ngf, nef, ncf)
        self.img_net1 = GET_IMAGE_G_noAct(ngf)
        self.img_net2 = GET_IMAGE_G_noAct(ngf)
        self.img_net3 = GET_IMAGE_G_noAct(ngf)

    def forward(self, LR, sent_emb, word_embs, mask, outmiddle=False):
        from . import fast, fmt_for, is_16bit
        if self.training and is_16bit():                        # bf16 training step: tcgen05 convs on blocked activations
            if outmiddle:
                raise NotImplementedError('outmiddle is an eval-time option')
            from . import train16
            return train16.g_sr_net_low(self, LR, sent_emb, word_embs, mask, [self.h_net1, self.h_net2, self.h_net3],
                                        [self.img_net1, self.img_net2, self.img_net3])
        if is_16bit() and fast.stage_supported(self.h_net1, LR.shape[2], LR.shape[3]):
            for st in (self.h_net1, self.h_net2, self.h_net3):
                st.att.applyMask(mask)                              # keep the stateful contract
            return fast.g_sr_net_low(self, LR, sent_emb, word_embs, mask, [self.h_net1, self.h_net2, self.h_net3],
                                     [self.img_net1, self.img_net2, self.img_net3], outmiddle, fmt_for('low'))
        fake_imgs, att_maps = [], []
        c_code, mu, logvar = self.ca_net(sent_emb)              # c_code unused, model.py:51
        h_code1, att0 = self.h_net1(None, LR, word_embs, mask)
        fake_imgs.append(self.img_net1(h_code1))
        if att0 is not None:
            att_maps.append(att0)
        h_code2, att1 = self.h_net2(h_code1, None, word_embs, mask)
        fake_imgs.append(self.img_net2(h_code2))
        if att1 is not None:
            att_maps.append(att1)
        h_code3, att2 = self.h_net3(h_code2, None, word_embs, mask)
        fake_imgs.append(self.img_net3(h_code3))
        if att2 is not None:
            att_maps.append(att2)
        if outmiddle:
            return fake_imgs, att_maps, mu, logvar, [h_code1, h_code2, h_code3]
        return fake_imgs, att_maps, mu, logvar


class _HighBase(nn.Module):
    """Shared trunk of NetG_high / NetG_highweight (model.py:133-160, :212-263)."""

    def _build(self, ngf, useAct=True):
        self.residual = nn.Sequential(*[ResBlock(channel_num=32) for _ in range(6)])
        self.upscale4x = upBlock(ngf, ngf)
        self.upscale2x = upBlock(ngf, ngf)
        self.upscale8x = upBlock(ngf, ngf)
        if useAct:
            self.conv_output = _ImageHead(conv5x5(ngf, 3), nn.Tanh())
        else:
            self.conv_output = _ImageHead(conv5x5(ngf, 3))
        self.convin = _ConvBnGlu(conv3x3(3, ngf * 2), nn.BatchNorm2d(ngf * 2), GLU())
        self.residual24 = _ConvBnGluConvBn(conv3x3(ngf, ngf * 2), nn.BatchNorm2d(ngf * 2), GLU(), conv3x3(ngf, ngf),
                                           nn.BatchNorm2d(ngf))
        self.residual48 = _ConvBnGluConvBn(conv3x3(ngf, ngf * 2), nn.BatchNorm2d(ngf * 2), GLU(), conv3x3(ngf, ngf),
                                           nn.BatchNorm2d(ngf))

    def _trunk(self, x, SRb, a, one, feats=None, x_sub=None):
        from . import fast, fmt_for, is_16bit
        if self.training and is_16bit():                        # bf16 training step (tgsr_b200/train16.py)
            if feats is not None:
                raise NotImplementedError('outmiddle is an eval-time option')
            from . import train16
            return train16.netg_high(self, x, x_sub, SRb, a, one,
                                     [(self.residual24, self.upscale4x), (self.residual48, self.upscale8x)])
        if is_16bit() and self.convin[0].weight.shape[0] == 64 and x.shape[2] % 16 == 0 \
                and x.shape[3] % 8 == 0:
            return fast.netg_high(self, x, x_sub, SRb, a, one,
                                  [(self.residual24, self.upscale4x), (self.residual48, self.upscale8x)], feats, fmt_for('high'))
        if x_sub is not None:
            x = torch.sub(x, x_sub)
        out = self.convin(x)
        out = self.residual(out)
        out = self.upscale2x(out)
        if feats is not None: feats.append(out)
        ims2 = self.conv_output(out, SRb[0], a, one)
        out = self.upscale4x(self.residual24(out))
        if feats is not None: feats.append(out)
        ims4 = self.conv_output(out, SRb[1], a, one)
        out = self.upscale8x(self.residual48(out))
        if feats is not None: feats.append(out)
        ims8 = self.conv_output(out, SRb[2], a, one)
        return [ims2, ims4, ims8]


class NetG_high(_HighBase):
    """model.py:133-209 (cat=False path; cat=True is an abandoned variant)."""

    def __init__(self, cat=False):
        super().__init__()
        if cat:
            raise NotImplementedError('NetG_high(cat=True) is out of scope (no shipped flow/checkpoint uses it)')
        self.cat = cat
        self._build(cfg.GAN.GF_DIM)

    def forward(self, LR, SRb, LRb, low=False, outmiddle=False):
        feats = [] if outmiddle else None
        ims = self._trunk(LRb, SRb, 1.0, 1.0, feats) if low else self._trunk(LR, SRb, 1.0, 1.0, feats, x_sub=LRb)
        return (ims, feats) if outmiddle else ims


class NetG_highweight(_HighBase):
    """model.py:212-298 with weightmap=False: out = one*tanh(conv5x5(h)) + a*SRb, a = 0.5, one = 1.
    In the reference ``a`` is ``nn.Parameter(...).cuda()`` -- a plain tensor, not registered, not in
    the state dict (model.py:246-248); it is the constant 0.5 here too."""

    def __init__(self, weightmap=False, low='lr-lrblur', useAct=True):
        super().__init__()
        if weightmap:
            raise NotImplementedError('weightmap=True is out of scope (trainer_objective.py:58 sets False)')
        self.low = low
        self.weightmap = weightmap
        self._build(cfg.GAN.GF_DIM, useAct)
        self.a = torch.tensor([0.5])
        self.one = torch.ones([1])

    def forward(self, LR, SRb, LRb):
        x_sub = None
        if self.low == 'lrblur':
            x = LRb
        elif self.low == 'lr-lrblur':
            x, x_sub = LR, LRb
        elif self.low == 'lr':
            x = LR
        else:
            raise ValueError(self.low)
        ims = self._trunk(x, SRb, 0.5, 1.0, x_sub=x_sub)
        dev = ims[0].device
        if self.a.device != dev:
            self.a, self.one = self.a.to(dev), self.one.to(dev)
        return ims, self.a, self.one
