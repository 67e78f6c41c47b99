"""DAMSM matching losses and MSE with the reference's function API (miscc/losses.py), on libtgsr_sm100 kernels.

``words_loss`` / ``sent_loss`` / ``MSE`` / ``generator_lossor_wordsentOnly`` keep the reference signatures and
return values; gradients flow to the image side (region features, cnn_code, generated images) through explicit
backward kernels (csrc/damsm.cu).  The text side (words_emb, sent_emb) is the frozen text encoder's output in the
training step and receives no gradient.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import call, no_grad_path, ptr
from .config import cfg

Tensor = torch.Tensor


def _f32(t: Tensor) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError('tgsr_b200 losses need CUDA tensors (no CPU path)')
    return t.detach().to(torch.float32).contiguous()


def _class_mask(class_ids, B: int, device) -> Optional[Tensor]:
    """losses.py:24-34 / :73-77: mask[i, j] = 1 where class_ids[j] == class_ids[i], j != i."""
    if class_ids is None:
        return None
    ids = np.asarray(class_ids)
    m = (ids[:, None] == ids[None, :]).astype(np.uint8)
    np.fill_diagonal(m, 0)
    return torch.from_numpy(m).to(device)


class _CE(object):
    """Both cross entropies over gamma3 * sim (losses.py:124-133, :47-59) and d loss / d sim."""

    @staticmethod
    def run(sim: Tensor, labels: Tensor, cmask: Optional[Tensor], gamma3: float):
        B = sim.shape[0]
        loss2 = torch.empty(2, device=sim.device, dtype=torch.float32)
        d0 = torch.empty_like(sim); d1 = torch.empty_like(sim)
        lab = labels.to(device=sim.device, dtype=torch.int64).contiguous()
        call('tgsr_damsm_ce_f32', sim.device, ptr(sim), ptr(cmask), ptr(lab), float(gamma3), ptr(loss2), ptr(d0), ptr(d1), B)
        return loss2, d0, d1


class WordsLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img_features: Tensor, words_emb: Tensor, labels: Tensor, cap_lens: Tensor, cmask: Optional[Tensor],
                g1: float, g2: float, g3: float):
        F_ = _f32(img_features); W_ = _f32(words_emb)
        B, ndf, ih, iw = F_.shape
        S, T = ih * iw, W_.shape[2]
        lens = cap_lens.to(device=F_.device, dtype=torch.int64).contiguous()
        dev = F_.device
        sim = torch.empty((B, B), device=dev, dtype=torch.float32)
        attn = torch.empty((B * B, T, S), device=dev, dtype=torch.float32)
        p = torch.empty((B * B, S, 24), device=dev, dtype=torch.float32)
        c = torch.empty((B * B, ndf, T), device=dev, dtype=torch.float32)
        call('tgsr_damsm_pairs_fwd_f32', dev, ptr(F_), ptr(W_), ptr(lens), float(g1), float(g2), 1e-8, ptr(sim), ptr(attn), ptr(p),
             ptr(c), B, B, ndf, S, T, 0)
        loss2, d0, d1 = _CE.run(sim, labels, cmask, g3)
        ctx.save_for_backward(F_, W_, lens, attn, p, c, d0, d1)
        ctx.g = (float(g1), float(g2))
        ctx.shape = tuple(img_features.shape)
        ctx.mark_non_differentiable(attn)
        return loss2[0], loss2[1], attn

    @staticmethod
    def backward(ctx, gl0, gl1, _gattn):
        F_, W_, lens, attn, p, c, d0, d1 = ctx.saved_tensors
        B, ndf, ih, iw = ctx.shape
        S, T = ih * iw, W_.shape[2]
        dev = F_.device
        z = torch.zeros((), device=dev, dtype=torch.float32)
        g0 = z if gl0 is None else gl0.to(torch.float32).contiguous()
        g1 = z if gl1 is None else gl1.to(torch.float32).contiguous()
        wT = W_.transpose(1, 2).contiguous()                      # [B,T,256]: data movement only
        dcT = torch.empty((B * B, 24, ndf), device=dev, dtype=torch.float32)
        draw = torch.empty((B * B, 24, S), device=dev, dtype=torch.float32)
        dF = torch.empty_like(F_)
        call('tgsr_damsm_pairs_bwd_f32', dev, ptr(F_), ptr(W_), ptr(wT), ptr(lens), ptr(attn), ptr(p), ptr(c), ptr(d0), ptr(d1),
             ptr(g0), ptr(g1), ctx.g[0], ctx.g[1], 1e-8, ptr(dcT), ptr(draw), ptr(dF), B, B, ndf, S, T)
        return dF.view(ctx.shape), None, None, None, None, None, None, None


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size):
    """losses.py:65-136.  img_features [B,256,17,17], words_emb [B,256,T] -> (loss0, loss1, att_maps) where
    att_maps[i] = attention of caption i on its own image, [1, cap_lens[i], 17, 17]."""
    B = batch_size
    sm = cfg.TRAIN.SMOOTH
    cmask = _class_mask(class_ids, B, img_features.device)
    if labels is None:
        raise NotImplementedError('words_loss(labels=None) returns no losses in the reference; not needed here')
    no_grad_path('words_loss(words_emb)', words_emb)          # only d/d(img_features) is built (frozen text encoder)
    loss0, loss1, attn = WordsLossFn.apply(img_features, words_emb, labels, cap_lens, cmask, sm.GAMMA1, sm.GAMMA2, sm.GAMMA3)
    return loss0, loss1, _AttMaps(attn, cap_lens, B, words_emb.shape[2], img_features.shape[2], img_features.shape[3])


class _AttMaps(object):
    """The att_maps list of words_loss (losses.py:93): element i = attention of caption i on its own image,
    [1, cap_lens[i], 17, 17].  Built lazily: reading cap_lens is a device->host sync that the training step (which
    discards the maps, losses.py:511) should not pay."""

    def __init__(self, attn, cap_lens, B, T, ih, iw):
        self._a, self._lens, self._B, self._T, self._ih, self._iw, self._list = attn, cap_lens, B, T, ih, iw, None

    def _build(self):
        if self._list is None:
            lens = [int(v) for v in self._lens.tolist()]
            B = self._B
            self._list = [self._a[i * B + i, :min(lens[i], self._T)].reshape(1, -1, self._ih, self._iw) for i in range(B)]
        return self._list

    def __len__(self):
        return self._B

    def __getitem__(self, i):
        return self._build()[i]

    def __iter__(self):
        return iter(self._build())


class SentLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cnn_code: Tensor, rnn_code: Tensor, labels: Tensor, cmask: Optional[Tensor], g3: float, eps: float):
        Cc = _f32(cnn_code); E = _f32(rnn_code)
        B, D = Cc.shape
        sim = torch.empty((B, B), device=Cc.device, dtype=torch.float32)
        call('tgsr_sent_sim_f32', Cc.device, ptr(Cc), ptr(E), float(eps), ptr(sim), B, D)
        loss2, d0, d1 = _CE.run(sim, labels, cmask, g3)
        ctx.save_for_backward(Cc, E, sim, d0, d1)
        ctx.eps = float(eps)
        return loss2[0], loss2[1]

    @staticmethod
    def backward(ctx, gl0, gl1):
        Cc, E, sim, d0, d1 = ctx.saved_tensors
        B, D = Cc.shape
        z = torch.zeros((), device=Cc.device, dtype=torch.float32)
        g0 = z if gl0 is None else gl0.to(torch.float32).contiguous()
        g1 = z if gl1 is None else gl1.to(torch.float32).contiguous()
        dC = torch.empty_like(Cc)
        call('tgsr_sent_bwd_f32', Cc.device, ptr(Cc), ptr(E), ptr(sim), ptr(d0), ptr(d1), ptr(g0), ptr(g1), ctx.eps, ptr(dC), B, D)
        return dC, None, None, None, None, None


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8):
    """losses.py:21-62 for 2-D codes [B,256]."""
    if cnn_code.dim() != 2:
        raise NotImplementedError('sent_loss: only [B, nef] codes (the SR training step) are implemented')
    cmask = _class_mask(class_ids, batch_size, cnn_code.device)
    no_grad_path('sent_loss(rnn_code)', rnn_code)              # only d/d(cnn_code) is built (frozen text encoder)
    return SentLossFn.apply(cnn_code, rnn_code, labels, cmask, cfg.TRAIN.SMOOTH.GAMMA3, eps)


class MseFn(torch.autograd.Function):
    """sum_k nn.MSELoss()(fake[k], label[k]) (losses.py:779-784) in one node."""

    @staticmethod
    def forward(ctx, n_pairs: int, *tensors):
        fakes, labels = tensors[:n_pairs], tensors[n_pairs:]
        dev = fakes[0].device
        loss = torch.zeros(1, device=dev, dtype=torch.float32)
        saved = []
        for f, l in zip(fakes, labels):
            f = _f32(f); l = _f32(l)
            if f.shape != l.shape:
                raise ValueError('MSE: shape mismatch %s vs %s' % (tuple(f.shape), tuple(l.shape)))
            call('tgsr_mse_f32', dev, ptr(f), ptr(l), ptr(loss), None, C.c_size_t(f.numel()))
            saved += [f, l]
        ctx.save_for_backward(*saved)
        ctx.n = n_pairs
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        g = g.to(torch.float32).contiguous()
        grads = []
        for k in range(ctx.n):
            f, l = ctx.saved_tensors[2 * k], ctx.saved_tensors[2 * k + 1]
            gr = torch.empty_like(f)
            call('tgsr_mse_bwd_f32', f.device, ptr(f), ptr(l), ptr(g), ptr(gr), C.c_size_t(f.numel()))
            grads.append(gr)
        return (None,) + tuple(grads) + (None,) * ctx.n


def MSE(fake, label):
    """losses.py:779-784."""
    return MseFn.apply(len(fake), *fake, *label)


class KlLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar):
        mu = _f32(mu); logvar = _f32(logvar)
        if mu.shape != logvar.shape:
            raise ValueError('KL_loss: shape mismatch %s vs %s' % (tuple(mu.shape), tuple(logvar.shape)))
        out = torch.empty(1, device=mu.device, dtype=torch.float32)
        call('tgsr_kl_loss_f32', mu.device, ptr(mu), ptr(logvar), ptr(out), C.c_size_t(mu.numel()))
        ctx.save_for_backward(mu, logvar)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        mu, logvar = ctx.saved_tensors
        g = g.to(torch.float32).reshape(1).contiguous()
        d_mu, d_lv = torch.empty_like(mu), torch.empty_like(logvar)
        call('tgsr_kl_loss_bwd_f32', mu.device, ptr(mu), ptr(logvar), ptr(g), ptr(d_mu), ptr(d_lv), C.c_size_t(mu.numel()))
        return d_mu, d_lv


def KL_loss(mu, logvar):
    """losses.py:806-810: -0.5 * mean(1 + logvar - mu^2 - exp(logvar))."""
    return KlLossFn.apply(mu, logvar)


class CosineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, dim, eps):
        x1 = _f32(x1); x2 = _f32(x2)
        if x1.shape != x2.shape:
            raise ValueError('cosine_similarity: shape mismatch %s vs %s' % (tuple(x1.shape), tuple(x2.shape)))
        dim = dim % x1.dim()
        outer = 1
        for s in x1.shape[:dim]:
            outer *= s
        inner = 1
        for s in x1.shape[dim + 1:]:
            inner *= s
        D = x1.shape[dim]
        oshape = tuple(x1.shape[:dim]) + tuple(x1.shape[dim + 1:])
        out = torch.empty(oshape, device=x1.device, dtype=torch.float32)
        stats = torch.empty((max(out.numel(), 1), 3), device=x1.device, dtype=torch.float32)
        call('tgsr_cosine_similarity_f32', x1.device, ptr(x1), ptr(x2), ptr(out), ptr(stats), C.c_size_t(outer), D, inner,
             C.c_float(eps))
        ctx.save_for_backward(x1, x2, stats)
        ctx.geom = (outer, D, inner, eps)
        return out

    @staticmethod
    def backward(ctx, g):
        x1, x2, stats = ctx.saved_tensors
        outer, D, inner, eps = ctx.geom
        g = g.to(torch.float32).contiguous()
        d1, d2 = torch.empty_like(x1), torch.empty_like(x2)
        call('tgsr_cosine_similarity_bwd_f32', x1.device, ptr(x1), ptr(x2), ptr(stats), ptr(g), ptr(d1), ptr(d2),
             C.c_size_t(outer), D, inner, C.c_float(eps))
        return d1, d2, None, None


def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """losses.py:12-18: cosine similarity along `dim`, squeezed like the reference's."""
    return CosineFn.apply(x1, x2, dim, eps).squeeze()


def generator_lossor_wordsentOnly(image_encoder, fake_imgs, words_embs, sent_emb, match_labels, cap_lens, class_ids, w=1, s=1):
    """losses.py:499-522: ranking losses on the last generated image through the (frozen) image encoder."""
    batch_size = fake_imgs[0].size(0)
    region_features, cnn_code = image_encoder(fake_imgs[-1])
    w_loss0, w_loss1, _ = words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size)
    w_loss = (w_loss0 + w_loss1) * cfg.TRAIN.SMOOTH.LAMBDA
    s_loss0, s_loss1 = sent_loss(cnn_code, sent_emb, match_labels, class_ids, batch_size)
    s_loss = (s_loss0 + s_loss1) * cfg.TRAIN.SMOOTH.LAMBDA
    errG_total = w * w_loss + s * s_loss
    logs = ''            # the reference formats w_loss/s_loss with .item() here: a device->host sync per step, left to the caller
    return errG_total, logs
