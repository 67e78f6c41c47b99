"""CPU oracle for the TGSR generator forward path (TEST INFRASTRUCTURE ONLY).

A functional fp32 restatement of the reference's eval-mode arithmetic.  Every
function takes an explicit ``state_dict`` (same keys as the reference modules)
plus a key prefix, and cites the reference file:line it follows.  The
arithmetic itself lives in PyTorch (the reference pins no version; this image
has torch 2.11.0+cu128), so the restatement uses ``torch.nn.functional``
primitives on CPU tensors in fp32 -- there are no nn.Module objects, no global
``cfg`` and no hidden state here.

Pinned against the reference's own modules by tests/golden/*.pt
(see oracle/make_golden.py and tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5  # nn.BatchNorm2d default, never overridden by the reference


# --------------------------------------------------------------------------
# Word-region attention
# --------------------------------------------------------------------------
def global_attention(inp: Tensor, context: Tensor, w_ctx: Tensor,
                     mask: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """GlobalAttentionGeneral.forward -- GlobalAttention.py:87-130.

    inp [B,idf,ih,iw], context [B,cdf,T], w_ctx [idf,cdf,1,1] (conv_context.weight),
    mask bool [B,T] or None.  Returns (weightedContext [B,idf,ih,iw], attn [B,T,ih,iw]).

    Mask quirk (GlobalAttention.py:109-112): the logits are laid out with row
    ``r = b*Q + q`` but the mask is ``mask.repeat(Q, 1)`` whose row r is
    ``mask[r % B]``.  Pixel q of image b is therefore masked with caption
    ``(b*Q + q) % B``.  Restated here with explicit index arithmetic.
    """
    B, idf, ih, iw = inp.shape
    Q = ih * iw
    T = context.shape[2]
    W = w_ctx.reshape(w_ctx.shape[0], w_ctx.shape[1])            # [idf, cdf]
    source = torch.einsum('ic,bct->bit', W, context)             # :100-102  [B, idf, T]
    target = inp.reshape(B, idf, Q)                              # :97
    logits = torch.einsum('biq,bit->bqt', target, source)        # :107  [B, Q, T]
    if mask is not None:
        rows = torch.arange(B * Q, device=inp.device)            # r = b*Q + q
        cap = rows % B                                           # :112 repeat(Q,1)
        m = mask.to(torch.bool)[cap].reshape(B, Q, T)
        logits = logits.masked_fill(m, float('-inf'))            # :113-116
    attn = torch.softmax(logits, dim=2)                          # :118 (dim=1 of [B*Q,T])
    attn_t = attn.transpose(1, 2).contiguous()                   # :122  [B, T, Q]
    wctx = torch.einsum('bit,btq->biq', source, attn_t)          # :126
    return wctx.reshape(B, idf, ih, iw), attn_t.reshape(B, T, ih, iw)


def func_attention(query: Tensor, context: Tensor, gamma1: float) -> Tuple[Tensor, Tensor]:
    """func_attention -- GlobalAttention.py:33-74 (DAMSM double softmax).

    query [B,ndf,Tq], context [B,ndf,ih,iw] -> (weightedContext [B,ndf,Tq], attn [B,Tq,ih,iw]).
    """
    B, ndf, Tq = query.shape
    ih, iw = context.shape[2], context.shape[3]
    S = ih * iw
    ctx = context.reshape(B, ndf, S)
    a = torch.einsum('bds,bdt->bst', ctx, query)                 # :51  [B,S,Tq]
    a = torch.softmax(a, dim=2)                                  # :54 over words
    a = a.transpose(1, 2)                                        # :59  [B,Tq,S]
    a = torch.softmax(a * gamma1, dim=2)                         # :62-63 over regions
    wctx = torch.einsum('bds,bts->bdt', ctx, a)                  # :70
    return wctx, a.reshape(B, Tq, ih, iw)


# --------------------------------------------------------------------------
# Generator building blocks
# --------------------------------------------------------------------------
def glu(x: Tensor) -> Tensor:
    """GLU -- util.py:45-53."""
    nc = x.shape[1]
    assert nc % 2 == 0, 'channels dont divide 2!'
    nc //= 2
    return x[:, :nc] * torch.sigmoid(x[:, nc:])


BN_TRAIN = False   # set by train_mode(): nn.BatchNorm2d in train() mode (batch statistics), for the training-step checks


class train_mode:
    """Context manager: the block functions below use batch statistics, as the reference modules do in train() mode."""

    def __enter__(self):
        global BN_TRAIN
        self._old, BN_TRAIN = BN_TRAIN, True

    def __exit__(self, *a):
        global BN_TRAIN
        BN_TRAIN = self._old


def _bn(sd: SD, p: str, x: Tensor) -> Tensor:
    """nn.BatchNorm2d / BatchNorm1d in eval mode (running statistics), or batch statistics under train_mode()."""
    if BN_TRAIN:
        return F.batch_norm(x, None, None, sd[p + '.weight'], sd[p + '.bias'], True, 0.0, BN_EPS)
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd[p + '.weight'], sd[p + '.bias'], False, 0.0, BN_EPS)


def _conv(sd: SD, p: str, x: Tensor, pad: int) -> Tensor:
    """conv3x3 / conv5x5, stride 1, no bias -- util.py:62-69."""
    return F.conv2d(x, sd[p + '.weight'], None, 1, pad)


def block3x3(sd: SD, p: str, x: Tensor) -> Tensor:
    """conv3x3 -> BatchNorm2d -> GLU (Block3x3_relu / im2f / convin) -- util.py:102-107, :741-744."""
    return glu(_bn(sd, p + '.1', _conv(sd, p + '.0', x, 1)))


def res_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """ResBlock(batchnorm=True) -- util.py:110-130."""
    y = glu(_bn(sd, p + '.block.1', _conv(sd, p + '.block.0', x, 1)))
    y = _bn(sd, p + '.block.4', _conv(sd, p + '.block.3', y, 1))
    return y + x


def conv_bn_glu_conv_bn(sd: SD, p: str, x: Tensor) -> Tensor:
    """residual24 / residual48: ResBlock body without the skip -- model.py:229-232."""
    y = glu(_bn(sd, p + '.1', _conv(sd, p + '.0', x, 1)))
    return _bn(sd, p + '.4', _conv(sd, p + '.3', y, 1))


def up_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """upBlock: nearest x2 -> conv3x3 -> BN -> GLU -- util.py:74-80."""
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    return glu(_bn(sd, p + '.2', _conv(sd, p + '.1', x, 1)))


def get_image(sd: SD, p: str, h: Tensor, act: bool) -> Tensor:
    """GET_IMAGE_G (tanh) util.py:894-905 / GET_IMAGE_G_noAct util.py:909-919."""
    y = _conv(sd, p + '.img.0', h, 1)
    return torch.tanh(y) if act else y


def ca_net(sd: SD, p: str, sent: Tensor, c_dim: int, eps: Optional[Tensor] = None):
    """CA_NET -- util.py:372-400.  ``eps`` replaces the reference's normal_() draw."""
    x = glu(F.linear(sent, sd[p + '.fc.weight'], sd[p + '.fc.bias']))   # :384
    mu, logvar = x[:, :c_dim], x[:, c_dim:]                               # :385-386
    if eps is None:
        eps = torch.zeros_like(mu)
    c_code = eps * torch.exp(0.5 * logvar) + mu                           # :389-395
    return c_code, mu, logvar


def init_stage_gimgup(sd: SD, p: str, LR: Tensor, words: Tensor, mask, r_num: int):
    """INIT_STAGE_GImgup.forward -- util.py:763-777."""
    h = block3x3(sd, p + '.im2f', LR)                                     # :770
    c, att = global_attention(h, words, sd[p + '.att.conv_context.weight'], mask)  # :772
    x = torch.cat((h, c), 1)                                              # :773
    for i in range(r_num):
        x = res_block(sd, f'{p}.residual.{i}', x)                         # :774
    return up_block(sd, p + '.upsample', x), att                          # :776


def next_stage_g(sd: SD, p: str, h: Tensor, words: Tensor, mask, r_num: int):
    """NEXT_STAGE_G.forward -- util.py:807-823."""
    c, att = global_attention(h, words, sd[p + '.att.conv_context.weight'], mask)  # :816
    x = torch.cat((h, c), 1)                                              # :817
    for i in range(r_num):
        x = res_block(sd, f'{p}.residual.{i}', x)                         # :818
    return up_block(sd, p + '.upsample', x), att                          # :821


def init_stage_g(sd: SD, p: str, z: Tensor, c: Tensor, ngf: int) -> Tensor:
    """INIT_STAGE_G.forward (AttnGAN noise stage) -- util.py:404-443."""
    x = torch.cat((c, z), 1)                                              # :430
    x = F.linear(x, sd[p + '.fc.0.weight'])
    x = glu(_bn(sd, p + '.fc.1', x))                                      # :432
    x = x.reshape(-1, ngf, 4, 4)                                          # :433
    for k in (1, 2, 3, 4):
        x = up_block(sd, f'{p}.upsample{k}', x)                           # :435-441
    return x


# --------------------------------------------------------------------------
# Whole nets
# --------------------------------------------------------------------------
def g_sr_net_low(sd: SD, LR: Tensor, sent: Tensor, words: Tensor, mask,
                 variant: str = 'x8', r_num: int = 2, c_dim: int = 100,
                 eps: Optional[Tensor] = None):
    """G_SR_NET_low.forward -- model.py:48-78 (x8) / models16.py:16-39 (x16).

    x16 shares one NEXT_STAGE_G for h_net2..4 and one GET_IMAGE_G (tanh) for all
    heads (models16.py:13-14); in a state_dict the shared tensors appear under
    every alias, so plain key lookups reproduce the sharing.
    """
    act = (variant == 'x16')
    n_next = 3 if variant == 'x16' else 2
    _, mu, logvar = ca_net(sd, 'ca_net', sent, c_dim, eps)               # c_code discarded, model.py:51
    imgs: List[Tensor] = []
    atts: List[Tensor] = []
    feats: List[Tensor] = []
    h, att = init_stage_gimgup(sd, 'h_net1', LR, words, mask, r_num)
    imgs.append(get_image(sd, 'img_net1', h, act)); atts.append(att); feats.append(h)
    for k in range(2, 2 + n_next):
        h, att = next_stage_g(sd, f'h_net{k}', h, words, mask, r_num)
        imgs.append(get_image(sd, f'img_net{k}', h, act)); atts.append(att); feats.append(h)
    return imgs, atts, mu, logvar, feats


def netg_high(sd: SD, LR: Tensor, SRb: Sequence[Tensor], LRb: Tensor, low: str = 'lr',
              a: float = 0.5, one: float = 1.0, variant: str = 'x8',
              x16_last: Tuple[str, str] = ('residual48', 'upscale8x')) -> List[Tensor]:
    """NetG_highweight.forward(weightmap=False) -- model.py:264-298;
    NetG_high.forward(cat=False) is the same graph with a=1 -- model.py:161-209.

    ``a`` is the constant 0.5: ``nn.Parameter(...).cuda()`` (model.py:246-247)
    is not a registered parameter.  low in {'lr','lrblur','lr-lrblur'} selects convin's input.
    x16 (models16.py:145-179) reproduces the reference's reuse of residual48/upscale8x for
    the last stage (x16_last default); its final add of SRb8 to a 256^2 tensor is a shape
    error in the reference so the oracle adds SRb[3] there instead.
    models16.NetG_high (models16.py:42-94) is x16_last=('residual816','upscale16x'), a=1.
    """
    if low == 'lrblur':
        x = LRb
    elif low == 'lr-lrblur':
        x = LR - LRb
    elif low == 'lr':
        x = LR
    else:
        raise ValueError(low)
    out = block3x3(sd, 'convin', x)                                       # :272
    for i in range(6):
        out = res_block(sd, f'residual.{i}', out)                         # :273
    ims = []
    out = up_block(sd, 'upscale2x', out)                                  # :275
    ims.append(one * torch.tanh(_conv(sd, 'conv_output.0', out, 2)) + a * SRb[0])   # :280
    out = up_block(sd, 'upscale4x', conv_bn_glu_conv_bn(sd, 'residual24', out))     # :283-284
    ims.append(one * torch.tanh(_conv(sd, 'conv_output.0', out, 2)) + a * SRb[1])   # :288
    out = up_block(sd, 'upscale8x', conv_bn_glu_conv_bn(sd, 'residual48', out))     # :291-292
    ims.append(one * torch.tanh(_conv(sd, 'conv_output.0', out, 2)) + a * SRb[2])   # :297
    if variant == 'x16':
        out = up_block(sd, x16_last[1], conv_bn_glu_conv_bn(sd, x16_last[0], out))   # models16.py:172-173 / :89-90
        ims.append(one * torch.tanh(_conv(sd, 'conv_output.0', out, 2)) + a * SRb[3])
    return ims


# --------------------------------------------------------------------------
# DAMSM matching losses (training step)
# --------------------------------------------------------------------------
def cosine_similarity(x1: Tensor, x2: Tensor, dim: int = 1, eps: float = 1e-8) -> Tensor:
    """miscc/losses.py:12-18."""
    w12 = torch.sum(x1 * x2, dim)
    w1 = torch.norm(x1, 2, dim)
    w2 = torch.norm(x2, 2, dim)
    return (w12 / (w1 * w2).clamp(min=eps)).squeeze()


def KL_loss(mu: Tensor, logvar: Tensor) -> Tensor:
    """miscc/losses.py:806-810."""
    KLD_element = mu.pow(2).add(logvar.exp()).mul(-1).add(1).add(logvar)
    return torch.mean(KLD_element).mul(-0.5)


def _class_masks(class_ids, B: int) -> Optional[Tensor]:
    """miscc/losses.py:24-34 / :73-77, :117-120."""
    if class_ids is None:
        return None
    ids = torch.as_tensor(class_ids)
    m = ids[:, None] == ids[None, :]
    m.fill_diagonal_(False)
    return m


def sent_loss(cnn_code: Tensor, rnn_code: Tensor, labels: Tensor, class_ids, gamma3: float, eps: float = 1e-8):
    """miscc/losses.py:21-62 for [B, nef] codes."""
    B = cnn_code.shape[0]
    c, r = cnn_code.unsqueeze(0), rnn_code.unsqueeze(0)                  # :37-39
    cn = torch.norm(c, 2, dim=2, keepdim=True); rn = torch.norm(r, 2, dim=2, keepdim=True)   # :42-43
    scores0 = torch.bmm(c, r.transpose(1, 2))                            # :45
    norm0 = torch.bmm(cn, rn.transpose(1, 2))                            # :46
    scores0 = (scores0 / norm0.clamp(min=eps) * gamma3).squeeze(0)       # :47-50
    masks = _class_masks(class_ids, B)
    if masks is not None:
        scores0.data.masked_fill_(masks, -float('inf'))                  # :51-55
    scores1 = scores0.transpose(0, 1)
    return F.cross_entropy(scores0, labels), F.cross_entropy(scores1, labels)   # :57-59


def words_loss(img_features: Tensor, words_emb: Tensor, labels: Tensor, cap_lens, class_ids, gamma1: float, gamma2: float,
               gamma3: float):
    """miscc/losses.py:65-136 (the per-caption Python loop, verbatim in structure)."""
    B = img_features.shape[0]
    att_maps, sims = [], []
    lens = [int(v) for v in (cap_lens.tolist() if torch.is_tensor(cap_lens) else cap_lens)]
    for i in range(B):
        n = lens[i]
        word = words_emb[i, :, :n].unsqueeze(0).contiguous().repeat(B, 1, 1)            # :79-83
        wei, attn = func_attention(word, img_features, gamma1)                          # :92
        att_maps.append(attn[i].unsqueeze(0).contiguous())                              # :93
        w2 = word.transpose(1, 2).contiguous().view(B * n, -1)                          # :95-98
        c2 = wei.transpose(1, 2).contiguous().view(B * n, -1)
        row_sim = cosine_similarity(w2, c2).view(B, n)                                  # :102-104
        row_sim = torch.log(torch.exp(row_sim * gamma2).sum(dim=1, keepdim=True))       # :107-109
        sims.append(row_sim)
    sim = torch.cat(sims, 1) * gamma3                                                   # :116, :124
    masks = _class_masks(class_ids, B)
    if masks is not None:
        sim.data.masked_fill_(masks, -float('inf'))                                     # :125-129
    sim1 = sim.transpose(0, 1)
    return F.cross_entropy(sim, labels), F.cross_entropy(sim1, labels), att_maps       # :131-136


# --------------------------------------------------------------------------
# Text encoder
# --------------------------------------------------------------------------
def _lstm_cell(x_proj: Tensor, h: Tensor, c: Tensor, w_hh: Tensor, H: int):
    """One nn.LSTM step; gate order i,f,g,o (torch.nn.LSTM docs)."""
    g = x_proj + h @ w_hh.t()
    i = torch.sigmoid(g[0 * H:1 * H]); f = torch.sigmoid(g[1 * H:2 * H])
    gg = torch.tanh(g[2 * H:3 * H]);   o = torch.sigmoid(g[3 * H:4 * H])
    c = f * c + i * gg
    h = o * torch.tanh(c)
    return h, c


def rnn_encoder(sd: SD, captions: Tensor, cap_lens: Tensor) -> Tuple[Tensor, Tensor]:
    """RNN_ENCODER.forward (LSTM, 1 layer, bidirectional, eval) -- util.py:233-260.

    captions int64 [B,Tcap]; cap_lens int64 [B] (sorted descending, datasets.py:75).
    Packed-sequence semantics restated per caption: the forward direction runs
    t=0..len-1, the reverse direction t=len-1..0, both from zero state
    (init_hidden, util.py:222-231); outputs beyond len are zero and the output
    width is max(cap_lens) (pad_packed_sequence, util.py:250).
    Returns words_emb [B, 2H, Tmax], sent_emb [B, 2H].
    """
    emb_w = sd['encoder.weight']
    H = sd['rnn.weight_hh_l0'].shape[1]
    B = captions.shape[0]
    lens = [int(v) for v in cap_lens.tolist()]
    Tmax = max(lens)
    words = torch.zeros(B, 2 * H, Tmax, dtype=torch.float32)
    sent = torch.zeros(B, 2 * H, dtype=torch.float32)
    for d, sfx in enumerate(('', '_reverse')):
        w_ih = sd['rnn.weight_ih_l0' + sfx]; w_hh = sd['rnn.weight_hh_l0' + sfx]
        bias = sd['rnn.bias_ih_l0' + sfx] + sd['rnn.bias_hh_l0' + sfx]
        for b in range(B):
            L = lens[b]
            x = emb_w[captions[b, :L]]                       # :236 (dropout is identity in eval)
            xp = x @ w_ih.t() + bias                         # [L, 4H]
            h = torch.zeros(H); c = torch.zeros(H)
            order = range(L) if d == 0 else range(L - 1, -1, -1)
            for t in order:
                h, c = _lstm_cell(xp[t], h, c, w_hh, H)
                words[b, d * H:(d + 1) * H, t] = h           # :253 transpose(1,2)
            sent[b, d * H:(d + 1) * H] = h                   # :256-259
    return words, sent


# --------------------------------------------------------------------------
# Synthetic inputs and image metrics
# --------------------------------------------------------------------------
from tgsr_b200.synth import synth_captions, synth_lr  # noqa: E402,F401  (shared synthetic-input generators)


def to_uint8(img: Tensor) -> Tensor:
    """trainer_objective.py:153-155 quantisation: round(clip((x+1)*127.5, 0, 255))."""
    return torch.round(torch.clamp((img + 1.0) * 127.5, 0, 255)).to(torch.uint8)


def psnr_u8(a: Tensor, b: Tensor) -> float:
    """psnr -- trainer_objective.py:177-181 (on uint8 images)."""
    diff = a.to(torch.float64) - b.to(torch.float64)
    rmse = math.sqrt(float((diff ** 2).mean()))
    return float('inf') if rmse == 0 else 20 * math.log10(255.0 / rmse)


def attention_word_stats(attn, thresh=0.0):
    """Word-axis reductions of attention maps [B,T,ih,iw] as the reference's visualisation computes them on the host:
    miscc/utils.py:122-125 `attn.max(dim=1, keepdim=True)` (values, indices) and :389-392
    `conf_score[j] = np.sum(one_map * (one_map > 2*thresh))` over the 3-channel replica of [max-map, word maps...].
    Pinned by executing the unmodified miscc/utils.py behind a skimage import shim: tests/golden/attn_stats.pt
    (oracle/make_golden.py gen_attn_stats records the max-map stack and the conf_score values of the reference's own calls)."""
    import numpy as np
    mx, am = attn.max(dim=1, keepdim=True)
    maps = torch.cat([mx, attn], 1)                                   # utils.py:125
    B, T1 = maps.shape[:2]
    conf = np.zeros((B, T1), np.float64)
    for b in range(B):
        rep = maps[b].view(-1, 1, *attn.shape[2:]).repeat(1, 3, 1, 1).numpy().transpose(0, 2, 3, 1)   # utils.py:127-130
        for j in range(T1):
            one_map = rep[j]
            conf[b, j] = np.sum(one_map * (one_map > (2. * thresh)))
    return mx, am, torch.from_numpy(conf)
