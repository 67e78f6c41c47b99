"""bf16 tensor-core execution of the generator modules (precision 'bf16').

The module objects in blocks.py / nets.py own the parameters; the functions here run their
forward on blocked-plane bf16 activations (tc.Blocked) so that consecutive layers chain in HBM
without layout conversions, channel concatenation is a plane offset, and every conv is a tcgen05
kernel.  Edge tensors stay in the reference's format: LR in and images / attention maps / mu /
logvar out are NCHW fp32.
"""
from __future__ import annotations

import os

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops, tc
from ._lib import ACT_GLU, ACT_NONE, ACT_TANH, call, ptr, tensor_key
from .tc import Blocked

Tensor = torch.Tensor


def _fold(bn):
    from .blocks import _folded_bn
    return _folded_bn(bn)


# ---------------------------------------------------------------- single layers
def conv_bn_act(x: Blocked, conv: nn.Conv2d, bn, act: int, residual: Optional[Blocked] = None, up2x: bool = False,
                out: Optional[Blocked] = None, reverse: bool = False, l2_flags: int = 0) -> Blocked:
    scale, shift = _fold(bn)
    return tc.conv3x3_tc(x, conv, scale, shift, act, residual=residual, up2x=up2x, out=out, reverse=reverse, l2_flags=l2_flags)


# TGSR_L2_POLICY (experiment, profiles/r2_experiments.md): 1 = inside a ResBlock the intermediate tensor is stored with L2
# evict_last and the second conv loads its (then dead) inputs with evict_first and stores evict_last; 2 = without the last hint
_L2P = int(os.environ.get('TGSR_L2_POLICY', '0'))


def conv_in(seq, x: Tensor, out: Blocked, x_sub: Optional[Tensor] = None) -> Blocked:
    """[conv3x3(3,2C), BN, GLU] from the NCHW fp32 image (util.py:741-744, model.py:228)."""
    conv = seq[0]
    bn = seq[1] if isinstance(seq[1], nn.BatchNorm2d) else None
    scale, shift = _fold(bn)
    x = x.contiguous()
    B, _, H, W = x.shape
    C = conv.weight.shape[0] // 2
    assert out.C == C and out.H == H and out.W == W
    xs = None if x_sub is None else x_sub.contiguous()
    call('tgsr_conv_in_blocked', x.device, ptr(x), ptr(xs), ptr(conv.weight.detach().contiguous()), ptr(scale), ptr(shift),
         ptr(out.buf), out.P, out.p0, B, H, W, C, out.fmt)
    return out


def res_block(x: Blocked, rb) -> Blocked:
    """ResBlock (util.py:110-130): two tcgen05 kernels, the skip add fused into the second."""
    blk = rb.block
    # Tile order alternates along the chain (see TGSR_CONV_REVERSE): the first conv walks backwards, so the second one,
    # walking forwards, starts where the intermediate tensor and the skip tensor were touched last.
    from ._lib import CONV_L2_IN_FIRST, CONV_L2_OUT_LAST
    y = conv_bn_act(x, blk[0], blk[1], ACT_GLU, reverse=True, l2_flags=CONV_L2_OUT_LAST if _L2P else 0)
    return conv_bn_act(y, blk[3], blk[4], ACT_NONE, residual=x,
                       l2_flags=(CONV_L2_IN_FIRST | (CONV_L2_OUT_LAST if _L2P == 1 else 0)) if _L2P else 0)


def body_noskip(x: Blocked, seq) -> Blocked:
    """residual24 / residual48 (model.py:229-232): conv-BN-GLU-conv-BN without skip."""
    y = conv_bn_act(x, seq[0], seq[1], ACT_GLU, reverse=True)
    return conv_bn_act(y, seq[3], seq[4], ACT_NONE)


def up_block(x: Blocked, seq, out_planes: Optional[int] = None, out: Optional[Blocked] = None) -> Blocked:
    """upBlock (util.py:74-80) as the 4-phase sub-pixel kernel; the output may be allocated with
    extra planes so the next stage's attention can append c_code in place."""
    conv, bn = seq[1], seq[2]
    C = conv.weight.shape[0] // 2
    if out is None:
        out = Blocked.empty(x.B, C, 2 * x.H, 2 * x.W, x.device, planes=out_planes, fmt=x.fmt)
    return conv_bn_act(x, conv, bn, ACT_GLU, up2x=True, out=out, reverse=True)


import os as _os
# 0 = off (default).  Measured on B200 (profiles/r2_experiments.md): with 2 / 4 chunks of the batch-32 step the layers'
# tensors fit in L2, yet every half-size kernel takes more than half the time (per-kernel fixed cost ~3-4 us) and the step
# slows from 0.72 to 0.76 / 0.83 ms -- the conv chain is not HBM-bound enough for the saved traffic to pay.
L2_CHUNK_BYTES = int(_os.environ.get('TGSR_L2_CHUNK_MB', '0')) << 20


def batch_chunks(B: int, bytes_per_image: int):
    """Split a batch so that the largest tensor a producer hands to its consumer stays L2-resident (126 MB L2 shared by
    reads, dirty writes and weights): the conv chain of a high-resolution stage streams ~3x fewer HBM bytes when each
    layer finds its input still in L2.  Returns [(b0, b1), ...] with equal power-of-two chunk counts."""
    n = 1
    while L2_CHUNK_BYTES > 0 and B % (2 * n) == 0 and (B // n) * bytes_per_image > L2_CHUNK_BYTES:
        n *= 2
    step = B // n
    return [(i * step, (i + 1) * step) for i in range(n)]


def _packed_head(conv: nn.Conv2d, fmt: int) -> Tensor:
    w = conv.weight
    key = (tensor_key(w), fmt)
    name = '_tgsr_packed_head_%d' % fmt
    cached = conv.__dict__.get(name)
    if cached is None or cached[0] != key:
        k = w.shape[2]
        packed = torch.empty(k * 16 * w.shape[1], device=w.device, dtype=tc._DTYPES[fmt])
        call('tgsr_pack_head', w.device, ptr(w.detach().contiguous()), ptr(packed), w.shape[1], k, fmt)
        cached = (key, packed)
        conv.__dict__[name] = cached
    return cached[1]


_NONFINITE = {}


def nonfinite_flag(device) -> Tensor:
    """Per-device int32 flag the image-head kernels OR a 1 into when a pre-activation sum is inf / NaN, i.e. when a 16-bit
    activation overflowed somewhere upstream (fp16 saturates at 65504).  Read it with `nonfinite_seen`."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device(device.type, torch.cuda.current_device())
    key = str(device)
    t = _NONFINITE.get(key)
    if t is None:
        t = _NONFINITE[key] = torch.zeros(1, device=device, dtype=torch.int32)
    return t


def nonfinite_seen(device, clear: bool = True) -> bool:
    """True when a 16-bit forward since the last call produced inf / NaN (one device->host read)."""
    t = nonfinite_flag(device)
    seen = bool(int(t.item()))
    if clear and seen:
        t.zero_()
    return seen


def image_head(x: Blocked, head_seq, addend: Optional[Tensor] = None, a: float = 0.0, one: float = 1.0,
               out: Optional[Tensor] = None) -> Tensor:
    """GET_IMAGE_G[_noAct].img / conv_output -> fp32 NCHW image (util.py:894-919, model.py:223-226,280)."""
    conv = head_seq[0]
    act = ACT_TANH if any(isinstance(m, nn.Tanh) for m in head_seq) else ACT_NONE
    k = conv.weight.shape[2]
    y = torch.empty((x.B, 3, x.H, x.W), device=x.device, dtype=torch.float32) if out is None else out
    assert y.is_contiguous() and tuple(y.shape) == (x.B, 3, x.H, x.W)
    ad = None
    if addend is not None:
        if tuple(addend.shape) != tuple(y.shape):
            raise ValueError('SRb shape %s != image shape %s' % (tuple(addend.shape), tuple(y.shape)))
        ad = addend.contiguous()
    call('tgsr_img_head_tc', x.device, ptr(x.buf), x.P, x.p0, ptr(_packed_head(conv, x.fmt)), ptr(ad), float(one), float(a),
         ptr(y), x.B, x.C, x.H, x.W, k, act, x.fmt, ptr(nonfinite_flag(x.device)))
    return y


def _sfrag_words(B: int, T: int) -> int:
    return 16 * (3 if T <= 24 else 4 if T <= 32 else 6) * 32 * B      # = tgsr_attn_sfrag_bytes_t(B, T) / 4


def attention_project(att, words: Tensor, fmt: int) -> Tensor:
    """S = conv_context(words) (GlobalAttention.py:99-101) as the tensor-core operand image of the blocked
    attention kernel; depends on the caption only, so a forward computes it once per stage up front."""
    B, cdf, T = words.shape
    w = att.conv_context.weight
    sfrag = torch.empty(_sfrag_words(B, T), device=words.device, dtype=torch.int32)
    words, wc = words.contiguous(), w.detach().contiguous()                  # raw pointers: keep both alive across the call
    call('tgsr_attn_project', words.device, ptr(words), ptr(wc), ptr(sfrag), B, w.shape[0], cdf, T, fmt)
    return sfrag


def attention_project_multi(atts, words: Tensor, fmt: int):
    """Operand images of up to three attention modules for the same captions, one launch."""
    B, cdf, T = words.shape
    ws = [a.conv_context.weight.detach().contiguous() for a in atts]
    sf = [torch.empty(_sfrag_words(B, T), device=words.device, dtype=torch.int32) for _ in atts]
    n = len(atts)
    p = lambda lst, k: ptr(lst[k]) if k < n else None
    call('tgsr_attn_project_multi', words.device, ptr(words.contiguous()), p(ws, 0), p(ws, 1), p(ws, 2), p(sf, 0), p(sf, 1), p(sf, 2),
         n, B, ws[0].shape[0], cdf, T, fmt)
    return sf


def attention_blocked_supported(idf: int, T: int, Q: int) -> bool:
    return idf == 32 and T <= 48 and Q % 32 == 0


def attention(h: Blocked, att, words: Tensor, mask: Optional[Tensor], out: Blocked, sfrag: Optional[Tensor] = None) -> Tensor:
    """GlobalAttentionGeneral on blocked activations; c_code goes to `out` (normally the next planes
    of h's own buffer).  Returns the fp32 attention map [B,T,ih,iw]."""
    B, T = words.shape[0], words.shape[2]
    idf, cdf = att.conv_context.weight.shape[0], att.conv_context.weight.shape[1]
    m8 = mask if mask is None or mask.dtype == torch.uint8 else mask.to(device=h.device, dtype=torch.uint8).contiguous()
    if attention_blocked_supported(idf, T, h.H * h.W):
        if sfrag is None:
            sfrag = attention_project(att, words, h.fmt)
        attn = torch.empty((B, T, h.H, h.W), device=h.device, dtype=torch.float32)
        call('tgsr_attn_fwd_blocked', h.device, ptr(h.buf), h.P, h.p0, ptr(sfrag), ptr(m8), ptr(out.buf), out.P, out.p0,
             ptr(attn), B, idf, h.H * h.W, T, h.fmt)
        return attn
    # outside the blocked kernel's envelope (long captions, odd sizes): fp32 kernel + layout conversion
    wctx, attn = ops.attention_fwd(tc.to_nchw(h), words, att.conv_context.weight.detach(), mask)
    tc.to_blocked(wctx, out=out)
    return attn


# ---------------------------------------------------------------- stages and nets
def stage_im2f(st, LR: Tensor, fmt: int = 0) -> Blocked:
    """im2f of INIT_STAGE_GImgup (util.py:741-744, :769) into the first half of a concat-sized buffer.  It needs nothing
    from the text encoder, so the eval pipeline runs it on its own stream next to the LSTM."""
    B, _, H, W = LR.shape
    cat = Blocked.empty(B, 2 * st.gf_dim, H, W, LR.device, fmt=fmt)
    return conv_in(st.im2f, LR, cat.slice(0, st.gf_dim))


def stage(st, h: Optional[Blocked], LR: Optional[Tensor], words: Tensor, mask, out_planes: Optional[int],
          fmt: int = 0, sfrag: Optional[Tensor] = None, head=None, head_stream=None):
    """INIT_STAGE_GImgup (LR given) / NEXT_STAGE_G (h given); util.py:763-777 / :807-823.
    `h` must be the first 32-channel view of a 64-channel buffer: c_code is appended in place.
    With `head` (the stage's image head, model.py:55-73) the stage also returns its image, and the conv chain behind the
    attention -- residual blocks, up-block, head -- runs per batch chunk so that it stays in L2 (`batch_chunks`).  The
    attention itself always sees the whole batch: its caption mask is indexed with the reference's row-order quirk
    (GlobalAttention.py:109-112), which depends on the batch size.
    `head_stream`: the image head is enqueued there instead of the current stream -- nothing downstream in G_SR_NET_low reads
    the image, so the (HBM-bound) head overlaps the next stage; the caller joins the stream where the image is consumed."""
    ngf = st.gf_dim
    if h is None:
        h = stage_im2f(st, LR, fmt)
        cat = Blocked(h.buf, 0, 2 * ngf)
    else:
        assert h.p0 == 0 and h.P * 8 >= 2 * ngf, 'h_code must live in a concat-sized buffer'
        cat = Blocked(h.buf, 0, 2 * ngf)
    att = attention(h, st.att, words, mask, out=cat.slice(ngf, ngf), sfrag=sfrag)
    B, H, W = cat.B, cat.H, cat.W
    up_conv = st.upsample[1]
    Cup = up_conv.weight.shape[0] // 2
    out = Blocked.empty(B, Cup, 2 * H, 2 * W, cat.device, planes=out_planes, fmt=cat.fmt)
    img = torch.empty((B, 3, 2 * H, 2 * W), device=cat.device, dtype=torch.float32) if head is not None else None
    for b0, b1 in batch_chunks(B, 4 * H * W * Cup * 2):                   # largest tensor of the chain: the up-block output
        x = cat.batch(b0, b1)
        for rb in st.residual:
            x = res_block(x, rb)
        o = up_block(x, st.upsample, out=out.batch(b0, b1))
        if head is not None:
            if head_stream is None:
                image_head(o, head.img, out=img[b0:b1])
            else:
                head_stream.wait_stream(torch.cuda.current_stream(cat.device))
                with torch.cuda.stream(head_stream):
                    image_head(o, head.img, out=img[b0:b1])
                out.buf.record_stream(head_stream); img.record_stream(head_stream)
    return out, att, img


def stage_supported(st, H: int, W: int) -> bool:
    return st.gf_dim == 32 and H % 16 == 0 and W % 8 == 0 and getattr(st, 'batchnorm', True)


def g_sr_net_low(net, LR: Tensor, sent: Tensor, words: Tensor, mask, stages, heads, outmiddle: bool, fmt: int = 0,
                 h0: Optional[Blocked] = None, ca=None, head_stream=None, head_events: Optional[list] = None):
    """G_SR_NET_low.forward (model.py:48-78) on blocked activations.  `h0`: stage_im2f(stages[0], LR) computed ahead;
    `ca`: callable returning ca_net(sent) = (c_code, mu, logvar), for callers that evaluate it off the critical path
    (c_code is unused by the SR nets, model.py:51-52).  `head_stream` + `head_events`: the image heads run on that stream and
    one event per image is appended to `head_events`; the caller must wait for event k (or the stream) before reading image k."""
    imgs: List[Tensor] = []
    atts: List[Tensor] = []
    feats: List[Blocked] = []
    if mask is not None and mask.dtype != torch.uint8:       # one conversion per forward, not one per stage
        mask = mask.to(device=LR.device, dtype=torch.uint8).contiguous()
    sfrags = text_side_work(stages, words, LR.shape[2] * LR.shape[3], fmt)
    c_code, mu, logvar = net.ca_net(sent) if ca is None else ca()
    h = h0
    n = len(stages)
    for i, (st, hd) in enumerate(zip(stages, heads)):
        planes = 8 if i + 1 < n else 4                     # leave room for the next stage's c_code
        h, att, img = stage(st, h, LR if (i == 0 and h0 is None) else None, words, mask, planes, fmt, sfrags[i], head=hd,
                            head_stream=head_stream)
        if head_stream is not None and head_events is not None:
            ev = torch.cuda.Event()
            ev.record(head_stream)
            head_events.append(ev)
        imgs.append(img)
        atts.append(att)
        feats.append(h)
    if outmiddle:
        return imgs, atts, mu, logvar, [tc.to_nchw(f) for f in feats]
    return imgs, atts, mu, logvar


def text_side_work(stages, words: Tensor, lr_pixels: int, fmt: int):
    """The word projections depend on the captions only: the operand images of all stages in one launch."""
    ok = [attention_blocked_supported(st.att.conv_context.weight.shape[0], words.shape[2], lr_pixels * 4 ** i)
          for i, st in enumerate(stages)]
    sfrags = [None] * len(stages)
    idx = [i for i, o in enumerate(ok) if o]
    for k in range(0, len(idx), 3):
        grp = idx[k:k + 3]
        for i, sf in zip(grp, attention_project_multi([stages[i].att for i in grp], words, fmt)):
            sfrags[i] = sf
    return sfrags


def netg_high_trunk(net, x: Tensor, x_sub: Optional[Tensor], fmt: int = 0) -> Blocked:
    """First, text-independent part of NetG_high / NetG_highweight (model.py:264-279): convin, six ResBlocks and
    upscale2x.  It needs nothing from the text encoder or G_SR_NET_low, so the eval pipeline runs it on a second
    stream next to them."""
    B, _, H, W = x.shape
    h = conv_in(net.convin, x, Blocked.empty(B, 32, H, W, x.device, fmt=fmt), x_sub)
    for rb in net.residual:
        h = res_block(h, rb)
    return up_block(h, net.upscale2x)


def netg_high_finish(net, h: Blocked, SRb, a: float, one: float, tail, feats=None, srb_events=None) -> List[Tensor]:
    """Rest of the trunk (model.py:280-298): image heads with the fused `one*y + a*SRb`, residual24/48 bodies and
    upscale4x/8x.  `tail` = [(body, upscale), ...] after the first upscale.  Each level's chain (body, up-block, head) runs
    per batch chunk so that it stays in L2 (`batch_chunks`)."""
    if feats is not None: feats.append(tc.to_nchw(h))
    cur = torch.cuda.current_stream(h.device)
    if srb_events:                                              # SRb[k] was produced on another stream (see g_sr_net_low)
        cur.wait_event(srb_events[0])
    ims = [image_head(h, net.conv_output, SRb[0], a, one)]
    for k, (body, up) in enumerate(tail):
        B, H, W = h.B, h.H, h.W
        Cup = up[1].weight.shape[0] // 2
        out = Blocked.empty(B, Cup, 2 * H, 2 * W, h.device, fmt=h.fmt)
        img = torch.empty((B, 3, 2 * H, 2 * W), device=h.device, dtype=torch.float32)
        srb = SRb[k + 1]
        if tuple(srb.shape) != tuple(img.shape):
            raise ValueError('SRb shape %s != image shape %s' % (tuple(srb.shape), tuple(img.shape)))
        srb = srb.contiguous()
        for ci, (b0, b1) in enumerate(batch_chunks(B, 4 * H * W * Cup * 2)):
            o = up_block(body_noskip(h.batch(b0, b1), body), up, out=out.batch(b0, b1))
            if srb_events and ci == 0:
                cur.wait_event(srb_events[k + 1])               # as late as possible: only this head reads SRb[k + 1]
            image_head(o, net.conv_output, srb[b0:b1], a, one, out=img[b0:b1])
        h = out
        if feats is not None: feats.append(tc.to_nchw(h))
        ims.append(img)
    return ims


def netg_high(net, x: Tensor, x_sub: Optional[Tensor], SRb, a: float, one: float, tail, feats=None,
              fmt: int = 0) -> List[Tensor]:
    """Trunk of NetG_high / NetG_highweight (model.py:264-298)."""
    return netg_high_finish(net, netg_high_trunk(net, x, x_sub, fmt), SRb, a, one, tail, feats)
