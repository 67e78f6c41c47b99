"""Generate tests/golden/*.pt by EXECUTING THE REFERENCE'S OWN MODULES (authoring container only).

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden
Requires /root/reference (read-only mount).  Each fixture stores the seeded
inputs, the weights (state_dict of the reference module) and the reference's
outputs, all fp32 CPU tensors.  The reference is imported unmodified via
oracle/ref_loader.py; nothing from it is copied into this repository.
"""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                      # noqa: E402
from oracle import tgsr_oracle as O                # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def randomize_bn(module, gen):
    """Give every BatchNorm non-trivial eval statistics (SURVEY.md 8d cfg 2)."""
    import torch.nn as nn
    for m in module.modules():
        if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)
            m.weight.data.copy_(1 + 0.02 * torch.randn(m.weight.shape, generator=gen))
            m.bias.data.copy_(0.05 * torch.randn(m.bias.shape, generator=gen))


def clone_sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def gen_attention(ns):
    cases = []
    g = torch.Generator().manual_seed(1001)
    specs = [  # B, idf, cdf, ih, iw, T, mask kind
        (1, 32, 256, 8, 8, 18, 'none'),
        (4, 32, 256, 8, 8, 18, 'ragged'),
        (3, 32, 256, 16, 8, 11, 'ragged'),
        (2, 32, 256, 16, 16, 18, 'equal'),
        (4, 64, 64, 8, 8, 32, 'ragged'),
        (2, 128, 32, 4, 8, 48, 'ragged'),
        (5, 32, 256, 6, 6, 7, 'ragged'),
    ]
    for (B, idf, cdf, ih, iw, T, mk) in specs:
        att = ns.GlobalAttention.GlobalAttentionGeneral(idf, cdf)
        att.conv_context.weight.data.copy_(torch.randn(idf, cdf, 1, 1, generator=g) * 0.1)
        inp = torch.randn(B, idf, ih, iw, generator=g) * 0.5      # logits O(3): fp32 rounding stays < 1e-5 rel
        ctx = torch.randn(B, cdf, T, generator=g) * 0.5
        if mk == 'none':
            mask = None
        elif mk == 'equal':
            mask = torch.zeros(B, T, dtype=torch.bool)
        else:
            lens = torch.randint(2, T + 1, (B,), generator=g)
            lens[0] = T
            mask = torch.arange(T)[None, :] >= lens[:, None]
        att.applyMask(mask)
        with torch.no_grad():
            wctx, attn = att(inp, ctx)
        cases.append(dict(inp=inp, ctx=ctx, w=att.conv_context.weight.data.clone(), mask=mask,
                          wctx=wctx.clone(), attn=attn.clone()))
    torch.save(cases, os.path.join(OUT, 'attention.pt'))


def gen_func_attention(ns):
    cases = []
    g = torch.Generator().manual_seed(1002)
    for (B, ndf, Tq, ih, iw, gamma) in [(2, 256, 18, 17, 17, 4.0), (3, 64, 7, 5, 6, 5.0)]:
        q = torch.randn(B, ndf, Tq, generator=g) * 0.3
        c = torch.randn(B, ndf, ih, iw, generator=g) * 0.3
        w, a = ns.GlobalAttention.func_attention(q, c, gamma)
        cases.append(dict(query=q, context=c, gamma1=gamma, wctx=w.clone(), attn=a.clone()))
    torch.save(cases, os.path.join(OUT, 'func_attention.pt'))


def gen_blocks(ns):
    u = ns.util
    g = torch.Generator().manual_seed(1003)
    torch.manual_seed(1003)
    out = {}
    with torch.no_grad():
        for name, mod, x in [
            ('block3x3_3_32', u.Block3x3_relu(3, 32), torch.randn(2, 3, 16, 16, generator=g)),
            ('block3x3_32_32', u.Block3x3_relu(32, 32), torch.randn(2, 32, 16, 8, generator=g)),
            ('resblock_64', u.ResBlock(64), torch.randn(2, 64, 16, 16, generator=g)),
            ('resblock_32', u.ResBlock(32), torch.randn(1, 32, 32, 32, generator=g)),
            ('upblock_64_32', u.upBlock(64, 32), torch.randn(2, 64, 16, 8, generator=g)),
            ('upblock_32_32', u.upBlock(32, 32), torch.randn(1, 32, 16, 16, generator=g)),
            ('get_image', u.GET_IMAGE_G(32), torch.randn(2, 32, 16, 16, generator=g)),
            ('get_image_noact', u.GET_IMAGE_G_noAct(32), torch.randn(2, 32, 16, 16, generator=g)),
        ]:
            randomize_bn(mod, g)
            mod.eval()
            out[name] = dict(sd=clone_sd(mod), x=x, y=mod(x).clone())
        ca = u.CA_NET().eval()
        s = torch.randn(3, 256, generator=g)
        x = ca.relu(ca.fc(s))
        out['ca_net'] = dict(sd=clone_sd(ca), x=s, mu=x[:, :100].clone(), logvar=x[:, 100:].clone())
        ig = u.INIT_STAGE_G(32, 100)
        randomize_bn(ig, g)
        ig.eval()
        z = torch.randn(2, 100, generator=g); c = torch.randn(2, 100, generator=g)
        out['init_stage_g'] = dict(sd=clone_sd(ig), z=z, c=c, y=ig(z, c).clone())
    torch.save(out, os.path.join(OUT, 'blocks.pt'))


def gen_rnn(ns):
    torch.manual_seed(100)                                     # test1.py:170
    enc = ns.util.RNN_ENCODER(41, nhidden=256).eval()
    caps_all = pickle.load(open(os.path.join(ref_loader.REF_ROOT, 'data/face/testset.pickle'), 'rb'))[0]
    order = sorted(range(len(caps_all)), key=lambda i: -len(caps_all[i]))[:8]
    lens = torch.tensor([len(caps_all[i]) for i in order])
    caps = torch.zeros(len(order), 18, dtype=torch.int64)
    for r, i in enumerate(order):
        caps[r, :len(caps_all[i])] = torch.tensor(caps_all[i])
    with torch.no_grad():
        words, sent = enc(caps, lens, enc.init_hidden(len(order)))
    torch.save(dict(sd=clone_sd(enc), captions=caps, cap_lens=lens, words=words.contiguous().clone(),
                    sent=sent.clone()), os.path.join(OUT, 'rnn_encoder.pt'))


def real_lr():
    """data/face/000155.png -> Resize(288) -> CenterCrop(256) -> Resize(32) -> Normalize(.5,.5)
    (datasets.py:1558-1563, :216)."""
    from PIL import Image
    from torchvision import transforms
    img = Image.open(os.path.join(ref_loader.REF_ROOT, 'data/face/000155.png')).convert('RGB')
    img = transforms.Compose([transforms.Resize(int(256 * 72 / 64)), transforms.CenterCrop(256)])(img)
    lr = transforms.Resize([32, 32])(img)
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    return norm(lr)[None], norm(img)[None]


def gen_face_x8(ns):
    """Shipped Checkpoint/face_S8 weights, real LR image, real caption indices (B=1 and B=3 ragged)."""
    ck = os.path.join(ref_loader.REF_ROOT, 'Checkpoint/face_S8')
    sdL = torch.load(os.path.join(ck, 'netG_epoch_7.pth'), map_location='cpu', weights_only=False)
    sdH = torch.load(os.path.join(ck, 'netGH_epoch_7.pth'), map_location='cpu', weights_only=False)
    sdL = {k: v.float().clone() if v.is_floating_point() else v.clone() for k, v in sdL.items()}
    sdH = {k: v.float().clone() if v.is_floating_point() else v.clone() for k, v in sdH.items()}
    torch.save(dict(netG=sdL, netGH=sdH), os.path.join(OUT, 'face_S8_weights.pt'))

    netGL = ns.model.G_SR_NET_low(); netGL.load_state_dict(sdL); netGL.eval()
    with ref_loader.cuda_noop():
        netGH = ns.model.NetG_highweight(weightmap=False, low='lr')
    netGH.load_state_dict(sdH); netGH.eval()
    rnn = torch.load(os.path.join(OUT, 'rnn_encoder.pt'))
    torch.manual_seed(100)
    enc = ns.util.RNN_ENCODER(41, nhidden=256).eval(); enc.load_state_dict(rnn['sd'])
    lr1, hr1 = real_lr()
    out = {}
    for tag, idx in (('b1', [0]), ('b3', [0, 2, 5])):
        caps = rnn['captions'][idx]; lens = rnn['cap_lens'][idx]
        B = len(idx)
        LR = lr1.repeat(B, 1, 1, 1).clone()
        if B > 1:   # perturb the copies so images differ
            g = torch.Generator().manual_seed(7)
            LR = (LR + 0.05 * torch.randn(LR.shape, generator=g)).clamp(-1, 1)
        with torch.no_grad():
            words, sent = enc(caps, lens, enc.init_hidden(B))
            mask = (caps == 0)[:, :words.size(2)]                 # trainer_objective.py:136-140
            imgsL, atts, mu, logvar = netGL(LR, sent, words, mask)
            fine, a, one = netGH(LR, imgsL, LR)
        assert float(a) == 0.5 and float(one) == 1.0
        if B == 1:     # full outputs; the 128^2 attention map is stored 4x-subsampled (size)
            out[tag] = dict(LR=LR, captions=caps, cap_lens=lens, words=words.contiguous().clone(), sent=sent.clone(),
                            mask=mask.clone(), imgsL=[t.clone() for t in imgsL],
                            atts=[atts[0].clone(), atts[1].clone()], att2_sub4=atts[2][:, :, ::4, ::4].clone(),
                            mu=mu.clone(), logvar=logvar.clone(), fine=[t.clone() for t in fine])
        else:          # ragged batch (mask quirk through the whole net); small slices only
            out[tag] = dict(LR=LR, captions=caps, cap_lens=lens, words=words.contiguous().clone(), sent=sent.clone(),
                            mask=mask.clone(), imgL0=imgsL[0].clone(), att0=atts[0].clone(),
                            att1_sub2=atts[1][:, :, ::2, ::2].clone(), fine0=fine[0].clone(),
                            fine2_sub4=fine[2][:, :, ::4, ::4].clone(), mu=mu.clone(), logvar=logvar.clone())
    out['hr'] = hr1
    torch.save(out, os.path.join(OUT, 'face_x8.pt'))


def gen_x16(ns):
    """models16 (x16, weight-shared stages, tanh heads).  netGL is built with GF_DIM=8 to keep the
    fixture small (it pins the WIRING of models16.py:5-39); NetG_high (models16.py:42-94) hard-codes
    32 channels and is fed synthetic SRb pyramids."""
    g = torch.Generator().manual_seed(1016)
    torch.manual_seed(100)
    ns.cfg.GAN.GF_DIM = 8
    try:
        netGL = ns.models16.G_SR_NET_low()
    finally:
        ns.cfg.GAN.GF_DIM = ref_loader.EVAL_CFG['GF_DIM']
    randomize_bn(netGL, g); netGL.eval()
    netGH = ns.models16.NetG_high(); randomize_bn(netGH, g); netGH.eval()
    LR = O.synth_lr(2, 16, seed=116)
    words = torch.randn(2, 256, 9, generator=g) * 0.2
    sent = torch.randn(2, 256, generator=g) * 0.2
    mask = torch.zeros(2, 9, dtype=torch.bool); mask[1, 6:] = True
    LRh = O.synth_lr(1, 8, seed=117)
    SRb = [O.synth_lr(1, 8 * 2 ** k, seed=118 + k) for k in range(1, 5)]
    with torch.no_grad():
        imgsL, atts, mu, logvar = netGL(LR, sent, words, mask)
        fine = netGH(LRh, SRb, LRh, low=True)
    torch.save(dict(sdL=netGL.state_dict(), sdH=netGH.state_dict(), LR=LR, words=words, sent=sent, mask=mask,
                    imgsL=[t.clone() for t in imgsL[:3]], imgL3_sub2=imgsL[3][:, :, ::2, ::2].clone(),
                    atts=[t.clone() for t in atts[:3]], att3_sub4=atts[3][:, :, ::4, ::4].clone(),
                    LRh=LRh, SRb=SRb, fine=[t.clone() for t in fine]),
               os.path.join(OUT, 'x16.pt'))


def gen_x16_gf32(ns):
    """models16 at PRODUCTION width (GF_DIM = 32; 16x16 -> 256x256): G_SR_NET_low, NetG_high and NetG_highweight.

    NetG_highweight x16 cannot finish its own forward on a real pyramid: its last stage adds SRb8 (128^2) to a 256^2
    tensor (models16.py:178).  To pin everything the reference DOES define -- including the reuse of residual48 /
    upscale8x for the last stage (:172-173) and which addend it takes -- the unmodified forward is executed with SRb[2]
    given as a per-image constant colour [B,3,1,1], which broadcasts against both the 128^2 and the 256^2 sums."""
    g = torch.Generator().manual_seed(1632)
    torch.manual_seed(132)
    netGL = ns.models16.G_SR_NET_low(); randomize_bn(netGL, g); netGL.eval()
    netGH = ns.models16.NetG_high(); randomize_bn(netGH, g); netGH.eval()
    with ref_loader.cuda_noop():
        netGHw = ns.models16.NetG_highweight(weightmap=False, low='lr')
    randomize_bn(netGHw, g); netGHw.eval()
    B, T = 2, 18
    LR = O.synth_lr(B, 16, seed=216)
    LRb = (LR + 0.05 * torch.randn(LR.shape, generator=g)).clamp(-1, 1)
    words = torch.randn(B, 256, T, generator=g) * 0.2
    sent = torch.randn(B, 256, generator=g) * 0.2
    mask = torch.zeros(B, T, dtype=torch.bool); mask[1, 11:] = True
    c8 = torch.rand(B, 3, 1, 1, generator=g) * 2 - 1
    with torch.no_grad():
        imgsL, atts, mu, logvar = netGL(LR, sent, words, mask)
        fine = netGH(LR, imgsL, LRb, low=False)                                # convin(LR - LRb)
        finew, a, one = netGHw(LR, [imgsL[0], imgsL[1], c8, imgsL[3]], LRb)
    assert float(a) == 0.5 and float(one) == 1.0
    sub = lambda t, k: t[:, :, ::k, ::k].clone()
    torch.save(dict(sdL=netGL.state_dict(), sdH=netGH.state_dict(), sdHw=netGHw.state_dict(), LR=LR, LRb=LRb, words=words,
                    sent=sent, mask=mask, c8=c8, mu=mu.clone(), logvar=logvar.clone(),
                    imgsL=[imgsL[0].clone(), imgsL[1].clone(), sub(imgsL[2], 2), sub(imgsL[3], 4)],
                    atts=[atts[0].clone(), sub(atts[1], 2), sub(atts[2], 4), sub(atts[3], 8)],
                    fine=[fine[0].clone(), fine[1].clone(), sub(fine[2], 2), sub(fine[3], 4)],
                    finew=[finew[0].clone(), finew[1].clone(), sub(finew[2], 2), sub(finew[3], 4)]),
               os.path.join(OUT, 'x16_gf32.pt'))


def gen_attn_stats(ns):
    """Pins ops.attention_word_stats / the oracle's restatement by EXECUTING the unmodified miscc/utils.py: build_super_images
    (the per-pixel max over words it prepends, :122-131) and build_super_imagesall (the thresholded confidence sums,
    :378-392).  Both keep those values in locals, so the module's `np` is wrapped for the duration of the calls to record the
    operands of `np.transpose` (the [max-map, word maps] stack) and the results of `np.sum` (conf_score); skimage (absent) is
    an import shim and the hard-coded Windows font falls back to PIL's default -- neither touches the recorded arithmetic."""
    import numpy as np
    shims = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', 'shims')
    if shims not in sys.path:
        sys.path.insert(0, shims)
    from PIL import ImageFont
    truetype = ImageFont.truetype
    default_font = ImageFont.load_default()
    ImageFont.truetype = lambda *a, **k: default_font
    from miscc import utils as U

    class NpProxy(object):
        def __init__(self):
            self.sums, self.transposed = [], []

        def __getattr__(self, name):
            return getattr(np, name)

        def sum(self, a, *args, **kw):
            r = np.sum(a, *args, **kw)
            self.sums.append(float(r))
            return r

        def transpose(self, a, *args, **kw):
            self.transposed.append(np.array(a, copy=True))
            return np.transpose(a, *args, **kw)

    g = torch.Generator().manual_seed(1404)
    B, T, s = 2, 18, 32
    attn = torch.softmax(torch.randn(B, T, s, s, generator=g) * 2.0, dim=1)
    attn[0, :, 0, 0] = 1.0 / T                                     # an exact tie
    lens = [18, 13]
    caps = torch.zeros(B, T, dtype=torch.int64)
    for b in range(B):
        caps[b, :lens[b]] = torch.randint(1, 40, (lens[b],), generator=g)
    imgs = torch.rand(B, 3, 256, 256, generator=g) * 2 - 1
    ixtoword = {i: 'w%d' % i for i in range(64)}
    proxy = NpProxy()
    U.np = proxy
    try:
        U.build_super_images(imgs.clone(), caps, ixtoword, [attn[b:b + 1] for b in range(B)], s, batch_size=B, max_word_num=T)
        stacks = [a for a in proxy.transposed if a.ndim == 4 and a.shape[1] == 3 and a.shape[0] == T + 1]
        assert len(stacks) == B
        max_maps = torch.from_numpy(np.stack([a[0, 0] for a in stacks]))          # [B, s, s]
        conf = []
        for b in range(B):
            proxy.sums = []
            U.build_super_imagesall(imgs[b:b + 1].clone(), caps[b:b + 1], [lens[b]], ixtoword, [attn[b, :lens[b]]], s)
            conf.append(torch.tensor(proxy.sums[:lens[b]], dtype=torch.float64))
            assert len(proxy.sums) >= lens[b]
    finally:
        U.np = np
        ImageFont.truetype = truetype
    torch.save(dict(attn=attn, lens=lens, max_maps=max_maps, conf=conf), os.path.join(OUT, 'attn_stats.pt'))


def gen_losses(ns):
    """words_loss / sent_loss / MSE of the UNMODIFIED miscc/losses.py (gamma 4/5/10, cfg/DAMSM/bird.yml:23-25) with
    autograd gradients to the image side: the pin for tgsr_b200/losses.py and the oracle's restatement."""
    from miscc import losses as L
    L.server = 1                                            # bool masks (uint8 masked_fill_ fails on torch 2.x)
    ns.cfg.TRAIN.SMOOTH.GAMMA1, ns.cfg.TRAIN.SMOOTH.GAMMA2, ns.cfg.TRAIN.SMOOTH.GAMMA3 = 4.0, 5.0, 10.0
    ns.cfg.TRAIN.SMOOTH.LAMBDA = 1.0
    import numpy as np
    cases = []
    g = torch.Generator().manual_seed(1007)
    for (B, T, ragged, with_cls) in [(3, 18, True, False), (4, 12, True, True), (2, 18, False, False)]:
        feats = (torch.randn(B, 256, 17, 17, generator=g) * 0.4).half().float().requires_grad_(True)   # fp16-exact: small fixture
        words = torch.randn(B, 256, T, generator=g) * 0.4
        cnn = (torch.randn(B, 256, generator=g) * 0.5).requires_grad_(True)
        sent = torch.randn(B, 256, generator=g) * 0.5
        lens = torch.randint(3, T + 1, (B,), generator=g) if ragged else torch.full((B,), T)
        lens = torch.sort(lens, descending=True).values
        labels = torch.arange(B)
        cls = np.array([0, 1, 0, 2][:B]) if with_cls else None
        w0, w1, att = L.words_loss(feats, words, labels, lens, cls, B)
        s0, s1 = L.sent_loss(cnn, sent, labels, cls, B)
        gw = torch.tensor([0.7, 1.3]); gs = torch.tensor([1.1, 0.6])
        (gw[0] * w0 + gw[1] * w1 + gs[0] * s0 + gs[1] * s1).backward()
        fake = [torch.randn(B, 3, 4 * 2 ** k, 4 * 2 ** k, generator=g).requires_grad_(True) for k in range(3)]
        real = [torch.randn(B, 3, 4 * 2 ** k, 4 * 2 ** k, generator=g) for k in range(3)]
        mse = L.MSE(fake, real)
        (1.7 * mse).backward()
        cases.append(dict(feats=feats.detach().half(), words=words, cnn=cnn.detach().clone(), sent=sent, lens=lens, labels=labels,
                          class_ids=None if cls is None else cls.tolist(), w0=w0.detach(), w1=w1.detach(), s0=s0.detach(),
                          s1=s1.detach(), att=[a.detach().clone() for a in att], gw=gw, gs=gs, d_feats=feats.grad.clone(),
                          d_cnn=cnn.grad.clone(), fake=[f.detach().clone() for f in fake], real=real, mse=mse.detach(),
                          d_fake=[f.grad.clone() for f in fake], gamma=(4.0, 5.0, 10.0)))
    torch.save(cases, os.path.join(OUT, 'losses.pt'))


def gen_train(ns):
    """The UNMODIFIED reference generator modules in train() mode (batch-statistics BatchNorm) on a small batch: outputs,
    an MSE loss over all six images and autograd gradients.  Pins the oracle's train_mode() and, on the GPU, the explicit
    backward kernels.  Weights are regenerated from the seed by the tests (state dicts are not stored)."""
    torch.manual_seed(4242)
    netGL = ns.model.G_SR_NET_low().train()
    with ref_loader.cuda_noop():
        netGH = ns.model.NetG_highweight(weightmap=False, low='lr')
    netGH.train()
    g = torch.Generator().manual_seed(4243)
    for m in list(netGL.modules()) + list(netGH.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.copy_(1 + 0.2 * torch.randn(m.weight.shape, generator=g)); m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    for net in (netGL, netGH):                     # fp16-exact weights: the fixture stores them as halves
        for p_ in net.parameters():
            p_.data.copy_(p_.data.half().float())
    half = lambda sd: {k: (v.half() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    sdL = half(netGL.state_dict()); sdH = half(netGH.state_dict())
    B, T = 2, 18
    LR = O.synth_lr(B, 16, seed=4244)
    words = torch.randn(B, 256, T, generator=g) * 0.5; sent = torch.randn(B, 256, generator=g) * 0.5
    mask = torch.zeros(B, T, dtype=torch.bool); mask[1, 11:] = True
    targets = [torch.randn(B, 3, 16 * 2 ** (k + 1), 16 * 2 ** (k + 1), generator=g) * 0.5 for k in range(3)]
    imgs, atts, mu, lv = netGL(LR, sent, words, mask)
    fine, a, one = netGH(LR, imgs, LR)
    loss = sum(torch.nn.functional.mse_loss(f, t) for f, t in zip(fine, targets)) + sum(torch.nn.functional.mse_loss(f, t) for f, t in zip(imgs, targets))
    loss.backward()
    pick = lambda net, names: {n: dict(net.named_parameters())[n].grad.clone() for n in names}
    gl_names = ['h_net1.im2f.0.weight', 'h_net1.im2f.1.weight', 'h_net1.att.conv_context.weight', 'h_net2.residual.1.block.3.weight',
                'h_net3.upsample.1.weight', 'h_net3.upsample.2.bias', 'img_net2.img.0.weight']
    gh_names = ['convin.0.weight', 'residual.3.block.0.weight', 'residual.5.block.4.weight', 'upscale4x.1.weight', 'residual48.1.bias',
                'conv_output.0.weight']
    torch.save(dict(sdL=sdL, sdH=sdH, LR=LR, words=words, sent=sent, mask=mask, targets=[t.half() for t in targets],
                    loss=loss.detach(), img0=imgs[0].detach().clone(), fine0=fine[0].detach().clone(),
                    fine2_sub4=fine[2].detach()[:, :, ::4, ::4].clone(), gradL=pick(netGL, gl_names), gradH=pick(netGH, gh_names),
                    run_mean=netGL.h_net1.im2f[1].running_mean.clone(), run_var=netGL.h_net1.im2f[1].running_var.clone()),
               os.path.join(OUT, 'train_step.pt'))


def gen_misc_losses(ns):
    """KL_loss and cosine_similarity of the UNMODIFIED miscc/losses.py (:806-810, :12-18) with autograd gradients."""
    from miscc import losses as L
    g = torch.Generator().manual_seed(1203)
    cases = []
    for shape, dim in [((4, 100), 1), ((3, 256, 7), 1), ((2, 5, 64), 2), ((6, 256), 1)]:
        x1 = torch.randn(shape, generator=g).requires_grad_(True)
        x2 = torch.randn(shape, generator=g).requires_grad_(True)
        if len(shape) == 2 and shape[0] == 6:
            x1.data[0] = 0.0                                   # zero vector: the eps clamp decides the value and the gradient
        c = L.cosine_similarity(x1, x2, dim)
        gc = torch.randn(c.shape, generator=g)
        (c * gc).sum().backward()
        cases.append(dict(kind='cos', x1=x1.detach().clone(), x2=x2.detach().clone(), dim=dim, out=c.detach().clone(), g=gc,
                          d1=x1.grad.clone(), d2=x2.grad.clone()))
    for shape in [(4, 100), (32, 100)]:
        mu = (torch.randn(shape, generator=g) * 0.7).requires_grad_(True)
        lv = (torch.randn(shape, generator=g) * 0.5).requires_grad_(True)
        k = L.KL_loss(mu, lv)
        (2.5 * k).backward()
        cases.append(dict(kind='kl', mu=mu.detach().clone(), logvar=lv.detach().clone(), out=k.detach().clone(), gscale=2.5,
                          d_mu=mu.grad.clone(), d_logvar=lv.grad.clone()))
    torch.save(cases, os.path.join(OUT, 'misc_losses.pt'))


def gen_pyramid(ns):
    """The reference's own `get_imgs_blur` (datasets.py:151-196), unmodified, on data/face/000155.png with the eval
    transform (Resize(288) -> CenterCrop(256), datasets.py:1558-1561) and norm (:1563); imsize = [32, 64, 128, 256]
    (BASE_SIZE 32, BRANCH_NUM 4).  datasets.py imports nltk at module scope for its caption tokenizer, which the image
    lacks; an empty stand-in module satisfies the import and is never called on this path.  The branch-0 tensors are
    stored as fp32 (pins ToTensor/Normalize); the larger ones as the uint8 they were computed from."""
    import types
    from torchvision import transforms
    if 'nltk' not in sys.modules:
        nltk = types.ModuleType('nltk'); tok = types.ModuleType('nltk.tokenize')
        tok.RegexpTokenizer = object; nltk.tokenize = tok
        sys.modules['nltk'] = nltk; sys.modules['nltk.tokenize'] = tok
    import datasets as D
    path = os.path.join(ref_loader.REF_ROOT, 'data/face/000155.png')
    transform = transforms.Compose([transforms.Resize(int(256 * 72 / 64)), transforms.CenterCrop(256)])
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
    imsize = [32, 64, 128, 256]
    ret, bic, retb, bicb = D.get_imgs_blur(path, imsize, None, transform, normalize=norm, cfg=ns.cfg)
    from PIL import Image
    hr = np.asarray(transform(Image.open(path).convert('RGB')))

    def u8(t):
        return np.round((t.numpy().transpose(1, 2, 0) * 0.5 + 0.5) * 255.0).astype(np.uint8)
    out = dict(hr=hr, imsize=np.asarray(imsize))
    for name, lst in (('ret', ret), ('bic', bic), ('retb', retb), ('bicb', bicb)):
        out[name + '0_f32'] = lst[0].numpy()
        for i in range(1, 4):
            out['%s%d_u8' % (name, i)] = u8(lst[i])
    np.savez_compressed(os.path.join(OUT, 'pyramid.npz'), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == 'misc':
        gen_misc_losses(ns)
        print('misc_losses.pt', os.path.getsize(os.path.join(OUT, 'misc_losses.pt')))
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'pyramid':
        gen_pyramid(ns)
        print('pyramid.npz', os.path.getsize(os.path.join(OUT, 'pyramid.npz')))
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'attn_stats':
        gen_attn_stats(ns)
        print('attn_stats.pt', os.path.getsize(os.path.join(OUT, 'attn_stats.pt')))
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'x16':
        gen_x16_gf32(ns)
        print('x16_gf32.pt', os.path.getsize(os.path.join(OUT, 'x16_gf32.pt')))
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'train':
        gen_train(ns)
        print('train_step.pt', os.path.getsize(os.path.join(OUT, 'train_step.pt')))
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'losses':
        gen_losses(ns)
        print('losses.pt', os.path.getsize(os.path.join(OUT, 'losses.pt')))
        return
    gen_attention(ns)
    gen_func_attention(ns)
    gen_blocks(ns)
    gen_rnn(ns)
    gen_face_x8(ns)
    gen_x16(ns)
    gen_x16_gf32(ns)
    gen_attn_stats(ns)
    gen_losses(ns)
    gen_train(ns)
    gen_pyramid(ns)
    gen_misc_losses(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
