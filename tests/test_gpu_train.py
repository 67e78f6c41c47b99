"""Training-mode parity (-m gpu): the explicit forward/backward kernels of the generator blocks (csrc/train_ops.cu,
csrc/attention_bwd.cu) against torch.autograd over the CPU oracle (the reference arithmetic with BatchNorm in train()
mode).  fp32; tolerances cover summation-order differences only."""
import pytest
import torch
import torch.nn.functional as F

import tgsr_b200
from oracle import tgsr_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _fp32():
    from tgsr_b200 import train
    tgsr_b200.set_eval_defaults()
    tgsr_b200.set_precision('fp32')
    train.set_conv_precision('fp32')             # (a GeneratorTrainStep built by an earlier test may have left 'tf32')
    yield
    train.set_conv_precision('fp32')


def close(a, b, rtol=2e-4, atol=2e-5):
    torch.testing.assert_close(a.detach().cpu(), b.detach(), rtol=rtol, atol=atol)


def _randomize(mod, seed):
    g = torch.Generator().manual_seed(seed)
    for m in mod.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.copy_(1 + 0.2 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def _oracle_sd(mod):
    return {k: (v.detach().cpu().clone().requires_grad_(True) if v.is_floating_point() else v.cpu().clone())
            for k, v in mod.state_dict().items()}


def _check_param_grads(mod, sd, rtol=1e-3, atol=2e-5, skip=()):
    n = 0
    for name, p in mod.named_parameters():
        if any(s in name for s in skip):
            continue
        ref = sd[name].grad
        if ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        scale = max(1.0, float(ref.abs().max()))
        torch.testing.assert_close(p.grad.cpu(), ref, rtol=rtol, atol=atol * scale, msg=lambda m, name=name: name + ': ' + m)
        n += 1
    assert n > 0


@pytest.mark.parametrize('C,H,W,B', [(32, 16, 16, 3), (64, 8, 24, 2)])
def test_resblock_train(C, H, W, B):
    from tgsr_b200 import blocks
    torch.manual_seed(C + H)
    m = blocks.ResBlock(C).to(DEV).train(); _randomize(m, 1)
    sd = _oracle_sd(m)
    rm0 = m.block[1].running_mean.clone()
    x = torch.randn(B, C, H, W) * 0.7
    xg = x.to(DEV).requires_grad_(True); xr = x.clone().requires_grad_(True)
    y = m(xg)
    with O.train_mode():
        yr = O.res_block(sd, '', xr) if False else None
    # the oracle's key prefix convention is '<p>.block.0'; ResBlock.state_dict keys are 'block.0.weight'
    sdp = {'r.' + k: v for k, v in sd.items()}
    with O.train_mode():
        yr = O.res_block(sdp, 'r', xr)
    close(y, yr)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(2))
    y.backward(gy.to(DEV)); yr.backward(gy)
    close(xg.grad, xr.grad, rtol=1e-3, atol=2e-5)
    _check_param_grads(m, sd)
    # running statistics follow nn.BatchNorm2d.train(): momentum 0.1, unbiased variance
    z = F.conv2d(x, sd['block.0.weight'].detach(), None, 1, 1)
    exp_rm = 0.9 * rm0.cpu() + 0.1 * z.mean((0, 2, 3))
    close(m.block[1].running_mean, exp_rm, rtol=1e-4, atol=1e-6)
    exp_rv = 0.9 * torch.ones(2 * C) + 0.1 * z.var((0, 2, 3), unbiased=True)
    close(m.block[1].running_var, exp_rv, rtol=1e-4, atol=1e-6)
    assert int(m.block[1].num_batches_tracked) == 1


def test_upblock_and_heads_train():
    from tgsr_b200 import blocks
    torch.manual_seed(3)
    up = blocks.upBlock(64, 32).to(DEV).train(); _randomize(up, 4)
    sd = {'u.' + k: v for k, v in _oracle_sd(up).items()}
    x = torch.randn(2, 64, 8, 12) * 0.5
    xg = x.to(DEV).requires_grad_(True); xr = x.clone().requires_grad_(True)
    y = up(xg)
    with O.train_mode():
        yr = O.up_block(sd, 'u', xr)
    close(y, yr)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
    y.backward(gy.to(DEV)); yr.backward(gy)
    close(xg.grad, xr.grad, rtol=1e-3, atol=2e-5)
    _check_param_grads(up, {k[2:]: v for k, v in sd.items()})
    # conv5x5 + tanh head with the fused one*y + a*SRb (model.py:223-226,280)
    head = blocks._ImageHead(blocks.conv5x5(32, 3), torch.nn.Tanh()).to(DEV).train()
    w = head[0].weight.detach().cpu().clone().requires_grad_(True)
    h = torch.randn(2, 32, 16, 20) * 0.5; srb = torch.randn(2, 3, 16, 20) * 0.3
    hg = h.to(DEV).requires_grad_(True); sg = srb.to(DEV).requires_grad_(True)
    hr = h.clone().requires_grad_(True); sr = srb.clone().requires_grad_(True)
    out = head(hg, sg, 0.5, 1.0)
    ref = 1.0 * torch.tanh(F.conv2d(hr, w, None, 1, 2)) + 0.5 * sr
    close(out, ref)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(6))
    out.backward(go.to(DEV)); ref.backward(go)
    close(hg.grad, hr.grad, rtol=1e-3, atol=2e-5); close(sg.grad, sr.grad); close(head[0].weight.grad, w.grad, rtol=1e-3, atol=5e-5)
    # conv3x3 head without activation (GET_IMAGE_G_noAct, util.py:909-919)
    g3 = blocks.GET_IMAGE_G_noAct(32).to(DEV).train()
    w3 = g3.img[0].weight.detach().cpu().clone().requires_grad_(True)
    hg2 = h.to(DEV).requires_grad_(True); hr2 = h.clone().requires_grad_(True)
    o3 = g3(hg2); r3 = F.conv2d(hr2, w3, None, 1, 1)
    close(o3, r3)
    o3.backward(go.to(DEV)); r3.backward(go)
    close(hg2.grad, hr2.grad, rtol=1e-3, atol=2e-5); close(g3.img[0].weight.grad, w3.grad, rtol=1e-3, atol=5e-5)


@pytest.mark.parametrize('B,idf,cdf,ih,iw,T,ragged', [(3, 32, 256, 16, 16, 18, True), (2, 32, 256, 9, 7, 18, False),
                                                       (2, 64, 64, 16, 8, 30, True)])
def test_attention_train(B, idf, cdf, ih, iw, T, ragged):
    g = torch.Generator().manual_seed(B + T)
    inp = torch.randn(B, idf, ih, iw, generator=g) * 0.5
    ctx = torch.randn(B, cdf, T, generator=g) * 0.5
    w = torch.randn(idf, cdf, 1, 1, generator=g) * (1.6 / cdf ** 0.5)
    mask = None
    if ragged:
        lens = torch.randint(2, T + 1, (B,), generator=g); lens[0] = T
        mask = torch.arange(T)[None] >= lens[:, None]
    att = tgsr_b200.GlobalAttentionGeneral(idf, cdf).to(DEV).train()
    att.conv_context.weight.data.copy_(w)
    att.applyMask(None if mask is None else mask.to(DEV))
    xg = inp.to(DEV).requires_grad_(True)
    wctx, attn = att(xg, ctx.to(DEV))
    xr = inp.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    ref_w, ref_a = O.global_attention(xr, ctx, wr, mask)
    close(wctx, ref_w, rtol=1e-4, atol=5e-5); close(attn, ref_a, rtol=1e-5, atol=1e-7)
    g1 = torch.randn(ref_w.shape, generator=g); g2 = torch.randn(ref_a.shape, generator=g)
    (wctx * g1.to(DEV)).sum().backward(retain_graph=True)
    (ref_w * g1).sum().backward(retain_graph=True)
    close(xg.grad, xr.grad, rtol=1e-3, atol=2e-5)
    close(att.conv_context.weight.grad, wr.grad, rtol=1e-3, atol=1e-4)
    # gradient through the returned attention maps as well
    xg.grad = None; att.conv_context.weight.grad = None; xr.grad = None; wr.grad = None
    ((wctx * g1.to(DEV)).sum() + (attn * g2.to(DEV)).sum()).backward()
    ((ref_w * g1).sum() + (ref_a * g2).sum()).backward()
    close(xg.grad, xr.grad, rtol=1e-3, atol=2e-5)
    close(att.conv_context.weight.grad, wr.grad, rtol=1e-3, atol=1e-4)


def test_generator_train_step_grads():
    """Whole G_SR_NET_low + NetG_highweight in train() mode on a 16x16 LR batch: outputs, the MSE loss against a target
    pyramid and every parameter gradient against autograd over the oracle."""
    from tgsr_b200 import nets
    torch.manual_seed(11)
    gl = nets.G_SR_NET_low().to(DEV).train(); gh = nets.NetG_highweight(low='lr').to(DEV).train()
    _randomize(gl, 12); _randomize(gh, 13)
    sdl, sdh = _oracle_sd(gl), _oracle_sd(gh)
    B, T = 2, 18
    LR = O.synth_lr(B, 16, seed=14)
    g = torch.Generator().manual_seed(15)
    words = torch.randn(B, 256, T, generator=g) * 0.5; sent = torch.randn(B, 256, generator=g) * 0.5
    mask = torch.zeros(B, T, dtype=torch.bool); mask[1, 12:] = True
    targets = [torch.randn(B, 3, 16 * 2 ** (k + 1), 16 * 2 ** (k + 1), generator=g) * 0.5 for k in range(3)]

    imgs, atts, mu, lv = gl(LR.to(DEV), sent.to(DEV), words.to(DEV), mask.to(DEV))
    fine, a, one = gh(LR.to(DEV), imgs, LR.to(DEV))
    loss = sum(F.mse_loss(f, t.to(DEV)) for f, t in zip(fine, targets)) + sum(F.mse_loss(f, t.to(DEV)) for f, t in zip(imgs, targets))
    loss.backward()

    with O.train_mode():
        r_imgs, r_atts, r_mu, r_lv, _ = O.g_sr_net_low(sdl, LR, sent, words, mask)
        r_fine = O.netg_high(sdh, LR, r_imgs, LR, low='lr')
    r_loss = sum(F.mse_loss(f, t) for f, t in zip(r_fine, targets)) + sum(F.mse_loss(f, t) for f, t in zip(r_imgs, targets))
    r_loss.backward()
    for k in range(3):
        close(imgs[k], r_imgs[k], rtol=1e-3, atol=1e-4); close(fine[k], r_fine[k], rtol=1e-3, atol=1e-4)
    close(loss, r_loss, rtol=1e-4, atol=1e-6)
    _check_param_grads(gl, sdl, rtol=5e-3, atol=2e-4, skip=('ca_net',))
    _check_param_grads(gh, sdh, rtol=5e-3, atol=2e-4)


def test_fused_adam_matches_torch():
    from tgsr_b200.train_step import FlatBucket, FusedAdam
    torch.manual_seed(21)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.Linear(7, 3)).to(DEV)
    ref = [p.detach().cpu().clone().requires_grad_(True) for p in net.parameters()]
    opt_ref = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999))
    bucket = FlatBucket([net]); opt = FusedAdam(bucket, 2e-4, (0.5, 0.999))
    g = torch.Generator().manual_seed(22)
    for step in range(4):
        bucket.zero_grad()
        for p, r in zip(net.parameters(), ref):
            gr = torch.randn(r.shape, generator=g)
            p.grad.add_(gr.to(DEV) * 2.0)           # as if two ranks had summed their gradients
            r.grad = gr.clone()
        opt.step(grad_scale=0.5); opt_ref.step()
        for p, r in zip(net.parameters(), ref):
            torch.testing.assert_close(p.detach().cpu(), r.detach(), rtol=1e-5, atol=1e-7)
    assert all(p.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())


def test_generator_train_step_runs_and_learns():
    """configs[4] at reduced batch: G forward (train BN) + DAMSM + MSE + backward + Adam; the loss goes down on a fixed batch."""
    from tgsr_b200 import nets
    from tgsr_b200.text_encoder import RNN_ENCODER
    from tgsr_b200.train_step import FrozenImageEncoder, GeneratorTrainStep
    torch.manual_seed(31)
    gl = nets.G_SR_NET_low().to(DEV); gh = nets.NetG_highweight(low='lr').to(DEV)
    enc = RNN_ENCODER(200, nhidden=256).to(DEV).eval()
    img = FrozenImageEncoder().to(DEV)
    ts = GeneratorTrainStep(gl, gh, enc, img, lr=2e-3)
    B = 4
    LR = O.synth_lr(B, 32, seed=32).to(DEV)
    real = [torch.nn.functional.interpolate(LR, scale_factor=2 ** (k + 1), mode='bicubic').clamp(-1, 1) for k in range(3)]
    caps, lens = O.synth_captions(B, 200, seed=33, words_num=18)
    caps, lens = caps.to(DEV), lens.to(DEV)
    outs = [ts.step(LR, real, caps, lens, 18) for _ in range(8)]
    hist = [float(o['total']) for o in outs]; mse = [float(o['mse']) for o in outs]
    assert all(torch.isfinite(torch.tensor(hist))), hist
    assert min(mse[-3:]) < mse[0], mse                # the reconstruction term falls on a fixed batch
    # the ranking term is noisy at batch 4 (and the BatchNorm reductions use atomics, so runs differ in the last bits):
    # require that training gets below the starting loss at some point, not at a particular step
    assert min(hist[1:]) < hist[0], hist
    assert ts.bucket.grad.abs().sum() > 0
    # the captured whole-step graph continues the same optimisation
    ts.capture(LR, real, caps, lens, 18, warmup=1)
    o2 = []
    for _ in range(4):                                  # the graph's outputs are static tensors: read each replay's value
        o = ts.step_graph(LR, real, caps, lens)
        o2.append((float(o['mse']), float(o['total'])))
    print('eager mse', mse, 'graph replays (mse, total)', o2)
    assert len({m for m, _ in o2}) > 1, 'replays do not advance the optimisation: %s' % (o2,)
    assert min(m for m, _ in o2) < mse[0] and all(t == t for _, t in o2), (o2, mse)


def test_train_step_from_images_matches_explicit_inputs():
    """step_from_images(uint8 batch) == step(LR, real) when LR / real come from the PIL-exact oracle pyramid
    (datasets.py:112-148): the device pyramid feeds the step bit-identical inputs, so the first-step losses agree."""
    from oracle import pil_pyramid as P
    from tgsr_b200 import nets
    from tgsr_b200.text_encoder import RNN_ENCODER
    from tgsr_b200.train_step import FrozenImageEncoder, GeneratorTrainStep
    B = 3
    hr = torch.nn.functional.interpolate(O.synth_lr(B, 32, seed=41), size=(256, 256), mode='bicubic').clamp(-1, 1)
    hr_u8 = ((hr + 1.0) * 127.5).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    ret, _, _, _ = P.get_imgs_blur(hr_u8.numpy(), [32, 64, 128, 256])
    caps, lens = O.synth_captions(B, 200, seed=42, words_num=18)
    caps, lens = caps.to(DEV), lens.to(DEV)
    losses = []
    for mode in ('explicit', 'images', 'graph'):
        torch.manual_seed(31)
        gl = nets.G_SR_NET_low().to(DEV); gh = nets.NetG_highweight(low='lr').to(DEV)
        enc = RNN_ENCODER(200, nhidden=256).to(DEV).eval()
        img = FrozenImageEncoder().to(DEV)
        ts = GeneratorTrainStep(gl, gh, enc, img, conv_precision='fp32')
        if mode == 'explicit':
            o = ts.step(torch.from_numpy(ret[0]).to(DEV), [torch.from_numpy(r).to(DEV) for r in ret[1:]], caps, lens, 18)
        elif mode == 'images':
            o = ts.step_from_images(hr_u8.to(DEV), caps, lens, 18)
        else:
            ts.capture_from_images(hr_u8.to(DEV), caps, lens, 18, warmup=1)
            assert ts.graph_img_kernels > 0
            continue
        losses.append({k: float(v) for k, v in o.items()})
    assert losses[0] == pytest.approx(losses[1], rel=1e-5), losses


@pytest.mark.parametrize('B,Cin,Cout,H,W,k,up', [(2, 64, 128, 32, 32, 3, False), (2, 64, 64, 16, 24, 3, True), (3, 32, 32, 20, 12, 3, False),
                                                 (2, 32, 64, 8, 8, 3, True), (2, 3, 64, 32, 32, 3, False), (2, 32, 3, 16, 16, 5, False)])
def test_tf32_convs_match_exact_kernels(B, Cin, Cout, H, W, k, up):
    """TF32 tensor-core forward / data gradient / weight gradient against the exact fp32 kernels (themselves checked against
    autograd above): operands rounded to 10-bit mantissas, fp32 accumulate -> 2e-3 of the output scale (PyTorch's own TF32
    convs differ from fp32 by the same amount)."""
    from tgsr_b200 import train
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = (torch.randn(B, Cin, H, W, generator=g) * 0.5).to(DEV)
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (1.0 / (Cin * k * k) ** 0.5)).to(DEV)
    Ho, Wo = (2 * H, 2 * W) if up else (H, W)
    dy = (torch.randn(B, Cout, Ho, Wo, generator=g) * 0.5).to(DEV)
    res = {}
    for mode in ('fp32', 'tf32'):
        train.set_conv_precision(mode)
        try:
            res[mode] = (train._conv_raw(x, w, up), train._conv_dgrad(dy, w, up), train._conv_wgrad(x, dy, w.shape, up))
        finally:
            train.set_conv_precision('fp32')
    for a, b in zip(res['fp32'], res['tf32']):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 2e-3 * max(1.0, float(a.abs().max())), (float((a - b).abs().max()), float(a.abs().max()))


def test_generator_train_grads_vs_reference_modules(golden):
    """The explicit backward kernels against gradients of the UNMODIFIED reference modules in train() mode
    (tests/golden/train_step.pt): loss, outputs, BatchNorm running statistics and a spread of parameter gradients."""
    from tgsr_b200 import nets
    c = golden('train_step.pt')
    f32 = lambda sd: {k: (v.float() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    gl = nets.G_SR_NET_low(); gh = nets.NetG_highweight(low='lr')
    gl.load_state_dict(f32(c['sdL']), strict=True); gh.load_state_dict(f32(c['sdH']), strict=True)
    # the stored state is the one AFTER the reference's step (running statistics updated once): reset them
    for m in list(gl.modules()) + list(gh.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
    gl.to(DEV).train(); gh.to(DEV).train()
    targets = [t.float().to(DEV) for t in c['targets']]
    imgs, atts, mu, lv = gl(c['LR'].to(DEV), c['sent'].to(DEV), c['words'].to(DEV), c['mask'].to(DEV))
    fine, a, one = gh(c['LR'].to(DEV), imgs, c['LR'].to(DEV))
    loss = sum(F.mse_loss(f, t) for f, t in zip(fine, targets)) + sum(F.mse_loss(f, t) for f, t in zip(imgs, targets))
    loss.backward()
    close(loss, c['loss'], rtol=1e-4, atol=1e-6)
    close(imgs[0], c['img0'], rtol=1e-3, atol=1e-4); close(fine[0], c['fine0'], rtol=1e-3, atol=1e-4)
    close(fine[2][:, :, ::4, ::4], c['fine2_sub4'], rtol=1e-3, atol=1e-4)
    close(gl.h_net1.im2f[1].running_mean, c['run_mean'], rtol=1e-4, atol=1e-6)
    close(gl.h_net1.im2f[1].running_var, c['run_var'], rtol=1e-4, atol=1e-6)
    for net, grads in ((gl, c['gradL']), (gh, c['gradH'])):
        params = dict(net.named_parameters())
        for name, g in grads.items():
            torch.testing.assert_close(params[name].grad.cpu(), g, rtol=5e-3, atol=1e-6 + 2e-4 * float(g.abs().max()),
                                       msg=lambda m, name=name: name + ': ' + m)


def test_eval_after_train_uses_fresh_weights():
    """Derived-parameter caches (folded BN, packed 16-bit operands, captured eval graphs) key on the weights epoch: the fused
    Adam and the BatchNorm running-stat update write through raw pointers, which bumps neither data_ptr nor _version.
    eval -> train steps -> eval must equal a fresh model loaded with the trained state dict."""
    from tgsr_b200 import nets
    from tgsr_b200.pipeline import SRPipeline
    from tgsr_b200.text_encoder import RNN_ENCODER
    from tgsr_b200.train_step import FrozenImageEncoder, GeneratorTrainStep
    torch.manual_seed(5)
    B = 4
    pipe = SRPipeline(200, device=DEV, seed=11, randomize_bn=True)
    LR = O.synth_lr(B, 32, seed=31).to(DEV)
    caps, lens = O.synth_captions(B, 200, seed=32, words_num=18)
    caps, lens = caps.to(DEV), lens.to(DEV)
    tgsr_b200.set_precision('mixed16')
    before = pipe.forward_graph(LR, caps, lens, 18)[0][-1].clone()           # warms every cache and captures a graph
    # two training steps on the same modules (fp32 training kernels)
    tgsr_b200.set_precision('fp32')
    img = FrozenImageEncoder().to(DEV)
    ts = GeneratorTrainStep(pipe.netGL, pipe.netGH, pipe.text_encoder, img, lr=2e-3, conv_precision='fp32')
    real = [torch.nn.functional.interpolate(LR, size=(s, s), mode='bicubic', align_corners=False).clamp(-1, 1) for s in (64, 128, 256)]
    for _ in range(2):
        ts.step(LR, real, caps, lens, 18)
    assert ts.opt.t == 2 and int(ts.opt.step_t.item()) == 2
    pipe.netGL.eval(); pipe.netGH.eval()
    tgsr_b200.set_precision('mixed16')
    after = pipe.forward_graph(LR, caps, lens, 18)[0][-1].clone()
    assert (after - before).abs().max() > 1e-3, 'training changed nothing?'
    fresh = SRPipeline(200, device=DEV, seed=11, randomize_bn=False)
    fresh.text_encoder.load_state_dict(pipe.text_encoder.state_dict())
    fresh.load_generators(pipe.netGL.state_dict(), pipe.netGH.state_dict())
    want = fresh.forward_device(LR, caps, lens)[0][-1]
    assert torch.equal(after, want), 'eval after training replayed stale folded / packed weights: max-abs %.3e' % \
        (after - want).abs().max().item()
    tgsr_b200.set_precision('fp32')


def test_adam_step_counter_eager_graph_mix():
    """The device step counter is the single source of truth: eager steps, a captured step and its replays, and eager steps
    again produce the same parameters as torch.optim.Adam over the same gradient sequence."""
    from tgsr_b200.train_step import FlatBucket, FusedAdam
    torch.manual_seed(3)
    net = torch.nn.Linear(16, 8).to(DEV)
    ref = [p.detach().cpu().clone().requires_grad_(True) for p in net.parameters()]
    opt_ref = torch.optim.Adam(ref, lr=1e-3, betas=(0.5, 0.999))
    bucket = FlatBucket([net]); opt = FusedAdam(bucket, 1e-3, (0.5, 0.999))
    gfix = [torch.randn(r.shape, generator=torch.Generator().manual_seed(7 + i)) for i, r in enumerate(ref)]

    def set_grads():
        for p, g in zip(net.parameters(), gfix):
            p.grad.copy_(g.to(DEV))
    set_grads(); opt.step()                                  # eager
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        set_grads(); opt.step()                              # warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        opt.step()
    for _ in range(3):
        g.replay(); opt.note_replay()
    set_grads(); opt.step()                                  # eager again
    torch.cuda.synchronize()
    assert opt.t == 6 and int(opt.step_t.item()) == 6
    for _ in range(6):
        for r, gr in zip(ref, gfix):
            r.grad = gr.clone()
        opt_ref.step()
    for p, r in zip(net.parameters(), ref):
        torch.testing.assert_close(p.detach().cpu(), r.detach(), rtol=2e-5, atol=1e-7)


def test_text_side_gradients_raise():
    """The drop-in losses / attention only carry the generator-side gradients; asking for a text-side gradient must raise
    instead of silently training nothing."""
    from tgsr_b200 import losses
    B, T = 4, 6
    feats = torch.randn(B, 256, 17, 17, device=DEV, requires_grad=True)
    words = torch.randn(B, 256, T, device=DEV, requires_grad=True)
    labels = torch.arange(B, device=DEV); lens = torch.full((B,), T, device=DEV)
    with pytest.raises(NotImplementedError, match='words_emb'):
        losses.words_loss(feats, words, labels, lens, None, B)
    losses.words_loss(feats, words.detach(), labels, lens, None, B)          # the supported call
    code = torch.randn(B, 256, device=DEV, requires_grad=True)
    with pytest.raises(NotImplementedError, match='rnn_code'):
        losses.sent_loss(code, torch.randn(B, 256, device=DEV, requires_grad=True), labels, None, B)
    att = tgsr_b200.GlobalAttentionGeneral(32, 256).to(DEV).train()
    with pytest.raises(NotImplementedError, match='context'):
        att(torch.randn(1, 32, 8, 8, device=DEV, requires_grad=True), torch.randn(1, 256, T, device=DEV, requires_grad=True))
