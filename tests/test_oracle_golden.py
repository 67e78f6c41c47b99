"""Pins the CPU oracle (oracle/tgsr_oracle.py) to outputs of the reference's own modules.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which imports the
unmodified reference from /root/reference.  CPU only; runs in seconds.
"""
import torch

from oracle import tgsr_oracle as O


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_attention_cases(golden):
    for c in golden('attention.pt'):
        wctx, attn = O.global_attention(c['inp'], c['ctx'], c['w'], c['mask'])
        close(attn, c['attn'], rtol=1e-5, atol=1e-7)
        close(wctx, c['wctx'], rtol=1e-4, atol=5e-5)
        # caption masking and the per-pixel argmax word index are bit-exact
        assert torch.equal(attn == 0, c['attn'] == 0)
        assert torch.equal(attn.argmax(1), c['attn'].argmax(1))


def test_attention_mask_quirk(golden):
    """GlobalAttention.py:109-112: pixel q of image b uses caption (b*Q+q) % B."""
    c = golden('attention.pt')[1]
    B, T, ih, iw = c['attn'].shape
    Q = ih * iw
    zero = (c['attn'] == 0).reshape(B, T, Q)
    for b in range(B):
        for q in (0, 1, 5, Q - 1):
            assert torch.equal(zero[b, :, q], c['mask'][(b * Q + q) % B])


def test_func_attention(golden):
    for c in golden('func_attention.pt'):
        w, a = O.func_attention(c['query'], c['context'], c['gamma1'])
        close(w, c['wctx']); close(a, c['attn'], atol=1e-7)


def test_blocks(golden):
    g = golden('blocks.pt')
    for name in ('block3x3_3_32', 'block3x3_32_32'):
        sd = {'b.' + k: v for k, v in g[name]['sd'].items()}
        close(O.block3x3(sd, 'b', g[name]['x']), g[name]['y'])
    for name in ('resblock_64', 'resblock_32'):
        sd = {'r.' + k: v for k, v in g[name]['sd'].items()}
        close(O.res_block(sd, 'r', g[name]['x']), g[name]['y'], atol=1e-5)
    for name in ('upblock_64_32', 'upblock_32_32'):
        sd = {'u.' + k: v for k, v in g[name]['sd'].items()}
        close(O.up_block(sd, 'u', g[name]['x']), g[name]['y'], atol=1e-5)
    sd = {'h.' + k: v for k, v in g['get_image']['sd'].items()}
    close(O.get_image(sd, 'h', g['get_image']['x'], True), g['get_image']['y'])
    sd = {'h.' + k: v for k, v in g['get_image_noact']['sd'].items()}
    close(O.get_image(sd, 'h', g['get_image_noact']['x'], False), g['get_image_noact']['y'])
    sd = {'ca.' + k: v for k, v in g['ca_net']['sd'].items()}
    _, mu, lv = O.ca_net(sd, 'ca', g['ca_net']['x'], 100)
    close(mu, g['ca_net']['mu']); close(lv, g['ca_net']['logvar'])
    sd = {'i.' + k: v for k, v in g['init_stage_g']['sd'].items()}
    close(O.init_stage_g(sd, 'i', g['init_stage_g']['z'], g['init_stage_g']['c'], 32), g['init_stage_g']['y'],
          atol=1e-5)


def test_rnn_encoder(golden):
    g = golden('rnn_encoder.pt')
    words, sent = O.rnn_encoder(g['sd'], g['captions'], g['cap_lens'])
    assert words.shape == g['words'].shape          # width = max(cap_lens), util.py:250
    close(words, g['words'], atol=2e-6); close(sent, g['sent'], atol=2e-6)
    # padded positions are exact zeros
    for b, L in enumerate(g['cap_lens'].tolist()):
        assert torch.count_nonzero(words[b, :, L:]) == 0


def test_face_x8_checkpoint(golden):
    w = golden('face_S8_weights.pt'); g = golden('face_x8.pt'); rnn = golden('rnn_encoder.pt')
    c = g['b1']
    words, sent = O.rnn_encoder(rnn['sd'], c['captions'], c['cap_lens'])
    close(words, c['words'], atol=2e-6)
    imgs, atts, mu, lv, _ = O.g_sr_net_low(w['netG'], c['LR'], c['sent'], c['words'], c['mask'])
    for a, b in zip(imgs, c['imgsL']):
        close(a, b, rtol=1e-4, atol=2e-5)
    close(atts[0], c['atts'][0], atol=1e-6); close(atts[1], c['atts'][1], rtol=1e-4, atol=1e-6)
    close(atts[2][:, :, ::4, ::4], c['att2_sub4'], rtol=1e-4, atol=1e-6)
    close(mu, c['mu']); close(lv, c['logvar'])
    fine = O.netg_high(w['netGH'], c['LR'], c['imgsL'], c['LR'], low='lr', a=0.5, one=1.0)
    for a, b in zip(fine, c['fine']):
        close(a, b, rtol=1e-4, atol=2e-5)
    # ragged batch: exercises the mask quirk through the whole net
    c = g['b3']
    imgs, atts, mu, lv, _ = O.g_sr_net_low(w['netG'], c['LR'], c['sent'], c['words'], c['mask'])
    close(imgs[0], c['imgL0'], rtol=1e-4, atol=2e-5); close(atts[0], c['att0'], atol=1e-6)
    assert torch.equal(atts[0] == 0, c['att0'] == 0)
    close(atts[1][:, :, ::2, ::2], c['att1_sub2'], rtol=1e-4, atol=1e-6)
    fine = O.netg_high(w['netGH'], c['LR'], imgs, c['LR'], low='lr')
    close(fine[0], c['fine0'], rtol=1e-4, atol=2e-5)
    close(fine[2][:, :, ::4, ::4], c['fine2_sub4'], rtol=1e-4, atol=5e-5)


def test_x16_wiring(golden):
    g = golden('x16.pt')
    imgs, atts, mu, lv, _ = O.g_sr_net_low(g['sdL'], g['LR'], g['sent'], g['words'], g['mask'], variant='x16')
    assert len(imgs) == 4 and len(atts) == 4
    for a, b in zip(imgs[:3], g['imgsL']):
        close(a, b, rtol=1e-4, atol=2e-5)
    close(imgs[3][:, :, ::2, ::2], g['imgL3_sub2'], rtol=1e-4, atol=2e-5)
    for a, b in zip(atts[:3], g['atts']):
        close(a, b, rtol=1e-4, atol=1e-6)
    close(atts[3][:, :, ::4, ::4], g['att3_sub4'], rtol=1e-4, atol=1e-6)
    fine = O.netg_high(g['sdH'], g['LRh'], g['SRb'], g['LRh'], low='lrblur', a=1.0, one=1.0, variant='x16',
                       x16_last=('residual816', 'upscale16x'))
    for a, b in zip(fine, g['fine']):
        close(a, b, rtol=1e-4, atol=2e-5)


def _x16_srb_w(g):
    """NetG_highweight x16 inputs of the fixture: SRb[2] is a per-image constant colour (see make_golden.gen_x16_gf32); the
    implementations take a full pyramid, so the constant is expanded to the 128^2 and 256^2 addends it was broadcast to."""
    c8 = g['c8']
    B = c8.shape[0]
    return [g['imgsL_full'][0], g['imgsL_full'][1], c8.expand(B, 3, 128, 128).contiguous(), c8.expand(B, 3, 256, 256).contiguous()]


def test_x16_production_width(golden):
    """models16 at GF_DIM = 32 (16 -> 256): the oracle against the unmodified reference's outputs for all three nets,
    including NetG_highweight's reuse of residual48 / upscale8x in the last stage (models16.py:172-173)."""
    g = dict(golden('x16_gf32.pt'))
    imgs, atts, mu, lv, _ = O.g_sr_net_low(g['sdL'], g['LR'], g['sent'], g['words'], g['mask'], variant='x16')
    g['imgsL_full'] = imgs
    close(mu, g['mu'], rtol=1e-4, atol=1e-6)
    for k, sub in enumerate((1, 1, 2, 4)):
        close(imgs[k][:, :, ::sub, ::sub], g['imgsL'][k], rtol=1e-4, atol=2e-5)
    for k, sub in enumerate((1, 2, 4, 8)):
        close(atts[k][:, :, ::sub, ::sub], g['atts'][k], rtol=1e-4, atol=1e-6)
    fine = O.netg_high(g['sdH'], g['LR'], imgs, g['LRb'], low='lr-lrblur', a=1.0, one=1.0, variant='x16',
                       x16_last=('residual816', 'upscale16x'))
    finew = O.netg_high(g['sdHw'], g['LR'], _x16_srb_w(g), g['LRb'], low='lr', a=0.5, one=1.0, variant='x16')
    for k, sub in enumerate((1, 1, 2, 4)):
        close(fine[k][:, :, ::sub, ::sub], g['fine'][k], rtol=1e-4, atol=5e-5)
        close(finew[k][:, :, ::sub, ::sub], g['finew'][k], rtol=1e-4, atol=5e-5)


def test_attention_word_stats_pinned(golden):
    """attention_word_stats against values recorded while the UNMODIFIED miscc/utils.py ran (build_super_images :122-131 for
    the max over words, build_super_imagesall :378-392 for the confidence sums)."""
    g = golden('attn_stats.pt')
    attn = g['attn']
    mx, am, _ = O.attention_word_stats(attn, 0.0)
    assert torch.equal(mx[:, 0], g['max_maps'])
    assert torch.equal(am[:, 0], attn.argmax(1))
    for b, n in enumerate(g['lens']):
        _, _, conf = O.attention_word_stats(attn[b:b + 1, :n], 2.0 / n)          # thresh = 2 / num_attn, mask0 = map > 2 * thresh
        close(conf[0, 1:], g['conf'][b], rtol=1e-12, atol=0)


def test_psnr_and_quantisation():
    x = torch.tensor([[-1.0, 0.0, 1.0, 1.2]])
    assert O.to_uint8(x).tolist() == [[0, 128, 255, 255]]       # round-half-even of 127.5 -> 128
    a = torch.zeros(4, 4, dtype=torch.uint8); b = a.clone(); b[0, 0] = 16
    import math
    assert abs(O.psnr_u8(a, b) - 20 * math.log10(255.0 / 4.0)) < 1e-9


def test_losses_oracle_matches_reference(golden):
    """The oracle's words_loss / sent_loss restatement against the UNMODIFIED miscc/losses.py outputs and autograd
    gradients (tests/golden/losses.pt, oracle/make_golden.py gen_losses)."""
    for c in golden('losses.pt'):
        g1, g2, g3 = c['gamma']
        feats = c['feats'].float().requires_grad_(True); cnn = c['cnn'].clone().requires_grad_(True)
        w0, w1, att = O.words_loss(feats, c['words'], c['labels'], c['lens'], c['class_ids'], g1, g2, g3)
        s0, s1 = O.sent_loss(cnn, c['sent'], c['labels'], c['class_ids'], g3)
        for a, b in ((w0, c['w0']), (w1, c['w1']), (s0, c['s0']), (s1, c['s1'])):
            torch.testing.assert_close(a.detach(), b, rtol=1e-5, atol=1e-6)
        for a, b in zip(att, c['att']):
            torch.testing.assert_close(a.detach(), b, rtol=1e-5, atol=1e-8)
        (c['gw'][0] * w0 + c['gw'][1] * w1 + c['gs'][0] * s0 + c['gs'][1] * s1).backward()
        torch.testing.assert_close(feats.grad, c['d_feats'], rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(cnn.grad, c['d_cnn'], rtol=1e-4, atol=1e-7)


def _train_case(golden):
    c = golden('train_step.pt')
    f32 = lambda sd: {k: (v.float() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    return c, f32(c['sdL']), f32(c['sdH'])


def test_train_mode_oracle_matches_reference(golden):
    """oracle.train_mode() (batch-statistics BatchNorm) + autograd against the UNMODIFIED reference modules in train():
    outputs, the 6-image MSE loss and parameter gradients (tests/golden/train_step.pt, oracle/make_golden.py gen_train)."""
    import torch.nn.functional as F
    c, sdL, sdH = _train_case(golden)
    for sd in (sdL, sdH):
        for k, v in sd.items():
            if v.is_floating_point():
                v.requires_grad_(True)
    targets = [t.float() for t in c['targets']]
    with O.train_mode():
        imgs, atts, mu, lv, _ = O.g_sr_net_low(sdL, c['LR'], c['sent'], c['words'], c['mask'])
        fine = O.netg_high(sdH, c['LR'], imgs, c['LR'], low='lr')
    loss = sum(F.mse_loss(f, t) for f, t in zip(fine, targets)) + sum(F.mse_loss(f, t) for f, t in zip(imgs, targets))
    loss.backward()
    torch.testing.assert_close(loss.detach(), c['loss'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(imgs[0].detach(), c['img0'], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(fine[0].detach(), c['fine0'], rtol=1e-4, atol=1e-5)
    for name, g in c['gradL'].items():
        torch.testing.assert_close(sdL[name].grad, g, rtol=2e-3, atol=1e-6 + 1e-4 * float(g.abs().max()))
    for name, g in c['gradH'].items():
        torch.testing.assert_close(sdH[name].grad, g, rtol=2e-3, atol=1e-6 + 1e-4 * float(g.abs().max()))


def test_misc_losses_oracle_matches_reference(golden):
    """KL_loss / cosine_similarity restatements against the unmodified miscc/losses.py outputs (tests/golden/misc_losses.pt)."""
    for c in golden('misc_losses.pt'):
        if c['kind'] == 'cos':
            close(O.cosine_similarity(c['x1'], c['x2'], c['dim']), c['out'], rtol=1e-6, atol=1e-7)
        else:
            close(O.KL_loss(c['mu'], c['logvar']), c['out'], rtol=1e-6, atol=1e-7)
