"""DAMSM matching losses and MSE (-m gpu): tgsr_b200.losses (csrc/damsm.cu) against the reference's own
miscc/losses.py outputs and autograd gradients (tests/golden/losses.pt) and against the oracle at batch 32."""
import pytest
import torch

import tgsr_b200
from oracle import tgsr_oracle as O
from tgsr_b200 import losses as L
from tgsr_b200.config import cfg

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _cfg():
    tgsr_b200.set_eval_defaults()
    tgsr_b200.set_precision('fp32')
    old = dict(cfg.TRAIN.SMOOTH)
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 4.0, 5.0, 10.0, 1.0
    yield
    for k, v in old.items():
        cfg.TRAIN.SMOOTH[k] = v


def test_losses_golden(golden):
    for c in golden('losses.pt'):
        B = c['feats'].shape[0]
        feats = c['feats'].float().to(DEV).requires_grad_(True); cnn = c['cnn'].to(DEV).requires_grad_(True)
        w0, w1, att = L.words_loss(feats, c['words'].to(DEV), c['labels'].to(DEV), c['lens'], c['class_ids'], B)
        s0, s1 = L.sent_loss(cnn, c['sent'].to(DEV), c['labels'].to(DEV), c['class_ids'], B)
        for a, b in ((w0, c['w0']), (w1, c['w1']), (s0, c['s0']), (s1, c['s1'])):
            torch.testing.assert_close(a.detach().cpu(), b, rtol=1e-4, atol=1e-5)
        assert len(att) == B
        for a, b in zip(att, c['att']):
            assert a.shape == b.shape
            torch.testing.assert_close(a.detach().cpu(), b, rtol=1e-4, atol=1e-7)
        gw, gs = c['gw'].to(DEV), c['gs'].to(DEV)
        (gw[0] * w0 + gw[1] * w1 + gs[0] * s0 + gs[1] * s1).backward()
        torch.testing.assert_close(feats.grad.cpu(), c['d_feats'], rtol=2e-3, atol=2e-7)
        torch.testing.assert_close(cnn.grad.cpu(), c['d_cnn'], rtol=1e-3, atol=1e-7)
        fake = [f.to(DEV).requires_grad_(True) for f in c['fake']]
        mse = L.MSE(fake, [r.to(DEV) for r in c['real']])
        torch.testing.assert_close(mse.detach().cpu(), c['mse'], rtol=1e-5, atol=1e-6)
        (1.7 * mse).backward()
        for f, d in zip(fake, c['d_fake']):
            torch.testing.assert_close(f.grad.cpu(), d, rtol=1e-5, atol=1e-8)


def test_words_loss_batch32_vs_oracle():
    """BASELINE configs[4] shape: per-GPU batch 32, 18-word captions, ragged lengths."""
    B, T = 32, 18
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(B, 256, 17, 17, generator=g) * 0.4; words = torch.randn(B, 256, T, generator=g) * 0.4
    lens = torch.sort(torch.randint(5, T + 1, (B,), generator=g), descending=True).values
    labels = torch.arange(B)
    fr = feats.clone().requires_grad_(True)
    r0, r1, ratt = O.words_loss(fr, words, labels, lens, None, 4.0, 5.0, 10.0)
    (r0 + r1).backward()
    fg = feats.to(DEV).requires_grad_(True)
    w0, w1, att = L.words_loss(fg, words.to(DEV), labels.to(DEV), lens, None, B)
    torch.testing.assert_close(w0.detach().cpu(), r0.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w1.detach().cpu(), r1.detach(), rtol=1e-4, atol=1e-5)
    (w0 + w1).backward()
    torch.testing.assert_close(fg.grad.cpu(), fr.grad, rtol=5e-3, atol=1e-7)
    for a, b in zip(att, ratt):
        torch.testing.assert_close(a.detach().cpu(), b.detach(), rtol=1e-4, atol=1e-7)


def test_kl_and_cosine_golden(golden):
    """KL_loss / cosine_similarity (csrc/loss_misc.cu) forward and backward against the reference's own outputs and
    autograd gradients (tests/golden/misc_losses.pt), including a zero vector where the eps clamp is active."""
    for c in golden('misc_losses.pt'):
        if c['kind'] == 'cos':
            x1 = c['x1'].to(DEV).requires_grad_(True); x2 = c['x2'].to(DEV).requires_grad_(True)
            out = L.cosine_similarity(x1, x2, c['dim'])
            assert out.shape == c['out'].shape
            torch.testing.assert_close(out.detach().cpu(), c['out'], rtol=1e-5, atol=1e-6)
            (out * c['g'].to(DEV)).sum().backward()
            torch.testing.assert_close(x1.grad.cpu(), c['d1'], rtol=1e-4, atol=1e-6)
            torch.testing.assert_close(x2.grad.cpu(), c['d2'], rtol=1e-4, atol=1e-6)
        else:
            mu = c['mu'].to(DEV).requires_grad_(True); lv = c['logvar'].to(DEV).requires_grad_(True)
            k = L.KL_loss(mu, lv)
            torch.testing.assert_close(k.detach().cpu(), c['out'], rtol=1e-5, atol=1e-6)
            (c['gscale'] * k).backward()
            torch.testing.assert_close(mu.grad.cpu(), c['d_mu'], rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(lv.grad.cpu(), c['d_logvar'], rtol=1e-5, atol=1e-7)
