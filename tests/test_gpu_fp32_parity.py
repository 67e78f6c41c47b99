"""GPU parity of the fp32 CUDA path against the CPU oracle and the committed golden fixtures.
Everything goes through the C ABI (tgsr_b200 modules -> ctypes -> libtgsr_sm100.so)."""
import pytest
import torch

import tgsr_b200
from oracle import tgsr_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.detach().cpu(), b, rtol=rtol, atol=atol)


@pytest.fixture(autouse=True)
def _cfg():
    tgsr_b200.set_eval_defaults()
    tgsr_b200.set_precision('fp32')
    yield


def test_attention_golden(golden):
    for c in golden('attention.pt'):
        idf, cdf = c['w'].shape[0], c['w'].shape[1]
        att = tgsr_b200.GlobalAttentionGeneral(idf, cdf).to(DEV)
        att.conv_context.weight.data.copy_(c['w'])
        att.applyMask(None if c['mask'] is None else c['mask'].to(DEV))
        wctx, attn = att(c['inp'].to(DEV), c['ctx'].to(DEV))
        close(attn, c['attn'], rtol=1e-5, atol=1e-7)          # north_star: 1e-5 relative on fp32 maps
        close(wctx, c['wctx'], rtol=1e-4, atol=5e-5)
        assert torch.equal(attn.cpu() == 0, c['attn'] == 0)   # caption masking bit-exact
        assert torch.equal(attn.cpu().argmax(1), c['attn'].argmax(1))


@pytest.mark.parametrize('B,idf,cdf,ih,iw,T,ragged', [
    (1, 32, 256, 32, 32, 18, False), (4, 32, 256, 64, 64, 18, True), (3, 32, 256, 17, 13, 18, True),
    (2, 64, 64, 32, 32, 32, True), (2, 128, 128, 16, 16, 48, True), (2, 256, 256, 8, 8, 18, False),
    (7, 32, 256, 16, 16, 5, True), (2, 32, 32, 9, 7, 33, True)])
def test_attention_vs_oracle(B, idf, cdf, ih, iw, T, ragged):
    g = torch.Generator().manual_seed(B * 1000 + idf + T)
    inp = torch.randn(B, idf, ih, iw, generator=g) * 0.5
    ctx = torch.randn(B, cdf, T, generator=g) * 0.5
    w = torch.randn(idf, cdf, 1, 1, generator=g) * (1.6 / cdf ** 0.5) * (32 / idf) ** 0.5
    mask = None
    if ragged:
        lens = torch.randint(1, T + 1, (B,), generator=g); lens[0] = T
        mask = torch.arange(T)[None] >= lens[:, None]
    ref_w, ref_a = O.global_attention(inp, ctx, w, mask)
    att = tgsr_b200.GlobalAttentionGeneral(idf, cdf).to(DEV)
    att.conv_context.weight.data.copy_(w)
    att.applyMask(None if mask is None else mask.to(DEV))
    wctx, attn = att(inp.to(DEV), ctx.to(DEV))
    close(attn, ref_a, rtol=1e-5, atol=1e-7)
    close(wctx, ref_w, rtol=1e-4, atol=5e-5)
    assert torch.equal(attn.cpu() == 0, ref_a == 0)
    # argmax: exclude near-ties (top-2 margin below fp32 noise) by construction
    top2 = ref_a.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-5
    assert torch.equal(attn.cpu().argmax(1)[safe], ref_a.argmax(1)[safe])
    close(attn.sum(1), torch.ones(B, ih, iw), rtol=0, atol=1e-5)          # rows sum to one


def test_attention_cuda_core_kernel_kept():
    """The CUDA-core fp32 kernel (used outside the tensor-core envelope) on a shape the tensor-core path would take."""
    from tgsr_b200 import ops
    g = torch.Generator().manual_seed(77)
    inp = torch.randn(2, 32, 16, 16, generator=g) * 0.5; ctx = torch.randn(2, 256, 18, generator=g) * 0.5
    w = torch.randn(32, 256, 1, 1, generator=g) * 0.1
    mask = torch.arange(18)[None] >= torch.tensor([18, 7])[:, None]
    ref_w, ref_a = O.global_attention(inp, ctx, w, mask)
    for tcores in (False, True):
        wctx, attn = ops.attention_fwd(inp.to(DEV), ctx.to(DEV), w.to(DEV), mask.to(DEV), tensor_cores=tcores)
        close(attn, ref_a, rtol=1e-5, atol=1e-7); close(wctx, ref_w, rtol=1e-4, atol=5e-5)
        assert torch.equal(attn.cpu() == 0, ref_a == 0)


def test_attention_errors():
    att = tgsr_b200.GlobalAttentionGeneral(32, 256).to(DEV)
    with pytest.raises(RuntimeError, match='unsupported'):
        att(torch.zeros(1, 32, 4, 4, device=DEV), torch.zeros(1, 256, 49, device=DEV))
    with pytest.raises(ValueError):
        att(torch.zeros(2, 32, 4, 4, device=DEV), torch.zeros(1, 256, 5, device=DEV))


def test_func_attention(golden):
    for c in golden('func_attention.pt'):
        w, a = tgsr_b200.func_attention(c['query'].to(DEV), c['context'].to(DEV), c['gamma1'])
        close(w, c['wctx'], rtol=1e-4, atol=1e-5); close(a, c['attn'], rtol=1e-4, atol=1e-7)


def test_blocks_golden(golden):
    from tgsr_b200 import blocks
    g = golden('blocks.pt')
    mk = {'block3x3_3_32': lambda: blocks.Block3x3_relu(3, 32), 'block3x3_32_32': lambda: blocks.Block3x3_relu(32, 32),
          'resblock_64': lambda: blocks.ResBlock(64), 'resblock_32': lambda: blocks.ResBlock(32),
          'upblock_64_32': lambda: blocks.upBlock(64, 32), 'upblock_32_32': lambda: blocks.upBlock(32, 32),
          'get_image': lambda: blocks.GET_IMAGE_G(32), 'get_image_noact': lambda: blocks.GET_IMAGE_G_noAct(32)}
    for name, ctor in mk.items():
        m = ctor(); m.load_state_dict(g[name]['sd'], strict=True); m.to(DEV).eval()
        close(m(g[name]['x'].to(DEV)), g[name]['y'], rtol=1e-4, atol=2e-5)
    ca = blocks.CA_NET(); ca.load_state_dict(g['ca_net']['sd'], strict=True); ca.to(DEV).eval()
    c, mu, lv = ca(g['ca_net']['x'].to(DEV))
    close(mu, g['ca_net']['mu'], rtol=1e-4, atol=1e-6); close(lv, g['ca_net']['logvar'], rtol=1e-4, atol=1e-6)
    assert c.shape == mu.shape and torch.isfinite(c).all()
    ig = blocks.INIT_STAGE_G(32, 100); ig.load_state_dict(g['init_stage_g']['sd'], strict=True); ig.to(DEV).eval()
    close(ig(g['init_stage_g']['z'].to(DEV), g['init_stage_g']['c'].to(DEV)), g['init_stage_g']['y'], rtol=1e-4,
          atol=2e-5)


def test_train_mode_needs_fp32():
    from tgsr_b200 import blocks
    m = blocks.ResBlock(64).to(DEV)          # training mode by default
    tgsr_b200.set_precision('mixed16')
    try:
        with pytest.raises(RuntimeError, match='fp32|eval'):      # no 16-bit training path: raises instead of computing something else
            m(torch.zeros(1, 64, 16, 8, device=DEV))
    finally:
        tgsr_b200.set_precision('fp32')


def test_rnn_encoder_golden(golden):
    from tgsr_b200.text_encoder import RNN_ENCODER
    g = golden('rnn_encoder.pt')
    enc = RNN_ENCODER(41, nhidden=256); enc.load_state_dict(g['sd'], strict=True); enc.to(DEV).eval()
    B = g['captions'].shape[0]
    words, sent = enc(g['captions'].to(DEV), g['cap_lens'].to(DEV), enc.init_hidden(B))
    assert words.shape == g['words'].shape
    close(words, g['words'], rtol=1e-4, atol=5e-6); close(sent, g['sent'], rtol=1e-4, atol=5e-6)
    for b, L in enumerate(g['cap_lens'].tolist()):
        assert torch.count_nonzero(words[b, :, L:]) == 0
    # unsorted lengths and a CPU length list are handled too (no pack_padded_sequence constraint)
    perm = torch.tensor([3, 0, 7, 1, 2, 6, 5, 4])
    w2, s2 = enc(g['captions'][perm].to(DEV), g['cap_lens'][perm].tolist(), enc.init_hidden(B))
    close(w2, g['words'][perm], rtol=1e-4, atol=5e-6); close(s2, g['sent'][perm], rtol=1e-4, atol=5e-6)


def test_rnn_encoder_vs_oracle_big_vocab():
    from tgsr_b200.text_encoder import RNN_ENCODER
    torch.manual_seed(5)
    enc = RNN_ENCODER(5450, nhidden=256).eval()
    caps, lens = O.synth_captions(19, 5450, seed=3, ragged=True)
    rw, rs = O.rnn_encoder(enc.state_dict(), caps, lens)
    enc.to(DEV)
    w, s = enc(caps.to(DEV), lens.to(DEV), enc.init_hidden(19))
    close(w, rw, rtol=1e-4, atol=5e-6); close(s, rs, rtol=1e-4, atol=5e-6)


def _face_nets(golden):
    from tgsr_b200 import nets
    w = golden('face_S8_weights.pt')
    gl = nets.G_SR_NET_low(); gl.load_state_dict(w['netG'], strict=True); gl.to(DEV).eval()
    gh = nets.NetG_highweight(weightmap=False, low='lr'); gh.load_state_dict(w['netGH'], strict=True); gh.to(DEV).eval()
    return gl, gh


def test_face_x8_checkpoint_fp32(golden):
    """Shipped checkpoint + real image fixture, whole eval path (trainer_objective.py:134-146)."""
    gl, gh = _face_nets(golden)
    g = golden('face_x8.pt')
    c = g['b1']
    imgs, atts, mu, lv = gl(c['LR'].to(DEV), c['sent'].to(DEV), c['words'].to(DEV), c['mask'].to(DEV))
    for a, b in zip(imgs, c['imgsL']):
        close(a, b, rtol=1e-3, atol=1e-4)
    close(atts[0], c['atts'][0], rtol=1e-4, atol=1e-6); close(atts[1], c['atts'][1], rtol=1e-3, atol=1e-6)
    close(atts[2][:, :, ::4, ::4], c['att2_sub4'], rtol=1e-3, atol=1e-6)
    close(mu, c['mu'], rtol=1e-4, atol=1e-6)
    fine, a, one = gh(c['LR'].to(DEV), imgs, c['LR'].to(DEV))
    assert float(a) == 0.5 and float(one) == 1.0
    for x, y in zip(fine, c['fine']):
        close(x, y, rtol=1e-3, atol=1e-4)
    sr, ref = O.to_uint8(fine[-1].cpu()), O.to_uint8(c['fine'][-1])
    assert (sr.int() - ref.int()).abs().max() <= 1
    c = g['b3']                                                   # ragged batch -> mask quirk
    imgs, atts, mu, lv = gl(c['LR'].to(DEV), c['sent'].to(DEV), c['words'].to(DEV), c['mask'].to(DEV))
    close(imgs[0], c['imgL0'], rtol=1e-3, atol=1e-4); close(atts[0], c['att0'], rtol=1e-4, atol=1e-6)
    assert torch.equal(atts[0].cpu() == 0, c['att0'] == 0)
    fine, _, _ = gh(c['LR'].to(DEV), imgs, c['LR'].to(DEV))
    close(fine[2][:, :, ::4, ::4], c['fine2_sub4'], rtol=1e-3, atol=2e-4)


def test_x16_nets_fp32(golden):
    from tgsr_b200 import nets16
    g = golden('x16.pt')
    tgsr_b200.cfg.GAN.GF_DIM = 8
    try:
        gl = nets16.G_SR_NET_low()
    finally:
        tgsr_b200.cfg.GAN.GF_DIM = 32
    gl.load_state_dict(g['sdL'], strict=True); gl.to(DEV).eval()
    imgs, atts, mu, lv = gl(g['LR'].to(DEV), g['sent'].to(DEV), g['words'].to(DEV), g['mask'].to(DEV))
    assert len(imgs) == 4
    for a, b in zip(imgs[:3], g['imgsL']):
        close(a, b, rtol=1e-3, atol=1e-4)
    for a, b in zip(atts[:3], g['atts']):
        close(a, b, rtol=1e-3, atol=1e-6)
    gh = nets16.NetG_high(); gh.load_state_dict(g['sdH'], strict=True); gh.to(DEV).eval()
    fine = gh(g['LRh'].to(DEV), [t.to(DEV) for t in g['SRb']], g['LRh'].to(DEV), low=True)
    for a, b in zip(fine, g['fine']):
        close(a, b, rtol=1e-3, atol=1e-4)


def _x16_nets(g):
    from tgsr_b200 import nets16
    gl = nets16.G_SR_NET_low(); gl.load_state_dict(g['sdL'], strict=True); gl.to(DEV).eval()
    gh = nets16.NetG_high(); gh.load_state_dict(g['sdH'], strict=True); gh.to(DEV).eval()
    ghw = nets16.NetG_highweight(weightmap=False, low='lr'); ghw.load_state_dict(g['sdHw'], strict=True); ghw.to(DEV).eval()
    return gl, gh, ghw


def run_x16_gf32(g):
    """All three models16 nets at GF_DIM = 32 on the fixture's inputs (shared by the fp32 and the 16-bit test)."""
    gl, gh, ghw = _x16_nets(g)
    LR, LRb = g['LR'].to(DEV), g['LRb'].to(DEV)
    imgs, atts, mu, lv = gl(LR, g['sent'].to(DEV), g['words'].to(DEV), g['mask'].to(DEV))
    fine = gh(LR, imgs, LRb, low=False)
    B = LR.shape[0]
    c8 = g['c8'].to(DEV)
    srb_w = [imgs[0], imgs[1], c8.expand(B, 3, 128, 128).contiguous(), c8.expand(B, 3, 256, 256).contiguous()]
    finew, a, one = ghw(LR, srb_w, LRb)
    assert float(a) == 0.5 and float(one) == 1.0 and 'a' in ghw.state_dict()       # registered here (models16.py:131)
    return imgs, atts, mu, fine, finew


def test_x16_production_width_fp32(golden):
    """models16 at GF_DIM = 32 against the unmodified reference (x16_gf32.pt), fp32 kernels."""
    g = golden('x16_gf32.pt')
    imgs, atts, mu, fine, finew = run_x16_gf32(g)
    close(mu, g['mu'], rtol=1e-4, atol=1e-6)
    for k, sub in enumerate((1, 1, 2, 4)):
        close(imgs[k][:, :, ::sub, ::sub], g['imgsL'][k], rtol=1e-3, atol=1e-4)
        close(fine[k][:, :, ::sub, ::sub], g['fine'][k], rtol=1e-3, atol=2e-4)
        close(finew[k][:, :, ::sub, ::sub], g['finew'][k], rtol=1e-3, atol=2e-4)
    for k, sub in enumerate((1, 2, 4, 8)):
        close(atts[k][:, :, ::sub, ::sub], g['atts'][k], rtol=1e-3, atol=1e-6)
        assert torch.equal(atts[k][:, :, ::sub, ::sub].cpu() == 0, g['atts'][k] == 0)


def test_conv_properties_full_size():
    """Size-independent checks at the production size (256^2 output): linearity of the conv
    path and equivalence of the fused upsample with an explicit nearest upsample."""
    from tgsr_b200 import ops
    g = torch.Generator().manual_seed(11)
    x1 = torch.randn(2, 64, 128, 128, generator=g).to(DEV); x2 = torch.randn(2, 64, 128, 128, generator=g).to(DEV)
    w = (torch.randn(64, 64, 3, 3, generator=g) * 0.05).to(DEV)
    y1 = ops.conv2d_fused(x1, w, None, None, 0, up2x=True); y2 = ops.conv2d_fused(x2, w, None, None, 0, up2x=True)
    y12 = ops.conv2d_fused(x1 + x2, w, None, None, 0, up2x=True)
    assert y12.shape == (2, 64, 256, 256)
    torch.testing.assert_close(y12, y1 + y2, rtol=1e-4, atol=1e-4)
    xu = x1.repeat_interleave(2, 2).repeat_interleave(2, 3).contiguous()
    torch.testing.assert_close(ops.conv2d_fused(xu, w, None, None, 0), y1, rtol=0, atol=0)


def test_attention_all_masked_caption_is_nan_like_torch():
    """A caption whose words are all masked gives NaN rows in the reference (softmax over -inf only); the kernels keep
    that behaviour at exactly the same pixels (row-order quirk included) and stay finite elsewhere."""
    from tgsr_b200 import fast, ops, tc
    g = torch.Generator().manual_seed(123)
    B, idf, cdf, ih, iw, T = 3, 32, 256, 16, 8, 6
    inp = torch.randn(B, idf, ih, iw, generator=g) * 0.5
    ctx = torch.randn(B, cdf, T, generator=g) * 0.5
    w = torch.randn(idf, cdf, 1, 1, generator=g) * 0.1
    mask = torch.zeros(B, T, dtype=torch.bool); mask[1, :] = True; mask[2, 3:] = True
    ref_w, ref_a = O.global_attention(inp, ctx, w, mask)
    assert torch.isnan(ref_a).any() and not torch.isnan(ref_a).all()
    for tcores in (True, False):
        wctx, attn = ops.attention_fwd(inp.to(DEV), ctx.to(DEV), w.to(DEV), mask.to(DEV), tensor_cores=tcores)
        assert torch.equal(torch.isnan(attn.cpu()), torch.isnan(ref_a)) and torch.equal(torch.isnan(wctx.cpu()), torch.isnan(ref_w))
        ok = ~torch.isnan(ref_a)
        close(attn.cpu()[ok], ref_a[ok], rtol=1e-5, atol=1e-7)
    # blocked 16-bit path
    tgsr_b200.set_precision('mixed16')
    try:
        att = tgsr_b200.GlobalAttentionGeneral(idf, cdf).to(DEV); att.conv_context.weight.data.copy_(w)
        xb = inp.to(torch.bfloat16).float()
        rw, ra = O.global_attention(xb, ctx, w, mask)
        cat = tc.Blocked.empty(B, 64, ih, iw, DEV)
        h = tc.to_blocked(xb.to(DEV), out=cat.slice(0, 32))
        a16 = fast.attention(h, att, ctx.to(DEV), mask.to(DEV), out=cat.slice(32, 32)).cpu()
        assert torch.equal(torch.isnan(a16), torch.isnan(ra))
        ok = ~torch.isnan(ra)
        torch.testing.assert_close(a16[ok], ra[ok], rtol=5e-5, atol=1e-7)
    finally:
        tgsr_b200.set_precision('fp32')


def test_attention_word_stats_pinned(golden):
    """The kernel against values recorded from the unmodified miscc/utils.py (tests/golden/attn_stats.pt)."""
    from tgsr_b200 import ops
    g = golden('attn_stats.pt')
    attn = g['attn']
    gmx, gam, _ = ops.attention_word_stats(attn.cuda(), 0.0)
    assert torch.equal(gmx.cpu()[:, 0], g['max_maps'])
    assert torch.equal(gam.cpu()[:, 0], attn.argmax(1))
    for b, n in enumerate(g['lens']):
        _, _, conf = ops.attention_word_stats(attn[b:b + 1, :n].contiguous().cuda(), 2.0 / n)
        torch.testing.assert_close(conf.cpu()[0, 1:], g['conf'][b], rtol=1e-6, atol=1e-9)


def test_attention_word_stats():
    """Per-pixel max / arg-max over words and thresholded confidence sums (miscc/utils.py:122-125, 389-392) against the oracle."""
    from tgsr_b200 import ops
    g = torch.Generator().manual_seed(5)
    for (B, T, s, thresh) in [(3, 18, 32, 0.02), (2, 7, 64, 0.0), (4, 18, 128, 1.0 / 18), (1, 48, 6, 0.01)]:
        attn = torch.softmax(torch.randn((B, T, s, s), generator=g) * 2.0, dim=1)
        attn[0, :, 0, 0] = 1.0 / T                              # exact ties: first index wins
        mx, am, conf = O.attention_word_stats(attn, thresh)
        gmx, gam, gconf = ops.attention_word_stats(attn.cuda(), thresh)
        assert torch.equal(gmx.cpu(), mx)
        assert torch.equal(gam.cpu(), am)
        torch.testing.assert_close(gconf.cpu(), conf, rtol=1e-5, atol=1e-6)
