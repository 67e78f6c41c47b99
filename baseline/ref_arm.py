"""The reference arm: cxm12/TGSR's own, unmodified modules (baseline/_ref/, see install_ref.py) driven through the body
of condGANTrainer.gen_exampleSRHL (trainer_objective.py:134-146): RNN_ENCODER -> mask -> G_SR_NET_low ->
NetG_highweight(weightmap=False, low='lr').

Used by bench.py only: `--impl reference` (CPU timing on the host cores, plus the informational torch-eager-on-GPU
leg) and the `cpu_baseline` leg of the product arm.  It imports nothing from tgsr_b200 except the synthetic-input
generators (tgsr_b200/synth.py); no kernel of the product runs here.
"""
from __future__ import annotations

import contextlib
import os
import statistics
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')
SHIMS = os.path.join(HERE, 'shims')

# cfg/eval_birdSR_attn2.yml:5-7,24-31 (== eval_faceSR_attn2.yml in every model dimension)
EVAL_CFG = dict(GF_DIM=32, DF_DIM=64, Z_DIM=100, R_NUM=2, EMBEDDING_DIM=256, BRANCH_NUM=4, BASE_SIZE=32)


def available() -> bool:
    return os.path.isfile(os.path.join(REF, 'GlobalAttention.py'))


def load():
    """Import the reference modules from baseline/_ref with the eval YAML values applied to their global cfg."""
    if not available():
        raise RuntimeError('baseline/_ref is empty: run `python baseline/install_ref.py` where /root/reference is mounted')
    for p in (SHIMS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import GlobalAttention, model, util          # noqa: E401  (the reference's own files)
    from miscc.config import cfg
    cfg.GAN.GF_DIM = EVAL_CFG['GF_DIM']; cfg.GAN.DF_DIM = EVAL_CFG['DF_DIM']
    cfg.GAN.Z_DIM = EVAL_CFG['Z_DIM']; cfg.GAN.R_NUM = EVAL_CFG['R_NUM']
    cfg.TEXT.EMBEDDING_DIM = EVAL_CFG['EMBEDDING_DIM']
    cfg.TREE.BRANCH_NUM = EVAL_CFG['BRANCH_NUM']; cfg.TREE.BASE_SIZE = EVAL_CFG['BASE_SIZE']
    GlobalAttention.server = 1     # bool masks: the uint8 masked_fill_ branch is a RuntimeError on torch 2.x
    return {'GlobalAttention': GlobalAttention, 'model': model, 'util': util, 'cfg': cfg}


@contextlib.contextmanager
def _cuda_noop():
    """NetG_highweight.__init__ calls .cuda() on `a` and `one` unconditionally (model.py:246-248).  For the CPU arm the
    call is neutralised during construction only; `a` stays the unregistered constant 0.5 either way."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self.detach().clone()
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def build(ref, ntoken: int, device: str = 'cpu', seed: int = 100):
    """The three nets of the eval flow from the reference's own constructors (trainer_objective.py:58-88), default torch
    init under `seed` (test1.py:170), BatchNorm statistics randomised as SURVEY.md 8d cfg 2 prescribes."""
    import torch
    import torch.nn as nn
    ref['cfg'].CUDA = device != 'cpu'
    torch.manual_seed(seed)
    enc = ref['model'].RNN_ENCODER(ntoken, nhidden=ref['cfg'].TEXT.EMBEDDING_DIM)
    gl = ref['model'].G_SR_NET_low()
    if device == 'cpu':
        with _cuda_noop():
            gh = ref['model'].NetG_highweight(weightmap=False, low='lr')
    else:
        gh = ref['model'].NetG_highweight(weightmap=False, low='lr')
    g = torch.Generator().manual_seed(1234 + seed)
    for mod in (gl, gh):
        for m in mod.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.data.copy_(1 + 0.02 * torch.randn(m.weight.shape, generator=g))
                m.bias.data.zero_()
    for m in (enc, gl, gh):
        m.to(device).eval()
    return enc, gl, gh


def forward(nets, LR, captions, cap_lens):
    """gen_exampleSRHL's per-batch body, trainer_objective.py:134-146."""
    import torch
    enc, gl, gh = nets
    with torch.no_grad():
        hidden = enc.init_hidden(captions.shape[0])
        words_embs, sent_emb = enc(captions, cap_lens, hidden)
        mask = (captions == 0)
        num_words = words_embs.size(2)
        if mask.size(1) > num_words:
            mask = mask[:, :num_words]
        fake_imgL, attention_maps, mu, logvar = gl(LR, sent_emb, words_embs, mask)
        fine_im, a, one = gh(LR, fake_imgL, LR)
    return fine_im, fake_imgL, attention_maps


def inputs(B: int, ntoken: int, words: int, device: str = 'cpu', seed: int = 101):
    from tgsr_b200 import synth
    LR = synth.synth_lr(B, 32, seed=seed).to(device)
    caps, lens = synth.synth_captions(B, ntoken, seed=seed + 1, words_num=words)
    return LR, caps.to(device), lens      # lengths stay on the host: RNN_ENCODER.forward does cap_lens.data.tolist()


def time_cpu(nets, B: int, ntoken: int, words: int, threads: int, warmup: int, iters: int):
    """(seconds per pass: list) of `iters` timed passes over one B-image batch on `threads` host threads."""
    import torch
    torch.set_num_threads(threads)
    LR, caps, lens = inputs(B, ntoken, words)
    for _ in range(warmup):
        forward(nets, LR, caps, lens)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        forward(nets, LR, caps, lens)
        ts.append(time.perf_counter() - t0)
    return ts


def cpu_protocol(nets, ntoken: int, words: int, cores: int, quick: bool = False):
    """SURVEY.md 8d "CPU reference timing": same process, B = 1 and B = 8, all host threads and one thread, 3 warm-ups and
    the median of >= 10 passes (the one-thread B = 8 row: 1 warm-up, 3 passes -- ~4 s per pass).  images/s = B / median."""
    rows = {}
    plan = [('all_threads_b1', 1, cores, 3, 10), ('all_threads_b8', 8, cores, 3, 10), ('one_thread_b1', 1, 1, 3, 10)]
    if not quick:
        plan.append(('one_thread_b8', 8, 1, 1, 3))
    for name, B, thr, w, n in plan:
        ts = time_cpu(nets, B, ntoken, words, thr, w, n)
        med = statistics.median(ts)
        rows[name] = {'images_per_s': B / med, 'ms_per_pass_median': 1e3 * med, 'batch': B, 'threads': thr, 'passes': n, 'warmup': w}
    return rows


def torch_eager_gpu(ref, ntoken: int, words: int, B: int, iters: int = 10):
    """Informational (SURVEY.md 8d "library comparison"): the unmodified reference modules under torch eager on the GPU
    (ATen / cuDNN / cuBLAS), fp32 with PyTorch's default TF32 convs and under bf16 autocast, device-resident inputs."""
    import torch
    nets = build(ref, ntoken, device='cuda')
    LR, caps, lens = inputs(B, ntoken, words, device='cuda')
    out = {}
    for name, ctx in (('fp32_tf32_convs', contextlib.nullcontext()), ('bf16_autocast', torch.autocast('cuda', dtype=torch.bfloat16))):
        with ctx:
            for _ in range(3):
                forward(nets, LR, caps, lens)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                forward(nets, LR, caps, lens)
            e.record(); e.synchronize()
        out[name] = {'value': B * iters / (s.elapsed_time(e) * 1e-3), 'unit': 'images/s', 'batch': B}
    return out


def torch_eager_train_gpu(ref, ntoken: int, words: int, B: int, iters: int = 5):
    """Informational: BASELINE.json configs[4] (generator training step) composed from the UNMODIFIED reference modules under
    torch eager + autograd on the GPU: G_SR_NET_low / NetG_highweight in train mode, the reference's own
    generator_lossor_wordsentOnly (DAMSM words / sentence losses: a Python loop over the batch, miscc/losses.py:499-522) and
    MSE (:779-784), Adam(2e-4, (0.5, 0.999)).  The image encoder is the reference's CNN_ENCODER (util.py:262-368) on a
    random-init Inception-v3 (its pretrained download is replaced by the constructor's own weights: no network here)."""
    import torch
    import torch.nn.functional as F
    import torchvision
    util, cfg = ref['util'], ref['cfg']
    from miscc import losses as ref_losses                       # the reference's file (baseline/_ref/miscc/losses.py)
    cfg.CUDA = True
    enc, gl, gh = build(ref, ntoken, device='cuda')
    gl.train(); gh.train()
    orig = util.model_zoo.load_url
    util.model_zoo.load_url = lambda url, *a, **k: torchvision.models.inception_v3(weights=None, aux_logits=True,
                                                                                   init_weights=False).state_dict()
    try:
        img_enc = util.CNN_ENCODER(cfg.TEXT.EMBEDDING_DIM)
    finally:
        util.model_zoo.load_url = orig
    img_enc = img_enc.cuda().eval()
    for p in img_enc.parameters():
        p.requires_grad = False
    LR, caps, lens = inputs(B, ntoken, words, device='cuda')
    real = [F.interpolate(LR, scale_factor=2 ** (k + 1), mode='bicubic').clamp(-1, 1) for k in range(3)]
    labels = torch.arange(B, device='cuda')
    params = [p for p in list(gl.parameters()) + list(gh.parameters()) if p.requires_grad]
    opt = torch.optim.Adam(params, lr=2e-4, betas=(0.5, 0.999))

    def step():
        with torch.no_grad():
            hidden = enc.init_hidden(B)
            words_embs, sent_emb = enc(caps, lens, hidden)
        mask = (caps == 0)[:, :words_embs.size(2)]
        fake_imgL, _att, _mu, _logvar = gl(LR, sent_emb, words_embs, mask)
        fine, _a, _one = gh(LR, fake_imgL, LR)
        errG, _ = ref_losses.generator_lossor_wordsentOnly(img_enc, fine, words_embs, sent_emb, labels, lens, None)
        mse = ref_losses.MSE(list(fine) + list(fake_imgL), list(real) + list(real))
        opt.zero_grad(set_to_none=True)
        (errG + mse).backward()
        opt.step()

    out = {}
    for name, ctx in (('fp32_tf32_convs', contextlib.nullcontext), ('bf16_autocast', lambda: torch.autocast('cuda', dtype=torch.bfloat16))):
        try:
            for _ in range(2):
                with ctx():
                    step()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                with ctx():
                    step()
            e.record(); e.synchronize()
            ms = s.elapsed_time(e) / iters
            out[name] = {'value': B / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms, 'batch': B}
        except Exception as ex:                                  # noqa: BLE001  (informational leg: report, do not fail the arm)
            out[name] = {'error': repr(ex)[:200]}
    return out
