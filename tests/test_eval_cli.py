"""Eval command line (tgsr_b200/eval_cli.py, the gen_exampleSRHL flow): host-side helpers on CPU, the whole flow on GPU."""
import os
import pickle

import numpy as np
import pytest
import torch

from tgsr_b200 import eval_cli as E


def _png(path, seed, size=(178, 218)):
    from PIL import Image
    rng = np.random.RandomState(seed)
    small = rng.randint(0, 255, (9, 7, 3)).astype(np.uint8)
    Image.fromarray(small).resize(size, Image.BICUBIC).save(path)


def test_caption_pickles_and_batching(tmp_path):
    caps = [[1, 2, 3], [4, 5, 6, 7, 8], [9]]
    ix = {0: '<end>', **{i: 'w%d' % i for i in range(1, 10)}}
    wi = {v: k for k, v in ix.items()}
    p3, p4 = tmp_path / 'c3.pickle', tmp_path / 'c4.pickle'
    pickle.dump([caps, ix, wi], open(p3, 'wb')); pickle.dump([[[1]], caps, ix, wi], open(p4, 'wb'))
    for p in (p3, p4):
        got, n = E.load_captions(str(p), 4)
        assert n == 10 and got == [[1, 2, 3], [4, 5, 6, 7], [9]]
    c, lens, order = E.batch_captions(caps, 6)
    assert order.tolist() == [1, 0, 2] and lens.tolist() == [5, 3, 1]
    assert c.tolist() == [[4, 5, 6, 7, 8, 0], [1, 2, 3, 0, 0, 0], [9, 0, 0, 0, 0, 0]]


def test_preprocess_shapes(tmp_path):
    p = str(tmp_path / 'a.png'); _png(p, 1)
    x = E.preprocess(p, 256, 32)
    assert x.shape == (1, 3, 32, 32) and x.dtype == torch.float32 and float(x.min()) >= -1 and float(x.max()) <= 1


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path, golden):
    """YAML + shipped-checkpoint weights + PNGs + caption pickle -> *_SR.png equal to the quantised pipeline output."""
    import tgsr_b200
    from PIL import Image
    from tgsr_b200 import ops
    from tgsr_b200.pipeline import SRPipeline
    w = golden('face_S8_weights.pt')
    torch.save(w['netG'], tmp_path / 'netG_epoch_7.pth'); torch.save(w['netGH'], tmp_path / 'netGH_epoch_7.pth')
    (tmp_path / 'eval.yml').write_text(
        "CONFIG_NAME: 'attn2'\nDATASET_NAME: 'face'\nTREE:\n  BRANCH_NUM: 4\n  BASE_SIZE: 32\nTRAIN:\n  FLAG: False\n  BATCH_SIZE: 1\n"
        "GAN:\n  DF_DIM: 64\n  GF_DIM: 32\n  Z_DIM: 100\n  R_NUM: 2\nTEXT:\n  EMBEDDING_DIM: 256\n  CAPTIONS_PER_IMAGE: 1\n  WORDS_NUM: 18\n")
    caps = [[1, 2, 17, 15, 4, 5, 6], [3, 4, 5], [2, 17, 3, 4, 7, 8, 19, 11, 12, 13]]
    ix = {i: 'w%d' % i for i in range(41)}
    pickle.dump([caps, ix, {v: k for k, v in ix.items()}], open(tmp_path / 'caps.pickle', 'wb'))
    imgs = []
    for k in range(3):
        p = str(tmp_path / ('im%d.png' % k)); _png(p, 10 + k); imgs.append(p)
    out = tmp_path / 'out'
    rc = E.main(['--cfg', str(tmp_path / 'eval.yml'), '--net-g', str(tmp_path / 'netG_epoch_7.pth'), '--captions',
                 str(tmp_path / 'caps.pickle'), '--out', str(out), '--precision', 'mixed16'] + imgs)
    assert rc == 0
    # the same flow through the Python API
    tgsr_b200.set_precision('mixed16')
    torch.manual_seed(100)
    pipe = SRPipeline(41, device='cuda:0', variant='x8', low='lr')
    pipe.load_generators(w['netG'], w['netGH'])
    LR = torch.cat([E.preprocess(p, 256, 32) for p in imgs])
    c, lens, order = E.batch_captions(caps, 18)
    fine, _, _ = pipe.forward_device(LR[order].cuda(), c.cuda(), lens.cuda())
    u8 = ops.quantize_u8(fine[-1]).cpu().numpy()
    for r, i in enumerate(order.tolist()):
        got = np.asarray(Image.open(out / ('im%d_SR.png' % i)))
        assert got.shape == (256, 256, 3) and np.array_equal(got, u8[r].transpose(1, 2, 0))
