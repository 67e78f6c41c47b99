"""Command-line eval flow: the body of test1.py / condGANTrainer.gen_exampleSRHL (trainer_objective.py:55-165) on the
B200 modules -- YAML config, shipped checkpoints, PNG in, `<name>_SR.png` out.

    python -m tgsr_b200.eval_cli --cfg cfg/eval_faceSR_attn2.yml --net-g Checkpoint/face_S8/netG_epoch_7.pth \\
        --captions data/face/testset.pickle --out out/ data/face/000155.png [more.png ...]

Preprocessing follows datasets.py:1558-1563,216 (Resize(imsize*72/64) -> CenterCrop(imsize) -> Resize(LR) ->
Normalize(.5,.5)); captions come from the reference's pickle ([train_caps, test_caps, ixtoword, wordtoix],
datasets.py:725-760, or the shipped 3-entry testset.pickle; caption k pairs with the k-th image) padded with 0 to WORDS_NUM and sorted by length as
prepare_data does.  The text encoder checkpoint (cfg.TRAIN.NET_E) is not shipped with the reference; without --net-e the
encoder is random-initialised under manual seed 100 like test1.py:170.  Output: uint8 PNG of fine_im[-1]
(trainer_objective.py:153-155, quantised on the device).
"""
from __future__ import annotations

import argparse
import os
import pickle
from typing import List, Tuple

import torch


def load_captions(path: str, words_num: int) -> Tuple[List[List[int]], int]:
    """-> (test captions as index lists, vocabulary size)."""
    with open(path, 'rb') as f:
        x = pickle.load(f)
    if len(x) == 4:                       # [train_captions, test_captions, ixtoword, wordtoix]  (datasets.py:741-745)
        test_caps, ixtoword = x[1], x[2]
    elif len(x) == 3:                     # the shipped data/face/testset.pickle: [captions, ixtoword, wordtoix]
        test_caps, ixtoword = x[0], x[1]
    else:
        raise ValueError('unrecognised caption pickle layout (%d entries)' % len(x))
    return [list(c)[:words_num] for c in test_caps], len(ixtoword)


def preprocess(path: str, imsize: int, lr_size: int) -> torch.Tensor:
    """PNG -> normalised LR tensor [1,3,lr,lr] (datasets.py:1558-1563, :216)."""
    from PIL import Image
    from torchvision import transforms
    img = Image.open(path).convert('RGB')
    img = transforms.Compose([transforms.Resize(int(imsize * 72 / 64)), transforms.CenterCrop(imsize)])(img)
    lr = transforms.Resize([lr_size, lr_size])(img)
    return transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])(lr)[None]


def batch_captions(caps: List[List[int]], words_num: int):
    """Zero-padded [B, words_num] int64, lengths, and the permutation that sorts by length descending
    (prepare_data, miscc/utils-style: the packed LSTM of the reference needs sorted lengths; ours does not, but the
    caption-mask row-order quirk of GlobalAttention.py:109-112 depends on batch order, so the order is kept identical)."""
    lens = torch.tensor([len(c) for c in caps], dtype=torch.int64)
    order = torch.sort(lens, descending=True, stable=True).indices
    out = torch.zeros(len(caps), words_num, dtype=torch.int64)
    for r, i in enumerate(order.tolist()):
        out[r, :len(caps[i])] = torch.tensor(caps[i], dtype=torch.int64)
    return out, lens[order], order


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('images', nargs='+')
    ap.add_argument('--cfg', required=True)
    ap.add_argument('--net-g', required=True, help='netG_*.pth (netGH_*.pth is looked up next to it, trainer_objective.py:93-94)')
    ap.add_argument('--net-e', default='', help='text encoder state dict (cfg.TRAIN.NET_E); random-init under seed 100 if absent')
    ap.add_argument('--captions', required=True)
    ap.add_argument('--out', required=True)
    ap.add_argument('--precision', default='mixed16', choices=['fp32', 'bf16', 'fp16', 'mixed16'])
    ap.add_argument('--device', default='cuda:0')
    a = ap.parse_args(argv)

    import tgsr_b200
    from tgsr_b200 import ops
    from tgsr_b200.config import cfg, cfg_from_file
    from tgsr_b200.pipeline import SRPipeline
    cfg_from_file(a.cfg)
    tgsr_b200.set_precision(a.precision)
    words_num = cfg.TEXT.WORDS_NUM
    caps, n_words = load_captions(a.captions, words_num)
    if len(caps) < len(a.images):
        raise SystemExit('%d images but only %d test captions in %s' % (len(a.images), len(caps), a.captions))
    variant = 'x8' if cfg.TREE.BRANCH_NUM == 4 else 'x16'
    lr_size = cfg.TREE.BASE_SIZE
    imsize = lr_size * (2 ** (cfg.TREE.BRANCH_NUM - 1))
    torch.manual_seed(100)                                               # test1.py:170
    pipe = SRPipeline(n_words, device=a.device, variant=variant, low='lr')
    sd = lambda p: torch.load(p, map_location='cpu', weights_only=False)
    pipe.load_generators(sd(a.net_g), sd(a.net_g.replace('netG', 'netGH')))
    if a.net_e:
        pipe.text_encoder.load_state_dict(sd(a.net_e), strict=True)
    LR = torch.cat([preprocess(p, imsize, lr_size) for p in a.images])
    captions, lens, order = batch_captions(caps[:len(a.images)], words_num)
    LR = LR[order]
    dev = torch.device(a.device)
    fine, _low, _att = pipe.forward_device(LR.to(dev), captions.to(dev), lens.to(dev))
    u8 = ops.quantize_u8(fine[-1]).cpu().numpy()                        # trainer_objective.py:153-155
    os.makedirs(a.out, exist_ok=True)
    from PIL import Image
    for r, i in enumerate(order.tolist()):
        name = os.path.splitext(os.path.basename(a.images[i]))[0]
        dst = os.path.join(a.out, '%s_SR.png' % name)
        Image.fromarray(u8[r].transpose(1, 2, 0)).save(dst)
        print('Save to:', dst)
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
