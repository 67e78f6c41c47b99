import sys; sys.path.insert(0, '.')
import torch, tgsr_b200
from tgsr_b200 import _lib
from tgsr_b200.text_encoder import RNN_ENCODER
from tgsr_b200 import synth as O
dev = torch.device('cuda')
tgsr_b200.set_eval_defaults()
enc = RNN_ENCODER(5450, nhidden=256).to(dev).eval()
caps, lens = O.synth_captions(32, 5450, seed=202, words_num=18)
caps, lens = caps.to(dev), lens.to(dev)
lib = _lib.load()
for i in range(4):
    n0 = lib.tgsr_launch_count()
    with torch.no_grad():
        enc._forward_tmax(caps, lens, 18)
    print('call', i, 'launches', lib.tgsr_launch_count() - n0)
    if i == 1:
        _lib.bump_weights_epoch(); print('bumped weights epoch')
    if i == 2:
        for p in enc.parameters(): p.requires_grad = False
