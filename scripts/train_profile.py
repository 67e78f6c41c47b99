"""Per-kernel time of one generator training step (torch.profiler / CUPTI activity records, eager launches, warm clocks).

    python scripts/train_profile.py [--precision bf16] [--batch 32]
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tgsr_b200
from tgsr_b200 import nets, synth as O
from tgsr_b200.text_encoder import RNN_ENCODER
from tgsr_b200.train_step import FrozenImageEncoder, GeneratorTrainStep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--batch', type=int, default=32)
    a = ap.parse_args()
    dev = torch.device('cuda')
    tgsr_b200.set_eval_defaults()
    from tgsr_b200.config import cfg
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3 = 4.0, 5.0, 10.0
    torch.manual_seed(100)
    gl = nets.G_SR_NET_low().to(dev); gh = nets.NetG_highweight(low='lr').to(dev)
    enc = RNN_ENCODER(5450, nhidden=256).to(dev).eval()
    img = FrozenImageEncoder().to(dev)
    ts = GeneratorTrainStep(gl, gh, enc, img, conv_precision=a.precision)
    B = a.batch
    LR = O.synth_lr(B, 32, seed=201).to(dev)
    real = [torch.nn.functional.interpolate(LR, scale_factor=2 ** (k + 1), mode='bicubic').clamp(-1, 1) for k in range(3)]
    caps, lens = O.synth_captions(B, 5450, seed=202, words_num=18)
    caps, lens = caps.to(dev), lens.to(dev)
    for _ in range(3):
        ts.step(LR, real, caps, lens, 18)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts.step(LR, real, caps, lens, 18)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    each = collections.defaultdict(list)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name.replace('(anonymous namespace)::', '').split('(')[0][:130]
            agg[k][0] += 1
            agg[k][1] += e.time_range.end - e.time_range.start
            each[k].append(e.time_range.end - e.time_range.start)
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if 'tgsr::' in k)
    print('kernels %d, sum of durations %.2f ms (tgsr:: kernels %.2f ms)' % (sum(v[0] for v in agg.values()), tot / 1e3, ours / 1e3))
    print('| ms | share | launches | kernel |\n|---:|---:|---:|---|')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print('| %.2f | %.1f%% | %d | `%s` |' % (t / 1e3, 100 * t / tot, c, k))
    print('\nlargest launches (us) of this library\'s kernels:')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        if 'tgsr::' in k:
            print('%-60s %s' % (k[:60], ' '.join('%.0f' % v for v in sorted(each[k], reverse=True)[:8])))


if __name__ == '__main__':
    main()
