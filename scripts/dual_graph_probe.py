"""Probe: two instances of the captured eval step replayed on two streams (step i+1 starts while step i drains) against one
instance replayed back to back.  Device-resident inputs, no L2 flush (the `sustained` protocol of bench.py)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tgsr_b200
from tgsr_b200 import synth as O
from tgsr_b200.pipeline import SRPipeline

dev = torch.device('cuda')
tgsr_b200.set_eval_defaults(); tgsr_b200.set_precision(tgsr_b200.DEFAULT_BENCH_PRECISION)
B, NT, W = 32, 5450, 18
pipes = [SRPipeline(NT, device=dev, variant='x8', low='lr', seed=100, randomize_bn=True) for _ in range(2)]
LR = O.synth_lr(B, 32, seed=101).to(dev)
caps, lens = O.synth_captions(B, NT, seed=102, words_num=W)
caps, lens = caps.to(dev), lens.to(dev)
ents = []
for p in pipes:
    p.forward_graph(LR, caps, lens, W)
    ents.append(list(p._graphs.values())[0])
torch.cuda.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run(n, dual):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    if not dual:
        for _ in range(n):
            ents[0]['graph'].replay()
    else:
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for i in range(n):
            with torch.cuda.stream(streams[i & 1]):
                ents[i & 1]['graph'].replay()
        for st in streams:
            cur.wait_stream(st)
    e.record(); e.synchronize()
    return B * n / (s.elapsed_time(e) * 1e-3)


for _ in range(2):
    print('single %.0f img/s   dual %.0f img/s' % (run(2000, False), run(2000, True)))
