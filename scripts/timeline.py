"""Kernel timeline of one CUDA-graph replay of the bench step (torch.profiler / CUPTI activity records).

Unlike an ncu launch list (serialised, cold), this is the step as bench.py times it: two captured streams,
programmatic dependent launch, warm clocks.  Prints start offset, duration and stream of every kernel of one replay and
the gaps between consecutive kernels on the critical stream.

    python scripts/timeline.py [--batch 32] [--precision mixed16] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tgsr_b200
from tgsr_b200 import synth as O
from tgsr_b200.pipeline import SRPipeline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--precision', default='mixed16')
    ap.add_argument('--variant', default='x8')
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    tgsr_b200.set_eval_defaults(); tgsr_b200.set_precision(a.precision)
    dev = torch.device('cuda')
    pipe = SRPipeline(5450, device=dev, variant=a.variant, seed=100, randomize_bn=True)
    B = a.batch
    size = 32 if a.variant == 'x8' else 16
    LR = O.synth_lr(B, size, seed=101).to(dev)
    caps, lens = O.synth_captions(B, 5450, seed=102, words_num=18)
    caps, lens = caps.to(dev), lens.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    pipe.forward_graph(LR, caps, lens, 18)
    LR, caps, lens = pipe.static_inputs(B, 18, 18, size)
    for _ in range(5):
        pipe.forward_graph(LR, caps, lens, 18)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            flush.zero_()
            pipe.forward_graph(LR, caps, lens, 18)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into replays at the big memset (flush)
    idx = [i for i, e in enumerate(evs) if 'fill' in e.name.lower() or 'memset' in e.name.lower()]
    start = idx[-1] + 1 if idx else 0
    step = evs[start:]
    t0 = step[0].time_range.start
    rows = []
    for e in step:
        rows.append({'name': e.name[:70], 'start_us': e.time_range.start - t0, 'dur_us': e.time_range.end - e.time_range.start,
                     'stream': getattr(e, 'stream', None) if hasattr(e, 'stream') else None})
    end = max(r['start_us'] + r['dur_us'] for r in rows)
    print('kernels %d, span %.1f us, sum of durations %.1f us' % (len(rows), end, sum(r['dur_us'] for r in rows)))
    for r in rows:
        print('%8.1f %7.1f  %s' % (r['start_us'], r['dur_us'], r['name']))
    if a.out:
        json.dump(rows, open(a.out, 'w'))


if __name__ == '__main__':
    main()
