"""x16 eval flow (models16.py: 16x16 -> 256x256, weight-shared stages) on one GPU.

    python scripts/bench_x16.py [--steps 20] [--warmup 5] [--batch 32] [--out profiles/r2_bench_x16.json]

Same timing rules as bench.py: one CUDA-graph replay per step, CUDA events, 256 MiB L2 flush between timed steps; `e2e`
goes through SRPipeline.submit / result with pinned host buffers.  Not a headline (BASELINE.json quotes the x8 flow): the
record shows that the second generator family runs on the same tcgen05 kernels at the same efficiency.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tgsr_b200
from tgsr_b200 import synth as O
from tgsr_b200.pipeline import SRPipeline

NTOKEN, WORDS = 5450, 18


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit('bench_x16.py: no CUDA device -- the product path has no CPU fallback')
    dev = torch.device('cuda', 0)
    tgsr_b200.set_eval_defaults()
    tgsr_b200.set_precision(tgsr_b200.DEFAULT_BENCH_PRECISION)
    B = a.batch
    pipe = SRPipeline(NTOKEN, device=dev, variant='x16', low='lr', seed=100, randomize_bn=True)
    packed_h = torch.empty(SRPipeline.packed_nbytes(B, WORDS, 16), dtype=torch.uint8).pin_memory()
    LR_h, caps_h, lens_h = SRPipeline.packed_views(packed_h, B, WORDS, 16)
    LR_h.copy_(O.synth_lr(B, 16, seed=101))
    c_, l_ = O.synth_captions(B, NTOKEN, seed=102, words_num=WORDS)
    caps_h.copy_(c_); lens_h.copy_(l_)
    pipe.forward_graph(LR_h.to(dev), caps_h.to(dev), lens_h.to(dev), WORDS)
    LR, caps, lens = pipe.static_inputs(B, WORDS, WORDS, 16)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)
    step = lambda: pipe.forward_graph(LR, caps, lens, WORDS)[0]
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for s, e in ev:
        flush.zero_()
        s.record(); step(); e.record()
    torch.cuda.synchronize()
    t_ms = sum(s.elapsed_time(e) for s, e in ev)

    def loop(n):
        prev = None
        for _ in range(n):
            tk = pipe.submit(LR_h, caps_h, lens_h, out='uint8', packed=packed_h)
            if prev is not None:
                pipe.result(prev)
            prev = tk
        return pipe.result(prev)

    loop(3)
    torch.cuda.synchronize()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record(); res = loop(a.steps); e2.record()
    torch.cuda.synchronize()
    line = {'metric': 'SR images/sec (eval fwd, x16 16->256, batch %d/GPU)' % B, 'value': B * a.steps / (t_ms * 1e-3),
            'unit': 'images/s', 'n_gpus': 1, 'steps': a.steps, 'warmup': max(a.warmup, 3), 'ms_per_step': t_ms / a.steps,
            'higher_is_better': True, 'dtype': 'bf16+f16 (16-bit tcgen05 operands, f32 accumulate)', 'data': 'synthetic',
            'config': {'workload': 'models16.py nets at GF_DIM %d: text encoder + G_SR_NET_low (4 weight-shared stages, 16->256) + '
                                   'NetG_highweight (16->256), batch %d, 18-word captions, ntoken %d, random-init, randomised BN stats'
                                   % (tgsr_b200.cfg.GAN.GF_DIM, B, NTOKEN),
                       'l2': 'flushed between timed iterations (256 MiB write)', 'launch': 'CUDA-graph replay'},
            'gpu_launches': pipe.last_graph_kernels * a.steps,
            'e2e': {'value': B * a.steps / (s2.elapsed_time(e2) * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': packed_h.numel(),
                    'd2h_bytes_per_step': res.numel() * res.element_size()}}
    print(json.dumps(line))
    if a.out:
        with open(os.path.join(ROOT, a.out), 'w') as f:
            json.dump(line, f, indent=1)


if __name__ == '__main__':
    main()
