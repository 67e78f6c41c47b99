"""Launch individual hot kernels at bench shapes (B=32) for ncu captures and quick CUDA-event timing.

    python scripts/probe_kernels.py conv128 | conv64res | up64 | head5 | head3 | attn | lstm  [--iters N]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tgsr_b200
from tgsr_b200 import fast, tc
from tgsr_b200._lib import ACT_GLU, ACT_NONE
from tgsr_b200.pipeline import SRPipeline


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('what')
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--flush', default='write', choices=['write', 'read', 'none'],
                    help='L2 state before each timed launch: 256 MiB memset (dirty lines), 256 MiB read (clean lines), nothing (warm)')
    a = ap.parse_args()
    tgsr_b200.set_eval_defaults(); tgsr_b200.set_precision('mixed16')
    dev = torch.device('cuda')
    pipe = SRPipeline(5450, device=dev, seed=100, randomize_bn=True)
    B, S = a.batch, a.size
    st = pipe.netGL.h_net3
    rb = st.residual[0].block
    flops = None
    nbytes = None
    if a.what == 'conv128':
        x = tc.to_blocked(torch.randn(B, 64, S, S, device=dev) * 0.5)
        out = tc.Blocked.empty(B, 64, S, S, dev)
        fn = lambda: fast.conv_bn_act(x, rb[0], rb[1], ACT_GLU, out=out)
        flops = 2 * 9 * 64 * 128 * S * S * B
    elif a.what == 'conv64res':
        x = tc.to_blocked(torch.randn(B, 64, S, S, device=dev) * 0.5)
        r = tc.to_blocked(torch.randn(B, 64, S, S, device=dev) * 0.5)
        out = tc.Blocked.empty(B, 64, S, S, dev)
        fn = lambda: fast.conv_bn_act(x, rb[3], rb[4], ACT_NONE, residual=r, out=out)
        flops = 2 * 9 * 64 * 64 * S * S * B
    elif a.what == 'up64':
        x = tc.to_blocked(torch.randn(B, 64, S, S, device=dev) * 0.5)
        out = tc.Blocked.empty(B, 32, 2 * S, 2 * S, dev)
        fn = lambda: fast.conv_bn_act(x, st.upsample[1], st.upsample[2], ACT_GLU, up2x=True, out=out)
        flops = 2 * 9 * 64 * 64 * 4 * S * S * B          # reference FLOPs (executed: 4/9 of it)
    elif a.what in ('body32', 'res32', 'up32'):      # NetG_high layers (fp16 operands) at 128x128
        gh = pipe.netGH
        hb = gh.residual48
        x = tc.to_blocked(torch.randn(B, 32, S, S, device=dev) * 0.5, fmt=1)
        r = tc.to_blocked(torch.randn(B, 32, S, S, device=dev) * 0.5, fmt=1)
        if a.what == 'body32':
            out = tc.Blocked.empty(B, 32, S, S, dev, fmt=1)
            fn = lambda: fast.conv_bn_act(x, hb[0], hb[1], ACT_GLU, out=out)
            flops = 2 * 9 * 32 * 64 * S * S * B
        elif a.what == 'res32':
            out = tc.Blocked.empty(B, 32, S, S, dev, fmt=1)
            fn = lambda: fast.conv_bn_act(x, hb[3], hb[4], ACT_NONE, residual=r, out=out)
            flops = 2 * 9 * 32 * 32 * S * S * B
        else:
            out = tc.Blocked.empty(B, 32, 2 * S, 2 * S, dev, fmt=1)
            fn = lambda: fast.conv_bn_act(x, gh.upscale8x[1], gh.upscale8x[2], ACT_GLU, up2x=True, out=out)
            flops = 2 * 9 * 32 * 64 * 4 * S * S * B
    elif a.what in ('head5', 'head3'):
        x = tc.to_blocked(torch.randn(B, 32, 2 * S, 2 * S, device=dev) * 0.5, fmt=1)
        seq = pipe.netGH.conv_output if a.what == 'head5' else pipe.netGL.img_net3.img
        add = torch.randn(B, 3, 2 * S, 2 * S, device=dev)
        fn = lambda: fast.image_head(x, seq, add, 0.5, 1.0)
        k = 5 if a.what == 'head5' else 3
        flops = 2 * k * k * 32 * 3 * 4 * S * S * B
    elif a.what in ('attn', 'attnproj'):
        # attn: the per-pixel kernel alone (word projection precomputed, as inside a forward);
        # attnproj: projection + per-pixel kernel (a stand-alone GlobalAttentionGeneral call)
        cat = tc.Blocked.empty(B, 64, S, S, dev); cat.buf.normal_(0, 0.5)
        ctx = torch.randn(B, 256, 18, device=dev) * 0.3
        mask = torch.zeros(B, 18, dtype=torch.uint8, device=dev)
        sfrag = fast.attention_project(st.att, ctx, cat.fmt) if a.what == 'attn' else None
        fn = lambda: fast.attention(cat.slice(0, 32), st.att, ctx, mask, out=cat.slice(32, 32), sfrag=sfrag)
        nbytes = B * S * S * (32 * 2 + 32 * 2 + 4 * 18) + B * 256 * 18 * 4 + 32 * 256 * 4 + B * 18
    elif a.what == 'attnf32':                   # the module-API attention (fp32 NCHW I/O, 3xTF32 tensor-core kernel)
        x = torch.randn(B, 32, S, S, device=dev) * 0.5
        ctx = torch.randn(B, 256, 18, device=dev) * 0.3
        tgsr_b200.set_precision('fp32')
        st.att.applyMask(torch.zeros(B, 18, dtype=torch.uint8, device=dev))
        fn = lambda: st.att(x, ctx)
        nbytes = B * S * S * (32 * 4 + 32 * 4 + 4 * 18) + B * 256 * 18 * 4 + 32 * 256 * 4 + B * 18
    elif a.what in ('convtf32', 'wgradtf32'):   # training convs: TF32 tensor-core forward / weight gradient, 64 -> 128
        from tgsr_b200 import train
        train.set_conv_precision('tf32')
        x = torch.randn(B, 64, S, S, device=dev) * 0.5
        w = torch.randn(128, 64, 3, 3, device=dev) * 0.05
        dy = torch.randn(B, 128, S, S, device=dev) * 0.5
        fn = (lambda: train._conv_raw(x, w, False)) if a.what == 'convtf32' else (lambda: train._conv_wgrad(x, dy, w.shape, False))
        flops = 2 * 9 * 64 * 128 * S * S * B
    elif a.what == 'null':                      # calibration: a trivially small kernel in the same harness
        from tgsr_b200 import ops
        v = torch.ones(64, device=dev)
        fn = lambda: ops.bn_fold(v, v, v, v, 1e-5)
    elif a.what == 'lstm':
        caps = torch.randint(1, 5450, (B, 18), device=dev); lens = torch.full((B,), 18, device=dev)
        fn = lambda: pipe.text_encoder._forward_tmax(caps, lens, 18)
    else:
        raise SystemExit('unknown kernel ' + a.what)
    # Device time without launch latency: two CUDA graphs, N x [L2 flush, op] and N x [L2 flush]; the
    # difference of their CUDA-event durations / N is the L2-cold duration of the op.
    flush = torch.empty(64 * 1024 * 1024, device=dev)          # 256 MiB > 126 MB L2
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    N = a.iters
    g_both, g_flush = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    sink = torch.zeros(1, device=dev)

    def do_flush():
        if a.flush == 'write':
            flush.zero_()
        elif a.flush == 'read':
            torch.sum(flush, dim=0, keepdim=True, out=sink)
    do_flush(); torch.cuda.synchronize()
    with torch.cuda.graph(g_both):
        for _ in range(N):
            do_flush(); fn()
    with torch.cuda.graph(g_flush):
        for _ in range(N):
            do_flush()

    def timed(g):
        best = 1e9
        for _ in range(3):
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record(); g.replay(); e.record(); e.synchronize()
            best = min(best, s.elapsed_time(e))
        return best
    timed(g_both); timed(g_flush)
    ms = (timed(g_both) - timed(g_flush)) / N
    msg = '%s B=%d S=%d: %.1f us/launch' % (a.what, B, S, ms * 1e3)
    if flops:
        msg += '  %.1f TFLOP/s (reference FLOPs)' % (flops / ms / 1e9)
    if nbytes:
        msg += '  %.0f GB/s (algorithmic bytes)' % (nbytes / ms / 1e6)
    print(msg)


if __name__ == '__main__':
    main()
