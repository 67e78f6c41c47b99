"""Attention microbenchmark sweep (BASELINE.json configs[2], SURVEY.md 8d cfg 3).

    python scripts/bench_attention.py [--quick] [--out profiles/attention_sweep.jsonl]

GlobalAttentionGeneral(idf, cdf): Q in {32^2..256^2} x T in {18,32,48} x idf=cdf in {32,64,128,256} plus the production
point (idf 32, cdf 256); batch chosen so that at least 256 MB move.  Variants: fp32 NCHW I/O (the reference API) and
16-bit blocked activations (the pipeline's internal path, idf 32).  func_attention at the DAMSM shape.
Every line: algorithmic bytes / L2-cold device time, as a fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tgsr_b200
from tgsr_b200 import fast, ops, tc
from tgsr_b200.timing import cold_time_ms, make_flush


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    return (json.load(open(p))['hbm_gbs'], 'measured') if os.path.isfile(p) else (6650.0, 'fallback')


def attn_bytes(B, Q, idf, cdf, T, e):
    return B * Q * (idf * e + idf * e + 4 * T) + B * cdf * T * 4 + idf * cdf * 4 + B * T


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--out', default=None)
    ap.add_argument('--min-mb', type=float, default=256.0)
    a = ap.parse_args()
    dev = torch.device('cuda')
    peak, src = hbm_peak()
    flush = make_flush(dev)
    g = torch.Generator(device='cpu').manual_seed(103)
    rows = []
    sizes = [32, 64, 128, 256]
    Ts = [18, 32, 48]
    dims = [(32, 256), (32, 32), (64, 64), (128, 128), (256, 256)]
    if a.quick:
        sizes, Ts, dims = [64, 128], [18], [(32, 256)]
    for idf, cdf in dims:
        w = (torch.randn(idf, cdf, generator=g) * 0.1).to(dev)
        att = tgsr_b200.GlobalAttentionGeneral(idf, cdf).to(dev)
        att.conv_context.weight.data.copy_(w.view(idf, cdf, 1, 1))
        for T in Ts:
            for S in sizes:
                Q = S * S
                for variant in ('fp32', 'blocked16'):
                    e = 4 if variant == 'fp32' else 2
                    if variant == 'blocked16' and not fast.attention_blocked_supported(idf, T, Q):
                        continue
                    per_img = attn_bytes(1, Q, idf, cdf, T, e)
                    B = max(1, min(4096, int(-(-a.min_mb * 1e6 // per_img))))
                    B = min(B, int(6e9 // (Q * (2 * idf * 4 + 4 * T))) or 1)       # keep allocations < ~6 GB
                    ctx = (torch.randn(B, cdf, T, generator=g) * 0.3).to(dev)
                    mask = torch.zeros(B, T, dtype=torch.uint8, device=dev)
                    if variant == 'fp32':
                        x = torch.randn(B, idf, S, S, device=dev) * 0.5
                        att.applyMask(mask)
                        fn = lambda: att(x, ctx)
                        kernel = 'attn_fwd_f32_kernel'
                    else:
                        cat = tc.Blocked.empty(B, 2 * idf, S, S, dev); cat.buf.normal_(0, 0.5)
                        fn = lambda: fast.attention(cat.slice(0, idf), att, ctx, mask, out=cat.slice(idf, idf))
                        kernel = 'attn_project_kernel + attn_blocked_kernel'
                    ms = cold_time_ms(fn, flush, iters=10)
                    nbytes = attn_bytes(B, Q, idf, cdf, T, e)
                    r = {'op': 'GlobalAttentionGeneral', 'variant': variant, 'kernel': kernel, 'B': B, 'Q': Q, 'T': T, 'idf': idf,
                         'cdf': cdf, 'bytes': nbytes, 'us': ms * 1e3, 'GBps': nbytes / ms / 1e6, 'frac_of_hbm_peak': nbytes / ms / 1e6 / peak,
                         'peak_GBps': peak, 'peak_source': src}
                    rows.append(r)
                    print(json.dumps(r), flush=True)
                    del ctx, mask
                    if variant == 'fp32':
                        del x
                    else:
                        del cat
                    torch.cuda.empty_cache()
    # func_attention (GlobalAttention.py:33-74) at the DAMSM shape: query [B,256,T], context [B,256,17,17], gamma1 4
    for B in ([32] if a.quick else [32, 256, 2048]):
        for T in ([18] if a.quick else [18, 32, 48]):
            q = (torch.randn(B, 256, T, generator=g)).to(dev); c = torch.randn(B, 256, 17, 17, device=dev)
            fn = lambda: ops.func_attention_fwd(q, c, 4.0)
            ms = cold_time_ms(fn, flush, iters=10)
            nbytes = B * (256 * 289 + 2 * 256 * T + T * 289) * 4
            r = {'op': 'func_attention', 'variant': 'fp32', 'kernel': 'damsm_pair_fwd_kernel (diagonal pairs)', 'B': B, 'Q': 289, 'T': T, 'idf': 256, 'cdf': 256,
                 'bytes': nbytes, 'us': ms * 1e3, 'GBps': nbytes / ms / 1e6, 'frac_of_hbm_peak': nbytes / ms / 1e6 / peak, 'peak_GBps': peak,
                 'peak_source': src}
            rows.append(r)
            print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, 'w') as f:
            for r in rows:
                f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
