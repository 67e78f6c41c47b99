"""Device-time measurement shared by bench.py and the probe scripts.

`cold_time_ms(fn, flush)` returns the L2-cold duration of one enqueue of `fn` without launch latency or
Python overhead: two CUDA graphs -- N x [L2-flushing memset, fn] and N x [memset] -- are replayed between
CUDA events on the launching stream and their difference is divided by N."""
from __future__ import annotations

import torch


def make_flush(device, mib: int = 256) -> torch.Tensor:
    """A buffer larger than the 126 MB L2; `flush.zero_()` evicts every cached line of the op under test."""
    return torch.empty(mib * 1024 * 1024 // 4, device=device, dtype=torch.float32)


def _best_replay_ms(g: torch.cuda.CUDAGraph, reps: int) -> float:
    best = float('inf')
    for _ in range(reps):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); e.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def cold_time_ms(fn, flush: torch.Tensor, iters: int = 20, warm: int = 3, reps: int = 4) -> float:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g_both, g_flush = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_both):
        for _ in range(iters):
            flush.zero_(); fn()
    with torch.cuda.graph(g_flush):
        for _ in range(iters):
            flush.zero_()
    _best_replay_ms(g_both, 1); _best_replay_ms(g_flush, 1)
    return (_best_replay_ms(g_both, reps) - _best_replay_ms(g_flush, reps)) / iters
