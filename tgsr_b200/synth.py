"""Synthetic inputs of the benches: seeded, host-side, image-like LR tensors and caption indices.
(Input generation only -- no reference arithmetic; the CPU oracle re-exports these so tests and benches share them.)"""
from __future__ import annotations

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def synth_lr(batch: int, size: int, seed: int, kind: str = 'smooth') -> Tensor:
    """'smooth': U(-1,1) at 4x4, bicubic to size, clamp; 'noise': iid U(-1,1)."""
    g = torch.Generator().manual_seed(seed)
    if kind == 'noise':
        return torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    base = torch.rand(batch, 3, 4, 4, generator=g) * 2 - 1
    return F.interpolate(base, size=(size, size), mode='bicubic', align_corners=False).clamp(-1, 1)


def synth_captions(batch: int, ntoken: int, seed: int, words_num: int = 18, ragged: bool = False):
    """iid tokens in [1,ntoken); lengths all words_num, or sorted-descending ragged 18..5; zeros beyond the length."""
    g = torch.Generator().manual_seed(seed)
    caps = torch.randint(1, ntoken, (batch, words_num), generator=g)
    if ragged:
        lens = torch.randint(5, words_num + 1, (batch,), generator=g)
        lens[0] = words_num
        lens = torch.sort(lens, descending=True).values
    else:
        lens = torch.full((batch,), words_num, dtype=torch.int64)
    for b in range(batch):
        caps[b, int(lens[b]):] = 0
    return caps, lens
