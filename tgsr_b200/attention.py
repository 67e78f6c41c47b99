"""Word-region attention modules with the reference's API (GlobalAttention.py).

``GlobalAttentionGeneral(idf, cdf)`` keeps the ``conv_context.weight`` state-dict key, the
stateful ``applyMask`` (GlobalAttention.py:84-85) and ``forward(input, context)`` ->
``(weightedContext, attn)``; the arithmetic is one fused sm_100a kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import no_grad_path


def conv1x1(in_planes, out_planes, bias=False):
    """Parameter holder for the 1x1 word projection (GlobalAttention.py:28-30 / util.py:56-59)."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=1, padding=0, bias=bias)


def func_attention(query, context, gamma1):
    """GlobalAttention.py:33-74.  query [B,ndf,Tq], context [B,ndf,ih,iw] -> (weightedContext, attn).  Forward only
    (the training step reaches it through the fused words_loss kernels, which carry the image-feature gradient)."""
    no_grad_path('func_attention', query, context)
    return ops.func_attention_fwd(query, context, gamma1)


class GlobalAttentionGeneral(nn.Module):
    """GlobalAttention.py:77-130."""

    def __init__(self, idf, cdf):
        super().__init__()
        self.conv_context = conv1x1(cdf, idf)
        self.sm = nn.Softmax(dim=1)      # kept for attribute parity; the softmax runs inside the kernel
        self.mask = None

    def applyMask(self, mask):
        self.mask = mask                 # batch x sourceL

    def forward(self, input, context):
        """input: batch x idf x ih x iw; context: batch x cdf x sourceL."""
        w = self.conv_context.weight
        if self.training and torch.is_grad_enabled() and (input.requires_grad or w.requires_grad):
            from .train import AttentionFn
            no_grad_path('GlobalAttentionGeneral(context)', context)     # word embeddings come from the frozen text encoder
            return AttentionFn.apply(input, w, context, self.mask)
        return ops.attention_fwd(input, context, w.detach(), self.mask)
