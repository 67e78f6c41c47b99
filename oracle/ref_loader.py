"""Import the UNMODIFIED reference modules from /root/reference (authoring container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing under tests/ -m gpu, smoke() or bench.py may call this; it is used by
oracle/make_golden.py (fixture generation) and by the CPU-only cross-check in
tests/test_oracle_golden.py when the reference happens to be mounted.
"""
from __future__ import annotations

import contextlib
import os
import sys

REF_ROOT = os.environ.get('TGSR_REFERENCE', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_shims')

# cfg/eval_faceSR_attn2.yml:5-7,24-31 (== eval_birdSR_attn2.yml in every model dimension)
EVAL_CFG = dict(GF_DIM=32, DF_DIM=64, Z_DIM=100, R_NUM=2, EMBEDDING_DIM=256, BRANCH_NUM=4, BASE_SIZE=32)


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, 'GlobalAttention.py'))


def load():
    """Returns a namespace with the reference modules, cfg set to the eval YAML values, CUDA off."""
    if not available():
        raise RuntimeError('reference tree not mounted at %s' % REF_ROOT)
    for p in (_SHIMS, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import GlobalAttention, util, model, models16          # noqa: E401
    from miscc.config import cfg
    cfg.CUDA = False
    cfg.GAN.GF_DIM = EVAL_CFG['GF_DIM']; cfg.GAN.DF_DIM = EVAL_CFG['DF_DIM']
    cfg.GAN.Z_DIM = EVAL_CFG['Z_DIM']; cfg.GAN.R_NUM = EVAL_CFG['R_NUM']
    cfg.TEXT.EMBEDDING_DIM = EVAL_CFG['EMBEDDING_DIM']
    cfg.TREE.BRANCH_NUM = EVAL_CFG['BRANCH_NUM']; cfg.TREE.BASE_SIZE = EVAL_CFG['BASE_SIZE']
    GlobalAttention.server = 1     # bool masks (uint8 masked_fill_ is a RuntimeError on torch 2.x)

    class NS:
        pass
    ns = NS()
    ns.GlobalAttention, ns.util, ns.model, ns.models16, ns.cfg = GlobalAttention, util, model, models16, cfg
    return ns


@contextlib.contextmanager
def cuda_noop():
    """NetG_highweight.__init__ calls .cuda() unconditionally (model.py:246-248).  On a GPU host
    ``nn.Parameter(...).cuda()`` yields a plain (unregistered) tensor, so ``a`` is the constant
    0.5 and is absent from the state_dict; the CPU stand-in returns a detached copy to match."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self.detach().clone()
    try:
        yield
    finally:
        torch.Tensor.cuda = orig
