"""Host side of the bf16 tensor-core path: blocked-plane activation views, cached packed weights
and the tcgen05 conv launcher.  All arithmetic is in libtgsr_sm100.so (csrc/conv_tc.cu)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from ._lib import ACT_GLU, ACT_NONE, FMT_BF16, FMT_F16, call, ptr, tensor_key

Tensor = torch.Tensor


_DTYPES = {FMT_BF16: torch.bfloat16, FMT_F16: torch.float16}


class Blocked:
    """A view of C channels inside a blocked 16-bit buffer [B, P, H, W, 8] starting at plane p0."""
    __slots__ = ('buf', 'p0', 'C')

    def __init__(self, buf: Tensor, p0: int, C: int):
        assert buf.dtype in (torch.bfloat16, torch.float16) and buf.dim() == 5 and buf.shape[4] == 8 and buf.is_contiguous()
        assert C % 8 == 0 and 0 <= p0 and p0 + C // 8 <= buf.shape[1]
        self.buf, self.p0, self.C = buf, p0, C

    @staticmethod
    def empty(B: int, C: int, H: int, W: int, device, planes: Optional[int] = None, p0: int = 0,
              fmt: int = FMT_BF16) -> 'Blocked':
        P = planes if planes is not None else C // 8
        return Blocked(torch.empty((B, P, H, W, 8), device=device, dtype=_DTYPES[fmt]), p0, C)

    @property
    def fmt(self): return FMT_F16 if self.buf.dtype == torch.float16 else FMT_BF16

    @property
    def B(self): return self.buf.shape[0]
    @property
    def P(self): return self.buf.shape[1]
    @property
    def H(self): return self.buf.shape[2]
    @property
    def W(self): return self.buf.shape[3]
    @property
    def device(self): return self.buf.device

    def slice(self, c0: int, C: int) -> 'Blocked':
        assert c0 % 8 == 0
        return Blocked(self.buf, self.p0 + c0 // 8, C)

    def batch(self, b0: int, b1: int) -> 'Blocked':
        """Images b0..b1-1 of this view (a contiguous slice of the buffer)."""
        return Blocked(self.buf[b0:b1], self.p0, self.C)


def to_blocked(x: Tensor, out: Optional[Blocked] = None, fmt: int = FMT_BF16) -> Blocked:
    """NCHW fp32 -> blocked 16-bit (round-to-nearest-even)."""
    x = x.contiguous()
    B, C, H, W = x.shape
    if out is None:
        out = Blocked.empty(B, C, H, W, x.device, fmt=fmt)
    call('tgsr_nchw_to_blocked', x.device, ptr(x), ptr(out.buf), B, C, H, W, out.P, out.p0, out.fmt)
    return out


def to_nchw(xb: Blocked) -> Tensor:
    y = torch.empty((xb.B, xb.C, xb.H, xb.W), device=xb.device, dtype=torch.float32)
    call('tgsr_blocked_to_nchw', xb.device, ptr(xb.buf), ptr(y), xb.B, xb.C, xb.H, xb.W, xb.P, xb.p0, xb.fmt)
    return y


def packed_weights(conv: nn.Conv2d, up2x: bool, fmt: int, scale: Optional[Tensor] = None, shift: Optional[Tensor] = None,
                   act: int = ACT_NONE) -> Tensor:
    """16-bit UMMA operand image of a conv3x3 weight with the folded BatchNorm inside: the scale is multiplied into
    the weights (fp32, one rounding), the shift becomes the operand of the per-tile bias MMA; cached per weight / BN
    version (not in state_dict)."""
    w = conv.weight
    key = (tensor_key(w), bool(up2x), fmt, act, None if scale is None else scale.data_ptr(),
           None if shift is None else shift.data_ptr())
    name = '_tgsr_packed_%d_%d_%d' % (int(up2x), fmt, act)
    cached = conv.__dict__.get(name)
    if cached is None or cached[0] != key:
        Cout, Cin = w.shape[0], w.shape[1]
        from . import _lib
        n = _lib.load().tgsr_packed_conv3x3_elems(Cout, Cin, int(up2x))
        packed = torch.empty(n, device=w.device, dtype=_DTYPES[fmt])
        call('tgsr_pack_conv3x3', w.device, ptr(w.detach().contiguous()), ptr(scale), ptr(shift), ptr(packed), Cout, Cin,
             int(up2x), act, fmt)
        cached = (key, packed, scale, shift)      # keep scale / shift alive so their data_ptr stays unique
        conv.__dict__[name] = cached
    return cached[1]


TC_SHAPES = {(64, 128, ACT_GLU, False), (64, 64, ACT_NONE, False), (32, 64, ACT_GLU, False),
             (32, 32, ACT_NONE, False), (64, 64, ACT_GLU, True), (32, 64, ACT_GLU, True)}


def tc_supported(Cin: int, Cout: int, act: int, up2x: bool, H: int, W: int) -> bool:
    return (Cin, Cout, act, bool(up2x)) in TC_SHAPES and H % 16 == 0 and W % 8 == 0


def conv3x3_tc(xb: Blocked, conv: nn.Conv2d, scale: Optional[Tensor], shift: Optional[Tensor], act: int,
               residual: Optional[Blocked] = None, up2x: bool = False, out: Optional[Blocked] = None,
               reverse: bool = False, l2_flags: int = 0) -> Blocked:
    Cout, Cin = conv.weight.shape[0], conv.weight.shape[1]
    assert Cin == xb.C, (Cin, xb.C)
    Cf = Cout // 2 if act == ACT_GLU else Cout
    Ho, Wo = (2 * xb.H, 2 * xb.W) if up2x else (xb.H, xb.W)
    if out is None:
        out = Blocked.empty(xb.B, Cf, Ho, Wo, xb.device, fmt=xb.fmt)
    assert out.fmt == xb.fmt and out.C == Cf and out.H == Ho and out.W == Wo and out.B == xb.B
    wp = packed_weights(conv, up2x, xb.fmt, scale, shift, act)
    rbuf, rP, rp0 = (residual.buf, residual.P, residual.p0) if residual is not None else (None, 0, 0)
    if residual is not None:
        assert residual.C == Cf and residual.H == Ho and residual.W == Wo and residual.fmt == xb.fmt
    call('tgsr_conv3x3_tc', xb.device, ptr(xb.buf), xb.P, xb.p0, ptr(wp), ptr(rbuf), rP, rp0,
         ptr(out.buf), out.P, out.p0, xb.B, Cin, xb.H, xb.W, Cout, int(up2x), act, xb.fmt, int(reverse) | l2_flags)
    return out
