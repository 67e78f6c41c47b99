"""tcgen05 conv kernel (bf16 blocked layout) vs a bf16-faithful CPU emulation built from the
oracle's fp32 conv: same bf16-rounded operands, fp32 accumulation, so only accumulation order
and the final bf16 rounding differ.  Goes through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

import tgsr_b200

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rounder(fmt):
    dt = torch.float16 if fmt else torch.bfloat16
    return lambda x: x.to(dt).to(torch.float32)


def emulate(x, w, scale, shift, glu, res, up2x, fmt=0):
    bf = rounder(fmt)
    xb = bf(x)
    w = w * scale[:, None, None, None]          # BN scale is folded into the weights before rounding
    if not up2x:
        z = F.conv2d(xb, bf(w), None, 1, 1)
    else:
        # four sub-pixel phases with fp32 pre-summed, bf16-rounded 2x2 weights (csrc/conv_tc.cu pack_up2x_kernel)
        B, C, H, W = x.shape
        z = torch.zeros(B, w.shape[0], 2 * H, 2 * W)
        rows = {0: {0: [0], 1: [1, 2]}, 1: {1: [0, 1], 2: [2]}}
        for py in (0, 1):
            for px in (0, 1):
                w3 = torch.zeros_like(w)
                for r, kys in rows[py].items():
                    for c, kxs in rows[px].items():
                        acc = torch.zeros(w.shape[0], w.shape[1])
                        for ky in kys:
                            for kx in kxs:
                                acc = acc + w[:, :, ky, kx]
                        w3[:, :, r, c] = acc
                z[:, :, py::2, px::2] = F.conv2d(xb, bf(w3), None, 1, 1)
    z = z + shift[None, :, None, None]
    if glu:
        n = z.shape[1] // 2
        z = z[:, :n] * torch.sigmoid(z[:, n:])
    if res is not None:
        z = z + bf(res)
    return z


CASES = [  # Cin, Cout, glu, up2x, residual, B, H, W
    (64, 128, True, False, False, 2, 32, 32),
    (64, 64, False, False, True, 2, 32, 16),
    (32, 64, True, False, False, 1, 16, 8),
    (32, 32, False, False, True, 3, 32, 32),
    (32, 32, False, False, False, 1, 16, 16),
    (64, 64, True, True, False, 2, 16, 16),
    (32, 64, True, True, False, 1, 32, 8),
    (64, 128, True, False, False, 5, 64, 64),     # more tiles than SMs -> pipeline wrap-around
    (64, 64, False, False, True, 5, 64, 64),      # skip-tile ring wrap-around
    (64, 128, True, False, False, 2, 128, 128),   # production resolutions of h_net3 / residual48 (per-kernel, 128x128)
    (64, 64, False, False, True, 2, 128, 128),
    (64, 64, True, True, False, 2, 128, 128),     # -> 256x256 output
    (32, 64, True, False, False, 2, 128, 128),
    (32, 32, False, False, True, 2, 128, 128),
    (32, 64, True, True, False, 2, 128, 128),
    (64, 64, False, False, True, 1, 48, 24),      # tile grid that is not a power of two (3 x 3 tiles)
]


@pytest.mark.parametrize('reverse', [False, True], ids=['fwd', 'rev'])
@pytest.mark.parametrize('fmt', [0, 1], ids=['bf16', 'fp16'])
@pytest.mark.parametrize('Cin,Cout,glu,up2x,use_res,B,H,W', CASES)
def test_conv_tc_matches_16bit_emulation(Cin, Cout, glu, up2x, use_res, B, H, W, fmt, reverse):
    from tgsr_b200 import tc
    from tgsr_b200._lib import ACT_GLU, ACT_NONE
    g = torch.Generator().manual_seed(Cin * 7 + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))
    scale = 1 + 0.1 * torch.randn(Cout, generator=g)
    shift = 0.1 * torch.randn(Cout, generator=g)
    Cf = Cout // 2 if glu else Cout
    res = torch.randn(B, Cf, H, W, generator=g) if use_res else None
    ref = emulate(x, w, scale, shift, glu, res, up2x, fmt)

    conv = torch.nn.Conv2d(Cin, Cout, 3, 1, 1, bias=False).to(DEV)
    conv.weight.data.copy_(w)
    xb = tc.to_blocked(x.to(DEV), fmt=fmt)
    rb = tc.to_blocked(res.to(DEV), fmt=fmt) if use_res else None
    yb = tc.conv3x3_tc(xb, conv, scale.to(DEV), shift.to(DEV), ACT_GLU if glu else ACT_NONE, residual=rb, up2x=up2x,
                       reverse=reverse)
    y = tc.to_nchw(yb).cpu()
    assert y.shape == ref.shape
    err = (y - ref).abs().max().item()
    tol = 2 ** (-11 if fmt else -8) * max(1.0, ref.abs().max().item()) + 1e-4     # one rounding of the output
    assert err <= tol, 'max-abs %.4g > %.4g' % (err, tol)


def test_blocked_roundtrip_and_concat_views():
    from tgsr_b200 import tc
    x = torch.randn(2, 32, 16, 8)
    buf = tc.Blocked.empty(2, 64, 16, 8, DEV)
    buf.buf.zero_()
    tc.to_blocked(x.to(DEV), out=buf.slice(32, 32))       # write into the second channel half (concat)
    back = tc.to_nchw(buf.slice(32, 32)).cpu()
    assert torch.equal(back, rounder(0)(x))
    assert torch.count_nonzero(tc.to_nchw(buf.slice(0, 32))) == 0


def test_conv_tc_rejects_unsupported():
    from tgsr_b200 import tc
    conv = torch.nn.Conv2d(16, 16, 3, 1, 1, bias=False).to(DEV)
    xb = tc.to_blocked(torch.zeros(1, 16, 16, 8, device=DEV))
    with pytest.raises(RuntimeError, match='unsupported'):
        tc.conv3x3_tc(xb, conv, None, None, 0)
