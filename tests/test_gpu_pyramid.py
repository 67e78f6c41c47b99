"""Input pyramid on the device (-m gpu): tgsr_b200.pyramid (csrc/pyramid.cu) through the C ABI against
  * the reference's own `get_imgs_blur` output on data/face/000155.png (tests/golden/pyramid.npz), bit-exact;
  * the oracle (oracle/pil_pyramid.py, pinned to Pillow) on seeded random batches, bit-exact, fused and generic kernels;
  * size-independent properties at the bench batch (constant images, aliasing of branch 0, value lattice)."""
import os

import numpy as np
import pytest
import torch

from oracle import pil_pyramid as P
from tgsr_b200 import pyramid as Y

pytestmark = pytest.mark.gpu
DEV = 'cuda'
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'pyramid.npz')


def _rand(B, H, W, C=3, seed=0, smooth=False):
    g = torch.Generator().manual_seed(seed)
    if not smooth:
        return torch.randint(0, 256, (B, H, W, C), generator=g, dtype=torch.uint8)
    low = torch.rand((B, C, max(H // 8, 2), max(W // 8, 2)), generator=g)
    up = torch.nn.functional.interpolate(low, size=(H, W), mode='bicubic', align_corners=False).clamp(0, 1)
    return (up * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def _u8(t):
    return np.round((np.transpose(t, (1, 2, 0)) * 0.5 + 0.5) * 255.0).astype(np.uint8)


def test_golden_reference_pyramid():
    g = np.load(GOLDEN)
    imsize = [int(s) for s in g['imsize']]
    hr = torch.from_numpy(g['hr'])[None].to(DEV)
    ret, bic, retb, bicb = Y.get_imgs_blur(hr, imsize)
    for name, lst in (('ret', ret), ('bic', bic), ('retb', retb), ('bicb', bicb)):
        assert len(lst) == len(imsize)
        assert np.array_equal(lst[0][0].cpu().numpy(), g[name + '0_f32']), name
        for i in range(1, len(imsize)):
            assert tuple(lst[i].shape) == (1, 3, imsize[i], imsize[i])
            assert np.array_equal(_u8(lst[i][0].cpu().numpy()), g['%s%d_u8' % (name, i)]), (name, i)


@pytest.mark.parametrize('imsize', [[32, 64, 128, 256], [16, 32, 64, 128], [64, 128], [128], [12, 24, 48, 96], [40, 80, 160]])
def test_fused_pyramid_vs_oracle(imsize):
    for smooth in (False, True):
        hr = _rand(5, imsize[-1], imsize[-1], seed=3 + len(imsize), smooth=smooth)
        want = P.get_imgs_blur(hr.numpy(), imsize)
        got = Y.get_imgs_blur(hr.to(DEV), imsize)
        for wl, gl in zip(want, got):
            assert len(wl) == len(gl)
            for w, g in zip(wl, gl):
                assert g.dtype == torch.float32 and np.array_equal(w, g.cpu().numpy())
        ret, bic = Y.get_imgs(hr.to(DEV), imsize)
        for w, g in zip(want[0] + want[1], ret + bic):
            assert np.array_equal(w, g.cpu().numpy())
        for w, g in zip(want[0], Y.real_pyramid(hr.to(DEV), imsize)):
            assert np.array_equal(w, g.cpu().numpy())


def test_generic_path_odd_sizes():
    # branch sizes that are not multiples of 4 take the one-op-at-a-time kernels
    imsize = [25, 50, 100]
    hr = _rand(3, 100, 100, seed=11)
    want = P.get_imgs_blur(hr.numpy(), imsize)
    got = Y.get_imgs_blur(hr.to(DEV), imsize)
    for wl, gl in zip(want, got):
        for w, g in zip(wl, gl):
            assert np.array_equal(w, g.cpu().numpy())


@pytest.mark.parametrize('shape,size', [((2, 288, 352, 3), 256), ((2, 97, 131, 3), (50, 77)), ((1, 32, 32, 3), (256, 256)),
                                        ((3, 64, 48, 1), (64, 96)), ((2, 40, 40, 3), (40, 40)), ((1, 5, 9, 3), (1, 3)),
                                        ((1, 512, 512, 3), (32, 32))])
def test_resize_vs_oracle(shape, size):
    a = _rand(*shape, seed=5)
    got = Y.resize(a.to(DEV), size).cpu().numpy()
    oh, ow = Y._size2(size, shape[1], shape[2])
    assert np.array_equal(P.resize_u8(a.numpy(), (oh, ow)), got)


@pytest.mark.parametrize('radius', [2.0, 1.0, 3.5])
def test_gaussian_blur_vs_oracle(radius):
    for shape in [(2, 64, 64, 3), (1, 97, 131, 3), (2, 256, 256, 1)]:
        a = _rand(*shape, seed=9)
        assert np.array_equal(P.gaussian_blur_u8(a.numpy(), radius), Y.gaussian_blur(a.to(DEV), radius).cpu().numpy())


def test_normalize_vs_oracle():
    a = torch.arange(256, dtype=torch.uint8).repeat(3)[: 8 * 8 * 3 * 4].reshape(4, 8, 8, 3).contiguous()
    assert np.array_equal(P.normalize_u8(a.numpy()), Y.normalize(a.to(DEV)).cpu().numpy())


def test_against_pillow_runtime():
    PIL = pytest.importorskip('PIL')
    from PIL import Image, ImageFilter
    a = _rand(1, 256, 256, seed=21, smooth=True)
    im = Image.fromarray(a[0].numpy())
    lr = im.resize((32, 32), Image.BILINEAR)
    assert np.array_equal(np.asarray(lr), Y.resize(a.to(DEV), (32, 32))[0].cpu().numpy())
    assert np.array_equal(np.asarray(lr.resize((128, 128), Image.BILINEAR).filter(ImageFilter.GaussianBlur(radius=2))),
                          Y.gaussian_blur(Y.resize(Y.resize(a.to(DEV), (32, 32)), (128, 128)))[0].cpu().numpy())


def test_properties_full_batch():
    B, imsize = 32, [32, 64, 128, 256]
    hr = _rand(B, 256, 256, seed=1, smooth=True).to(DEV)
    hr[0] = 77                                              # a constant image stays constant through every stage
    ret, bic, retb, bicb = Y.get_imgs_blur(hr, imsize)
    const = (np.float32(77) / np.float32(255) - np.float32(0.5)) / np.float32(0.5)
    for lst in (ret, bic, retb, bicb):
        for t in lst:
            assert t.shape[0] == B and bool((t[0] == float(const)).all())
            q = (t * 0.5 + 0.5) * 255.0                     # every value is one of the 256 normalised byte values
            assert float((q - q.round()).abs().max()) < 1e-3 and float(t.min()) >= -1.0 and float(t.max()) <= 1.0
    assert bic[0] is ret[0] and bicb[0] is retb[0]
    assert torch.equal(ret[3], Y.normalize(hr))             # last branch is the image itself
    # fused and generic kernels agree at the full size
    assert torch.equal(retb[2], Y.normalize(Y.gaussian_blur(Y.resize(hr, (128, 128)))))
    assert torch.equal(bicb[3], Y.normalize(Y.gaussian_blur(Y.resize(Y.resize(hr, (32, 32)), (256, 256)))))
    # shard independence: the first half of the batch alone gives the same bytes
    ret2, _, _, bicb2 = Y.get_imgs_blur(hr[:16].contiguous(), imsize)
    assert torch.equal(ret2[1], ret[1][:16]) and torch.equal(bicb2[3], bicb[3][:16])


def test_errors():
    a = _rand(1, 64, 64)
    with pytest.raises(RuntimeError):
        Y.get_imgs_blur(a, [16, 32, 64])                    # CPU tensor: no CPU path
    with pytest.raises(TypeError):
        Y.resize(a.to(DEV).float(), (32, 32))
    with pytest.raises(ValueError):
        Y.get_imgs_blur(a.to(DEV), [16, 32])                # last branch must be the image itself
    with pytest.raises(ValueError):
        Y.get_imgs_blur(a.to(DEV), [4, 8, 16, 32, 64])
