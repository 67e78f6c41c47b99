"""Input pyramid, CPU side: oracle/pil_pyramid.py (the numpy restatement of Pillow's resize / GaussianBlur and of
ToTensor+Normalize) against (1) the output of the reference's own `get_imgs_blur` (datasets.py:151-196) on
data/face/000155.png, stored in tests/golden/pyramid.npz by oracle/make_golden.py pyramid, and (2) Pillow itself at run
time on random images of odd sizes (Pillow is the third-party library the reference calls; pinned 12.2.0 in this image)."""
import os

import numpy as np
import pytest

from oracle import pil_pyramid as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'pyramid.npz')


def _u8(t):
    return np.round((np.transpose(t, (1, 2, 0)) * 0.5 + 0.5) * 255.0).astype(np.uint8)


def test_oracle_matches_reference_get_imgs_blur():
    g = np.load(GOLDEN)
    imsize = [int(s) for s in g['imsize']]
    ret, bic, retb, bicb = P.get_imgs_blur(g['hr'], imsize)
    for name, lst in (('ret', ret), ('bic', bic), ('retb', retb), ('bicb', bicb)):
        assert lst[0].dtype == np.float32
        assert np.array_equal(lst[0], g[name + '0_f32']), name       # bit-exact fp32, incl. ToTensor/Normalize rounding
        for i in range(1, len(imsize)):
            assert np.array_equal(_u8(lst[i]), g['%s%d_u8' % (name, i)]), (name, i)


def test_oracle_matches_pillow_random():
    PIL = pytest.importorskip('PIL')
    from PIL import Image, ImageFilter
    rng = np.random.default_rng(7)
    for (h, w) in [(256, 256), (97, 131), (32, 32), (5, 9)]:
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        im = Image.fromarray(a)
        for (oh, ow) in [(32, 32), (64, 64), (h, w), (50, 77), (300, 288), (1, 3)]:
            ref = np.asarray(im.resize((ow, oh), Image.BILINEAR))
            assert np.array_equal(ref, P.resize_u8(a, (oh, ow))), (h, w, oh, ow)
        for rad in (2, 1, 3.5):
            if min(h, w) < 2 * int(rad) + 4:
                continue                      # the oracle restates the edgeA <= edgeB branch (radius smaller than the line)
            ref = np.asarray(im.filter(ImageFilter.GaussianBlur(radius=rad)))
            assert np.array_equal(ref, P.gaussian_blur_u8(a, rad)), (h, w, rad)


def test_box_parameters():
    # GaussianBlur(radius=2): three box blurs of radius 1.375 -> integer radius 1 and the two 8.24 weights
    assert P.gaussian_box_params(2.0) == (1, 4473924, 1677722)
    r, ww, fw = P.gaussian_box_params(2.0)
    assert (2 * r + 1) * ww + 2 * fw <= 1 << 24 and (1 << 24) - ((2 * r + 1) * ww + 2 * fw) < 4


def test_resample_coeffs_sum_to_one():
    for (i, o) in [(256, 32), (256, 64), (256, 128), (32, 64), (32, 256), (288, 256), (131, 50)]:
        bounds, kk, ks = P.resample_coeffs(i, o)
        assert kk.shape == (o, ks)
        assert np.all(np.abs(kk.sum(1) - (1 << P.PRECISION_BITS)) <= ks)
        assert np.all(bounds[:, 0] >= 0) and np.all(bounds[:, 0] + bounds[:, 1] <= i)
