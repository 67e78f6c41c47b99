"""Input pyramid on the device: the image side of the reference's datasets.py, batched.

The reference prepares every sample on the host with PIL inside DataLoader workers: `get_imgs_blur`
(datasets.py:151-196) / `get_imgs` (:112-148) / `get_imgsexampletestblur` (:236-279) resize the cropped HR image to each
branch size (`transforms.Resize`), upsample the LR image back ("bic"), blur both (`ImageFilter.GaussianBlur(radius=2)`)
and normalise (`ToTensor` + `Normalize(.5,.5)`, :1563).  At the rates the generator runs on a B200 (tens of thousands of
images per second) that host work is the bottleneck, so here the same arithmetic runs in CUDA (csrc/pyramid.cu) on uint8
batches `[B, H, W, 3]` that are already decoded and cropped (decode / crop / flip stay with the caller).  Results are
byte-identical to Pillow's; the functions keep the reference's names and return structure.

No CPU path: every function raises if the tensors are not on a CUDA device or the library is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import call, ptr

Tensor = torch.Tensor


def _u8(img: Tensor) -> Tensor:
    if not img.is_cuda:
        raise RuntimeError('tgsr_b200.pyramid: images must be CUDA tensors (no CPU path)')
    if img.dtype != torch.uint8 or img.dim() != 4:
        raise TypeError('tgsr_b200.pyramid: expected a uint8 [B, H, W, C] batch, got %s %s' % (img.dtype, tuple(img.shape)))
    return img if img.is_contiguous() else img.contiguous()


def _size2(size, h: int, w: int) -> Tuple[int, int]:
    """torchvision Resize semantics: an int resizes the smaller edge to `size` keeping the aspect ratio."""
    if isinstance(size, int):
        if h <= w:
            return size, int(size * w / h)
        return int(size * h / w), size
    oh, ow = size
    return int(oh), int(ow)


def resize(img: Tensor, size) -> Tensor:
    """`transforms.Resize(size)(pil_img)` for a uint8 [B, H, W, C] batch -> uint8 [B, oh, ow, C] (PIL bilinear with
    antialiasing; datasets.py:130,139,145)."""
    img = _u8(img)
    B, H, W, Cn = img.shape
    oh, ow = _size2(size, H, W)
    out = torch.empty((B, oh, ow, Cn), device=img.device, dtype=torch.uint8)
    lib = _lib.load()
    ws = torch.empty(max(int(lib.tgsr_resize_u8_ws_bytes(B, H, W, oh, ow, Cn)), 16), device=img.device, dtype=torch.uint8)
    call('tgsr_resize_u8', img.device, ptr(img), ptr(out), ptr(ws), B, H, W, oh, ow, Cn)
    return out


def gaussian_blur(img: Tensor, radius: float = 2.0) -> Tensor:
    """`pil_img.filter(ImageFilter.GaussianBlur(radius))` for a uint8 [B, H, W, C] batch (datasets.py:186,193)."""
    img = _u8(img)
    B, H, W, Cn = img.shape
    out = torch.empty_like(img)
    tmp = torch.empty_like(img)
    call('tgsr_gaussian_blur_u8', img.device, ptr(img), ptr(out), ptr(tmp), B, H, W, Cn, C.c_float(radius), 3)
    return out


def normalize(img: Tensor) -> Tensor:
    """`ToTensor()` + `Normalize((.5,.5,.5),(.5,.5,.5))` (datasets.py:1563): uint8 [B,H,W,C] -> fp32 [B,C,H,W] in [-1,1]."""
    img = _u8(img)
    B, H, W, Cn = img.shape
    out = torch.empty((B, Cn, H, W), device=img.device, dtype=torch.float32)
    call('tgsr_normalize_u8_f32', img.device, ptr(img), ptr(out), B, H, W, Cn)
    return out


def _levels(src: Tensor, sizes: Sequence[Tuple[int, int]], want_out: Sequence[bool], want_blur: Sequence[bool],
            want_u8: Sequence[bool], radius: float):
    """One fused launch: per level (normalised, normalised-blurred, resized bytes), each None unless wanted."""
    B, H, W, Cn = src.shape
    n = len(sizes)
    dev = src.device
    outs = [torch.empty((B, Cn, oh, ow), device=dev, dtype=torch.float32) if w else None for (oh, ow), w in zip(sizes, want_out)]
    blurs = [torch.empty((B, Cn, oh, ow), device=dev, dtype=torch.float32) if w else None for (oh, ow), w in zip(sizes, want_blur)]
    u8s = [torch.empty((B, oh, ow, Cn), device=dev, dtype=torch.uint8) if w else None for (oh, ow), w in zip(sizes, want_u8)]

    def parr(ts):
        return (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts])
    oh_a = (C.c_int * n)(*[s[0] for s in sizes])
    ow_a = (C.c_int * n)(*[s[1] for s in sizes])
    call('tgsr_pyramid_levels_u8', dev, ptr(src), B, H, W, Cn, n, oh_a, ow_a, parr(outs), parr(blurs), parr(u8s),
         C.c_float(radius))
    return outs, blurs, u8s


def _pyramid(img: Tensor, imsize: Sequence[int], blur: bool, radius: float, want_bic: bool = True):
    img = _u8(img)
    B, H, W, Cn = img.shape
    sizes = [_size2(int(s), H, W) for s in imsize]
    n = len(sizes)
    if n < 1 or n > 4:
        raise ValueError('between 1 and 4 branch sizes are supported, got %d' % n)
    if sizes[-1] != (H, W):
        # the reference uses the transformed image itself as the last branch (datasets.py:140-141,180-181)
        raise ValueError('the last branch size %s must equal the image size %s' % (sizes[-1], (H, W)))
    lib = _lib.load()
    lr_hw = sizes[0]
    fused = all(lib.tgsr_pyramid_supported(H, W, oh, ow, int(blur)) and
                lib.tgsr_pyramid_supported(lr_hw[0], lr_hw[1], oh, ow, int(blur)) for oh, ow in sizes)
    if not fused:
        # generic kernels, one op at a time (odd sizes, very large planes)
        lr = resize(img, lr_hw)
        ret, bic, retb, bicb = [], [], [], []
        for i, s in enumerate(sizes):
            re_img = resize(img, s) if i < n - 1 else img
            ret.append(normalize(re_img))
            if blur:
                retb.append(normalize(gaussian_blur(re_img, radius)))
            if want_bic:
                bicim = resize(lr, s)
                bic.append(normalize(bicim))
                if blur:
                    bicb.append(normalize(gaussian_blur(bicim, radius)))
        return ret, bic, retb, bicb
    # launch 1: every branch of the HR image, largest plane first; branch 0 also leaves the LR bytes for launch 2
    order = list(range(n - 1, -1, -1))
    outs, blurs, u8s = _levels(img, [sizes[i] for i in order], [True] * n, [blur] * n, [want_bic and i == 0 for i in order], radius)
    ret = [None] * n
    retb = [None] * n
    for k, i in enumerate(order):
        ret[i], retb[i] = outs[k], blurs[k]
    if not want_bic:
        return ret, [], (retb if blur else []), []
    lr = u8s[order.index(0)]
    # launch 2: the LR image resized to branches 1.. ("bic"); branch 0 of bic is the LR image itself
    bic = [ret[0]] + [None] * (n - 1)
    bicb = [retb[0]] + [None] * (n - 1)
    if n > 1:
        order2 = list(range(n - 1, 0, -1))
        outs2, blurs2, _ = _levels(lr, [sizes[i] for i in order2], [True] * (n - 1), [blur] * (n - 1), [False] * (n - 1), radius)
        for k, i in enumerate(order2):
            bic[i], bicb[i] = outs2[k], blurs2[k]
    return ret, bic, (retb if blur else []), (bicb if blur else [])


def get_imgs_blur(img: Tensor, imsize: Sequence[int], radius: float = 2.0
                  ) -> Tuple[List[Tensor], List[Tensor], List[Tensor], List[Tensor]]:
    """datasets.py:151-196 after the crop/transform, for a uint8 [B, H, W, 3] CUDA batch with H == W == imsize[-1]:
    -> (ret, bic, retb, bicb), lists over the branches of normalised fp32 [B, 3, s, s]:
       ret[i]  = Resize(imsize[i])(img) (the image itself for the last branch), retb[i] = GaussianBlur(2) of it,
       bic[i]  = Resize(imsize[i])(Resize(imsize[0])(img)),                    bicb[i] = GaussianBlur(2) of it.
    bic[0] / bicb[0] are the same tensors as ret[0] / retb[0] (Resize to the same size returns a copy of the image)."""
    return _pyramid(img, imsize, True, radius)


def real_pyramid(img: Tensor, imsize: Sequence[int]) -> List[Tensor]:
    """Only `ret` of get_imgs (the LR input and the MSE targets of a training step): one launch, no "bic" branch."""
    return _pyramid(img, imsize, False, 2.0, want_bic=False)[0]


def get_imgs(img: Tensor, imsize: Sequence[int]) -> Tuple[List[Tensor], List[Tensor]]:
    """datasets.py:112-148 after the crop/transform -> (ret, bic)."""
    ret, bic, _, _ = _pyramid(img, imsize, False, 2.0)
    return ret, bic
