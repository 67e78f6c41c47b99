"""Input-pyramid microbenchmark: tgsr_b200.pyramid.get_imgs_blur on a batch of decoded
256x256 uint8 images, branches [32, 64, 128, 256], against the HBM roofline, with the reference's host path (PIL, one core
per DataLoader worker, datasets.py:151-196) timed beside it.  Prints one JSON line.

Algorithmic bytes per image: the HR image read once (256*256*3 B) + every distinct fp32 output written once:
ret, retb over 4 branches and bic, bicb over branches 1..3 (branch 0 aliases ret/retb)
= 196,608 + 4 * 3 * ((2*1024) + 2*2*(4096 + 16384 + 65536)) B = 4.35 MB."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgsr_b200 import pyramid as Y                       # noqa: E402
from tgsr_b200.timing import cold_time_ms                # noqa: E402


def pil_get_imgs_blur(a, imsize):
    from PIL import Image, ImageFilter
    img = Image.fromarray(a)
    lr = img.resize((imsize[0], imsize[0]), Image.BILINEAR)
    out = []
    for i, s in enumerate(imsize):
        re_img = img.resize((s, s), Image.BILINEAR) if i < len(imsize) - 1 else img
        bic = lr.resize((s, s), Image.BILINEAR)
        for im in (re_img, re_img.filter(ImageFilter.GaussianBlur(radius=2)), bic, bic.filter(ImageFilter.GaussianBlur(radius=2))):
            t = torch.from_numpy(np.asarray(im)).permute(2, 0, 1).float().div(255)
            out.append(t.sub_(0.5).div_(0.5))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--cpu-images', type=int, default=64)
    args = ap.parse_args()
    imsize = [32, 64, 128, 256]
    B = args.batch
    g = torch.Generator().manual_seed(100)
    low = torch.rand((B, 3, 32, 32), generator=g)
    hr = (torch.nn.functional.interpolate(low, size=(256, 256), mode='bicubic').clamp(0, 1) * 255).round().to(torch.uint8)
    hr = hr.permute(0, 2, 3, 1).contiguous()
    hr_d = hr.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ms = cold_time_ms(lambda: Y.get_imgs_blur(hr_d, imsize), flush, iters=args.iters)
    px = sum(s * s for s in imsize)
    bytes_img = 256 * 256 * 3 + 4 * 3 * (2 * imsize[0] ** 2 + 4 * (px - imsize[0] ** 2))
    peak = None
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) as f:
            mp = json.load(f)
        peak = float(mp.get('hbm_gbs') or 0) or None
    except Exception:
        pass
    peak = peak or 6553.0
    gbps = bytes_img * B / (ms * 1e-3) / 1e9
    t0 = time.perf_counter()
    n = 0
    while n < args.cpu_images:
        pil_get_imgs_blur(hr[n % B].numpy(), imsize); n += 1
    cpu = n / (time.perf_counter() - t0)
    print(json.dumps(dict(metric='input_pyramid_images_per_sec', value=round(B / (ms * 1e-3), 1), unit='img/s', batch=B,
                          ms_per_batch=round(ms, 4), launches_per_batch=2,
                          roofline=dict(bound='hbm', achieved=round(gbps, 1), peak=peak, unit='GB/s', frac=round(gbps / peak, 4),
                                        bytes_per_image=bytes_img),
                          cpu_baseline=dict(value=round(cpu, 1), unit='img/s', cores=1, kind='reference',
                                            sample='%d images through PIL resize/GaussianBlur + ToTensor/Normalize' % n),
                          config=dict(workload='get_imgs_blur 256x256 uint8 -> branches [32,64,128,256]', l2='flushed between launches'))))


if __name__ == '__main__':
    main()
