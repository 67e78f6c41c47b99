import numpy as np


def pyramid_expand(image, upscale=2, sigma=None, **_kw):
    """Nearest-neighbour upscale (the real function also smooths with a Gaussian of `sigma`): shape-compatible, enough for
    the visualisation canvas the unmodified build_super_images* assemble."""
    k = int(upscale)
    out = np.repeat(np.repeat(image, k, axis=0), k, axis=1)
    return out.astype(np.float64)


def resize(image, output_shape, **_kw):
    """Nearest-neighbour resize to `output_shape` (rows, cols) (the real function interpolates and anti-aliases)."""
    h, w = int(output_shape[0]), int(output_shape[1])
    ys = (np.arange(h) * image.shape[0] // h).astype(np.int64)
    xs = (np.arange(w) * image.shape[1] // w).astype(np.int64)
    return image[ys][:, xs].astype(np.float64)
