"""Import stand-in for scikit-image (absent from this image): miscc/utils.py:12 imports skimage.transform for the attention
VISUALISATION only (pyramid_expand of the per-word maps); nothing on the hot path or in the SR image depends on it."""
from . import transform  # noqa: F401
