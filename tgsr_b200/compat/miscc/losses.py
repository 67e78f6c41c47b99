"""Alias of the reference's miscc/losses.py for the module-level drop-in."""
from tgsr_b200.losses import (KL_loss, MSE, cosine_similarity, generator_lossor_wordsentOnly, sent_loss,  # noqa: F401
                              words_loss)
from tgsr_b200.attention import func_attention  # noqa: F401
