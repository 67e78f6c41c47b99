from util import *  # noqa: F401,F403  (the reference's model.py re-exports util, model.py:3)
from tgsr_b200.nets import G_SR_NET_low, NetG_high, NetG_highweight  # noqa: F401


class G_SR_NET_low_stage1(nn.Module):
    """model.py:81-130: an abandoned single-stage variant the eval driver imports by name (trainer_objective.py:8) but only
    builds when its hard-coded `stage1` flag is set (trainer_objective.py:56,76-77).  Out of scope."""

    def __init__(self):
        super().__init__()
        raise NotImplementedError('G_SR_NET_low_stage1 is out of scope (no shipped flow or checkpoint uses it)')
