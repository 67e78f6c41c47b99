from util import *  # noqa: F401,F403  (the reference's model.py re-exports util, model.py:3)
from tgsr_b200.nets import G_SR_NET_low, NetG_high, NetG_highweight  # noqa: F401
