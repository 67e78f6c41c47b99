from util import *  # noqa: F401,F403
from tgsr_b200.nets16 import G_SR_NET_low, NetG_high, NetG_highweight  # noqa: F401
