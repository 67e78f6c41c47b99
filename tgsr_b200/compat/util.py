import torch  # noqa: F401  (the reference's util.py exports these names through `from util import *`, model.py:3)
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
from torch.autograd import Variable  # noqa: F401
from tgsr_b200.attention import GlobalAttentionGeneral as ATT_NET  # noqa: F401
from tgsr_b200.blocks import (CA_NET, GET_IMAGE_G, GET_IMAGE_G_noAct, GLU, INIT_STAGE_G, INIT_STAGE_GImgup,  # noqa: F401
                              NEXT_STAGE_G, Block3x3_relu, ResBlock, conv1x1, conv3x3, conv5x5, upBlock, upBlocknoBN)
from tgsr_b200.config import cfg  # noqa: F401
from tgsr_b200.text_encoder import RNN_ENCODER  # noqa: F401
