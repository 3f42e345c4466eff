import torch.nn as nn

from .. import functional as F


class KLLoss(nn.Module):
    """modules/loss.py:8-10"""

    def forward(self, x, y):
        return F.kl_loss(x, y)
