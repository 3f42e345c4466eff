import torch.nn as nn


class SE3d(nn.Module):
    """Squeeze-and-excite over a [B,C,R,R,R] grid (reference: modules/se.py:6-17):
    mean over the voxels -> Linear(C, C/8) -> ReLU -> Linear(C/8, C) -> Sigmoid -> channel scale."""

    def __init__(self, channel, reduction=8):
        super().__init__()
        hidden = channel // reduction
        self.fc = nn.Sequential(nn.Linear(channel, hidden, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(hidden, channel, bias=False), nn.Sigmoid())

    def gate(self, pooled):
        return self.fc(pooled)

    def forward(self, inputs):
        b, c = inputs.shape[:2]
        pooled = inputs.mean(-1).mean(-1).mean(-1)
        return inputs * self.fc(pooled).view(b, c, 1, 1, 1)
