import torch.nn as nn


class SharedMLP(nn.Module):
    """Stack of (1x1 conv, BatchNorm, ReLU) over points (reference: modules/shared_mlp.py:6-33).
    `dim` selects Conv1d/BatchNorm1d (1) or Conv2d/BatchNorm2d (2); the parameters live in
    `self.layers` with the reference's indices (layers.{0,1}, layers.{3,4}, ...), so released
    checkpoints load unchanged.  A tuple/list input has its first element transformed and the
    rest passed through."""

    _BLOCKS = {1: (nn.Conv1d, nn.BatchNorm1d), 2: (nn.Conv2d, nn.BatchNorm2d)}

    def __init__(self, in_channels, out_channels, dim=1):
        super().__init__()
        if dim not in self._BLOCKS:
            raise ValueError
        conv, bn = self._BLOCKS[dim]
        widths = list(out_channels) if isinstance(out_channels, (list, tuple)) else [out_channels]
        seq = []
        prev = in_channels
        for w in widths:
            seq += [conv(prev, w, 1), bn(w), nn.ReLU(True)]
            prev = w
        self.layers = nn.Sequential(*seq)

    def forward(self, inputs):
        if isinstance(inputs, (list, tuple)):
            head, *rest = inputs
            return (self.layers(head), *rest)
        return self.layers(inputs)
