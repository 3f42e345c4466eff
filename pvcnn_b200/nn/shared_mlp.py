import os

import torch.nn as nn


def native_mlp_enabled():
    """PVCNN_B200_MLP=torch selects the stock torch layers (cuDNN + ATen) as the in-repo comparison arm."""
    return os.environ.get("PVCNN_B200_MLP", "native").lower() != "torch"


class SharedMLP(nn.Module):
    """Stack of (1x1 conv, BatchNorm, ReLU) over points (reference: modules/shared_mlp.py:6-33).
    `dim` selects Conv1d/BatchNorm1d (1) or Conv2d/BatchNorm2d (2); the parameters live in
    `self.layers` with the reference's indices (layers.{0,1}, layers.{3,4}, ...), so released
    checkpoints load unchanged.  A tuple/list input has its first element transformed and the
    rest passed through.

    Execution on a CUDA tensor does NOT go through those sub-modules' forward(): every layer runs as one
    tcgen05 GEMM (igemm_conv_kernel) with the BatchNorm statistics / apply / ReLU passes and the whole backward in
    our own kernels (pvcnn_b200/mlp.py -> pvcnn_mlp_layer_forward/backward), on channels-last rows.  The sub-modules
    only own the parameters and running statistics."""

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

    def _run(self, x):
        if x.is_cuda and native_mlp_enabled():
            from .. import mlp
            if mlp.native_supported(self.layers):
                return mlp.shared_mlp_forward(self.layers, x)
        return self.layers(x)

    def forward(self, inputs):
        if isinstance(inputs, (list, tuple)):
            head, *rest = inputs
            return (self._run(head), *rest)
        return self._run(inputs)
