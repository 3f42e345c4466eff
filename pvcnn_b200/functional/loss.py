"""Pure-torch losses of the reference surface (modules/functional/loss.py:7-17).  Not on the hot
path; provided so that `modules.functional` is complete."""
import torch
import torch.nn.functional as TF


def kl_loss(x, y):
    """KL(softmax(x) || softmax(y)) averaged over the batch; x is treated as a constant
    (modules/functional/loss.py:7-10)."""
    p = TF.softmax(x.detach(), dim=1)
    log_q = TF.log_softmax(y, dim=1)
    return (p * (p.log() - log_q)).sum(dim=1).mean()


def huber_loss(error, delta):
    """Mean Huber penalty with threshold `delta` (modules/functional/loss.py:13-17)."""
    a = error.abs()
    q = a.clamp(max=delta)
    return (0.5 * q * q + delta * (a - q)).mean()
