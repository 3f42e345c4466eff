"""Loss modules of the reference's `modules` package.  Pure torch, not on the accelerated path; present so that
`from modules import KLLoss` (train_dml.py:100) keeps working on the drop-in package."""
from torch import nn

from ..functional.loss import kl_loss


class KLLoss(nn.Module):
    """Mutual-learning criterion (reference: modules/loss.py:8-10): KL(softmax(target) || softmax(logits)) with the
    first argument detached, averaged over the batch.  Stateless: no parameters, no buffers."""

    def forward(self, target_logits, logits):
        return kl_loss(target_logits, logits)

    def extra_repr(self):
        return "reduction=batch-mean, target detached"
