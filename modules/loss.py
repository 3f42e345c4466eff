"""`modules.loss` alias (reference: modules/loss.py)."""
from pvcnn_b200.nn.loss import *  # noqa: F401,F403
from pvcnn_b200.nn import loss as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_") and isinstance(getattr(_impl, n), type)]
