"""`modules.shared_mlp` alias (reference: modules/shared_mlp.py)."""
from pvcnn_b200.nn.shared_mlp import *  # noqa: F401,F403
from pvcnn_b200.nn import shared_mlp as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_") and isinstance(getattr(_impl, n), type)]
