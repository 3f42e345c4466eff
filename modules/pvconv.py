"""`modules.pvconv` alias (reference: modules/pvconv.py)."""
from pvcnn_b200.nn.pvconv import *  # noqa: F401,F403
from pvcnn_b200.nn import pvconv as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_") and isinstance(getattr(_impl, n), type)]
