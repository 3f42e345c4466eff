"""`modules.ball_query` alias (reference: modules/ball_query.py)."""
from pvcnn_b200.nn.ball_query import *  # noqa: F401,F403
from pvcnn_b200.nn import ball_query as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_") and isinstance(getattr(_impl, n), type)]
