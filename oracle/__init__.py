"""CPU oracle for the PVConv hot path -- TEST INFRASTRUCTURE, never the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  See oracle/pvcnn_oracle.c for the per-kernel reference citations.
"""
from .oracle import *  # noqa: F401,F403
