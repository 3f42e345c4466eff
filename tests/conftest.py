import importlib.util
import os
import sys
import time

_T0 = time.time()   # the driver's GPU-test step (python start-up to exit) is killed at 1200 s: see _GPU_BUDGET_S below

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from util import host_threads  # noqa: E402

# The CPU references (torch fp64 conv3d / matmul, the OpenMP C oracle) default to one thread per HARDWARE thread of the
# node; inside a container with a smaller CPU quota that oversubscription is catastrophic for torch's CPU conv3d (measured
# by bench.py's cpu_baseline on a GPU box: 4.2 s/step with 128 threads vs 0.16 s with 8-32).  Cap them once, before torch
# and libgomp read the environment; results do not depend on it (fp64 references, 1e-5 tolerances).
os.environ.setdefault("OMP_NUM_THREADS", str(host_threads()))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    try:
        import torch
        torch.set_num_threads(host_threads())
    except Exception:  # noqa: BLE001
        pass


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# Safety net, not a feature: the round-end driver kills `pytest -m gpu` at 1200 s (GPUTEST_r01.json: step timeout_s 1200;
# round 1 took 820 s for 107 tests, this suite has 222).  A killed run reports nothing, so past this many seconds since
# start-up the REMAINING gpu tests are skipped with an explicit reason (visible in the summary as skips, never as passes).
# PVCNN_TEST_BUDGET_S=0 disables it (local full runs).
_GPU_BUDGET_S = float(os.environ.get("PVCNN_TEST_BUDGET_S", "960"))


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_runtest_setup(item):
    if _GPU_BUDGET_S > 0 and "gpu" in item.keywords and time.time() - _T0 > _GPU_BUDGET_S:
        pytest.skip("GPU suite time budget (%d s since start-up) exhausted; the driver kills the step at 1200 s -- run this "
                    "test alone or with PVCNN_TEST_BUDGET_S=0" % _GPU_BUDGET_S)


_REF = None


def load_ref_backend():
    """The UNMODIFIED reference extension built by oracle/build_ref.py (checker only)."""
    global _REF
    if _REF is None:
        import torch  # noqa: F401  (libtorch symbols must be loaded first)
        path = os.path.join(ROOT, "oracle", "_ref", "_pvcnn_backend.so")
        if not os.path.exists(path):
            pytest.skip("oracle/_ref/_pvcnn_backend.so not built (run python oracle/build_ref.py)")
        spec = importlib.util.spec_from_file_location("_pvcnn_backend", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _REF = mod
    return _REF


@pytest.fixture(scope="session")
def ref_backend():
    return load_ref_backend()
