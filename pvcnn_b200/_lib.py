"""ctypes binding of the C ABI declared in include/pvcnn_b200.h.

There is deliberately NO fallback: if the shared library is missing or a call fails, this
module raises.  The oracle under oracle/ is never imported from here.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpvcnn_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "pvcnn_b200.h")

_lib = None


class PvcnnError(RuntimeError):
    pass


def header_symbols():
    """Names of every function declared in include/pvcnn_b200.h (used by the ABI test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pvcnn_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PvcnnError(
                "libpvcnn_b200.so not built (%s). Run `python -m pvcnn_b200.build`; there is no CPU "
                "or PyTorch fallback for the hot path." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.pvcnn_build_info.restype = ctypes.c_char_p
        lib.pvcnn_launch_count.restype = ctypes.c_ulonglong
        if lib.pvcnn_abi_version() != 1:
            raise PvcnnError("ABI version mismatch")
        _lib = lib
    return _lib


def launch_count():
    return int(load().pvcnn_launch_count())


_FN = {}
_VOID0 = ctypes.c_void_p(0)


def call(name, *args, device=None):
    """Invoke `name(*args, stream)` on the current CUDA stream of `device`; raise on error.
    (Host-side cost matters for the small point ops, which are launch-bound: function lookup is cached and the
    argument conversion is a single pass.)"""
    import torch
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(load(), name)
    Tensor = torch.Tensor
    cargs = []
    for a in args:
        if isinstance(a, Tensor):
            if device is None:
                device = a.device
            cargs.append(ctypes.c_void_p(a.data_ptr()))
        elif a is None:
            cargs.append(_VOID0)
        elif isinstance(a, float):
            cargs.append(ctypes.c_float(a))
        elif isinstance(a, (bool, int)):
            cargs.append(ctypes.c_int(int(a)))
        else:
            cargs.append(a)
    if device is None or device.type != "cuda":
        raise PvcnnError("%s: tensors must live on a CUDA device (the hot path has no CPU implementation)" % name)
    cur = torch.cuda.current_device()
    index = device.index if device.index is not None else cur
    if index == cur:   # common case: no device switch
        rc = fn(*cargs, ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(index)))
    else:
        with torch.cuda.device(index):
            rc = fn(*cargs, ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(index)))
    if rc != 0:
        raise PvcnnError("%s failed with code %d" % (name, rc))
