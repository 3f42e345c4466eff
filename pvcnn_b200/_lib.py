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


def _arg(a):
    import torch
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, float):
        return ctypes.c_float(a)
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, int):
        return ctypes.c_int(a)
    return a


def _raw_stream(index):
    import torch
    get = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    return get(index) if get is not None else torch.cuda.current_stream(index).cuda_stream


def call(name, *args, device=None):
    """Invoke `name(*args, stream)` on the current CUDA stream of `device`; raise on error."""
    import torch
    fn = getattr(load(), name)
    if device is None:
        for a in args:
            if isinstance(a, torch.Tensor):
                device = a.device
                break
    if device is None or device.type != "cuda":
        raise PvcnnError("%s: tensors must live on a CUDA device (the hot path has no CPU implementation)" % name)
    index = device.index if device.index is not None else torch.cuda.current_device()
    cargs = [_arg(a) for a in args]
    if index == torch.cuda.current_device():   # common case: no device switch, ~10 us less host overhead per call
        rc = fn(*cargs, ctypes.c_void_p(_raw_stream(index)))
    else:
        with torch.cuda.device(index):
            rc = fn(*cargs, ctypes.c_void_p(_raw_stream(index)))
    if rc != 0:
        raise PvcnnError("%s failed with code %d" % (name, rc))
