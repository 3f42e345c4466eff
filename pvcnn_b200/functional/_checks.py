"""Argument validation mirroring the reference's CHECK_* macros
(modules/functional/src/utils.hpp:7-18): same conditions, same RuntimeError messages."""
import torch


def check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(name + " must be a CUDA tensor")


def check_contiguous(t, name):
    if not t.is_contiguous():
        raise RuntimeError(name + " must be a contiguous tensor")


def check_float(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError(name + " must be a float tensor")


def check_int(t, name):
    if t.dtype != torch.int32:
        raise RuntimeError(name + " must be an int tensor")


def check_f(t, name):
    check_cuda(t, name); check_contiguous(t, name); check_float(t, name)


def check_i(t, name):
    check_cuda(t, name); check_contiguous(t, name); check_int(t, name)
