"""Tensor-level wrappers of the dense tcgen05 primitives (pvcnn_igemm_conv & friends).
Channels-last activations [B, X, Y, Z, C] (or [M, C] for plain GEMMs)."""
import torch

from . import _lib


def split_tf32(x, want_hi=True):
    """(hi, lo) with hi = trunc_tf32(x) (or x itself when want_hi=False: the tensor core truncates)."""
    x = x.contiguous()
    lo = torch.empty_like(x)
    hi = torch.empty_like(x) if want_hi else None
    _lib.call("pvcnn_split_tf32", _LL(x.numel()), x, hi, lo)
    return (hi if want_hi else x), lo


def _LL(v):
    import ctypes
    return ctypes.c_longlong(int(v))


def prep_weight(w, mode=0):
    """w: [cout, cin, *k] conv weight -> (hi, lo) [ntaps, rows, ld] with rows=cout (mode 0, forward)
    or rows=cin (mode 1, data gradient: taps flipped, K=cout)."""
    w = w.detach().contiguous().float()
    cout, cin = w.shape[:2]
    ntaps = w[0, 0].numel()
    kdim = cin if mode == 0 else cout
    rows = cout if mode == 0 else cin
    ld = ((kdim + 31) // 32) * 32
    hi = torch.empty((ntaps, rows, ld), dtype=torch.float32, device=w.device)
    lo = torch.empty_like(hi)
    _lib.call("pvcnn_conv_weight_prep", cout, cin, ntaps, mode, ld, w, hi, lo)
    return hi, lo


def igemm_conv(a_hi, a_lo, w_hi, w_lo, bias, npass=3, cout=None):
    """a_*: [B,X,Y,Z,C] channels-last (C % 4 == 0); w_*: [ntaps, cout, ld] from prep_weight.
    Returns fp32 [B,X,Y,Z,cout_pad4]."""
    nb, sx, sy, sz, c = a_hi.shape
    ntaps, rows, ldw = w_hi.shape
    cout = rows if cout is None else cout
    ldo = ((cout + 3) // 4) * 4
    out = torch.empty((nb, sx, sy, sz, ldo), dtype=torch.float32, device=a_hi.device)
    _lib.call("pvcnn_igemm_conv", nb, sx, sy, sz, c, cout, ntaps, a_hi, a_lo if npass > 1 else None, c, w_hi,
              w_lo if npass > 1 else None, ldw, bias, out, ldo, npass)
    return out


def conv_wgrad(x_hi, x_lo, g_hi, g_lo, cin, cout, ntaps, npass=3):
    """x: layer input [B,X,Y,Z,Cx], g: output gradient [B,X,Y,Z,Cg] -> dW [cout, cin, ntaps] (torch layout)."""
    nb, sx, sy, sz, ldx = x_hi.shape
    ldg = g_hi.shape[-1]
    dw = torch.empty((cout, cin, ntaps), dtype=torch.float32, device=x_hi.device)
    _lib.call("pvcnn_conv_wgrad", nb, sx, sy, sz, cin, cout, ntaps, x_hi, x_lo if npass > 1 else None, ldx, g_hi,
              g_lo if npass > 1 else None, ldg, dw, npass)
    return dw
