"""Bring-up + parity tests of the tcgen05 implicit-GEMM convolution (pvcnn_igemm_conv)."""
import numpy as np
import pytest
import torch

from pvcnn_b200 import dense
from util import rng

pytestmark = pytest.mark.gpu


def trunc_tf32(t):
    return (t.view(torch.int32) & -8192).view(torch.float32)


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("m,k,n", [(128, 32, 64), (1024, 64, 64), (4096, 96, 128), (65536, 64, 64), (300, 40, 24),
                                   (2048, 64, 200)])
def test_gemm_exact_tf32_operands(m, k, n):
    """Operands already exact in tf32 -> a single pass must reproduce the fp64 product to fp32
    accumulation error.  Validates descriptors, swizzle, TMA boxes, TMEM addressing."""
    torch.manual_seed(0)
    kp = (k + 3) // 4 * 4
    a = torch.zeros(m, kp, device="cuda")
    a[:, :k] = trunc_tf32(torch.randn(m, k, device="cuda"))
    w = trunc_tf32(torch.randn(n, k, 1, device="cuda"))
    bias = torch.randn(n, device="cuda")
    w_hi, w_lo = dense.prep_weight(w)
    out = dense.igemm_conv(a.view(1, 1, 1, m, kp), None, w_hi, None, bias, npass=1)
    ref = a[:, :k].double() @ w[:, :, 0].double().t() + bias.double()
    assert relerr(out.view(m, -1)[:, :n], ref) < 2e-6


def test_hw_rounding_probe(capsys):
    """Informational: does kind::tf32 truncate or round raw fp32 operands?"""
    torch.manual_seed(1)
    m, k, n = 1024, 64, 64
    a = torch.randn(m, k, device="cuda")
    w = torch.randn(n, k, 1, device="cuda")
    ld = 64
    w_raw = w[:, :, 0].contiguous().view(1, n, ld)
    out = dense.igemm_conv(a.view(1, 1, 1, m, k), None, w_raw, None, None, npass=1)
    out = out.view(m, n)
    ref_trunc = trunc_tf32(a).double() @ trunc_tf32(w[:, :, 0].contiguous()).double().t()
    a_rn = trunc_tf32((a.view(torch.int32) + 4096).view(torch.float32))
    w_rn = trunc_tf32((w[:, :, 0].contiguous().view(torch.int32) + 4096).view(torch.float32))
    ref_rn = a_rn.double() @ w_rn.double().t()
    ref = a.double() @ w[:, :, 0].double().t()
    with capsys.disabled():
        print("\n[tf32 probe] err vs trunc-model %.3e | vs round-model %.3e | vs exact %.3e"
              % (relerr(out, ref_trunc), relerr(out, ref_rn), relerr(out, ref)))


@pytest.mark.parametrize("npass,tol", [(1, 3e-3), (3, 1e-5)])
@pytest.mark.parametrize("b,r,cin,cout", [(2, 8, 16, 16), (1, 16, 64, 64), (2, 32, 64, 64), (1, 12, 64, 128),
                                          (1, 16, 9, 64), (1, 8, 128, 256), (2, 12, 4, 64), (1, 16, 128, 128),
                                          (1, 8, 256, 256), (1, 32, 32, 48), (1, 12, 64, 64), (1, 20, 24, 40)])
def test_conv3d_forward(npass, tol, b, r, cin, cout):
    torch.manual_seed(2)
    cp = (cin + 3) // 4 * 4
    x = torch.randn(b, cin, r, r, r, device="cuda")
    conv = torch.nn.Conv3d(cin, cout, 3, padding=1).cuda()
    ref = torch.nn.functional.conv3d(x.double().cpu(), conv.weight.double().cpu(), conv.bias.double().cpu(), padding=1)
    xcl = torch.zeros(b, r, r, r, cp, device="cuda")
    xcl[..., :cin] = x.permute(0, 2, 3, 4, 1)
    a_hi, a_lo = dense.split_tf32(xcl, want_hi=False)
    w_hi, w_lo = dense.prep_weight(conv.weight)
    out = dense.igemm_conv(a_hi, a_lo, w_hi, w_lo, conv.bias.detach(), npass=npass)
    got = out[..., :cout].permute(0, 4, 1, 2, 3).cpu()
    assert relerr(got, ref) < tol


@pytest.mark.parametrize("b,r,cin,cout", [(1, 8, 16, 32), (1, 16, 64, 64), (1, 12, 64, 128), (1, 8, 256, 256),
                                          (2, 32, 9, 64)])
def test_conv3d_dgrad(b, r, cin, cout):
    torch.manual_seed(3)
    x = torch.randn(b, cin, r, r, r, dtype=torch.float64, requires_grad=True)
    conv = torch.nn.Conv3d(cin, cout, 3, padding=1).double()
    gy = torch.randn(b, cout, r, r, r, dtype=torch.float64)
    conv(x).backward(gy)
    gcl = gy.float().permute(0, 2, 3, 4, 1).contiguous().cuda()
    g_hi, g_lo = dense.split_tf32(gcl, want_hi=False)
    w_hi, w_lo = dense.prep_weight(conv.weight.float().cuda(), mode=1)
    out = dense.igemm_conv(g_hi, g_lo, w_hi, w_lo, None, npass=3)
    got = out[..., :cin].permute(0, 4, 1, 2, 3).cpu()
    assert relerr(got, x.grad) < 1e-5


def test_conv_linearity_full_size():
    """Metric-size property test (B=16, R=32, C=64): conv(a*x1 + x2) == a*conv(x1) + conv(x2) without
    bias -- size-independent check where the CPU oracle is too slow."""
    torch.manual_seed(4)
    b, r, c = 16, 32, 64
    x1 = torch.randn(b, r, r, r, c, device="cuda")
    x2 = torch.randn(b, r, r, r, c, device="cuda")
    w = torch.randn(c, c, 3, 3, 3, device="cuda") * 0.05
    w_hi, w_lo = dense.prep_weight(w)

    def conv(t):
        hi, lo = dense.split_tf32(t, want_hi=False)
        return dense.igemm_conv(hi, lo, w_hi, w_lo, None, npass=3)

    lhs = conv(2.5 * x1 + x2)
    rhs = 2.5 * conv(x1) + conv(x2)
    assert relerr(lhs, rhs) < 1e-5
    # spot-check 64 random voxels against an fp64 evaluation
    idx = torch.randint(1, r - 1, (64, 3))
    y = conv(x1)
    for (i, j, k) in idx.tolist():
        patch = x1[3, i - 1:i + 2, j - 1:j + 2, k - 1:k + 2, :].double()      # [3,3,3,C]
        ref = torch.einsum("xyzc,ocxyz->o", patch, w.double())
        assert relerr(y[3, i, j, k], ref) < 2e-5


@pytest.mark.parametrize("npass,tol", [(1, 3e-3), (3, 2e-5)])
@pytest.mark.parametrize("b,r,cin,cout", [(1, 8, 16, 16), (2, 16, 64, 64), (1, 12, 64, 128), (1, 16, 9, 64),
                                          (4, 32, 64, 64), (2, 8, 256, 256), (1, 8, 128, 192)])
def test_conv3d_wgrad(npass, tol, b, r, cin, cout):
    torch.manual_seed(5)
    x = torch.randn(b, cin, r, r, r, device="cuda")
    gy = torch.randn(b, cout, r, r, r, device="cuda")
    conv = torch.nn.Conv3d(cin, cout, 3, padding=1).double()
    xd = x.double().cpu()
    conv(xd).backward(gy.double().cpu())
    ref = conv.weight.grad.reshape(cout, cin, 27)
    cp = (cin + 3) // 4 * 4
    xcl = torch.zeros(b, r, r, r, cp, device="cuda")
    xcl[..., :cin] = x.permute(0, 2, 3, 4, 1)
    gcl = gy.permute(0, 2, 3, 4, 1).contiguous()
    x_hi, x_lo = dense.split_tf32(xcl, want_hi=False)
    g_hi, g_lo = dense.split_tf32(gcl, want_hi=False)
    dw = dense.conv_wgrad(x_hi, x_lo, g_hi, g_lo, cin, cout, 27, npass=npass)
    assert relerr(dw.cpu(), ref) < tol


@pytest.mark.parametrize("m,cin,cout", [(4096, 64, 64), (65536, 64, 64), (2048, 16, 32), (5000, 9, 64), (4096, 1472, 512),
                                        (3000, 128, 1024)])
def test_pointwise_wgrad(m, cin, cout):
    torch.manual_seed(6)
    cp = (cin + 3) // 4 * 4
    x = torch.zeros(m, cp, device="cuda")
    x[:, :cin] = torch.randn(m, cin, device="cuda")
    g = torch.randn(m, cout, device="cuda")
    ref = g.double().t() @ x[:, :cin].double()
    x_hi, x_lo = dense.split_tf32(x.view(1, 1, 1, m, cp), want_hi=False)
    g_hi, g_lo = dense.split_tf32(g.view(1, 1, 1, m, cout), want_hi=False)
    dw = dense.conv_wgrad(x_hi, x_lo, g_hi, g_lo, cin, cout, 1, npass=3)
    assert relerr(dw[:, :, 0], ref) < 1e-5
