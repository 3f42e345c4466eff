"""GPU parity: every stand-alone op through the C ABI (pvcnn_b200.functional.backend ->
libpvcnn_b200.so) against (a) the CPU oracle and (b) the reference's own CUDA kernels
(oracle/_ref/_pvcnn_backend.so, built unmodified from /root/reference).
Bar: integer / index outputs bit-exact; fp32 outputs within 1e-5 relative (BASELINE.json)."""
import numpy as np
import pytest
import torch

import oracle
from pvcnn_b200.functional import backend as B
from pvcnn_b200 import functional as F
from util import rng, s3dis_like_coords, surface_coords, degenerate_coords, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


COORD_GENS = {"s3dis": s3dis_like_coords, "surface": surface_coords, "degenerate": degenerate_coords}


VOX_CASES = [(2, 1024, 8, True, 0.0), (16, 4096, 32, True, 0.0), (3, 777, 12, False, 0.0), (4, 2048, 16, True, 1e-15),
             (32, 1024, 12, True, 0.0), (8, 8192, 32, True, 0.0), (5, 999, 12, True, 0.0)]


@pytest.mark.parametrize("mode", ["exact", "aten"])
@pytest.mark.parametrize("b,n,r,normalize,eps", VOX_CASES)
def test_voxelize_coords_bit_exact_vs_reference_program(mode, b, n, r, normalize, eps, monkeypatch):
    """North-star: integer voxel indices bit-exact with the reference.  Compared with the LITERAL tensor program of
    modules/voxelization.py:17-24 run by torch on the same GPU (cfg 1, the metric config, R=12 cases)."""
    from util import reference_voxelization
    monkeypatch.setenv("PVCNN_B200_VOX", mode)
    g = rng(10)
    for dist in ("s3dis", "surface"):
        c = COORD_GENS[dist](g, b, n) - (0.4 if not normalize else 0.0)
        ct = cu(c)
        nc_ref, vc_ref = reference_voxelization(ct, r, normalize, eps)
        nc, vc = F.voxelize_coords(ct, r, normalize, eps)
        assert torch.equal(vc, vc_ref)
        assert torch.equal(nc, nc_ref)


@pytest.mark.parametrize("b,n,r,normalize,eps", VOX_CASES[:4])
def test_voxelize_coords_vs_oracle(b, n, r, normalize, eps):
    from util import device_mean
    g = rng(10)
    c = s3dis_like_coords(g, b, n) - (0.4 if not normalize else 0.0)
    nc0, vc0 = oracle.voxelize_coords(c, r, normalize, eps, mean=device_mean(c))
    nc, vc = F.voxelize_coords(cu(c), r, normalize, eps)
    assert np.array_equal(npy(vc), vc0)          # integer voxel indices: bit-exact
    assert np.array_equal(npy(nc), nc0)          # same arithmetic -> identical floats


def test_voxelize_coords_fused_variant(monkeypatch):
    """PVCNN_B200_VOX=fused: single kernel, fp64 mean (round-1 definition) == oracle's mean="fp64"."""
    monkeypatch.setenv("PVCNN_B200_VOX", "fused")
    g = rng(10)
    c = s3dis_like_coords(g, 4, 2048)
    nc0, vc0 = oracle.voxelize_coords(c, 16, True, 0.0, mean="fp64")
    nc, vc = F.voxelize_coords(cu(c), 16, True, 0.0)
    assert np.array_equal(npy(vc), vc0) and np.array_equal(npy(nc), nc0)


@pytest.mark.parametrize("dist", ["s3dis", "surface", "degenerate"])
@pytest.mark.parametrize("b,c,n,r", [(2, 16, 1024, 8), (4, 64, 4096, 32), (3, 9, 1000, 12), (1, 1, 33, 4)])
def test_avg_voxelize(dist, b, c, n, r, ref_backend):
    g = rng(11)
    co = COORD_GENS[dist](g, b, n)
    f = g.standard_normal((b, c, n), dtype=np.float32)
    _, vc = oracle.voxelize_coords(co, r)
    out0, ind0, cnt0 = oracle.avg_voxelize(f, vc, r)
    out, ind, cnt = B.avg_voxelize_forward(cu(f), cu(vc), r)
    assert np.array_equal(npy(ind), ind0) and np.array_equal(npy(cnt), cnt0)
    assert rel_err(npy(out), out0) < TOL
    # reference CUDA kernels
    ro, ri, rc = ref_backend.avg_voxelize_forward(cu(f), cu(vc), r)
    torch.cuda.synchronize()
    assert torch.equal(ri, ind) and torch.equal(rc, cnt)
    assert rel_err(npy(out), npy(ro)) < TOL
    gy = g.standard_normal((b, c, r ** 3), dtype=np.float32)
    gx0 = oracle.avg_voxelize_grad(gy, ind0, cnt0)
    gx = B.avg_voxelize_backward(cu(gy), ind, cnt)
    assert np.array_equal(npy(gx), gx0)
    assert rel_err(npy(gx), npy(ref_backend.avg_voxelize_backward(cu(gy), ri, rc))) < TOL


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("b,c,n,r", [(2, 16, 1024, 8), (4, 64, 4096, 32), (3, 9, 1000, 12)])
def test_trilinear_devoxelize(training, b, c, n, r, ref_backend):
    g = rng(12)
    co = s3dis_like_coords(g, b, n)
    nc, _ = oracle.voxelize_coords(co, r)
    nc[:, :, :7] = np.float32(r - 1)
    nc[:, :, 7:14] = np.floor(nc[:, :, 7:14])
    grid = g.standard_normal((b, c, r ** 3), dtype=np.float32)
    o0, i0, w0 = oracle.trilinear_devoxelize(grid, nc, r, training)
    o, i, w = B.trilinear_devoxelize_forward(r, training, cu(nc), cu(grid))
    assert rel_err(npy(o), o0) < TOL
    ro, ri, rw = ref_backend.trilinear_devoxelize_forward(r, training, cu(nc), cu(grid))
    assert rel_err(npy(o), npy(ro)) < TOL
    if training:
        assert np.array_equal(npy(i), i0) and np.array_equal(npy(w), w0)
        assert torch.equal(i, ri) and torch.equal(w, rw)
        gy = g.standard_normal((b, c, n), dtype=np.float32)
        gx0 = oracle.trilinear_devoxelize_grad(gy, i0, w0, r)
        gx = B.trilinear_devoxelize_backward(cu(gy), i, w, r)
        assert rel_err(npy(gx), gx0) < TOL
        assert rel_err(npy(gx), npy(ref_backend.trilinear_devoxelize_backward(cu(gy), ri, rw, r))) < TOL
    else:
        assert i.numel() == 1 and w.numel() == 1


@pytest.mark.parametrize("b,n,m,radius,u", [(2, 1024, 256, 0.2, 32), (8, 8192, 1024, 0.1, 32), (3, 500, 77, 0.4, 5),
                                            (2, 64, 64, 10.0, 64), (2, 300, 40, 1e-4, 8)])
def test_ball_query(b, n, m, radius, u, ref_backend):
    g = rng(13)
    p = g.random((b, 3, n), dtype=np.float32)
    ce = p[:, :, g.permutation(n)[:m]].copy()
    if radius == 1e-4:
        ce += 5.0  # nothing in range -> all-zero rows
    out0 = oracle.ball_query(ce, p, radius, u)
    out = B.ball_query(cu(ce), cu(p), radius, u)
    assert np.array_equal(npy(out), out0)
    assert torch.equal(out, ref_backend.ball_query(cu(ce), cu(p), radius, u))


@pytest.mark.parametrize("b,c,n,m,u", [(2, 35, 1024, 256, 32), (8, 32, 8192, 1024, 32), (1, 3, 50, 7, 3)])
def test_grouping(b, c, n, m, u, ref_backend):
    g = rng(14)
    f = g.standard_normal((b, c, n), dtype=np.float32)
    idx = g.integers(0, n, size=(b, m, u)).astype(np.int32)
    idx[:, :, u // 2:] = idx[:, :, :1]  # first-hit padding pattern -> duplicate destinations
    out = B.grouping_forward(cu(f), cu(idx))
    assert np.array_equal(npy(out), oracle.grouping(f, idx))
    assert torch.equal(out, ref_backend.grouping_forward(cu(f), cu(idx)))
    gy = g.standard_normal((b, c, m, u), dtype=np.float32)
    gx = B.grouping_backward(cu(gy), cu(idx), n)
    assert rel_err(npy(gx), oracle.grouping_grad(gy, idx, n)) < TOL
    assert rel_err(npy(gx), npy(ref_backend.grouping_backward(cu(gy), cu(idx), n))) < TOL


@pytest.mark.parametrize("b,c,n,m", [(2, 3, 1024, 256), (8, 64, 8192, 1024)])
def test_gather(b, c, n, m, ref_backend):
    g = rng(15)
    f = g.standard_normal((b, c, n), dtype=np.float32)
    idx = g.integers(0, n, size=(b, m)).astype(np.int32)
    out = B.gather_features_forward(cu(f), cu(idx))
    assert np.array_equal(npy(out), oracle.gather(f, idx))
    assert torch.equal(out, ref_backend.gather_features_forward(cu(f), cu(idx)))
    gy = g.standard_normal((b, c, m), dtype=np.float32)
    gx = B.gather_features_backward(cu(gy), cu(idx), n)
    assert rel_err(npy(gx), oracle.gather_grad(gy, idx, n)) < TOL
    assert rel_err(npy(gx), npy(ref_backend.gather_features_backward(cu(gy), cu(idx), n))) < TOL


@pytest.mark.parametrize("b,n,m", [(2, 1024, 256), (8, 8192, 1024), (3, 700, 64), (2, 300, 300), (1, 5000, 16),
                                   (1, 20000, 32), (4, 4096, 512), (2, 6000, 300), (2, 16384, 200), (3, 12000, 64)])
def test_fps(b, n, m, ref_backend):
    g = rng(16)
    co = g.random((b, 3, n), dtype=np.float32)
    if n in (700, 6000):  # quantised coordinates -> many exact distance ties
        co = np.round(co * 4) / 4
    idx0 = oracle.furthest_point_sampling(co, m)
    idx = B.furthest_point_sampling(cu(co), m)
    assert np.array_equal(npy(idx), idx0)
    assert torch.equal(idx, ref_backend.furthest_point_sampling(cu(co), m))
    # public API returns coordinates
    cc = F.furthest_point_sample(cu(co), m)
    assert np.array_equal(npy(cc), oracle.gather(co, idx0))


@pytest.mark.parametrize("b,c,n,m", [(2, 16, 1024, 256), (8, 64, 8192, 1024), (2, 5, 100, 2), (1, 4, 3000, 2500)])
def test_three_nn(b, c, n, m, ref_backend):
    g = rng(17)
    p = g.random((b, 3, n), dtype=np.float32)
    ce = g.random((b, 3, m), dtype=np.float32)
    f = g.standard_normal((b, c, m), dtype=np.float32)
    idx0, w0 = oracle.three_nn(p, ce)
    out0 = oracle.three_nn_interpolate(f, idx0, w0)
    out, idx, w = B.three_nearest_neighbors_interpolate_forward(cu(p), cu(ce), cu(f))
    assert np.array_equal(npy(idx), idx0)
    assert rel_err(npy(w), w0) < TOL and rel_err(npy(out), out0) < TOL
    ro, ri, rw = ref_backend.three_nearest_neighbors_interpolate_forward(cu(p), cu(ce), cu(f))
    assert torch.equal(idx, ri)
    assert rel_err(npy(w), npy(rw)) < TOL and rel_err(npy(out), npy(ro)) < TOL
    gy = g.standard_normal((b, c, n), dtype=np.float32)
    gx = B.three_nearest_neighbors_interpolate_backward(cu(gy), idx, w, m)
    assert rel_err(npy(gx), oracle.three_nn_interpolate_grad(gy, idx0, w0, m)) < TOL
    assert rel_err(npy(gx), npy(ref_backend.three_nearest_neighbors_interpolate_backward(cu(gy), ri, rw, m))) < TOL


def test_error_behaviour():
    """utils.hpp:7-18 semantics: RuntimeError with the reference's messages."""
    f = torch.zeros(1, 2, 8)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        B.avg_voxelize_forward(f, torch.zeros(1, 3, 8, dtype=torch.int32), 2)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        B.avg_voxelize_forward(f.cuda(), torch.zeros(1, 3, 8).cuda(), 2)
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        B.avg_voxelize_forward(torch.zeros(1, 8, 2).cuda().transpose(1, 2), torch.zeros(1, 3, 8, dtype=torch.int32).cuda(), 2)


def test_autograd_wrappers():
    g = rng(18)
    b, c, n, r = 2, 8, 512, 8
    co = s3dis_like_coords(g, b, n)
    f = torch.from_numpy(g.standard_normal((b, c, n), dtype=np.float32)).cuda().requires_grad_(True)
    nc, vc = F.voxelize_coords(cu(co), r)
    grid = F.avg_voxelize(f, vc, r)
    assert grid.shape == (b, c, r, r, r)
    out = F.trilinear_devoxelize(grid, nc, r, True)
    out.sum().backward()
    assert f.grad is not None and f.grad.shape == f.shape and torch.isfinite(f.grad).all()


@pytest.mark.parametrize("b,c,n,m,u", [(2, 32, 1024, 256, 32), (8, 9, 8192, 1024, 32), (1, 0, 50, 7, 3), (2, 5, 300, 33, 16)])
def test_group_concat(b, c, n, m, u):
    """Fused BallQuery grouping (modules/ball_query.py:16-30) == grouping(coords) - centres ++ grouping(features):
    forward bit-exact (copies and one fp32 subtraction), backward to features / point coords / centre coords."""
    g = rng(21)
    p = s3dis_like_coords(g, b, n)
    ce = np.ascontiguousarray(p[:, :, :m])
    f = g.standard_normal((b, c, n), dtype=np.float32) if c else None
    idx = oracle.ball_query(ce, p, 0.3, u)
    out = B.group_concat_forward(cu(p), cu(ce), None if f is None else cu(f), cu(idx))
    assert out.shape == (b, 3 + c, m, u)
    assert np.array_equal(npy(out), oracle.group_concat(p, ce, f, idx))
    # the composition the reference performs, on our stand-alone ops
    rel = F.grouping(cu(p), cu(idx)) - cu(ce).unsqueeze(-1)
    comp = rel if f is None else torch.cat([rel, F.grouping(cu(f), cu(idx))], dim=1)
    assert torch.equal(out, comp)
    gy = g.standard_normal((b, 3 + c, m, u), dtype=np.float32)
    gf, gp, gc = B.group_concat_backward(cu(gy), cu(idx), n, need_features=True, need_points=True, need_centers=True)
    of, op, oc = oracle.group_concat_grad(gy, idx, n)
    if c:
        assert rel_err(npy(gf), of) < TOL
    else:
        assert gf is None
    assert rel_err(npy(gp), op) < TOL
    assert rel_err(npy(gc), oc) < TOL


def test_ball_query_module_autograd():
    """BallQuery module (fused grouping) against the reference's tensor sequence under autograd."""
    import modules
    g = rng(22)
    b, c, n, m, u = 2, 16, 2048, 256, 32
    p = cu(s3dis_like_coords(g, b, n)).requires_grad_(True)
    ce = p.detach()[:, :, :m].clone().requires_grad_(True)
    f = cu(g.standard_normal((b, c, n), dtype=np.float32)).requires_grad_(True)
    gy = cu(g.standard_normal((b, 3 + c, m, u), dtype=np.float32))
    out = modules.BallQuery(0.25, u)(p, ce, f)
    out.backward(gy)
    got = [t.grad.clone() for t in (p, ce, f)]
    for t in (p, ce, f):
        t.grad = None
    idx = F.ball_query(ce.detach(), p.detach(), 0.25, u)
    ref = torch.cat([F.grouping(p, idx) - ce.unsqueeze(-1), F.grouping(f, idx)], dim=1)
    assert torch.equal(out, ref)
    ref.backward(gy)
    for a, t in zip(got, (p, ce, f)):
        assert rel_err(npy(a), npy(t.grad)) < TOL
