"""Generates tests/golden/ref_ops_golden.npz by running the UNMODIFIED reference CUDA kernels
(oracle/_ref/_pvcnn_backend.so, compiled from /root/reference by oracle/build_ref.py) on seeded
inputs.  Must run on a GPU box:  python tests/golden/make_golden.py gpurun_out/ref_ops_golden.npz
The committed .npz pins the CPU oracle to real reference outputs (tests/test_golden_cpu.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.ref_gpu import backend  # noqa: E402
from util import rng, s3dis_like_coords, surface_coords  # noqa: E402


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main(path):
    be = backend()
    g = rng(100)
    out = {}
    # --- config 1 of BASELINE.json: B=2 N=1024 C=16 R=8
    b, c, n, r = 2, 16, 1024, 8
    co = s3dis_like_coords(g, b, n)
    t = torch.from_numpy(co).cuda()
    nc = t - t.mean(2, keepdim=True)
    nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0) + 0.5
    nc = torch.clamp(nc * r, 0, r - 1)
    vc = torch.round(nc).to(torch.int32)
    f = g.standard_normal((b, c, n), dtype=np.float32)
    o, ind, cnt = be.avg_voxelize_forward(cu(f), vc.contiguous(), r)
    gy = g.standard_normal((b, c, r ** 3), dtype=np.float32)
    gx = be.avg_voxelize_backward(cu(gy), ind, cnt)
    out.update(vox_coords_in=co, vox_norm=nc.cpu().numpy(), vox_vc=vc.cpu().numpy(), vox_feat=f,
               vox_out=o.cpu().numpy(), vox_ind=ind.cpu().numpy(), vox_cnt=cnt.cpu().numpy(), vox_gy=gy,
               vox_gx=gx.cpu().numpy())
    grid = g.standard_normal((b, c, r ** 3), dtype=np.float32)
    do, di, dw = be.trilinear_devoxelize_forward(r, True, nc.contiguous(), cu(grid))
    dgy = g.standard_normal((b, c, n), dtype=np.float32)
    dgx = be.trilinear_devoxelize_backward(cu(dgy), di, dw, r)
    out.update(devox_grid=grid, devox_out=do.cpu().numpy(), devox_inds=di.cpu().numpy(), devox_wgts=dw.cpu().numpy(),
               devox_gy=dgy, devox_gx=dgx.cpu().numpy())
    # --- surface distribution (heavy voxel sharing), small
    co2 = surface_coords(g, 1, 512)
    out["surf_coords"] = co2
    # --- PointNet++ ops
    bn, nn_, m, u = 2, 600, 96, 16
    p = g.random((bn, 3, nn_), dtype=np.float32)
    p[1] = np.round(p[1] * 8) / 8  # exact ties for FPS / ball query
    fidx = be.furthest_point_sampling(cu(p), m)
    centers = be.gather_features_forward(cu(p), fidx)
    bq = be.ball_query(centers, cu(p), 0.25, u)
    pf = g.standard_normal((bn, 7, nn_), dtype=np.float32)
    grp = be.grouping_forward(cu(pf), bq)
    ggy = g.standard_normal((bn, 7, m, u), dtype=np.float32)
    ggx = be.grouping_backward(cu(ggy), bq, nn_)
    cf = g.standard_normal((bn, 5, m), dtype=np.float32)
    io, ii, iw = be.three_nearest_neighbors_interpolate_forward(cu(p), centers, cu(cf))
    igy = g.standard_normal((bn, 5, nn_), dtype=np.float32)
    igx = be.three_nearest_neighbors_interpolate_backward(cu(igy), ii, iw, m)
    out.update(pn_points=p, pn_fps=fidx.cpu().numpy(), pn_centers=centers.cpu().numpy(), pn_bq=bq.cpu().numpy(),
               pn_feat=pf, pn_group=grp.cpu().numpy(), pn_ggy=ggy, pn_ggx=ggx.cpu().numpy(), pn_cf=cf,
               pn_interp=io.cpu().numpy(), pn_iidx=ii.cpu().numpy(), pn_iw=iw.cpu().numpy(), pn_igy=igy,
               pn_igx=igx.cpu().numpy())
    torch.cuda.synchronize()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_ops_golden.npz")
