"""Shared input generators for the parity tests (seeded, config-by-config)."""
import os

import numpy as np

SEED = 1588147245  # the reference's own seed (configs/__init__.py:3)


def rng(extra=0):
    return np.random.default_rng(SEED + extra)


def s3dis_like_coords(g, b, n):
    """coords ~ U(0,1.5) x U(0,1.5) x U(0,3.0)  (S3DIS block, data/s3dis/prepare_data.py:89)"""
    c = g.random((b, 3, n), dtype=np.float32)
    c[:, 0] *= 1.5
    c[:, 1] *= 1.5
    c[:, 2] *= 3.0
    return c


def surface_coords(g, b, n):
    """points on three axis-aligned planes: heavy voxel sharing"""
    c = s3dis_like_coords(g, b, n)
    which = g.integers(0, 3, size=(b, n))
    for a in range(3):
        c[:, a][which == a] = 0.25
    return c


def degenerate_coords(g, b, n):
    """all points in <= 8 voxels"""
    c = np.zeros((b, 3, n), np.float32)
    c += (g.integers(0, 2, size=(b, 3, n)) * 1.0).astype(np.float32)
    c += g.random((b, 3, n), dtype=np.float32) * 1e-3
    return c


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def device_mean(coords):
    """`coords.mean(2, keepdim=True)` (modules/voxelization.py:18) evaluated by torch on the device under test -> [B,3].
    The summation order of that reduction belongs to torch; oracle and product both take its result."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float32))
    if torch.cuda.is_available():
        t = t.cuda()
    return t.mean(2, keepdim=True).cpu().numpy().reshape(-1, 3)


def reference_voxelization(coords_t, r, normalize=True, eps=0.0):
    """The literal tensor program of modules/voxelization.py:17-24 on whatever device `coords_t` lives."""
    import torch
    coords_t = coords_t.detach()
    norm_coords = coords_t - coords_t.mean(2, keepdim=True)
    if normalize:
        norm_coords = norm_coords / (norm_coords.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps) + 0.5
    else:
        norm_coords = (norm_coords + 1) / 2.0
    norm_coords = torch.clamp(norm_coords * r, 0, r - 1)
    return norm_coords, torch.round(norm_coords).to(torch.int32)


def host_threads(cap=32):
    """Threads for the CPU references: the scheduler affinity and the cgroup CPU quota bound what this process can really
    use (os.cpu_count() reports the whole node), and more than ~32 does not help the sizes tested here."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return max(1, min(n, cap))
