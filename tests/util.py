"""Shared input generators for the parity tests (seeded, config-by-config)."""
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
