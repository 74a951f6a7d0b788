"""Synthetic NTU-shaped point-cloud sequences (host side, numpy).

The reference trains on NTU-RGB+D 3DV clouds produced by generate_data/generate_NTU.py, which we do
not have.  That generator normalises each cloud by its y-extent (generate_NTU.py:231-247), so
y spans a unit interval and x/z are narrower, and stores 4 channels: xyz + a motion/appearance
value (:259-260).  This module draws clouds with that value distribution:

    y ~ U(-0.5, 0.5),  x ~ 0.45*U(-0.5, 0.5),  z ~ 0.30*U(-0.5, 0.5),  c ~ U(-0.5, 0.5)

optionally concentrated on a 15-blob "skeleton" mixture, and optionally resampled with replacement
(what the live loader does, cn3D_data_set.py:287), which creates exact duplicate points.
"""
import numpy as np


def make_sequences(B, G, N, D=4, seed=1, skeleton=False, resample=False, dtype=np.float32):
    """Returns points (B, G, N, D), the layout the reference DataLoader yields
    (cn3d_train_motion_GL.py:224-225)."""
    rng = np.random.default_rng(seed)
    scale = np.array([0.45, 1.0, 0.30], dtype=np.float64)
    if skeleton:
        joints = (rng.random((B, 1, 15, 3)) - 0.5) * scale
        joints = joints + 0.02 * rng.standard_normal((B, G, 15, 3))
        pick = rng.integers(0, 15, size=(B, G, N))
        xyz = np.take_along_axis(joints, pick[..., None].repeat(3, -1), axis=2)
        xyz = xyz + 0.04 * rng.standard_normal((B, G, N, 3)) * scale
        xyz = np.clip(xyz, -0.5 * scale, 0.5 * scale)
    else:
        xyz = (rng.random((B, G, N, 3)) - 0.5) * scale
    feat = rng.random((B, G, N, D - 3)) - 0.5
    pts = np.concatenate([xyz, feat], axis=-1)
    if resample:
        take = rng.integers(0, N, size=(B, G, N))
        pts = np.take_along_axis(pts, take[..., None].repeat(D, -1), axis=2)
    return np.ascontiguousarray(pts.astype(dtype))


def g_major(points_bgnd):
    """(B,G,N,D) -> (G*B, N, D): row g*B+b is view g of sample b (cn3d_train_motion_GL.py:225-226)."""
    B, G, N, D = points_bgnd.shape
    return np.ascontiguousarray(points_bgnd.transpose(1, 0, 2, 3).reshape(G * B, N, D))


def view_order(G, seed=1):
    """The circle loss's view permutation, drawn the way the reference does
    (np.random.shuffle of arange(G), cn3d_train_motion_GL.py:297-298) but from a private RNG."""
    order = np.arange(G)
    np.random.RandomState(seed).shuffle(order)
    return order
