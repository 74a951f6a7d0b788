"""Oracle: farthest-point sampling (test infrastructure; see oracle/__init__.py).

Restates `NTU_RGBD_new.farthest_point_sampling_fast` (reference
training_code/cn3D_data_set.py:675-694; identical copies at cn3d_data_load.py:301-320 and
generate_data/generate_NTU.py:299-318) and `fps_sample_data` (cn3D_data_set.py:665-672).

Semantics pinned by tests/golden/fps_*.npz (generated from the reference function itself):
  * squared distance is ((dx*dx + dy*dy) + dz*dz) in the input dtype (numpy row-sum of 3 terms);
  * pick i (i >= 1) is the FIRST index attaining max(min_dist) (np.argmax);
  * min_dist is refreshed after every pick except the last one (cn3D_data_set.py:689);
  * the first pick is an argument here -- the reference draws it from the global numpy RNG
    (cn3D_data_set.py:679).
"""
import numpy as np


def farthest_point_sampling(pc, sample_num, start):
    """pc (N,3) float32/float64 -> (sample_num,) int32 pick order."""
    pc = np.asarray(pc)
    n = pc.shape[0]
    picks = np.empty(sample_num, dtype=np.int32)
    picks[0] = start
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]

    def sqdist_to(j):
        dx = x - x[j]
        dy = y - y[j]
        dz = z - z[j]
        return (dx * dx + dy * dy) + dz * dz

    best = sqdist_to(start)
    for i in range(1, sample_num):
        # first index attaining the maximum (what a strict '>' scan from index 0 keeps)
        j = int(np.flatnonzero(best == best.max())[0])
        picks[i] = j
        if i < sample_num - 1:
            best = np.minimum(best, sqdist_to(j))
    return picks


def fps_reorder_indices(picks, n):
    """Row permutation of cn3D_data_set.py:669-671: picks first, the rest ascending, truncated to n rows
    (`new_idx[:NUM_POINT]`: repeated picks -- a cloud with fewer than m distinct points -- leave more than
    n - m unpicked rows)."""
    mask = np.ones(n, dtype=bool)
    mask[picks] = False
    return np.concatenate([np.asarray(picks, dtype=np.int64), np.flatnonzero(mask)])[:n]


def fps_sample_data(points, sample_num, starts):
    """points (V,N,D): returns a reordered copy, FPS picks in rows [0, sample_num) of each cloud
    (cn3D_data_set.py:665-672; the reference writes in place)."""
    points = np.asarray(points)
    out = np.empty_like(points)
    for v in range(points.shape[0]):
        picks = farthest_point_sampling(points[v, :, 0:3], sample_num, int(starts[v]))
        out[v] = points[v, fps_reorder_indices(picks, points.shape[1])]
    return out
