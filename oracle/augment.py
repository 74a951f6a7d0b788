"""CPU restatement of the training-view augmentation (TEST INFRASTRUCTURE ONLY -- never on the product path).

Follows /root/reference/training_code/cn3D_data_set.py:
  * get_temporal_augment_data   :654-663   rows with a non-zero temporal channel, resampled with replacement
  * get_data_train              :285-350   the ten views of one sequence
  * reverse_transform           :708-713   round to f32, x -> -x, jitter again, round to f32
  * rotate_trans                :734-748   round to f32, rotate about y by (u - 0.5) * 0.8 * pi, round to f32
  * jitter_point_cloud          :765-776   xyz += clip(sigma * z, +-clip), computed in f64
  * NTU_RGBD_new.__getitem__    :105-121   call order (and with it the order the numpy RNG is consumed in)

The reference draws from the global numpy RNG; here every draw is an explicit argument (`Draws`), and
`record_draws()` replays the reference's consumption order on a seeded RandomState so that the restatement can
be pinned against the reference itself (tests/golden/augment.npz, written by tests/golden/make_golden.py).

Pinned: yes -- tests/test_oracle_golden.py::test_augment_matches_reference.
"""
from dataclasses import dataclass

import numpy as np

SIGMA, CLIP = 0.01, 0.05

# view recipe: (source, channel copied to column 3, nonzero_only, jitter, mirror, rotate)
SRC_POINTS, SRC_KEY, SRC_RES1, SRC_RES2 = 0, 1, 2, 3
GET_DATA_TRAIN = [
    (SRC_POINTS, 3, 0, 0, 0, 0),   # raw_p
    (SRC_POINTS, 3, 0, 1, 1, 0),   # rev_p
    (SRC_KEY, 3, 0, 1, 0, 0),      # ke1_p
    (SRC_KEY, 3, 0, 1, 1, 0),      # ke2_p
    (SRC_POINTS, 3, 0, 1, 0, 1),   # ro1_p
    (SRC_POINTS, 3, 0, 1, 0, 1),   # ro2_p
    (SRC_POINTS, 4, 1, 0, 0, 0),   # ti1_p = get_temporal_augment_data(points, 4)
    (SRC_POINTS, 7, 1, 0, 0, 0),   # ti2_p = get_temporal_augment_data(points, 7)
    (SRC_RES1, 3, 0, 0, 0, 0),     # rs1_p
    (SRC_RES2, 3, 0, 0, 0, 0),     # rs2_p
]


@dataclass
class Draws:
    idx: np.ndarray       # (G, N) int32   resample indices (into the non-zero rows for nonzero_only views)
    noise: np.ndarray     # (G, 2, N, 3) f64 standard normals: [0] jitter, [1] the mirror's second jitter
    angle_u: np.ndarray   # (G,) f64 uniforms in [0, 1)


def nonzero_rows(src, channel):
    return np.nonzero(src[:, channel] != 0)[0]


def record_draws(rs, sources, N, recipes=GET_DATA_TRAIN):
    """Replay the order in which __getitem__ (:116-119) + get_data_train (:287-318) consume the numpy RNG `rs`."""
    G = len(recipes)
    d = Draws(np.zeros((G, N), np.int32), np.zeros((G, 2, N, 3)), np.zeros(G))
    order = [g for g, r in enumerate(recipes) if r[2]] + [g for g, r in enumerate(recipes) if not r[2]]
    for g in order:
        s, ch, nz, jit, mir, rot = recipes[g]
        count = len(nonzero_rows(sources[s], ch)) if nz else sources[s].shape[0]
        d.idx[g] = rs.randint(0, count, N)
        if jit:
            d.noise[g, 0] = rs.randn(1, N, 3)[0]
        if mir:
            d.noise[g, 1] = rs.randn(1, N, 3)[0]
        if rot:
            d.angle_u[g] = rs.rand()
    return d


def _jitter(xyz, z):
    return np.clip(SIGMA * z, -CLIP, CLIP) + xyz


def make_views(sources, draws, recipes=GET_DATA_TRAIN):
    """sources: list of (P_s, C_s) arrays for ONE sequence -> (G, N, 4) float32 views."""
    G, N = draws.idx.shape
    out = np.empty((G, N, 4), np.float32)
    for g, (s, ch, nz, jit, mir, rot) in enumerate(recipes):
        src = np.asarray(sources[s], np.float64)
        rows = draws.idx[g].astype(np.int64)
        if nz:
            rows = nonzero_rows(src, ch)[rows]
        xyz = src[rows, :3].copy()
        if jit:
            xyz = _jitter(xyz, draws.noise[g, 0])
        if mir:
            xyz = xyz.astype(np.float32)
            xyz[:, 0] = -xyz[:, 0]
            xyz = _jitter(xyz.astype(np.float64), draws.noise[g, 1])
        if rot:
            xyz = xyz.astype(np.float32).astype(np.float64)
            a = (draws.angle_u[g] - 0.5) * np.pi * 0.8
            c, sn = np.cos(a), np.sin(a)
            x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
            xyz = np.stack([x * c + y * 0.0 + z * (-sn), x * 0.0 + y * 1.0 + z * 0.0, x * sn + y * 0.0 + z * c], 1)
        out[g, :, :3] = xyz.astype(np.float32)
        out[g, :, 3] = src[rows, ch].astype(np.float32)
    return out
