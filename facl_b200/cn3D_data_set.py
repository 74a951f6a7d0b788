"""Device-side counterpart of the reference loader's per-sample work, `training_code/cn3D_data_set.py`
(SURVEY section 8 f1): the ten training views of a sequence and the FPS reordering, computed for a whole batch
on the GPU from source clouds that stay resident in HBM.  No CPU path -- CPU tensors raise.

  ViewAugmenter.get_data_train       NTU_RGBD_new.__getitem__ + get_data_train + get_temporal_augment_data
                                     (:105-121, :285-350, :654-663; helpers :708-713, :734-748, :765-776)
  farthest_point_sampling_fast       :675-694  (batched; the first pick is an argument, the reference draws it from
                                               numpy's global RNG)
  fps_sample_data                    :665-672

Differences a maintainer has to know about:
  * the reference draws from numpy's global Mersenne Twister inside 16 loader processes; here the draws are either
    passed in (`Draws`, bit-identical output to numpy for the same draws) or generated on the device by a
    counter-based Philox keyed by (seed, step) -- same distributions, different stream;
  * sources are float32 in HBM (the .npy files are float64 holding values that the train script casts to float32
    after augmentation, cn3d_train_motion_GL.py:228); the xyz arithmetic itself is done in float64 like numpy's;
  * the output is already G-major float32 (G*B, N, 4), i.e. the tensor `data1` of cn3d_train_motion_GL.py:226-228.
"""
import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib, ops
from ._lib import check, lib, ptr, require_cuda, stream_ptr

NUM_POINT = 512                                   # cn3D_data_set.py:24
SRC_POINTS, SRC_KEY, SRC_RES1, SRC_RES2 = 0, 1, 2, 3

# (source, channel -> column 3, nonzero_only, jitter, mirror, rotate): the views of get_data_train, :287-348
GET_DATA_TRAIN = [
    (SRC_POINTS, 3, 0, 0, 0, 0),   # raw_p
    (SRC_POINTS, 3, 0, 1, 1, 0),   # rev_p
    (SRC_KEY, 3, 0, 1, 0, 0),      # ke1_p
    (SRC_KEY, 3, 0, 1, 1, 0),      # ke2_p
    (SRC_POINTS, 3, 0, 1, 0, 1),   # ro1_p
    (SRC_POINTS, 3, 0, 1, 0, 1),   # ro2_p
    (SRC_POINTS, 4, 1, 0, 0, 0),   # ti1_p (time_seg2)
    (SRC_POINTS, 7, 1, 0, 0, 0),   # ti2_p (time_seg4)
    (SRC_RES1, 3, 0, 0, 0, 0),     # rs1_p
    (SRC_RES2, 3, 0, 0, 0, 0),     # rs2_p
]


@dataclass
class Ragged:
    """B variable-length clouds back to back: rows (sum P_b, C) fp32 cuda, offsets (B+1) int32 cuda."""
    rows: torch.Tensor
    offsets: torch.Tensor
    max_rows: int

    @staticmethod
    def from_list(clouds, device="cuda"):
        lens = [int(c.shape[0]) for c in clouds]
        rows = torch.cat([torch.as_tensor(c, dtype=torch.float32) for c in clouds]).to(device).contiguous()
        off = torch.tensor([0] + lens, dtype=torch.int64).cumsum(0).to(torch.int32).to(device)
        return Ragged(rows, off, max(lens))


@dataclass
class Draws:
    """Explicit random draws (all CUDA): idx (B,G,N) int32, noise (B,G,2,N,3) f64, angle_u (B,G) f64."""
    idx: torch.Tensor
    noise: torch.Tensor
    angle_u: torch.Tensor


class ViewAugmenter:
    def __init__(self, num_point=NUM_POINT, recipes=GET_DATA_TRAIN, sigma=0.01, clip=0.05, seed=1):
        self.N, self.recipes, self.sigma, self.clip, self.seed = num_point, list(recipes), sigma, clip, seed
        self.step = 0

    def get_data_train(self, sources, draws=None, g_major=True, want_rows=False):
        """sources: list of Ragged (points, key_points, res_points_1, res_points_2 for GET_DATA_TRAIN), one entry per
        sequence in each.  -> views (G*B, N, 4) fp32 [g-major] or (B, G, N, 4); optionally the source rows drawn."""
        G, N = len(self.recipes), self.N
        B = int(sources[0].offsets.numel()) - 1
        dev = sources[0].rows.device
        src = (_lib.PointSource * len(sources))()
        for i, s in enumerate(sources):
            require_cuda(s.rows, "source rows")
            require_cuda(s.offsets, "source offsets", torch.int32)
            assert s.rows.is_contiguous() and s.offsets.numel() == B + 1
            src[i] = _lib.PointSource(s.rows.data_ptr(), s.offsets.data_ptr(), s.rows.shape[1])
        rec = (_lib.ViewRecipe * G)(*[_lib.ViewRecipe(*r) for r in self.recipes])
        out = torch.empty((G * B, N, 4) if g_major else (B, G, N, 4), dtype=torch.float32, device=dev)
        rows = torch.empty((B, G, N), dtype=torch.int32, device=dev) if want_rows else None
        a = _lib.AugmentArgs()
        a.B, a.G, a.N, a.n_sources = B, G, N, len(sources)
        a.sources, a.recipes = src, rec
        a.sigma, a.clip = self.sigma, self.clip
        if draws is not None:
            idx = require_cuda(draws.idx, "idx", torch.int32).contiguous()
            noise = require_cuda(draws.noise, "noise", torch.float64).contiguous()
            ang = require_cuda(draws.angle_u, "angle_u", torch.float64).contiguous()
            assert idx.shape == (B, G, N) and noise.shape == (B, G, 2, N, 3) and ang.shape == (B, G)
            a.idx, a.noise, a.angle_u = idx.data_ptr(), noise.data_ptr(), ang.data_ptr()
        a.seed, a.step = self.seed, self.step
        a.g_major = int(g_major)
        a.max_rows = max(s.max_rows for s in sources)
        a.out, a.out_rows = out.data_ptr(), (rows.data_ptr() if want_rows else None)
        check(lib().facl_augment_views(C.byref(a), stream_ptr()), "facl_augment_views")
        self.step += 1
        return (out, rows) if want_rows else out


def farthest_point_sampling_fast(pc, sample_num, start_idx):
    """pc (V,N,3+) fp32 cuda, start_idx (V) int32 -> (V, sample_num) int32 (reference :675-694 returns (m,1) per cloud)."""
    return ops.fps(pc, sample_num, start_idx)


def fps_sample_data(points_xyzc, sample_num_level1, sample_num_level2, start_idx):
    """reference :665-672: FPS picks first, the remaining rows in ascending index.  Returns a new tensor."""
    picks = ops.fps(points_xyzc, sample_num_level1, start_idx)
    return ops.fps_reorder(points_xyzc, picks)
