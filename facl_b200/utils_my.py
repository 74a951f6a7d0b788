"""Drop-in for the reference module `training_code/utils_my.py` (the functions on the hot path).

Same names, argument meaning, return shapes and `opt` side effects as the reference; the work is done by
libfacl_b200.so on the GPU (no CPU path -- CPU tensors raise).

  group_points_3DV / _2048 / _nums / group_points   reference utils_my.py:255-291 / :7-42 / :293-328 / :217-253
  group_points_2 / group_points_2_3DV               reference utils_my.py:332-356 / :358-381 (level-2 set abstraction)
  global_contrast / circle_contrast / Info_NCE      reference utils_my.py:53-83 / :85-116 / :200-213
  CLD_Loss / grouping / KMeans                      reference utils_my.py:152-198 (facl_b200/heads.py)
"""
import numpy as np
import torch

from . import _lib, ops
from .losses import contrast_losses, info_nce_logits_cuda
from .heads import CLD_Loss, KMeans, grouping  # noqa: F401  (reference utils_my.py:152-198; disabled in the scripts by cld_if = 0)


def _group(points, S, K, r2, N=None):
    _lib.require_cuda(points, "points")
    M = points.shape[0]
    pts = points.reshape(M, -1, points.shape[-1]) if N is None else points.reshape(M, N, -1)
    rows, _ = ops.group_points_raw(pts, S, K, r2, want_idx=False)
    # reference return views (utils_my.py:283-284): xt (M,D,S,K) over the [M][S][K][D] buffer, yt (M,3,S,1)
    inputs_level1 = rows.permute(0, 3, 1, 2)
    centre = pts[:, 0:S, 0:3].contiguous()
    inputs_level1_center = centre.view(-1, 1, S, 3).transpose(1, 3)
    return inputs_level1, inputs_level1_center


def group_points_3DV(points, opt):
    """reference utils_my.py:255-291.  Side effects kept: overwrites opt.INPUT_FEATURE_NUM / knn_K / ball_radius
    (:259-261), which makes the CLI radius and the per-step radius jitter of the train script dead."""
    opt.INPUT_FEATURE_NUM = points.shape[-1]
    opt.knn_K = 64
    opt.ball_radius = 0.06
    return _group(points, opt.sample_num_level1, opt.knn_K, opt.ball_radius, N=opt.SAMPLE_NUM)


def group_points_3DV_2048(points, knn_K, sample_num_level1, SAMPLE_NUM=2048):
    """reference utils_my.py:7-42 (squared radius hard-coded to 0.16, :13)."""
    return _group(points, sample_num_level1, knn_K, 0.16)


def group_points_3DV_nums(points, opt, sample_num_level1, knn_K):
    """reference utils_my.py:293-328 (sets opt.ball_radius = 0.06, :299)."""
    opt.INPUT_FEATURE_NUM = points.shape[-1]
    opt.ball_radius = 0.06
    return _group(points, sample_num_level1, knn_K, opt.ball_radius, N=opt.SAMPLE_NUM)


def group_points(points, opt):
    """reference utils_my.py:217-253 (sets opt.knn_K = 64, opt.ball_radius = 0.14)."""
    opt.knn_K = 64
    opt.ball_radius = 0.14
    opt.INPUT_FEATURE_NUM = points.shape[-1]
    return _group(points, opt.sample_num_level1, opt.knn_K, opt.ball_radius, N=opt.SAMPLE_NUM)


def group_points_2(points, sample_num_level1, sample_num_level2, knn_K, ball_radius):
    """reference utils_my.py:332-356.  points (B, 3+C, S1) channel-first.  As in the reference, knn_K is overridden
    to 64 (:335) and `ball_radius` is compared against SQUARED distances as given (:345).
    -> inputs_level2 (B, 3+C, S2, 64), inputs_level2_center (B, 3, S2, 1)."""
    out, _ = ops.group_level2(points, sample_num_level2, 64, float(ball_radius))
    return out, points[:, 0:3, 0:sample_num_level2].unsqueeze(3)


def group_points_2_3DV(points, sample_num_level1, sample_num_level2, knn_K, ball_radius):
    """reference utils_my.py:358-381: knn_K = 32 and ball_radius = 0.11 are hard-coded (:361-362)."""
    out, _ = ops.group_level2(points, sample_num_level2, 32, 0.11)
    return out, points[:, 0:3, 0:sample_num_level2].unsqueeze(3)


def global_contrast(num_crop, x_global, x, opt, criterion=None):
    """reference utils_my.py:53-83.  `criterion` must be CrossEntropyLoss with mean reduction (the only one the
    reference passes, cn3d_train_motion_GL.py:179); it is accepted for signature parity and not called."""
    loss_c, _ = contrast_losses(x, x_global, num_crop, opt.batchSize, order=None, want_global=True, want_circle=False)
    return loss_c


def circle_contrast(num_crop, x, batchSize, criterion=None, order=None):
    """reference utils_my.py:85-116.  The view order is drawn like the reference does (np.random.shuffle of
    arange(num_crop) from numpy's global RNG, :96-97) unless `order` is given."""
    if order is None:
        order = np.arange(0, num_crop, 1)
        np.random.shuffle(order)
    _, loss_circle = contrast_losses(x, None, num_crop, batchSize, order=order, want_global=False, want_circle=True)
    return loss_circle


def Info_NCE(x, opt):
    """reference utils_my.py:200-213 (unused by the live scripts): logits (B, 1+4B) and zero labels."""
    return info_nce_logits_cuda(x, opt.batchSize)
