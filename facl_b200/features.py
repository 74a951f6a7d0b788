"""Feature-file wire format between extraction and the linear probe.

  save_single_feature   training_code/extract_motion_feature.py:217-221: one `<name>.npy` per video holding the
                        (num_crop * 512,) float32 row [view 0 | ... | view G-1 | sequence-level], numpy .npy v1.0
  load_sample           linear_classify/dataset_of_lin.py:103-105: concatenate(motion file, appearance file)
  load_batch            a batch of those as one pinned host array, ready for an async H2D copy
"""
import os

import numpy as np
import torch


def feature_rows(feature, num_crop=11):
    """(num_crop * B, 512) device or host tensor [x ; x_global] -> (B, num_crop * 512) host float32 array."""
    if isinstance(feature, torch.Tensor):
        B = feature.shape[0] // num_crop
        feature = feature.detach().reshape(num_crop, B, 512).permute(1, 0, 2).reshape(B, num_crop * 512)   # on the device
        host = torch.empty(feature.shape, dtype=torch.float32, pin_memory=feature.is_cuda)
        host.copy_(feature)
        return host.numpy()
    return np.asarray(feature).reshape(num_crop, -1, 512).transpose(1, 0, 2).reshape(-1, num_crop * 512)


def save_single_feature(feature, save_path, name, num_crop=11):
    rows = feature_rows(feature, num_crop)
    for batch_i in range(rows.shape[0]):
        np.save(save_path + name[batch_i] + '.npy', rows[batch_i])


def load_sample(motion_path, appearance_path):
    return np.concatenate((np.load(motion_path), np.load(appearance_path)), 0)


def load_batch(pairs, pin=True):
    """pairs: list of (motion_path, appearance_path) -> (len(pairs), F) float32 host tensor (pinned if possible)."""
    first = load_sample(*pairs[0])
    out = torch.empty((len(pairs), first.shape[0]), dtype=torch.float32)
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    out[0] = torch.from_numpy(first)
    for i, p in enumerate(pairs[1:], 1):
        out[i] = torch.from_numpy(load_sample(*p))
    return out
