"""Oracle: linear probe on the extracted features and the feature-file layout (test infrastructure only).

Follows
  * training_code/extract_motion_feature.py:217-221  `save_single_feature`: the (num_crop * B, 512) encoder output
    (views g-major, then the sequence-level block) -> one (num_crop * 512,) float32 .npy per video;
  * linear_classify/dataset_of_lin.py:103-105  a sample = concatenate(motion file, appearance file) -> (22 * 512,);
  * linear_classify/fc_model.py:12-25          `Final_FC`: F.normalize(x, p=2, dim=1) -> Linear(22 * 512, 120),
                                               weight ~ N(0, 0.01), bias 0;
  * linear_classify/linercls.py:86-124         CrossEntropyLoss (mean), Adam(lr 0.005, betas (0.5, 0.999), eps 1e-6),
                                               top-1 accuracy in percent (:158-172).
Pinned: tests/golden/probe.npz, written by running the reference's Final_FC / save_single_feature / accuracy.
"""
import numpy as np
import torch

from .train_step import adam_update


def feature_rows(feature, num_crop=11):
    """extract_motion_feature.py:218: (num_crop * B, 512) -> (B, num_crop * 512), one row per video."""
    feature = np.asarray(feature)
    return feature.reshape(num_crop, -1, 512).transpose(1, 0, 2).reshape(-1, num_crop * 512)


def final_fc(x, weight, bias):
    """fc_model.py:21-25."""
    xn = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return xn @ weight.t() + bias


def probe_step(sd, x, labels, adam_state, lr=0.005):
    """One iteration of linercls.py:109-122.  sd = {"fc.weight", "fc.bias"} (updated in place).
    -> dict(loss, top1 [percent], logits, grads)."""
    w = sd["fc.weight"].clone().requires_grad_(True)
    b = sd["fc.bias"].clone().requires_grad_(True)
    logits = final_fc(x, w, b)
    lse = torch.logsumexp(logits, dim=1)
    loss = (lse - logits.gather(1, labels[:, None])[:, 0]).mean()
    loss.backward()
    grads = {"fc.weight": w.grad.detach(), "fc.bias": b.grad.detach()}
    top1 = float((logits.argmax(dim=1) == labels).float().sum() * (100.0 / x.shape[0]))
    adam_update(sd, grads, adam_state, lr=lr)
    return dict(loss=float(loss.detach()), top1=top1, logits=logits.detach(), grads=grads)
