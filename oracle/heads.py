"""Oracle: the disabled loss heads (test infrastructure; see oracle/__init__.py).

Restates, in torch-CPU ops with autograd, what the reference does in
  cn3d_model_conbag.py:391-425   distributed_sinkhorn, shoot_infs
  cn3d_train_motion_GL.py:236-262 the inline SwAV block (swa_if = 0; queue_length = 0 at :186, so `queue` is None)
  utils_my.py:152-198            CLD_Loss, grouping, KMeans (identical copies at cn3d_train_motion_GL.py:36-70)
Pinned by tests/golden/heads.npz, written from the reference functions themselves (tests/golden/make_golden.py: gen_heads).
"""
import numpy as np
import torch


def shoot_infs(t):
    """cn3d_model_conbag.py:409-425"""
    t = t.clone()
    mask = torch.isinf(t)
    if bool(mask.any()):
        t[mask] = 0
        t[mask] = t.max()
    return t


def distributed_sinkhorn(Q, nmb_iters):
    """cn3d_model_conbag.py:391-406; Q (K, B) -> (B, K)"""
    with torch.no_grad():
        Q = shoot_infs(Q)
        Q = Q / torch.sum(Q)
        r = torch.ones(Q.shape[0], dtype=Q.dtype) / Q.shape[0]
        c = torch.ones(Q.shape[1], dtype=Q.dtype) / Q.shape[1]
        for _ in range(nmb_iters):
            u = shoot_infs(r / torch.sum(Q, dim=1))
            Q = Q * u.unsqueeze(1)
            Q = Q * (c / torch.sum(Q, dim=0)).unsqueeze(0)
        return (Q / torch.sum(Q, dim=0, keepdim=True)).t().float()


def swav_loss(code, num_crop, B, epsilon=0.03, temperature=0.1, iters=3):
    """cn3d_train_motion_GL.py:240-261 with queue = None"""
    loss_swa = 0
    for crop_id in range(num_crop - 1):
        with torch.no_grad():
            po = code[B * crop_id: B * (crop_id + 1), :] / epsilon
            q = distributed_sinkhorn(torch.exp(po).t(), iters)[-B:]
        subloss = 0
        for v in np.delete(np.arange(num_crop - 1), crop_id):
            p = torch.softmax(code[B * v: B * (v + 1)] / temperature, dim=1)
            subloss = subloss - torch.mean(torch.sum(q * torch.log(p), dim=1))
        loss_swa = loss_swa + subloss
    return loss_swa / (num_crop - 1)


def kmeans(x, K=10, Niters=10):
    """utils_my.py:180-198: -> (labels of the last assignment, centroids after the last update); differentiable in x through
    the last scatter-add, as in the reference."""
    N, D = x.shape
    c = x[:K, :].clone()
    for _ in range(Niters):
        D_ij = ((x[:, None, :] - c[None, :, :]) ** 2).sum(-1)
        cl = D_ij.argmin(dim=1).long().view(-1)
        counts = torch.bincount(cl, minlength=K).clamp_min(1)          # empty clusters divide by 1 (:191-193)
        c = torch.zeros(K, D, dtype=x.dtype).index_add_(0, cl, x) / counts.to(x.dtype).unsqueeze(1)
    return cl, c


def grouping(f1, f2, T, k_eigen, clusters, num_iters):
    """utils_my.py:165-178"""
    ce = torch.nn.CrossEntropyLoss()
    l1, c1 = kmeans(f1, clusters, num_iters)
    l2, c2 = kmeans(f2, clusters, num_iters)
    loss = ce(torch.mm(f1, c2.t()) / T, l2)
    return (loss + ce(torch.mm(f2, c1.t()) / T, l1)) / 2


def cld_loss(num_crop, x_nor, B):
    """utils_my.py:152-162"""
    total = 0
    for i in range(num_crop - 4):
        total = total + grouping(x_nor[i * B: (i + 3) * B], x_nor[(i + 1) * B: (i + 4) * B], 0.05, 10, 60, 5)
    return total
