"""The loss heads the reference scripts carry but switch off by constants, on libfacl_b200.

  distributed_sinkhorn, shoot_infs   reference training_code/cn3d_model_conbag.py:391-425
  swav_loss                          the inline block of training_code/cn3d_train_motion_GL.py:236-262 (swa_if = 0, queue disabled)
  KMeans, grouping, CLD_Loss         reference training_code/utils_my.py:152-198 (= cn3d_train_motion_GL.py:36-70, cld_if = 0)
  normalized_code                    x -> (x_nor, code) of cn3d_model_conbag.py:231-232 WITH autograd (the encoder returns them detached,
                                     because no live reference loss reads them)

Same names, argument meaning and return values as the reference.  The work runs on the GPU through the C ABI: facl_sinkhorn,
facl_soft_xent, facl_kmeans, facl_softmax_xent, facl_l2_normalize and the tensor-core GEMM facl_gemm_tc (mapping / affinity
products and their gradients); torch is used for tensor plumbing and for the autograd graph only.  CUDA tensors only.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, lib, ptr, stream_ptr


class _MatmulNT(torch.autograd.Function):
    """out = a @ b.T on facl_gemm_tc (bf16x3 split products, fp32 accumulation); a (R, K), b (S, K) -> (R, S)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        R, K = a.shape
        S = b.shape[0]
        out = torch.empty((R, S), dtype=torch.float32, device=a.device)
        ops.gemm_tc(R, S, K, nsplit=3, a=dict(src0=a, ld=K), b_mode=ops.B_ROWMAJOR, b=dict(src0=b, ld=K), out_mode=ops.OUT_CHMAJOR,
                    out=out, ldo=S)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        dout = dout.contiguous()
        R, K = a.shape
        S = b.shape[0]
        da = torch.empty_like(a)          # da = dout @ b      : D[r][k] = sum_s dout[r][s] b[s][k]
        ops.gemm_tc(R, K, S, nsplit=3, a=dict(src0=dout, ld=S), b_mode=ops.B_CHMAJOR, b=dict(src0=b, ld=K), out_mode=ops.OUT_CHMAJOR,
                    out=da, ldo=K)
        dout_t = dout.t().contiguous()
        db = torch.empty_like(b)          # db = dout.T @ a    : D[s][k] = sum_r dout[r][s] a[r][k]
        ops.gemm_tc(S, K, R, nsplit=3, a=dict(src0=dout_t, ld=R), b_mode=ops.B_CHMAJOR, b=dict(src0=a, ld=K), out_mode=ops.OUT_CHMAJOR,
                    out=db, ldo=K)
        return da, db


def matmul_nt(a, b):
    _lib.require_cuda(a, "a")
    _lib.require_cuda(b, "b")
    return _MatmulNT.apply(a, b)


class _L2Normalize(torch.autograd.Function):
    """F.normalize(x, p=2, dim=1) (eps 1e-12) on facl_l2_normalize; backward dx = (dy - y (y . dy)) / max(|x|, eps)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib().facl_l2_normalize(ptr(x), x.shape[0], x.shape[1], ptr(y), stream_ptr()), "facl_l2_normalize")
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        nrm = x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        return (dy - y * (y * dy).sum(dim=1, keepdim=True)) / nrm


def normalized_code(x, mapping_weight):
    """x (M, 512), mapping.weight (64, 512) -> (x_nor, code) as cn3d_model_conbag.py:231-232, differentiable."""
    _lib.require_cuda(x, "x")
    x_nor = _L2Normalize.apply(x)
    return x_nor, matmul_nt(x_nor, mapping_weight)


# ------------------------------------------------------------------------------------------------ SwAV / Sinkhorn
def shoot_infs(inp_tensor):
    """reference cn3d_model_conbag.py:409-425: infinities -> the maximum of the tensor with them zeroed (in place)."""
    mask = torch.isinf(inp_tensor)
    if bool(mask.any()):
        inp_tensor[mask] = 0
        inp_tensor[mask] = inp_tensor.max()
    return inp_tensor


def distributed_sinkhorn(Q, nmb_iters):
    """reference cn3d_model_conbag.py:391-406.  Q (K, B) fp32 CUDA -> (B, K); no gradient (the reference runs it under no_grad)."""
    _lib.require_cuda(Q, "Q")
    with torch.no_grad():
        Q = Q.contiguous()
        K, B = Q.shape
        out = torch.empty((B, K), dtype=torch.float32, device=Q.device)
        check(lib().facl_sinkhorn(ptr(Q), K, B, int(nmb_iters), ptr(out), stream_ptr()), "facl_sinkhorn")
    return out


class _SoftXent(torch.autograd.Function):
    """-mean_r sum_k q[r][k] log softmax(scale * logits[r])[k]; gradient w.r.t. logits only (q is a constant target)."""

    @staticmethod
    def forward(ctx, logits, q, scale):
        logits, q = logits.contiguous(), q.contiguous()
        rows, K = logits.shape
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        dlog = torch.empty_like(logits)
        check(lib().facl_soft_xent(ptr(logits), ptr(q), rows, K, float(scale), ptr(loss), ptr(dlog), stream_ptr()), "facl_soft_xent")
        ctx.save_for_backward(dlog)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dlog,) = ctx.saved_tensors
        return dlog * g, None, None


def swav_loss(code, num_crop, batchSize, epsilon=0.03, temperature=0.1, sinkhorn_iters=3):
    """The SwAV block of cn3d_train_motion_GL.py:236-262 with the queue disabled (queue_length = 0, :186): for every view
    crop_id < num_crop - 1 the Sinkhorn assignment q of its codes is the soft target of every OTHER view v < num_crop - 1;
    loss_swa = sum_crop sum_v -mean(sum(q * log softmax(code_v / 0.1))) / (num_crop - 1)."""
    _lib.require_cuda(code, "code")
    B = batchSize
    loss_swa = 0
    for crop_id in range(num_crop - 1):
        with torch.no_grad():
            po = code[B * crop_id: B * (crop_id + 1), :] / epsilon                  # :253
            q = distributed_sinkhorn(torch.exp(po).t().contiguous(), sinkhorn_iters)[-B:]   # :255-256
        subloss = 0
        for v in np.delete(np.arange(num_crop - 1), crop_id):
            subloss = subloss + _SoftXent.apply(code[B * v: B * (v + 1)], q, 1.0 / temperature)
        loss_swa = loss_swa + subloss
    return loss_swa / (num_crop - 1)


# ------------------------------------------------------------------------------------------------ CLD / k-means
class _KMeans(torch.autograd.Function):
    """labels, centroids of utils_my.KMeans.  The centroids are differentiable in x exactly as in the reference, where the last
    `scatter_add_ / count` is on the autograd tape: d x[n] = d c[label[n]] / count[label[n]] (the initial `x[:K].clone()` is
    overwritten by the first update and carries no gradient)."""

    @staticmethod
    def forward(ctx, x, K, Niters):
        x = x.contiguous()
        N, D = x.shape
        labels = torch.empty(N, dtype=torch.int32, device=x.device)
        cent = torch.empty((K, D), dtype=torch.float32, device=x.device)
        counts = torch.empty(K, dtype=torch.int32, device=x.device)
        check(lib().facl_kmeans(ptr(x), N, D, int(K), int(Niters), ptr(labels), ptr(cent), ptr(counts), stream_ptr()), "facl_kmeans")
        labels = labels.long()
        ctx.save_for_backward(labels, counts)
        ctx.mark_non_differentiable(labels)
        return labels, cent

    @staticmethod
    def backward(ctx, _dlabels, dcent):
        labels, counts = ctx.saved_tensors
        return (dcent / counts.to(dcent.dtype).unsqueeze(1))[labels], None, None


def KMeans(x, K=10, Niters=10, verbose=False):
    """reference utils_my.py:180-198 -> (cl (N,) int64, c (K, D))."""
    _lib.require_cuda(x, "x")
    return _KMeans.apply(x, K, Niters)


class _HardXent(torch.autograd.Function):
    """CrossEntropyLoss(mean) of logits against integer labels on facl_softmax_xent."""

    @staticmethod
    def forward(ctx, logits, labels):
        logits = logits.contiguous()
        rows, K = logits.shape
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        dlog_t = torch.empty((K, rows), dtype=torch.float32, device=logits.device)
        lab = labels.to(torch.int32).contiguous()
        check(lib().facl_softmax_xent(ptr(logits), ptr(lab), rows, K, ptr(loss), ptr(dlog_t), None, None, stream_ptr()),
              "facl_softmax_xent")
        ctx.save_for_backward(dlog_t)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dlog_t,) = ctx.saved_tensors
        return dlog_t.t() * g, None


def grouping(features_groupDis1, features_groupDis2, T, k_eigen, clusters, num_iters):
    """reference utils_my.py:165-178: cross-level discrimination between two feature sets through each other's k-means centroids."""
    cluster_label1, centroids1 = KMeans(features_groupDis1, clusters, num_iters)
    cluster_label2, centroids2 = KMeans(features_groupDis2, clusters, num_iters)
    affnity1 = matmul_nt(features_groupDis1, centroids2) / T
    CLD_loss = _HardXent.apply(affnity1, cluster_label2)
    affnity2 = matmul_nt(features_groupDis2, centroids1) / T
    return (CLD_loss + _HardXent.apply(affnity2, cluster_label1)) / 2


def CLD_Loss(epoch, num_crop, x_nor, opt):
    """reference utils_my.py:152-162 (three-view windows, temperature 0.05, 60 clusters, 5 k-means rounds)."""
    loss_CLD = 0
    CLD_start = 0
    if epoch >= CLD_start:
        for CLD_i in range(num_crop - 4):
            loss_CLD = loss_CLD + grouping(x_nor[CLD_i * opt.batchSize: (CLD_i + 3) * opt.batchSize],
                                           x_nor[(CLD_i + 1) * opt.batchSize: (CLD_i + 4) * opt.batchSize], 0.05, 10, 60, 5)
    return loss_CLD
