"""Host side of the contrastive losses: autograd.Function over the C ABI call facl_contrast_losses."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, stream_ptr

# The losses run the bf16x3 split products in BOTH modes: they are < 1 % of the step's FLOPs, and their logits are un-normalised
# dot products of magnitude ~500 with temperature 1 (utils_my.py:72-82) -- single bf16 products would move a logit by > 1.
_PRECISION = {"fp32": 3, "bf16": 3, "bf16_fast": 3}
precision = "fp32"      # module-level switch used by utils_my.global_contrast / circle_contrast

_ws_cache = {}


def _workspace(G, B, Cdim, device):
    key = (G, B, Cdim, device)
    ws = _ws_cache.get(key)
    if ws is None:
        _ws_cache.clear()
        ws = torch.empty(lib().facl_contrast_workspace_bytes(G, B, 1, Cdim), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


class ContrastLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, x_global, G, B, order_dev, want_global, want_circle, nsplit):
        x = x.contiguous()
        M, Cdim = x.shape
        if M != G * B:
            raise _lib.FaclError(f"x has {M} rows, expected num_crop*batchSize = {G * B}")
        dev = x.device
        xg = x_global.contiguous() if x_global is not None else None
        loss = torch.empty(2, dtype=torch.float32, device=dev)
        # the two losses may receive different upstream gradients, so their x-gradients are kept apart: one call each
        ws = _workspace(G, B, Cdim, dev)
        p = lambda t: None if t is None else t.data_ptr()
        dxg_part = dxg = dxc = None
        if want_global:
            dxg_part = torch.empty_like(x)
            dxg = torch.empty((B, Cdim), dtype=torch.float32, device=dev)
            lg = torch.empty(2, dtype=torch.float32, device=dev)
            check(lib().facl_contrast_losses(p(x), p(xg), None, G, B, B, 0, Cdim, None, 1, 0, nsplit, ws.data_ptr(),
                                             lg.data_ptr(), p(dxg_part), p(dxg), p(dxg_part), stream_ptr()),
                  "facl_contrast_losses")
            loss[0:1].copy_(lg[0:1])
        else:
            loss[0:1].zero_()
        if want_circle:
            dxc = torch.empty_like(x)
            lc = torch.empty(2, dtype=torch.float32, device=dev)
            check(lib().facl_contrast_losses(p(x), None, None, G, B, B, 0, Cdim, p(order_dev), 0, 1, nsplit, ws.data_ptr(),
                                             lc.data_ptr(), p(dxc), None, p(dxc), stream_ptr()), "facl_contrast_losses")
            loss[1:2].copy_(lc[1:2])
        else:
            loss[1:2].zero_()
        ctx.grads = (dxg_part, dxg, dxc)
        ctx.has_xg = x_global is not None
        return loss[0], loss[1]

    @staticmethod
    def backward(ctx, g_global, g_circle):
        dxg_part, dxg, dxc = ctx.grads
        dx = None
        if dxg_part is not None and g_global is not None:
            dx = dxg_part * g_global
        if dxc is not None and g_circle is not None:
            dx = dxc * g_circle if dx is None else dx + dxc * g_circle
        dg = dxg * g_global if (dxg is not None and g_global is not None) else None
        return dx, (dg if ctx.has_xg else None), None, None, None, None, None, None


def contrast_losses(x, x_global, num_crop, batch_size, order=None, want_global=True, want_circle=True, prec=None):
    """(loss_global, loss_circle) as 0-dim CUDA tensors with autograd (either may be a constant 0)."""
    _lib.require_cuda(x, "x")
    if want_global:
        _lib.require_cuda(x_global, "x_global")
    order_dev = None
    if want_circle:
        order_dev = torch.as_tensor(np.asarray(order, dtype=np.int32), device=x.device)
    nsplit = _PRECISION[prec or precision]
    return ContrastLossFunction.apply(x, x_global if want_global else None, int(num_crop), int(batch_size), order_dev,
                                      bool(want_global), bool(want_circle), nsplit)


class _SimilarityFunction(torch.autograd.Function):
    """sims = x @ x.T on the tcgen05 GEMM (bf16x3 split, fp32 accumulation), differentiable."""

    @staticmethod
    def forward(ctx, x):
        from . import ops
        x = x.contiguous()
        R, Cd = x.shape
        out = torch.empty((R, R), dtype=torch.float32, device=x.device)
        ops.gemm_tc(R, R, Cd, nsplit=3, a=dict(src0=x, ld=Cd), b_mode=ops.B_ROWMAJOR, b=dict(src0=x, ld=Cd),
                    out_mode=ops.OUT_CHMAJOR, out=out, ldo=R)
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dsims):
        from . import ops
        (x,) = ctx.saved_tensors
        R, Cd = x.shape
        dsym = (dsims + dsims.t()).contiguous()
        dx = torch.empty_like(x)
        ops.gemm_tc(R, Cd, R, nsplit=3, a=dict(src0=dsym, ld=R), b_mode=ops.B_CHMAJOR, b=dict(src0=x, ld=Cd),
                    out_mode=ops.OUT_CHMAJOR, out=dx, ldo=Cd)
        return dx


def info_nce_logits_cuda(x, batch_size):
    """Two-view logits of reference utils_my.py:200-213 (not on the live path): the similarity
    GEMM (and its gradient) run on libfacl_b200's tensor-core kernel, the masking / concatenation is tensor plumbing."""
    _lib.require_cuda(x, "x")
    B = batch_size
    if x.shape[0] != 2 * B:
        raise _lib.FaclError(f"Info_NCE expects the two views of {B} samples, got {x.shape[0]} rows")   # :209-210 need (B, 2B)
    n = torch.arange(B, device=x.device)[:, None]
    j = torch.arange(2 * B, device=x.device)[None, :]
    mask = ((j % B) != n).to(x.dtype)
    sims = _SimilarityFunction.apply(x)
    pos = (x[0:B] * x[B:2 * B]).sum(dim=1, keepdim=True)
    logits = torch.cat([pos, sims[0:B] * mask, sims[B:2 * B] * mask], dim=1)
    return logits, torch.zeros(B, dtype=torch.long, device=x.device)
