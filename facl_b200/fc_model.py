"""Drop-in for the reference module `linear_classify/fc_model.py`: the linear probe trained on
the extracted features.  Same constructor, state-dict keys (`fc.weight`, `fc.bias`) and initialisation as the
reference (:12-19); forward = F.normalize(x, p=2, dim=1) -> fc (:21-25), computed by libfacl_b200.so
(facl_l2_normalize + the tcgen05 GEMM).  CUDA only."""
import torch
import torch.nn as nn

from . import ops
from ._lib import check, lib, ptr, require_cuda, stream_ptr


def l2_normalize(x):
    require_cuda(x, "x")
    x = x.contiguous()
    out = torch.empty_like(x)
    check(lib().facl_l2_normalize(ptr(x), x.shape[0], x.shape[1], ptr(out), stream_ptr()), "facl_l2_normalize")
    return out


def linear_forward(xn, weight, bias, nsplit=3):
    """logits (rows, classes) = xn @ weight.T + bias on the tensor cores.  The output is a single 128 x 256 tile, so the
    long reduction (F = 11 264) is split over 16 CTAs that accumulate atomically onto the broadcast bias."""
    rows, F = xn.shape
    Cc = weight.shape[0]
    out = bias.expand(rows, Cc).contiguous() if bias is not None else torch.zeros((rows, Cc), dtype=torch.float32, device=xn.device)
    ops.gemm_tc(rows, Cc, F, nsplit=nsplit, a=dict(src0=xn, ld=F), b_mode=ops.B_ROWMAJOR, b=dict(src0=weight, ld=F),
                ksplit=min(16, (F + 63) // 64), out_mode=ops.OUT_ATOMIC, out=out, ldo=Cc)
    return out


def linear_wgrad(dlogits_t, xn, nsplit=3, out=None):
    """dW (classes, F) = dlogits.T @ xn; dlogits_t is (classes, rows) contiguous."""
    Cc, rows = dlogits_t.shape
    F = xn.shape[1]
    if out is None:
        out = torch.empty((Cc, F), dtype=torch.float32, device=xn.device)
    ops.gemm_tc(Cc, F, rows, nsplit=nsplit, a=dict(src0=dlogits_t, ld=rows), b_mode=ops.B_CHMAJOR, b=dict(src0=xn, ld=F),
                out_mode=ops.OUT_CHMAJOR, out=out, ldo=F)
    return out


class _FinalFCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        if x.requires_grad:
            raise NotImplementedError("Final_FC: the features are data (linercls.py:109); no gradient w.r.t. x")
        xn = l2_normalize(x)
        ctx.save_for_backward(xn)
        return linear_forward(xn, weight.contiguous(), bias.contiguous())

    @staticmethod
    def backward(ctx, dlogits):
        (xn,) = ctx.saved_tensors
        dW = linear_wgrad(dlogits.t().contiguous(), xn)
        return None, dW, dlogits.sum(dim=0)


class Final_FC(nn.Module):
    def __init__(self, input_dim=512, gost=11 + 11, num_class=120):
        super().__init__()
        self.fc = nn.Linear(input_dim * gost * 1, num_class)
        self.fc.weight.data.normal_(mean=0.0, std=0.01)
        self.fc.bias.data.zero_()

    def forward(self, x):
        require_cuda(x, "x")
        return _FinalFCFunction.apply(x, self.fc.weight, self.fc.bias)
