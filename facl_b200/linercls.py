"""The training / evaluation loop body of the reference's `linear_classify/linercls.py` (:86-124, :136-147) as one
fused device sequence per batch: normalise -> logits GEMM -> softmax cross-entropy (+ gradient, bias gradient,
top-1 hits) -> weight-gradient GEMM -> Adam(lr, betas (0.5, 0.999), eps 1e-6).  No host synchronisation inside a step;
loss and accuracy are read when asked for."""
import torch

from . import fc_model
from ._lib import check, lib, ptr, require_cuda, stream_ptr
from .optim import Adam


class ProbeTrainer:
    def __init__(self, netR=None, learning_rate=0.005, device="cuda"):
        self.netR = (netR if netR is not None else fc_model.Final_FC()).to(device)
        w, b = self.netR.fc.weight, self.netR.fc.bias
        w.grad, b.grad = torch.zeros_like(w), torch.zeros_like(b)
        self.optimizer = Adam([w, b], lr=learning_rate, betas=(0.5, 0.999), eps=1e-6)    # linercls.py:92
        self._acc = torch.zeros(2, dtype=torch.float32, device=device)                    # loss
        self._hits = torch.zeros(1, dtype=torch.int32, device=device)
        self._seen = 0

    def set_epoch(self, epoch, base_lr=None):
        """StepLR(step_size=5, gamma=0.7) stepped with the epoch number (linercls.py:93,118)."""
        base = base_lr if base_lr is not None else self.optimizer.defaults["lr"]
        for g in self.optimizer.param_groups:
            g["lr"] = base * 0.7 ** (epoch // 5)

    def _forward(self, features, labels, train):
        require_cuda(features, "features")
        labels = require_cuda(labels, "labels", None).to(torch.int32).contiguous()
        w, b = self.netR.fc.weight, self.netR.fc.bias
        xn = fc_model.l2_normalize(features)
        logits = fc_model.linear_forward(xn, w.data, b.data)
        rows, Cc = logits.shape
        dlt = torch.empty((Cc, rows), dtype=torch.float32, device=logits.device) if train else None
        if train:
            b.grad.zero_()
        check(lib().facl_softmax_xent(ptr(logits), ptr(labels), rows, Cc, ptr(self._acc), ptr(dlt), ptr(b.grad) if train else None,
                                      ptr(self._hits), stream_ptr()), "facl_softmax_xent")
        self._seen += rows
        return xn, logits, dlt

    def step(self, features, labels):
        """One iteration of linercls.py:109-122.  Returns the logits (device tensor)."""
        xn, logits, dlt = self._forward(features, labels, True)
        fc_model.linear_wgrad(dlt, xn, out=self.netR.fc.weight.grad)
        self.optimizer.step()
        return logits

    @torch.no_grad()
    def evaluate(self, features, labels):
        """linercls.py:139-147."""
        return self._forward(features, labels, False)[1]

    def pop_meters(self):
        """-> (sum of the per-batch mean losses, top-1 accuracy in percent) since the last call; synchronises."""
        loss_sigma = float(self._acc[0])
        top1 = 100.0 * float(self._hits[0]) / max(self._seen, 1)
        self._acc.zero_()
        self._hits.zero_()
        self._seen = 0
        return loss_sigma, top1
