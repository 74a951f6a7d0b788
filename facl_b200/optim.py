"""Adam on the GPU through the C ABI (facl_adam_step): torch.optim.Adam semantics (no weight decay, no amsgrad),
the optimiser the reference builds at cn3d_train_motion_GL.py:180 -- Adam(lr=3e-4, betas=(0.5, 0.999), eps=1e-6).

One kernel launch updates every parameter: the (param, grad, exp_avg, exp_avg_sq, numel) records live in a small
device table that is rebuilt only when a gradient tensor moved."""
import struct

import torch

from ._lib import check, lib, stream_ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=3e-4, betas=(0.5, 0.999), eps=1e-6):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._table = None
        self._table_key = None
        self._step = 0

    # The step count lives in one Python int on the hot path (the fused train step increments it without touching the
    # per-parameter state); state_dict() / load_state_dict() carry it as torch.optim.Adam does -- state[p]["step"], a
    # float32 scalar tensor per parameter -- so a resumed run keeps its bias correction and the state is interchangeable
    # with torch.optim.Adam(lr, betas, eps) (cn3d_train_motion_GL.py:180).
    def state_dict(self):
        for st in self.state.values():
            if st:
                st["step"] = torch.tensor(float(self._step), dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [int(st["step"]) for st in self.state.values() if st and "step" in st]
        self._step = max(steps) if steps else 0
        self._table_key = None              # the moment tensors were replaced: rebuild the device table

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closure is not supported")
        self._step += 1
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            recs, key = [], []
            for p in live:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("facl_b200.optim.Adam updates fp32 CUDA parameters only")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                recs.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()))
                key.append(recs[-1])
            key = tuple(key)
            if key != self._table_key:
                blob = b"".join(struct.pack("<QQQQq", *r) for r in recs)
                host = torch.frombuffer(bytearray(blob), dtype=torch.uint8).pin_memory()
                self._table = host.to(live[0].device, non_blocking=True)
                self._host = host           # keep the pinned staging buffer alive until the copy has run
                self._table_key = key
            b1, b2 = group["betas"]
            check(lib().facl_adam_step(self._table.data_ptr(), len(recs), float(group["lr"]), float(b1), float(b2),
                                       float(group["eps"]), self._step, stream_ptr()), "facl_adam_step")
        return None
