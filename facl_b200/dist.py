"""Multi-GPU data parallelism for the training step: one process per GPU (torchrun), NCCL over NVLink.

The reference is single-GPU (its nn.DataParallel is pinned to one device, cn3d_train_motion_GL.py:33,176; the only
torch.distributed code, concat_all_gather at cn3d_model_conbag.py:559-570, is dead).  The sharding follows
SURVEY.md section 8e:

  * the SEQUENCE axis B is sharded (never the flattened G*B axis: every view of a sequence must stay on one rank for
    the sequence max-pool, cn3d_model_conbag.py:225); rank r owns samples [r*Bl, (r+1)*Bl) and encodes them locally
    with per-rank BatchNorm statistics (what nn.DataParallel does);
  * forward exchange: all-gather of the per-rank view embeddings x_r (G*Bl, 512) -> keys (R*G*Bl, 512), rank-major.
    Every rank evaluates the two losses for ITS anchors against ALL keys (global negatives);
  * backward exchange: the key-side gradient dkeys (R*G*Bl, 512) is sum-reduce-scattered back to the owners, so the
    gradient flows through the gather (the reference's helper is @no_grad);
  * parameter gradients (+ the two loss values) are summed by ONE all-reduce over a flat buffer, then every rank
    applies the same Adam step.  BatchNorm running statistics stay rank-local; rank 0's are the ones to checkpoint.
"""
import numpy as np
import torch
import torch.distributed as dist

from .train import FusedTrainStep


def key_index(g, n, G, Bl):
    """Rank-major key row of view g of global sample n (see csrc/losses.cu `Idx`)."""
    return (n // Bl) * (G * Bl) + g * Bl + (n % Bl)


def reference_order_from_keys(keys, G, B, Bl):
    """keys (R*G*Bl, C) rank-major -> the reference's row order (row g*B + n), for checks against the oracle."""
    idx = [key_index(g, n, G, Bl) for g in range(G) for n in range(B)]
    return keys[torch.as_tensor(idx, device=keys.device)]


class DistributedFusedTrainStep(FusedTrainStep):
    """FusedTrainStep with the batch sharded over the ranks of the default process group.
    `B` is the PER-RANK batch; the losses see the global batch B * world_size."""

    def __init__(self, trainer, B, G, N, r2=0.06):
        super().__init__(trainer, B, G, N, r2=r2)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        M = G * B
        dev = trainer.device
        self.keys = torch.empty((self.world * M, 512), dtype=torch.float32, device=dev)
        self.dkeys = torch.empty((self.world * M, 512), dtype=torch.float32, device=dev)
        self.dkeys_loc = torch.empty((M, 512), dtype=torch.float32, device=dev)
        self.loss_ws = torch.empty(self._lib.lib().facl_contrast_workspace_bytes(G, B, self.world, 512), dtype=torch.uint8,
                                   device=dev)
        a = self.args
        a.loss_ws = self.loss_ws.data_ptr()
        a.B_global, a.sample_offset = B * self.world, B * self.rank
        a.keys, a.dkeys, a.dx_extra = self.keys.data_ptr(), self.dkeys.data_ptr(), self.dkeys_loc.data_ptr()
        # weights must start identical on every rank
        if self.world > 1:
            for p in trainer.netR.parameters():
                dist.broadcast(p.data, src=0)

    def _call(self, phases):
        self.args.phases = phases
        self._lib.check(self._lib.lib().facl_train_step(self.C.byref(self.args), self._lib.stream_ptr()), "facl_train_step")

    def step(self, batch, order=None, want_host_loss=False, next_batch=None):
        dev_copy, pf_slot = self._take_prefetched(batch)
        if dev_copy is not None:
            batch = dev_copy
        tr = self.tr
        G = self.shape[1]
        if order is None:                                   # same seed on every rank -> same permutation
            order = np.arange(0, G, 1)
            tr.rng.shuffle(order)
        opt = tr.optimizer
        opt._step += 1
        slot = self.order_ring[opt._step % len(self.order_ring)]
        slot.copy_(torch.from_numpy(np.asarray(order, dtype=np.int32)))
        self.order_dev.copy_(slot, non_blocking=True)
        a = self.args
        if batch.is_cuda:
            a.points_bgnd, a.points_host = batch.data_ptr(), None
        else:
            a.points_bgnd, a.points_host = None, batch.data_ptr()
        a.lr, a.step = float(opt.param_groups[0]["lr"]), opt._step
        a.loss_host = self.loss_host.data_ptr() if want_host_loss else None
        self._call(1)                                                          # forward
        if self.world > 1:
            dist.all_gather_into_tensor(self.keys, self.x)
        else:
            self.keys.copy_(self.x)
        self._call(2)                                                          # losses, dx / dkeys
        if self.world > 1:
            dist.reduce_scatter_tensor(self.dkeys_loc, self.dkeys, op=dist.ReduceOp.SUM)
        else:
            self.dkeys_loc.copy_(self.dkeys)
        self._call(4)                                                          # backward
        self.flat[-4:-1].copy_(self.loss2)
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.loss2.copy_(self.flat[-4:-1])
        self._call(8)                                                          # Adam, loss D2H
        self._release_prefetched(pf_slot)
        if next_batch is not None:
            self.prefetch(next_batch)
        self.pending_bn_steps += 1
        return self.loss2
