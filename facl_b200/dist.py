"""Multi-GPU data parallelism for the training step: one process per GPU (torchrun), NCCL over NVLink.

The reference is single-GPU (its nn.DataParallel is pinned to one device, cn3d_train_motion_GL.py:33,176; the only
torch.distributed code, concat_all_gather at cn3d_model_conbag.py:559-570, is dead).  The sharding follows
SURVEY.md section 8e:

  * the SEQUENCE axis B is sharded (never the flattened G*B axis: every view of a sequence must stay on one rank for
    the sequence max-pool, cn3d_model_conbag.py:225); rank r owns samples [r*Bl, (r+1)*Bl) and encodes them locally
    with per-rank BatchNorm statistics (what nn.DataParallel does);
  * forward exchange: all-gather of the per-rank view embeddings x_r (G*Bl, 512) -> keys (R*G*Bl, 512), rank-major.
    Every rank evaluates the two losses for ITS anchors against ALL keys (global negatives).  The gather is asynchronous:
    the head on the sequence features (x_global, which no other rank needs) runs beside it;
  * backward exchange: the key-side gradient dkeys (R*G*Bl, 512) is sum-reduce-scattered back to the owners, so the
    gradient flows through the gather (the reference's helper is @no_grad).  Asynchronous as well: the sequence half of the
    head backward needs dx_global only and runs beside it;
  * parameter gradients (+ the loss values) live in one flat buffer, reduced as TWO buckets: everything except
    net3DV_1's gradients (99 % of the bytes) is all-reduced asynchronously as soon as the net3DV_3 backward has produced
    it and hides under the net3DV_1 backward; the small rest follows.  Then every rank applies the same Adam step.
    BatchNorm running statistics stay rank-local; rank 0's are the ones to checkpoint.
"""
import numpy as np
import torch
import torch.distributed as dist

from .train import FusedTrainStep

PHASE_FORWARD, PHASE_LOSS, PHASE_BACKWARD, PHASE_UPDATE, PHASE_BACKWARD_HEAD, PHASE_BACKWARD_L1 = 1, 2, 4, 8, 16, 32   # FACL_PHASE_*
PHASE_FORWARD_X, PHASE_FORWARD_G, PHASE_BACKWARD_HEAD_G, PHASE_BACKWARD_HEAD_X = 64, 128, 256, 512


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
        self.overlap = True          # bucketed all-reduce, large bucket under the net3DV_1 backward (False: one all-reduce at the end)
        self._tl = None
        import os
        self.max_ahead = int(os.environ.get("FACL_MAX_AHEAD", "2"))   # steps the host may run ahead of the GPU (0: unbounded)
        self._inflight = []
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

    # ---- optional timeline: CUDA events on the compute stream at every phase / collective boundary (bench.py --dist-timeline) ----
    PHASE_NAMES = ["forward to x", "head(x_global) + wait all_gather(x)", "losses", "head backward (sequence half) + wait reduce_scatter(dkeys)",
                   "backward head (cloud half) + net3DV_3", "backward net3DV_1", "all_reduce(small) + wait(big)", "adam + loss D2H"]

    def enable_timeline(self, on=True):
        self._tl = [] if on else None

    def _mark(self):
        if getattr(self, "_tl", None) is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._cur.append(ev)

    def timeline_ms(self):
        """Average milliseconds per phase over the recorded steps (synchronises)."""
        torch.cuda.synchronize()
        n = len(self._tl)
        acc = [0.0] * len(self.PHASE_NAMES)
        for evs in self._tl:
            for i in range(len(self.PHASE_NAMES)):
                acc[i] += evs[i].elapsed_time(evs[i + 1])
        return {name: acc[i] / max(n, 1) for i, name in enumerate(self.PHASE_NAMES)}

    def _call(self, phases):
        self.args.phases = phases
        self._lib.check(self._lib.lib().facl_train_step(self.C.byref(self.args), self._lib.stream_ptr()), "facl_train_step")

    def step(self, batch, order=None, want_host_loss=False, next_batch=None):
        multi = self.world > 1
        if multi and self.max_ahead > 0:
            # bounded run-ahead: the host may be at most `max_ahead` steps in front of the GPU.  Measured on 8 GPUs: a free-running
            # host loop (every rank's Python thread issuing NCCL calls and ~100 launches per step without ever blocking) made the
            # step 0.3 ms SLOWER than a loop that reads the loss back every step; NCCL's host-side threads compete with it.
            while len(self._inflight) >= self.max_ahead:
                self._inflight.pop(0).synchronize()
        pf_slot = self._begin_step(batch, order, want_host_loss)
        self._cur = []
        self._mark()
        self._call(PHASE_FORWARD_X)                                             # ... -> cloud embeddings x
        self._mark()
        if multi:
            ag = dist.all_gather_into_tensor(self.keys, self.x, async_op=True)  # NCCL's stream; the sequence head runs beside it
        else:
            self.keys.copy_(self.x)
        self._call(PHASE_FORWARD_G)                                             # head on the sequence features -> x_global
        if multi:
            ag.wait()
        self._mark()
        self._call(PHASE_LOSS)                                                  # losses, dx / dkeys
        self._mark()
        if multi:
            rs = dist.reduce_scatter_tensor(self.dkeys_loc, self.dkeys, op=dist.ReduceOp.SUM, async_op=True)
        else:
            self.dkeys_loc.copy_(self.dkeys)
        self.flat[0:3].copy_(self.loss2)                                        # this rank's loss shares ride in the small bucket
        self._call(PHASE_BACKWARD_HEAD_G)                                       # sequence half of the head: needs dx_global only
        if multi:
            rs.wait()
        self._mark()
        self._call(PHASE_BACKWARD_HEAD_X)                                       # dx += dkeys_loc; cloud half + net3DV_3: 99 % of the gradient bytes are final
        big = None
        if multi and self.overlap:
            # asynchronous: NCCL's stream waits for the kernels issued so far and reduces the large bucket while the
            # net3DV_1 backward (passes C / D, 2.2 ms) runs on the compute stream
            big = dist.all_reduce(self.flat[self.flat_l1_end:], op=dist.ReduceOp.SUM, async_op=True)
        self._mark()
        self._call(PHASE_BACKWARD_L1)
        self._mark()
        if multi:
            if big is not None:
                dist.all_reduce(self.flat[:self.flat_l1_end], op=dist.ReduceOp.SUM)   # loss values + net3DV_1 gradients (86 KB)
                big.wait()                                                      # compute stream waits for the large bucket
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)                # one bucket, after the whole backward
        self.loss2.copy_(self.flat[0:3])
        self._mark()
        self._call(PHASE_UPDATE)                                                # Adam, loss D2H
        self._mark()
        if getattr(self, "_tl", None) is not None:
            self._tl.append(self._cur)
        if multi and self.max_ahead > 0:
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
        self._release_prefetched(pf_slot)
        if next_batch is not None:
            self.prefetch(next_batch)
        self.pending_bn_steps += 1
        return self.loss2
