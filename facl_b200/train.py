"""The training-step call sequence of the reference scripts, on the sm_100a library.

`TrainStep` re-issues the loop body of reference training_code/cn3d_train_motion_GL.py:224-335
(cn3d_train_apperance_GL.py is the same file up to three constants) through the reference-shaped API of this
package -- group_points_3DV -> netR(xt, yt) -> global loss -> circle loss -> zero_grad / backward / Adam step --
with synthetic batches standing in for the NTU DataLoader.
"""
import types

import numpy as np
import torch

from . import cn3d_model_conbag as MODELL
from . import losses as _losses
from . import utils_my
from .optim import Adam


def default_opt(batchSize=64, SAMPLE_NUM=2048, sample_num_level1=64, knn_K=64):
    """The argparse defaults of cn3d_train_motion_GL.py:77-135 that the hot path reads."""
    return types.SimpleNamespace(batchSize=batchSize, INPUT_FEATURE_NUM=4, temperal_num=3, pooling="concatenation",
                                 SAMPLE_NUM=SAMPLE_NUM, Num_Class=512, knn_K=knn_K, sample_num_level1=sample_num_level1,
                                 sample_num_level2=64, ball_radius=0.16, ball_radius2=0.25, learning_rate=0.0003)


class TrainStep:
    """One object = model + optimiser + the per-step call sequence.

    step(out_points, order=None):  out_points is the DataLoader tensor (B, G, N, 4) -- a CUDA tensor, or a (pinned)
    host tensor that is copied to the device first (cn3d_train_motion_GL.py:228).  Returns the loss as a 0-dim CUDA
    tensor; nothing synchronises unless the caller reads it."""

    def __init__(self, opt=None, num_crop=10, precision="fp32", radius2=None, device="cuda", seed=1, model=None):
        self.opt = opt or default_opt()
        self.num_crop = num_crop
        self.precision = precision
        self.radius2 = radius2
        self.device = torch.device(device)
        if model is None:
            torch.manual_seed(seed)                                   # cn3d_train_motion_GL.py:142-144
            model = MODELL.PointNet_Plus_fine(self.opt, gost=num_crop, sample_num_level1=self.opt.sample_num_level1,
                                              knn_K=self.opt.knn_K)
        self.netR = model.to(self.device)
        self.netR.precision = precision
        self.netR.train()
        self.optimizer = Adam(self.netR.parameters(), lr=self.opt.learning_rate, betas=(0.5, 0.999), eps=1e-06)
        self.base_lr = self.opt.learning_rate
        self.rng = np.random.RandomState(seed)

    def set_epoch(self, epoch):
        """StepLR(step_size=4, gamma=0.7) stepped with the epoch number (cn3d_train_motion_GL.py:181,333)."""
        for g in self.optimizer.param_groups:
            g["lr"] = self.base_lr * (0.7 ** (epoch // 4))

    def group(self, data1):
        if self.radius2 is None:
            return utils_my.group_points_3DV(data1, self.opt)                                   # r2 = 0.06
        return utils_my._group(data1, self.opt.sample_num_level1, self.opt.knn_K, self.radius2)  # e.g. 0.16 (_2048)

    def step(self, out_points, order=None):
        B, G, N, D = out_points.shape
        if not out_points.is_cuda:
            out_points = out_points.to(self.device, non_blocking=True)
        # G-major flatten on the device (the reference permutes on the host, :225-226)
        data1 = out_points.permute(1, 0, 2, 3).reshape(-1, N, D).to(torch.float32)
        xt, yt = self.group(data1)
        x, code, x_nor, x_global = self.netR(xt, yt, 1)
        if order is None:
            order = np.arange(0, G, 1)
            self.rng.shuffle(order)                                                              # :297-298
        loss_c, loss_circle = _losses.contrast_losses(x, x_global, G, B, order=order, prec=self.precision)
        loss = loss_circle + loss_c                                                              # :329
        self.optimizer.zero_grad(set_to_none=False)
        loss.backward()
        self.optimizer.step()
        return loss.detach()


def extract_features(netR, opt, out_points, radius2=None):
    """Forward-only feature extraction of extract_motion_feature.py:171-184,217-221: returns the (B, (G+1)*512)
    array the reference writes per video (num_crop + 1 blocks of 512)."""
    B, G, N, D = out_points.shape
    netR.eval()
    with torch.no_grad():
        data1 = out_points.to(next(netR.parameters()).device).permute(1, 0, 2, 3).reshape(-1, N, D).to(torch.float32)
        if radius2 is None:
            xt, yt = utils_my.group_points_3DV(data1, opt)
        else:
            xt, yt = utils_my._group(data1, opt.sample_num_level1, opt.knn_K, radius2)
        x, _, _, x_global = netR(xt, yt)
        feat = torch.cat((x, x_global), dim=0)
    return feat.reshape(G + 1, B, 512).permute(1, 0, 2).reshape(B, (G + 1) * 512)


class DevicePrefetcher:
    """Wraps the DataLoader of the reference loop (`for i, data in enumerate(train_loader)`, cn3d_train_motion_GL.py:223): yields
    the same items with the (B, G, N, 4) batch already on the device, while the host->device copy of the NEXT batch runs on a side
    stream under the current step (the reference copies synchronously, `.type(FloatTensor).cuda()` at :228, which leaves the GPU
    idle for ~1.7 ms per 42 MB batch).  Items may be tensors or tuples whose first element is the batch."""

    def __init__(self, loader, device="cuda"):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = None                 # two persistent device buffers: no allocator traffic on the hot path
        self.free = [None, None]          # event after the last compute-stream use of each slot

    def _start(self, item, k):
        batch = item[0] if isinstance(item, (tuple, list)) else item
        if batch.is_cuda:
            return item, batch.to(torch.float32), None
        if not batch.is_pinned():
            batch = batch.pin_memory()
        if self.slots is None or self.slots[0].shape != batch.shape:
            self.slots = [torch.empty(batch.shape, dtype=torch.float32, device=self.device) for _ in range(2)]
            self.free = [None, None]
        with torch.cuda.stream(self.stream):
            if self.free[k] is not None:
                self.stream.wait_event(self.free[k])
            self.slots[k].copy_(batch, non_blocking=True)      # converts float64 loader tensors on the fly (:228)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return item, self.slots[k], ev

    def __iter__(self):
        it = iter(self.loader)
        k = 0
        try:
            nxt = self._start(next(it), k)
        except StopIteration:
            return
        while nxt is not None:
            item, dev, ev = nxt
            cur_slot = k
            k ^= 1
            try:
                nxt = self._start(next(it), k)
            except StopIteration:
                nxt = None
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
            yield (dev, *item[1:]) if isinstance(item, (tuple, list)) else dev
            # the consumer has issued its work on `dev`: the slot may be overwritten once that work has run
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream())
            self.free[cur_slot] = done


class FusedTrainStep:
    """The same step as TrainStep.step, issued as ONE C-ABI call (facl_train_step) on persistent buffers: no torch
    autograd graph, no per-step allocation, gradients written straight into the tensors bound to `p.grad`.
    `step(batch)` takes the DataLoader-shaped (B, G, N, 4) fp32 batch either as a CUDA tensor or as a PINNED host
    tensor (then the H2D copy is part of the call) and returns a device tensor [loss_global, loss_circle, loss]."""

    def __init__(self, trainer, B, G, N, r2=0.06):
        import ctypes as C
        from . import _lib
        from .encoder_rt import _dims, _params_struct
        self.C, self._lib = C, _lib
        self.tr = trainer
        net = trainer.netR
        dev = trainer.device
        S, K = net.sample_num_level1, net.knn_K
        M = G * B
        self.shape = (B, G, N, 4)
        self.dims = _dims(M, S, K, G, trainer.precision, True, net._flags(True, True))
        self.ws = net._workspace(self.dims, dev, True)
        params = net._param_list()
        self.params_struct = _params_struct([p.detach() for p in params], net._bn_buffers())
        # one flat gradient buffer: [4 floats for the loss values | net3DV_1's gradients (12 tensors) | everything else], so a
        # sharded run reduces it as two contiguous buckets (dist.py): the large tail as soon as the net3DV_3 backward is
        # done, the small head after the net3DV_1 backward
        sizes = [p.numel() for p in params[:30]]
        self.flat = torch.zeros(sum(sizes) + 4, dtype=torch.float32, device=dev)
        self.flat_l1_end = 4 + sum(sizes[:12])
        self.grads, off = [], 4
        for p, n in zip(params[:30], sizes):
            self.grads.append(self.flat[off: off + n].view_as(p))
            off += n
        for p, g in zip(params[:30], self.grads):
            p.grad = g
        gs = _lib.EncoderGrads()
        for l in range(7):
            gs.dw[l], gs.db[l], gs.dgamma[l], gs.dbeta[l] = (t.data_ptr() for t in self.grads[4 * l: 4 * l + 4])
        gs.dfc3_w, gs.dfc3_b = self.grads[28].data_ptr(), self.grads[29].data_ptr()
        self.grads_struct = gs
        self._params30 = params[:30]
        self._adam_key = None
        f32 = dict(dtype=torch.float32, device=dev)
        self.staging = torch.empty((B, G, N, 4), **f32)
        self.clouds = torch.empty((M, N, 4), **f32)
        self.xt = torch.empty((M, S, K, 4), **f32)
        self.centres = torch.empty((M * S, 3), **f32)
        self.x = torch.empty((M, 512), **f32)
        self.xg = torch.empty((B, 512), **f32)
        self.dx = torch.empty((M, 512), **f32)
        self.dxg = torch.empty((B, 512), **f32)
        self.loss2 = torch.zeros(3, **f32)
        self.loss_ws = torch.empty(_lib.lib().facl_contrast_workspace_bytes(G, B, 1, 512), dtype=torch.uint8, device=dev)
        self.order_dev = torch.zeros(G, dtype=torch.int32, device=dev)
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.r2 = r2
        self.pending_bn_steps = 0
        self._copy_stream, self._pf_bufs, self._pf_free, self._pending, self._pf_slot = None, None, None, None, 0
        a = _lib.TrainStepArgs()
        a.dims, a.params, a.grads = C.pointer(self.dims), C.pointer(self.params_struct), C.pointer(self.grads_struct)
        a.enc_buffers = self.ws.table
        a.N, a.r2 = N, r2
        a.staging, a.clouds, a.xt, a.centres = (t.data_ptr() for t in (self.staging, self.clouds, self.xt, self.centres))
        a.x, a.x_global, a.order, a.loss_ws, a.loss2 = (t.data_ptr() for t in (self.x, self.xg, self.order_dev, self.loss_ws,
                                                                                 self.loss2))
        a.dx, a.dx_global = self.dx.data_ptr(), self.dxg.data_ptr()
        a.beta1, a.beta2, a.eps = 0.5, 0.999, 1e-6
        a.order_by_value = 1
        self.args = a
        self._bind_adam_state()

    def _bind_adam_state(self):
        """(Re)build the device table of (param, grad, exp_avg, exp_avg_sq, numel) records.  The optimiser's moment tensors are
        replaced by `optimizer.load_state_dict` (checkpoint resume): the table is rebuilt whenever they moved."""
        import struct
        opt = self.tr.optimizer
        recs, key = [], []
        for p, g in zip(self._params30, self.grads):
            st = opt.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
            key.append((st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()))
            recs.append(struct.pack("<QQQQq", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                                    st["exp_avg_sq"].data_ptr(), p.numel()))
        self.adam_table = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(self.staging.device)
        self._adam_key = key
        self.args.adam_table, self.args.adam_ntensors = self.adam_table.data_ptr(), 30

    def _adam_state_moved(self):
        opt = self.tr.optimizer
        for p, k in zip(self._params30, self._adam_key):
            st = opt.state[p]
            if "exp_avg" not in st or (st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()) != k:
                return True
        return False

    def _begin_step(self, batch, order, want_host_loss):
        """Argument block of one step: batch pointer, the view permutation BY VALUE, learning rate, step count."""
        if tuple(batch.shape) != self.shape or batch.dtype != torch.float32:
            raise self._lib.FaclError(f"batch must be float32 {self.shape}")
        dev_copy, pf_slot = self._take_prefetched(batch)
        if dev_copy is not None:
            batch = dev_copy
        tr = self.tr
        G = self.shape[1]
        if order is None:                                    # cn3d_train_motion_GL.py:297-298 (same seed on every rank -> same draw)
            order = np.arange(0, G, 1)
            tr.rng.shuffle(order)
        order = [int(v) for v in order]
        if sorted(order) != list(range(G)):
            raise self._lib.FaclError(f"order must be a permutation of range({G})")
        opt = tr.optimizer
        if self._adam_state_moved():
            self._bind_adam_state()
        opt._step += 1
        a = self.args
        for i, v in enumerate(order):
            a.order_vals[i] = v
        if batch.is_cuda:
            a.points_bgnd, a.points_host = batch.data_ptr(), None
        else:
            if not batch.is_pinned():
                raise self._lib.FaclError("host batches must be pinned (torch.Tensor.pin_memory)")
            a.points_bgnd, a.points_host = None, batch.data_ptr()
        a.lr = float(opt.param_groups[0]["lr"])
        a.step = opt._step
        a.loss_host = self.loss_host.data_ptr() if want_host_loss else None
        return pf_slot

    def prefetch(self, batch):
        """Start the host->device copy of a PINNED (B, G, N, 4) batch on a side stream, so that it overlaps the step that is
        running; the next step(batch) picks the device copy up (what a prefetching DataLoader wrapper does)."""
        if batch.is_cuda:
            return
        if not batch.is_pinned():
            raise self._lib.FaclError("host batches must be pinned (torch.Tensor.pin_memory)")
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.staging.device)
            self._pf_bufs = [torch.empty_like(self.staging) for _ in range(2)]
            self._pf_free = [None, None]
        self._pf_slot ^= 1
        slot = self._pf_slot
        if self._pf_free[slot] is not None:                       # the step that last read this slot has been issued: wait for it
            self._copy_stream.wait_event(self._pf_free[slot])
        with torch.cuda.stream(self._copy_stream):
            self._pf_bufs[slot].copy_(batch, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._pending = (batch.data_ptr(), slot, ev)

    def _take_prefetched(self, batch):
        """Device copy of `batch` if prefetch() was called for it, else None."""
        if self._pending is None or batch.is_cuda or self._pending[0] != batch.data_ptr():
            return None, None
        _, slot, ev = self._pending
        self._pending = None
        torch.cuda.current_stream().wait_event(ev)
        return self._pf_bufs[slot], slot

    def _release_prefetched(self, slot):
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._pf_free[slot] = ev

    def step(self, batch, order=None, want_host_loss=False, next_batch=None):
        pf_slot = self._begin_step(batch, order, want_host_loss)
        self.args.phases = 0
        self._lib.check(self._lib.lib().facl_train_step(self.C.byref(self.args), self._lib.stream_ptr()), "facl_train_step")
        self._release_prefetched(pf_slot)
        if next_batch is not None:
            self.prefetch(next_batch)
        self.pending_bn_steps += 1
        return self.loss2

    def flush_counters(self):
        """num_batches_tracked is bookkeeping only (momentum is fixed): applied lazily, off the hot path."""
        if self.pending_bn_steps:
            with torch.no_grad():
                for i, (_, bn) in enumerate(self.tr.netR._layers()):
                    bn.num_batches_tracked += (2 if i == 6 else 1) * self.pending_bn_steps
            self.pending_bn_steps = 0
