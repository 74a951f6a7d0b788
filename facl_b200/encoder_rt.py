"""Host-side runtime for the encoder: owns the work buffers (torch tensors) and calls the C ABI
(facl_encoder_forward / facl_encoder_backward) through an autograd.Function."""
import ctypes as C

import torch

from . import _lib
from ._lib import EncoderDims, EncoderGrads, EncoderParams, check, lib, stream_ptr

# order of the parameter tensors handed to the autograd.Function: 7 x (w, b, gamma, beta), fc3_w, fc3_b, map_w
LAYER_PREFIXES = [("net3DV_1", 0, 1), ("net3DV_1", 3, 4), ("net3DV_1", 6, 7),
                  ("net3DV_3", 0, 1), ("net3DV_3", 3, 4), ("net3DV_3", 6, 7), ("netR_FC", 0, 1)]
NSPLIT = {"fp32": 3, "bf16": 1, "bf16_fast": 1}


class EncoderWorkspace:
    """Named device buffers for one problem size.  Buffers flagged backward-only by the library are only
    allocated when `backward` is requested."""

    def __init__(self, dims, device, backward):
        L = lib()
        self.key = (dims.M, dims.S, dims.K, dims.G, bool(backward), dims.flags)
        n = L.facl_encoder_num_buffers()
        self.tensors = {}
        self.table = (C.c_void_p * n)()
        for i in range(n):
            name = L.facl_encoder_buffer_name(i).decode()
            if L.facl_encoder_buffer_backward_only(i) and not backward:
                self.table[i] = None
                continue
            nbytes = L.facl_encoder_buffer_bytes(i, C.byref(dims))
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self.tensors[name] = t
            self.table[i] = t.data_ptr()
        self.generation = 0

    def view(self, name, shape, dtype=torch.float32):
        """Typed view of a buffer (tests / debugging)."""
        t = self.tensors[name]
        numel = 1
        for s in shape:
            numel *= s
        return t[: numel * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)


def _dims(M, S, K, G, precision, training, flags=0):
    d = EncoderDims()
    d.M, d.S, d.K, d.G, d.nsplit, d.training, d.flags = M, S, K, G, NSPLIT[precision], int(bool(training)), int(flags)
    return d


def _params_struct(params, buffers):
    """params: list of 31 tensors (see LAYER_PREFIXES); buffers: 7 x (running_mean, running_var)."""
    p = EncoderParams()
    for l in range(7):
        w, b, g, be = params[4 * l: 4 * l + 4]
        rm, rv = buffers[2 * l: 2 * l + 2]
        p.layer[l].w, p.layer[l].b, p.layer[l].gamma, p.layer[l].beta = w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr()
        p.layer[l].running_mean, p.layer[l].running_var = rm.data_ptr(), rv.data_ptr()
    p.fc3_w, p.fc3_b, p.map_w = params[28].data_ptr(), params[29].data_ptr(), params[30].data_ptr()
    return p


class EncoderFunction(torch.autograd.Function):
    """forward(rows, centres, owner, *params) -> (x, code, x_nor, x_global); backward -> parameter gradients."""

    @staticmethod
    def forward(ctx, rows, centres, owner, *params):
        M, S, K, _ = rows.shape
        G = owner.gost
        training = owner.training
        need_bwd = owner._need_bwd        # decided by the caller: grad mode is always off inside Function.forward
        dims = _dims(M, S, K, G, owner.precision, training, owner._flags(training, need_bwd))
        ws = owner._workspace(dims, rows.device, need_bwd)
        bufs = owner._bn_buffers()
        ps = _params_struct([p.detach() for p in params], bufs)
        B = M // G
        x = torch.empty((M, 512), dtype=torch.float32, device=rows.device)
        xg = torch.empty((B, 512), dtype=torch.float32, device=rows.device)
        x_nor = torch.empty_like(x)
        code = torch.empty((M, 64), dtype=torch.float32, device=rows.device)
        check(lib().facl_encoder_forward(C.byref(dims), C.byref(ps), rows.data_ptr(), centres.data_ptr(), ws.table,
                                         x.data_ptr(), xg.data_ptr(), x_nor.data_ptr(), code.data_ptr(), stream_ptr()),
              "facl_encoder_forward")
        ws.generation += 1
        ctx.dims, ctx.ws, ctx.gen, ctx.owner = dims, ws, ws.generation, owner
        ctx.save_for_backward(rows, *params)
        ctx.mark_non_differentiable(code, x_nor)
        return x, code, x_nor, xg

    @staticmethod
    def backward(ctx, dx, dcode, dx_nor, dxg):
        rows, *params = ctx.saved_tensors
        ws = ctx.ws
        if ws.generation != ctx.gen:
            raise _lib.FaclError("encoder work buffers were overwritten by a later forward before this backward ran")
        if ws.table[lib().facl_encoder_num_buffers() - 1] is None:
            raise _lib.FaclError("forward ran without gradient buffers (no parameter required grad)")
        bufs = ctx.owner._bn_buffers()
        ps = _params_struct([p.detach() for p in params], bufs)
        grads = [torch.empty_like(p) for p in params[:30]]
        g = EncoderGrads()
        for l in range(7):
            g.dw[l], g.db[l], g.dgamma[l], g.dbeta[l] = (t.data_ptr() for t in grads[4 * l: 4 * l + 4])
        g.dfc3_w, g.dfc3_b = grads[28].data_ptr(), grads[29].data_ptr()
        dxp = dx.contiguous().data_ptr() if dx is not None else None
        dxgp = dxg.contiguous().data_ptr() if dxg is not None else None
        check(lib().facl_encoder_backward(C.byref(ctx.dims), C.byref(ps), rows.data_ptr(), ws.table, dxp, dxgp, C.byref(g),
                                          stream_ptr()), "facl_encoder_backward")
        # mapping.weight only feeds `code`, which is marked non-differentiable (no live reference loss uses it)
        return (None, None, None, *grads, None)
