"""Thin tensor-level wrappers over the C ABI (one Python function per exported kernel launcher).

Everything here takes and returns CUDA tensors; there is no CPU path.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Gemm, Operand, check, lib, ptr, require_cuda, stream_ptr

A_PACKED, A_ROWMAJOR, A_IMAGE = 0, 1, 2
B_ROWMAJOR, B_CHMAJOR, B_XT4, B_IMAGE_MN, B_IMAGE_K = 1, 2, 3, 4, 5
OUT_NONE, OUT_CHMAJOR, OUT_ROWMAJOR, OUT_ATOMIC = 0, 1, 2, 3


def fps(points, m, start_idx):
    """points (V,N,D) fp32 cuda, start_idx (V) int32 cuda -> picks (V,m) int32.
    C ABI: facl_fps (replaces cn3D_data_set.py:675-694)."""
    require_cuda(points, "points")
    points = points.contiguous()
    V, N, D = points.shape
    start_idx = require_cuda(start_idx, "start_idx", torch.int32).contiguous()
    out = torch.empty((V, m), dtype=torch.int32, device=points.device)
    check(lib().facl_fps(ptr(points), V, N, D, ptr(start_idx), m, ptr(out), stream_ptr()), "facl_fps")
    return out


def fps_reorder(points, picks):
    """C ABI: facl_fps_reorder (replaces cn3D_data_set.py:665-672)."""
    require_cuda(points, "points")
    points = points.contiguous()
    V, N, D = points.shape
    picks = require_cuda(picks, "picks", torch.int32).contiguous()
    out = torch.empty_like(points)
    check(lib().facl_fps_reorder(ptr(points), V, N, D, ptr(picks), picks.shape[1], ptr(out), stream_ptr()),
          "facl_fps_reorder")
    return out


def group_points_raw(points, S, K, r2, want_idx=True):
    """points (M,N,D) fp32 cuda -> rows (M,S,K,D) fp32, idx (M,S,K) int32 | None.
    C ABI: facl_group_points (replaces utils_my.py:255-291 and copies)."""
    require_cuda(points, "points")
    points = points.contiguous()
    M, N, D = points.shape
    rows = torch.empty((M, S, K, D), dtype=torch.float32, device=points.device)
    idx = torch.empty((M, S, K), dtype=torch.int32, device=points.device) if want_idx else None
    check(lib().facl_group_points(ptr(points), M, N, D, S, K, float(r2), ptr(rows), ptr(idx), stream_ptr()),
          "facl_group_points")
    return rows, idx


def group_level2(feats, S2, K, r2, want_idx=False):
    """feats (M,C,N1) fp32 cuda channel-first -> out (M,C,S2,K), idx (M,S2,K) int32 | None.
    C ABI: facl_group_level2 (replaces utils_my.py:332-381)."""
    require_cuda(feats, "points")
    feats = feats.contiguous()
    M, Cc, N1 = feats.shape
    out = torch.empty((M, Cc, S2, K), dtype=torch.float32, device=feats.device)
    idx = torch.empty((M, S2, K), dtype=torch.int32, device=feats.device) if want_idx else None
    scratch = torch.empty(lib().facl_group_level2_scratch_bytes(M, N1, S2, K), dtype=torch.uint8, device=feats.device)
    check(lib().facl_group_level2(ptr(feats), M, Cc, N1, S2, K, float(r2), ptr(out), ptr(idx), ptr(scratch), stream_ptr()),
          "facl_group_level2")
    return out, idx


def pack_weight(w, rows, cols, stride_m, stride_k, out=None):
    """fp32 matrix view A[m][k] = w.flat[m*stride_m + k*stride_k] -> bf16 hi/lo tile image (uint8 tensor)."""
    require_cuda(w, "w")
    nbytes = lib().facl_packed_weight_bytes(rows, cols)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    assert out.numel() >= nbytes
    check(lib().facl_pack_weight(ptr(w), stride_m, stride_k, rows, cols, ptr(out), stream_ptr()), "facl_pack_weight")
    return out


def _operand(src0=None, src1=None, ld=0, s0=None, s1=None, s2=None, lo=None):
    o = Operand()
    o.src0, o.src1, o.ld = _p(src0), _p(src1), int(ld)
    o.s0, o.s1, o.s2, o.lo = _p(s0), _p(s1), _p(s2), _p(lo)
    return o


def _p(t):
    return None if t is None else t.data_ptr()


def stat_partials(Md, Nd):
    return lib().facl_gemm_stat_partials(Md, Nd)


class Image:
    """A bf16 hi/lo activation image on the device (C ABI: facl_act_image); see include/facl_b200.h."""

    def __init__(self, C_, R, device):
        half = lib().facl_act_image_half_bytes(C_, R)
        self.buf = torch.empty(2 * half, dtype=torch.uint8, device=device)
        self.C, self.R = C_, R
        self.struct = _lib.ActImage(self.buf.data_ptr(), self.buf.data_ptr() + half, ((C_ + 63) // 64) * 8, (R + 63) // 64)


def act_image(C_, R, nsplit, *, ld1=0, pool_arg=None, pool=0, **operand):
    """fp32 channel-major activation(s) -> Image (BN / ReLU / BN-backward transform applied on the way)."""
    src = _operand(**operand)
    dev = operand["src0"].device
    img = Image(C_, R, dev)
    check(lib().facl_act_image(C.byref(src), int(ld1), C_, R, _p(pool_arg), pool, nsplit, C.byref(img.struct), stream_ptr()),
          "facl_act_image")
    return img


def gemm_tc(Md, Nd, Kd, *, nsplit, b_mode, b=None, a_packed=None, a=None, a_img=None, b_img=None, ksplit=1, bias=None, out_mode=OUT_NONE, out=None,
            ldo=0, zin=None, ldz=0, zs0=None, zs2=None, stats=None, pool=0, pool_sign=None, pool_out=None,
            pool_arg=None, ldp=0):
    """D[m,n] = sum_k A[m,k] B[n,k] on tcgen05 with fused prologue / epilogue (C ABI: facl_gemm_tc).
    `a` / `b` are dicts of _operand() keyword arguments."""
    g = Gemm()
    g.Md, g.Nd, g.Kd, g.nsplit = Md, Nd, Kd, nsplit
    if a_packed is not None:
        g.a_mode, g.a_packed, g.a_packed_kblocks = A_PACKED, a_packed.data_ptr(), (Kd + 63) // 64
        g.a = _operand()
    elif a_img is not None:
        g.a_mode, g.a_img, g.a = A_IMAGE, a_img.struct, _operand()
    else:
        g.a_mode, g.a = A_ROWMAJOR, _operand(**a)
    g.b_mode, g.b = b_mode, _operand(**(b or {}))
    if b_img is not None:
        g.b_img = b_img.struct
    g.ksplit = ksplit
    g.bias, g.out_mode, g.out, g.ldo = _p(bias), out_mode, _p(out), int(ldo)
    g.zin, g.ldz, g.zs0, g.zs2 = _p(zin), int(ldz), _p(zs0), _p(zs2)
    g.stats, g.pool, g.pool_sign, g.pool_out, g.pool_arg, g.ldp = _p(stats), pool, _p(pool_sign), _p(pool_out), \
        _p(pool_arg), int(ldp)
    check(lib().facl_gemm_tc(C.byref(g), stream_ptr()), "facl_gemm_tc")
