"""ctypes binding of libfacl_b200.so (the C ABI declared in include/facl_b200.h).

There is no CPU fallback: if the shared library is missing, importing a product module that needs it
raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C facl_b200/csrc`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# FACL_LIB_PATH: load a diagnostic build of the same library instead (e.g. libfacl_b200_prof.so, `make -C facl_b200/csrc prof`)
LIB_PATH = os.environ.get("FACL_LIB_PATH") or os.path.join(_HERE, "libfacl_b200.so")

_lib = None


class FaclError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("src0", C.c_void_p), ("src1", C.c_void_p), ("ld", C.c_longlong),
                ("s0", C.c_void_p), ("s1", C.c_void_p), ("s2", C.c_void_p), ("lo", C.c_void_p)]


class ActImage(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("cgs", C.c_int), ("rbs", C.c_int)]


class Gemm(C.Structure):
    _fields_ = [("Md", C.c_int), ("Nd", C.c_int), ("Kd", C.c_int), ("nsplit", C.c_int),
                ("a_mode", C.c_int), ("b_mode", C.c_int),
                ("a_packed", C.c_void_p), ("a_packed_kblocks", C.c_int),
                ("a", Operand), ("b", Operand),
                ("ksplit", C.c_int),
                ("bias", C.c_void_p), ("out_mode", C.c_int), ("out", C.c_void_p), ("ldo", C.c_longlong),
                ("zin", C.c_void_p), ("ldz", C.c_longlong), ("zs0", C.c_void_p), ("zs2", C.c_void_p),
                ("stats", C.c_void_p), ("pool", C.c_int), ("pool_sign", C.c_void_p), ("pool_out", C.c_void_p),
                ("pool_arg", C.c_void_p), ("ldp", C.c_longlong), ("a_img", ActImage), ("b_img", ActImage)]


NUM_BN_LAYERS = 7


class Layer(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p)]


class EncoderParams(C.Structure):
    _fields_ = [("layer", Layer * NUM_BN_LAYERS), ("fc3_w", C.c_void_p), ("fc3_b", C.c_void_p), ("map_w", C.c_void_p)]


class EncoderGrads(C.Structure):
    _fields_ = [("dw", C.c_void_p * NUM_BN_LAYERS), ("db", C.c_void_p * NUM_BN_LAYERS),
                ("dgamma", C.c_void_p * NUM_BN_LAYERS), ("dbeta", C.c_void_p * NUM_BN_LAYERS),
                ("dfc3_w", C.c_void_p), ("dfc3_b", C.c_void_p)]


class EncoderDims(C.Structure):
    _fields_ = [("M", C.c_int), ("S", C.c_int), ("K", C.c_int), ("G", C.c_int), ("nsplit", C.c_int),
                ("training", C.c_int), ("flags", C.c_int)]


ENC_FUSED_L1 = 1


class TrainStepArgs(C.Structure):
    _fields_ = [("dims", C.POINTER(EncoderDims)), ("params", C.POINTER(EncoderParams)), ("grads", C.POINTER(EncoderGrads)),
                ("enc_buffers", C.POINTER(C.c_void_p)), ("N", C.c_int), ("r2", C.c_float),
                ("points_bgnd", C.c_void_p), ("points_host", C.c_void_p), ("staging", C.c_void_p),
                ("clouds", C.c_void_p), ("xt", C.c_void_p), ("centres", C.c_void_p), ("x", C.c_void_p),
                ("x_global", C.c_void_p), ("order", C.c_void_p), ("loss_ws", C.c_void_p), ("loss2", C.c_void_p),
                ("dx", C.c_void_p), ("dx_global", C.c_void_p), ("adam_table", C.c_void_p), ("adam_ntensors", C.c_int),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("step", C.c_int),
                ("loss_host", C.c_void_p), ("phases", C.c_int), ("B_global", C.c_int), ("sample_offset", C.c_int),
                ("keys", C.c_void_p), ("dkeys", C.c_void_p), ("dx_extra", C.c_void_p),
                ("order_by_value", C.c_int), ("order_vals", C.c_int * 256)]


class PointSource(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("offsets", C.c_void_p), ("C", C.c_int)]


class ViewRecipe(C.Structure):
    _fields_ = [("source", C.c_int), ("channel", C.c_int), ("nonzero_only", C.c_int), ("jitter", C.c_int),
                ("mirror", C.c_int), ("rotate", C.c_int)]


class AugmentArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("G", C.c_int), ("N", C.c_int), ("n_sources", C.c_int),
                ("sources", C.POINTER(PointSource)), ("recipes", C.POINTER(ViewRecipe)),
                ("sigma", C.c_double), ("clip", C.c_double),
                ("idx", C.c_void_p), ("noise", C.c_void_p), ("angle_u", C.c_void_p),
                ("seed", C.c_ulonglong), ("step", C.c_ulonglong), ("g_major", C.c_int), ("max_rows", C.c_int),
                ("out", C.c_void_p), ("out_rows", C.c_void_p)]


_I, _LL, _P, _F, _SZ = C.c_int, C.c_longlong, C.c_void_p, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/facl_b200.h declares (tests check this)
SIGNATURES = {
    "facl_version": (C.c_char_p, []),
    "facl_error_string": (C.c_char_p, [_I]),
    "facl_fps": (_I, [_P, _I, _I, _I, _P, _I, _P, _P]),
    "facl_fps_reorder": (_I, [_P, _I, _I, _I, _P, _I, _P, _P]),
    "facl_group_points": (_I, [_P, _I, _I, _I, _I, _I, _F, _P, _P, _P]),
    "facl_group_level2_scratch_bytes": (_SZ, [_I, _I, _I, _I]),
    "facl_group_level2": (_I, [_P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "facl_l2_normalize": (_I, [_P, _I, _I, _P, _P]),
    "facl_softmax_xent": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "facl_augment_views": (_I, [C.POINTER(AugmentArgs), _P]),
    "facl_sinkhorn": (_I, [_P, _I, _I, _I, _P, _P]),
    "facl_soft_xent": (_I, [_P, _P, _I, _I, _F, _P, _P, _P]),
    "facl_kmeans": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "facl_packed_weight_bytes": (_SZ, [_I, _I]),
    "facl_pack_weight": (_I, [_P, _LL, _LL, _I, _I, _P, _P]),
    "facl_gemm_stat_partials": (_I, [_I, _I]),
    "facl_gemm_tc": (_I, [C.POINTER(Gemm), _P]),
    "facl_act_image_half_bytes": (_SZ, [_I, _LL]),
    "facl_act_image": (_I, [C.POINTER(Operand), _LL, _I, _LL, _P, _I, _I, C.POINTER(ActImage), _P]),
    "facl_encoder_num_buffers": (_I, []),
    "facl_encoder_buffer_name": (C.c_char_p, [_I]),
    "facl_encoder_buffer_bytes": (_SZ, [_I, C.POINTER(EncoderDims)]),
    "facl_encoder_buffer_backward_only": (_I, [_I]),
    "facl_encoder_forward": (_I, [C.POINTER(EncoderDims), C.POINTER(EncoderParams), _P, _P, C.POINTER(C.c_void_p),
                                  _P, _P, _P, _P, _P]),
    "facl_encoder_backward": (_I, [C.POINTER(EncoderDims), C.POINTER(EncoderParams), _P, C.POINTER(C.c_void_p),
                                   _P, _P, C.POINTER(EncoderGrads), _P]),
    "facl_adam_step": (_I, [_P, _I, _F, _F, _F, _F, _I, _P]),
    "facl_gmajor": (_I, [_P, _P, _I, _I, _I, _P]),
    "facl_train_step": (_I, [C.POINTER(TrainStepArgs), _P]),
    "facl_timing_enable": (None, [_I]),
    "facl_timing_collect": (_I, [_P, _P, _I]),
    "facl_launch_count": (C.c_longlong, []),
    "facl_debug_l1_dump": (None, [_P, _P]),
    "facl_contrast_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "facl_contrast_losses": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
}


def lib():
    """The loaded library (loads on first use; raises FaclError when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FaclError(
                f"{LIB_PATH} not found: the CUDA library has not been built "
                "(run `make -C facl_b200/csrc`); facl_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().facl_error_string(int(code)).decode()
        raise FaclError(f"{what or 'facl call'} failed: {msg} (cudaError {code})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name="tensor", dtype=torch.float32):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise FaclError(f"{name} must be a CUDA tensor: facl_b200 has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise FaclError(f"{name} must be {dtype}, got {t.dtype}")
    return t
