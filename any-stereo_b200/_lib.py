"""ctypes binding of the C-ABI library ``csrc/libanystereo_b200.so`` (see include/anystereo_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libanystereo_b200.so")

AS_MAX_LEVELS = 8
AS_MAX_SRC = 4
DTYPE_F32, DTYPE_F16, DTYPE_F64 = 0, 1, 2
CORR_FP32_SIMT, CORR_BF16X3, CORR_BF16 = 0, 1, 2
LAYOUT_NHWC, LAYOUT_NCHW = 0, 1
EPI_BIAS, EPI_BIAS_RELU, EPI_GRU_ZR, EPI_GRU_Q = 0, 1, 2, 3
UEPI_RELU_SPLIT, UEPI_MOTION, UEPI_GRU_ZR, UEPI_GRU_Q, UEPI_DISPHEAD, UEPI_LINEAR_F32 = 0, 1, 2, 3, 4, 5

_vp = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_sz = C.c_size_t
_pp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int)


class ConvSrc(C.Structure):
    _fields_ = [("ptr", _vp), ("channels", _i), ("pitch", _i), ("layout", _i)]


class UmmaSrc(C.Structure):
    _fields_ = [("hi", _vp), ("lo", _vp), ("channels", _i)]


class ConvUmmaDesc(C.Structure):
    _fields_ = [
        ("B", _i), ("H", _i), ("W", _i), ("KH", _i), ("KW", _i), ("Cout", _i), ("num_src", _i),
        ("src", UmmaSrc * 3),
        ("w_hi", _vp), ("w_lo", _vp), ("nsplit", _i), ("epilogue", _i),
        ("bias", _vp), ("ctx", _vp), ("ctx_pitch", _i), ("h", _vp), ("z", _vp),
        ("out_f32", _vp), ("out_hi", _vp), ("out_lo", _vp),
        ("out_pitch", _i), ("out_coff", _i), ("cout_valid", _i),
        ("disp", _vp), ("w2", _vp), ("u", _vp),
    ]


class LiifQueryDesc(C.Structure):
    _fields_ = [
        ("n_in", _i), ("P", _vp * 3), ("h", _i * 3), ("w", _i * 3),
        ("coords", _vp), ("B", _i), ("Q", _i), ("wc", _vp),
        ("w2_hi", _vp), ("w2_lo", _vp), ("w3_hi", _vp), ("w3_lo", _vp), ("w4_hi", _vp), ("w4_lo", _vp),
        ("b2", _vp), ("b3", _vp), ("b4", _vp), ("nsplit", _i),
        ("disp", _vp), ("disp_scale", _vp), ("hd", _i), ("wd", _i),
        ("logits", _vp), ("out", _vp),
    ]


class WgradUmmaDesc(C.Structure):
    _fields_ = [
        ("B", _i), ("H", _i), ("W", _i), ("KH", _i), ("KW", _i), ("Cout", _i), ("num_src", _i),
        ("src", UmmaSrc * 3), ("dy_hi", _vp), ("dy_lo", _vp), ("Wp", _i), ("nsplit", _i), ("ws", _vp), ("dw_acc", _vp),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", _i), ("H", _i), ("W", _i), ("KH", _i), ("KW", _i), ("Cout", _i), ("num_src", _i),
        ("src", ConvSrc * AS_MAX_SRC),
        ("weight", _vp), ("bias", _vp), ("epilogue", _i),
        ("out", _vp), ("out_pitch", _i), ("out_coff", _i), ("out_layout", _i),
        ("ctx", _vp), ("ctx_pitch", _i), ("h", _vp), ("z", _vp), ("save", _vp),
    ]


# name -> (restype, argtypes); mirrors include/anystereo_b200.h one to one
SIGNATURES = {
    "as_abi_version": (_i, []),
    "as_error_string": (C.c_char_p, [_i]),
    "as_compiled_sm": (_i, []),
    "as_sampler_fwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_sampler_bwd": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_corr1d_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "as_corr1d_build": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _pp, _ip, _i, _vp, _sz, _vp]),
    "as_pool1d_halve": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp]),
    "as_pool1d_halve_bwd_acc": (_i, [_vp, _vp, _ll, _i, _i, _i, _vp]),
    "as_corr1d_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "as_geo_pyramid_build": (_i, [_vp, _i, _i, _i, _i, _i, _i, _pp, _vp]),
    "as_geo_pyramid_bwd": (_i, [_pp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "as_corr_lookup_fwd": (_i, [_pp, _ip, _ip, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "as_corr_lookup_bwd": (_i, [_pp, _ip, _ip, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "as_geo_lookup_fwd": (_i, [_pp, _i, _i, _pp, _ip, _ip, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "as_geo_lookup_bwd": (_i, [_pp, _i, _i, _pp, _ip, _ip, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "as_geo_lookup_convc1": (_i, [_pp, _i, _i, _pp, _ip, _ip, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i,
                                  _vp]),
    "as_geo_lookup_convc1_tap": (_i, [_pp, _i, _i, _pp, _ip, _ip, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i,
                                      _vp]),
    "as_isu_affinity": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "as_liif_query": (_i, [C.POINTER(LiifQueryDesc), _vp]),
    "as_context_upsample_multiscale": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "as_context_upsample_multiscale_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll, _vp]),
    "as_liif_layer1_fwd": (_i, [_i, _pp, _pp, _ip, _ip, _vp, _vp, _vp, _i, _i, _ll, _vp]),
    "as_liif_layer1_bwd": (_i, [_i, _pp, _pp, _ip, _ip, _vp, _vp, _vp, _pp, _pp, _vp, _i, _i, _ll, _vp]),
    "as_nearest_gather_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _ll, _vp]),
    "as_nearest_gather_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _ll, _vp]),
    "as_init_disparity": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "as_disparity_regression": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "as_convd1_umma": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_set_operand_format": (_i, [_i]),
    "as_get_operand_format": (_i, []),
    "as_corr_lookup_convc1": (_i, [_pp, _ip, _ip, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "as_lookup_taps": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "as_gwc_build_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_gwc_build_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_gwc_corr_stem_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, C.c_float, _vp]),
    "as_instnorm_workspace_bytes": (C.c_size_t, [_i, _i]),
    "as_instnorm_nhwc": (_i, [_vp, _vp, _vp, _vp, C.c_size_t, _i, _ll, _i, C.c_float, _i, _vp]),
    "as_pack_conv_weight": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "as_conv2d_fp32": (_i, [C.POINTER(ConvDesc), _vp]),
    "as_pool2x_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "as_interp_bilinear_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_add_f32": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "as_nchw_to_nhwc_bias": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_conv2d_umma": (_i, [C.POINTER(ConvUmmaDesc), _vp]),
    "as_pack_conv_weight_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_split_f32": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "as_nchw_to_nhwc_split": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "as_pool2x_nhwc_split": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "as_interp_bilinear_nhwc_split": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "as_convd1_split": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "as_disp_delta": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "as_pack_conv_weight_dgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "as_conv2d_wgrad_fp32": (_i, [C.POINTER(ConvDesc), _vp, _i, _i, _vp, _vp, _vp]),
    "as_relu_bwd": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _ll, _i, _vp]),
    "as_gru_bwd_gates1": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "as_gru_bwd_gates2": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "as_conv_epilogue_fp32": (_i, [_vp, _i, _ll, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "as_convd1_fp32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "as_convd1_wgrad_fp32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "as_bias_grad_fp32": (_i, [_vp, _i, _i, _ll, _vp, _vp]),
    "as_transpose_split": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "as_conv2d_wgrad_umma": (_i, [C.POINTER(WgradUmmaDesc), _vp]),
    "as_add_slice": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _vp]),
    "as_pool2x_nhwc_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "as_interp_bilinear_nhwc_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None
launch_count = 0  # kernels-launching ABI calls made by this process (bench.py reports it)


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "anystereo_b200: CUDA library %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is deliberately no CPU/PyTorch fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


FMT_BF16, FMT_F16, FMT_F16F8 = 0, 1, 2
_operand_format = FMT_BF16


def operand_format() -> int:
    return _operand_format


def set_operand_format(fmt: int):
    """16-bit operand format of the tensor-core kernels (process-wide, see as_set_operand_format)."""
    global _operand_format
    if fmt != _operand_format:
        check(lib().as_set_operand_format(fmt), "as_set_operand_format")
        _operand_format = fmt


class operand_format_scope:
    """Run a block with a given operand format and restore the previous one."""

    def __init__(self, fmt):
        self.fmt = fmt

    def __enter__(self):
        self.prev = _operand_format
        set_operand_format(self.fmt)

    def __exit__(self, *exc):
        set_operand_format(self.prev)


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().as_error_string(rc).decode()
        raise RuntimeError("anystereo_b200 %s failed (%d): %s" % (what, rc, msg))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def int_array(vals):
    arr = (C.c_int * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def call(name: str, *args):
    """Invoke an int-returning entry point on the current device/stream and raise on error."""
    global launch_count
    rc = getattr(lib(), name)(*args)
    launch_count += 1
    check(rc, name)


def require_cuda(t: torch.Tensor, name: str, dtype=None, contiguous=True):
    """Boundary checks of the reference's native path (sampler/sampler.cpp:20-22) plus dtype."""
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if contiguous and not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("%s must have dtype %s (got %s)" % (name, dtype, t.dtype))
    return t


def version_of(t) -> int:
    """``t._version`` for cache keys.  Inference tensors (created under torch.inference_mode()) do not track a
    version counter and raise on access: they key on identity alone (callers also hold a weak reference), version -1.
    In-place writes to such a tensor, like writes through ``param.data``, are invisible to the caches -- call
    ``BasicMultiUpdateBlock.invalidate_weights()`` / ``reset_caches()`` after them."""
    if torch.is_inference(t):
        return -1
    return t._version
