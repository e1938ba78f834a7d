"""ctypes binding of libdxmi_b200.so (include/dxmi_b200.h).

The shared library is the product; this module only loads it and mirrors the C structs.  There is no Python /
PyTorch fallback: if the library is missing, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdxmi_b200.so")

ARCH_DDPM_UNET, ARCH_ADM_UNET, ARCH_IGEBM_V2 = 0, 1, 2
F32, F16, I64 = 0, 1, 2
ACT_NONE, ACT_LRELU02, ACT_SILU = 0, 1, 2


class ArchDesc(C.Structure):
    _fields_ = [
        ("arch", C.c_int),
        ("resolution", C.c_int),
        ("in_channels", C.c_int),
        ("out_channels", C.c_int),
        ("ch", C.c_int),
        ("n_levels", C.c_int),
        ("ch_mult", C.c_int * 8),
        ("num_res_blocks", C.c_int),
        ("n_attn", C.c_int),
        ("attn_resolutions", C.c_int * 8),
        ("num_classes", C.c_int),
        ("num_head_channels", C.c_int),
        ("num_heads", C.c_int),
        ("use_scale_shift_norm", C.c_int),
        ("resblock_updown", C.c_int),
        ("learn_out_scale", C.c_int),
        ("precision", C.c_int),
    ]


class OptTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_longlong), ("lr", C.c_float), ("reserved", C.c_int)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("a_ptr", C.c_void_p * 3),
        ("a_C", C.c_int * 3),
        ("a_ld", C.c_int * 3),
        ("N", C.c_int),
        ("H", C.c_int),
        ("W", C.c_int),
        ("nseg", C.c_int),
        ("seg_src", C.c_int * 3),
        ("seg_taps", C.c_int * 3),
        ("stride", C.c_int),
        ("out_H", C.c_int),
        ("out_W", C.c_int),
        ("b_ptr", C.c_void_p),
        ("b_rows", C.c_int),
        ("b_ld", C.c_longlong),
        ("b_batch_stride", C.c_longlong),
        ("batch", C.c_int),
        ("a_batched", C.c_int),
        ("b_batched", C.c_int),
        ("out", C.c_void_p),
        ("ldo", C.c_int),
        ("out_batch_stride", C.c_longlong),
        ("out_fp32", C.c_int),
        ("bias", C.c_void_p),
        ("bias_along_m", C.c_int),
        ("rowvec", C.c_void_p),
        ("ldrv", C.c_int),
        ("rows_per_image", C.c_int),
        ("residual", C.c_void_p),
        ("ldr", C.c_int),
        ("res_batch_stride", C.c_longlong),
        ("act", C.c_int),
        ("alpha", C.c_float),
        ("softmax", C.c_int),
        ("block_n", C.c_int),
        ("out_nchw", C.c_int),
        ("gn_stats", C.c_void_p),
        ("gn_seg", C.c_int),
        ("gn_halo_P", C.c_int),
        ("gate", C.c_void_p),
        ("ldg", C.c_int),
        ("up2", C.c_int),
    ]


_lib = None

# name -> (restype, argtypes); also the list of symbols the header declares (checked by the CPU tests)
_VP, _I, _LL, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float
SYMBOLS = {
    "dxmi_create": (_I, [C.POINTER(ArchDesc), _I, C.POINTER(_VP)]),
    "dxmi_destroy": (None, [_VP]),
    "dxmi_bind_weight": (_I, [_VP, C.c_char_p, _VP, _I, C.POINTER(C.c_int64), _I]),
    "dxmi_num_weights": (_I, [_VP]),
    "dxmi_weight_key": (_I, [_VP, _I, C.c_char_p, _I]),
    "dxmi_weight_shape": (_I, [_VP, _I, C.POINTER(C.c_int64), C.POINTER(_I)]),
    "dxmi_finalize": (_I, [_VP, _VP]),
    "dxmi_repack": (_I, [_VP, _VP]),
    "dxmi_unet_forward": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dxmi_value_forward": (_I, [_VP, _VP, _VP, _I, _VP]),
    "dxmi_var_step": (_I, [_VP] * 10 + [_I, _I, _VP]),
    "dxmi_edm_step": (_I, [_VP] * 6 + [_I, _I, _VP]),
    "dxmi_var_rollout": (_I, [_VP, C.POINTER(_F), _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dxmi_edm_rollout": (_I, [_VP, C.POINTER(_F), _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "dxmi_quantize_u8": (_I, [_VP, _VP, _LL, _VP]),
    "dxmi_running_cost_fwd": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP]),
    "dxmi_running_cost_bwd": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP]),
    "dxmi_opt_chunk_elems": (_I, []),
    "dxmi_opt_grad_norm": (_I, [_VP, _VP, _VP, _I, _F, _VP, _VP, _I, _VP]),
    "dxmi_opt_adam_step": (_I, [_VP, _VP, _VP, _I, _VP, _F, _F, _F, _I, _I, _VP]),
    "dxmi_op_conv_gemm": (_I, [C.POINTER(GemmDesc), _VP]),
    "dxmi_op_pack_conv_weight": (_I, [_VP, _I, _I, _I, _I, _I, _I, _I, _VP, _LL, _LL, _VP]),
    "dxmi_op_group_norm": (_I, [_VP, _I, _I, _VP, _I, _I, _I, _I, _I, _F, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP]),
    "dxmi_op_gn_ws_floats": (_I, [_I, _I, _I]),
    "dxmi_op_halo_tiles_per_image": (_I, [_I, _I]),
    "dxmi_op_attention": (_I, [_VP, _LL, _I, _I, _VP, _VP, _I, _I, _I, _I, _F, _VP]),
    "dxmi_bind_grad": (_I, [_VP, C.c_char_p, _VP]),
    "dxmi_unet_forward_train": (_I, [_VP, _VP, _VP, _VP, _F, C.c_ulonglong, _I, _VP]),
    "dxmi_adm_forward_train": (_I, [_VP, _VP, _VP, _VP, _VP, _F, C.c_ulonglong, _I, _VP]),
    "dxmi_op_dropout_mask": (_I, [_VP, _LL, _F, C.c_ulonglong, C.c_uint, _VP]),
    "dxmi_unet_backward": (_I, [_VP, _VP, _VP, _VP, _I, _VP]),
    "dxmi_value_forward_train": (_I, [_VP, _VP, _VP, _I, _VP]),
    "dxmi_value_backward": (_I, [_VP, _VP, _VP, _VP, _I, _VP]),
    "dxmi_op_gn_bwd_ws_floats": (_LL, [_I, _I, _I]),
    "dxmi_op_group_norm_bwd": (_I, [_VP, _I, _VP, _I, _VP, _VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "dxmi_op_pack_conv_weight_up2": (_I, [_VP, _I, _I, _I, _VP, _VP]),
    "dxmi_op_pack_conv_weight_dgrad": (_I, [_VP, _I, _I, _I, _I, _VP, _LL, _LL, _VP]),
    "dxmi_op_wgrad_ws_floats": (_LL, [_I, _I, _I, _I, _I, _I]),
    "dxmi_op_conv_wgrad": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _I, _VP, _I, _I, _F, _VP, _VP]),
    "dxmi_last_error": (C.c_char_p, []),
    "dxmi_set_option": (_I, [C.c_char_p, _I]),
    "dxmi_set_debug_buffer": (_I, [_VP]),
    "dxmi_set_timing_dump": (_I, [C.c_char_p]),
    "dxmi_launch_count": (_LL, []),
    "dxmi_gemm_timing": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_LL)]),
    "dxmi_aux_timing": (_I, [_I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_LL)]),
    "dxmi_plan_gemm_flops": (C.c_double, [_VP, _I]),
    "dxmi_workspace_bytes": (C.c_size_t, [_VP, _I]),
}


def lib():
    """Load (once) and return the C-ABI library. Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(the DxMI B200 path has no CPU / PyTorch fallback)"
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error():
    return lib().dxmi_last_error().decode()


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libdxmi_b200 {what} failed (rc={rc}): {last_error()}")


def stream_ptr(device=None):
    """Current torch stream of `device` (a torch.device / index / tensor; default: the current device)."""
    import torch

    if device is not None and hasattr(device, "device"):
        device = device.device
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
