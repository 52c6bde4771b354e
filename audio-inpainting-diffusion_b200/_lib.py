"""ctypes binding of libaid_b200.so (include/aid_b200.h).

The product path has no CPU fallback: importing this module without the compiled CUDA library raises.
Build it with `python __graft_entry__.py` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaid_b200.so")
MAX_OCTS = 16


class AidConfig(C.Structure):
    _fields_ = [
        ("num_octs", C.c_int32), ("bins_per_oct", C.c_int32), ("audio_len", C.c_int32), ("window_kind", C.c_int32),
        ("sample_rate", C.c_double), ("beta", C.c_double),
        ("emb_dim", C.c_int32), ("num_heads", C.c_int32),
        ("Ns", C.c_int32 * MAX_OCTS), ("num_dils", C.c_int32 * MAX_OCTS), ("attention_layers", C.c_int32 * (MAX_OCTS + 1)),
        ("num_bottleneck_layers", C.c_int32), ("conv_mode", C.c_int32),
    ]


class AidError(RuntimeError):
    pass


_P = C.c_void_p
_SIGS = {
    "aid_create": (C.c_int, [C.POINTER(AidConfig), C.c_int, C.POINTER(_P)]),
    "aid_destroy": (None, [_P]),
    "aid_last_error": (C.c_char_p, [_P]),
    "aid_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "aid_num_weights": (C.c_int, [_P]),
    "aid_weight_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "aid_finalize": (C.c_int, [_P]),
    "aid_workspace_bytes": (C.c_int, [_P, C.c_int, C.POINTER(C.c_size_t)]),
    "aid_unet_forward": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, C.c_float, C.c_float, C.c_float, _P, C.c_size_t, _P]),
    "aid_unet_forward_ds": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "aid_edm_step_ds": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_int, _P, _P, _P, _P, _P]),
    "aid_philox_normal": (C.c_int, [_P, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_int, _P, _P]),
    "aid_sched_select": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "aid_vjp_workspace_bytes": (C.c_int, [_P, C.c_int, C.POINTER(C.c_size_t)]),
    "aid_unet_forward_tape": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, C.c_float, C.c_float, C.c_float, _P, C.c_size_t, _P]),
    "aid_unet_backward": (C.c_int, [_P, _P, _P, _P]),
    "aid_cqt_layout": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "aid_cqt_fwd": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    "aid_cqt_bwd": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    "aid_hpf_dc": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    "aid_cqt_plan": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P, _P, _P, _P, _P]),
    "aid_cqt_workspace_bytes": (C.c_int, [_P, C.c_int, C.POINTER(C.c_size_t)]),
    "aid_edm_add_noise": (C.c_int, [_P, _P, C.c_float, C.c_int64, _P]),
    "aid_edm_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_int, _P, _P, _P, _P, _P]),
    "aid_spectral_mask": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, _P]),
    "aid_op_conv2d": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                _P, _P, _P, C.c_float, C.c_float, _P, _P, C.c_int, _P]),
    "aid_op_groupnorm_act": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "aid_op_resample": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "aid_op_attention": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "aid_op_resample_adj": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "aid_op_groupnorm_act_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "aid_op_attention_bwd": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t, _P]),
    "aid_op_conv2d_bwd_input": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "aid_cqt_fwd_vjp": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    "aid_cqt_bwd_vjp": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    "aid_op_attention_mode": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "aid_op_embedding": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "aid_debug_time_conv2d": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        _P, _P, C.c_float, _P, _P, C.c_int, C.POINTER(C.c_float)]),
    "aid_debug_tc2_operands": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "aid_debug_tc2_profile": (C.c_int, [C.POINTER(C.c_uint64)]),
    "aid_debug_time_gn_tc2": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "aid_debug_dilated_layer": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_float, C.c_int, _P, _P,
                                          C.POINTER(C.c_float)]),
    "aid_debug_init_block": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, _P, _P, C.POINTER(C.c_float)]),
    "aid_debug_out_block": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P, C.POINTER(C.c_float)]),
    "aid_debug_fusion": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "aid_profile": (C.c_int, [_P, C.c_int]),
    "aid_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "aid_debug_probe": (C.c_int, [_P, C.c_char_p, _P]),
    "aid_debug_saturation": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "aid_launch_count": (C.c_uint64, []),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AidError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                           "(run `python __graft_entry__.py`); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def check(rc, handle=None):
    if rc != 0:
        msg = lib().aid_last_error(handle)
        raise AidError(f"aid status {rc}: {msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device (or host) pointer of a torch tensor, or NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())
