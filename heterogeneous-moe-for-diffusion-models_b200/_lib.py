"""ctypes binding of libhdmoe_b200.so (the C ABI declared in include/hdmoe_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HDMOE_B200_LIB: another build of the same library (kernel A/B measurements in tools/); the product path is the in-tree one
LIB_PATH = os.environ.get("HDMOE_B200_LIB") or os.path.join(_HERE, "lib", "libhdmoe_b200.so")

F32, BF16 = 0, 1
MAX_EXPERTS, MAX_TOPK = 64, 8
WLAYOUT_SAME, WLAYOUT_TAPS, WLAYOUT_TAPS_T = 0, 1, 2

_p, _i, _f, _i64, _sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t


class WprepDesc(C.Structure):
    _fields_ = [("w", _p), ("w_hat", _p), ("w_hat2", _p), ("gain_ptr", _p), ("active", _p), ("gain", _f),
                ("rows", C.c_int32), ("fan_in", C.c_int32), ("cin", C.c_int32), ("taps", C.c_int32),
                ("cin_pad", C.c_int32), ("cin_rows", C.c_int32), ("cout_pad", C.c_int32), ("out_dtype", C.c_int32),
                ("layout", C.c_int32), ("layout2", C.c_int32), ("block_start", C.c_int32)]


class WprepBwdDesc(C.Structure):
    _fields_ = [("w", _p), ("d_w_hat", _p), ("gain_ptr", _p), ("d_w", _p), ("d_gain", _p), ("gain", _f),
                ("rows", C.c_int32), ("fan_in", C.c_int32), ("cin", C.c_int32), ("taps", C.c_int32),
                ("cin_pad", C.c_int32), ("layout", C.c_int32), ("block_start", C.c_int32)]


class OptTensorDesc(C.Structure):
    _fields_ = [("p", _p), ("g", _p), ("m", _p), ("v", _p), ("numel", C.c_int64), ("lr", _f), ("weight_decay", _f),
                ("chunk_start", C.c_int32), ("pad", C.c_int32)]


MAX_MASK_EXPERTS = 16


class MaskGenDesc(C.Structure):       # hdmoe_maskgen_t
    _fields_ = [("centers", _f * MAX_MASK_EXPERTS), ("p_mean", _f), ("p_std", _f), ("bandwidth", _f),
                ("n_experts", C.c_int32), ("min_active", C.c_int32)]


# name -> (restype, argtypes); mirrors include/hdmoe_b200.h one to one
PROTOTYPES = {
    "hdmoe_version": (_i, []),
    "hdmoe_last_error": (C.c_char_p, []),
    "hdmoe_launch_count": (_i64, []),
    "hdmoe_router_gate_workspace_bytes": (_sz, [_i, _i]),
    "hdmoe_router_gate_fwd": (_i, [_p, _p, _p, _p, _f, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hdmoe_router_gate_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "hdmoe_dispatch_plan_workspace_bytes": (_sz, [_i, _i]),
    "hdmoe_dispatch_plan": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hdmoe_dispatch_plan_topk": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hdmoe_permute_rows": (_i, [_p, _p, _p, _i, _p, _p, _i, _p]),
    "hdmoe_combine_rows": (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i64, _p]),
    "hdmoe_combine_rows_bwd": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i64, _p, _p, _p]),
    "hdmoe_edm_precond_in": (_i, [_p, _p, _i, _f, _p, _i, _i64, _i64, _p]),
    "hdmoe_edm_precond_out": (_i, [_p, _i, _p, _i, _p, _i, _f, _p, _i64, _i64, _p]),
    "hdmoe_edm_precond_out_bwd": (_i, [_p, _p, _i, _f, _p, _i, _p, _i, _i64, _i64, _p]),
    "hdmoe_edm_precond_in_bwd": (_i, [_p, _i, _p, _i, _f, _p, _i64, _i64, _p]),
    "hdmoe_edm_heun_pre": (_i, [_p, _p, _f, _f, _f, _p, _p, _i, _i64, _p]),
    "hdmoe_edm_heun_euler": (_i, [_p, _p, _i, _p, _p, _i, _f, _f, _f, _f, _p, _p, _p, _i64, _p]),
    "hdmoe_edm_heun_correct": (_i, [_p, _p, _p, _i, _p, _p, _i, _f, _f, _f, _f, _p, _p, _i64, _p]),
    "hdmoe_wprep_fwd": (_i, [_p, _p, _i, _i, _p]),
    "hdmoe_gconv2_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i64, _p, _p, _i, _p, _p, _p, _i, _p, _f, _f, _p]),
    "hdmoe_gconv_wgrad": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i64, _p, _p, _i, _p, _p, _p]),
    "hdmoe_nhwc_pixnorm_silu_fwd": (_i, [_p, _p, _p, _i64, _i, _p]),
    "hdmoe_nhwc_pixnorm_silu_bwd": (_i, [_p, _p, _p, _p, _i64, _i, _p]),
    "hdmoe_nhwc_gain_silu_fwd": (_i, [_p, _p, _p, _i64, _i64, _i, _p]),
    "hdmoe_nhwc_gain_silu_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i, _p]),
    "hdmoe_nhwc_axpby": (_i, [_p, _p, _f, _f, _p, _i64, _p]),
    "hdmoe_nhwc_scale2": (_i, [_p, _f, _f, _p, _p, _i64, _p]),
    "hdmoe_nhwc_cat": (_i, [_p, _p, _f, _f, _i, _i, _p, _i64, _p]),
    "hdmoe_nhwc_split": (_i, [_p, _f, _f, _i, _i, _p, _p, _i64, _p]),
    "hdmoe_nchw_to_nhwc": (_i, [_p, _p, _i64, _i, _i, _i64, _i, _p]),
    "hdmoe_nhwc_to_nchw": (_i, [_p, _p, _i64, _i, _i, _i64, _p]),
    "hdmoe_lin32_wgrad": (_i, [_p, _p, _p, _i64, _p]),
    "hdmoe_vit_block_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _i, _p, _p]),
    "hdmoe_vit_block_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "hdmoe_gn1_relu_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "hdmoe_gn1_relu_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "hdmoe_gn1_relu_fwd_t": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _f, _i, _p]),
    "hdmoe_gn1_relu_bwd_t": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "hdmoe_attn_d4_tc_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "hdmoe_attn_d4_tc_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "hdmoe_wprep_fwd_resident": (_i, [_p, _i, _i, _i, _p]),
    "hdmoe_wprep_bwd_multi_resident": (_i, [_p, _i, _i, _p]),
    "hdmoe_wprep_bwd_multi": (_i, [_p, _p, _i, _p]),
    "hdmoe_wprep_bwd": (_i, [_p, _p, _p, _f, _i, _i, _p, _p, _p]),
    "hdmoe_trunk_swap_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i64, _p]),
    "hdmoe_trunk_swap_slices": (_i, []),
    "hdmoe_trunk_swap_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _p]),
    "hdmoe_trunk_gate_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _f, _f, _f, _f, _p]),
    "hdmoe_trunk_gate_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _f, _f, _f, _f, _p]),
    "hdmoe_analytic_scaling": (_i, [_p, _f, _f, _p, _i, _p]),
    "hdmoe_scale_pair_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i64, _p]),
    "hdmoe_scale_pair_tiles": (_i, [_i64]),
    "hdmoe_scale_pair_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i64, _p]),
    "hdmoe_scaling_router_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _f, _p, _i, _i, _p]),
    "hdmoe_scaling_router_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "hdmoe_sqerr_rows": (_i, [_p, _p, _p, _i, _i64, _p]),
    "hdmoe_sqerr_rows_bwd": (_i, [_p, _p, _p, _p, _i, _i64, _p]),
    "hdmoe_peer_barrier": (_i, [_p, _p, _p, _i, _i, _p]),
    "hdmoe_peer_pull": (_i, [_p, _p, _i64, _i, _i, _p]),
    "hdmoe_train_inputs": (_i, [_p, _p, _p, _p, _i, _i64, C.POINTER(MaskGenDesc), _p, C.POINTER(MaskGenDesc), _p, _p]),
    "hdmoe_optim_chunk_elems": (_i, []),
    "hdmoe_adamw_step": (_i, [_p, _i, _i, _p, _p, _p, _f, _f, _f, _f, _i, _p]),
}

_lib = None


def lib():
    """The loaded shared library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                               "hdmoe_b200 has no CPU or eager fallback.")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(_lib, name)      # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().hdmoe_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().hdmoe_launch_count())
