"""ctypes binding of csrc/libmdhs_b200.so (the C ABI declared in include/mdhs_b200.h).

There is deliberately no fallback: if the library is missing or a kernel reports an error the
call raises, so a GPU box can never silently run on another code path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmdhs_b200.so")
_lib = None


class MdhsError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int64), ("a_mn_major", ctypes.c_int32),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_int64), ("b_mn_major", ctypes.c_int32),
        ("D", ctypes.c_void_p), ("ldd", ctypes.c_int64), ("d_dtype", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("aux_out", ctypes.c_void_p), ("ld_aux_out", ctypes.c_int64),
        ("aux_in", ctypes.c_void_p), ("ld_aux_in", ctypes.c_int64),
        ("act", ctypes.c_int32), ("dact", ctypes.c_int32),
        ("residual", ctypes.c_void_p), ("ldr", ctypes.c_int64), ("r_dtype", ctypes.c_int32),
        ("split_k", ctypes.c_int32), ("bn_hint", ctypes.c_int32),
        ("colsum", ctypes.c_void_p), ("colsumsq", ctypes.c_void_p),
        ("dropout_p", ctypes.c_float), ("dropout_seed", ctypes.c_uint64),
        ("conv_mode", ctypes.c_int32), ("cN", ctypes.c_int32), ("cH", ctypes.c_int32), ("cW", ctypes.c_int32),
        ("cC", ctypes.c_int32), ("cR", ctypes.c_int32), ("cS", ctypes.c_int32), ("c_stride", ctypes.c_int32),
        ("c_pad", ctypes.c_int32),
        ("stat_x", ctypes.c_void_p), ("ld_stat_x", ctypes.c_int64), ("stat_mean", ctypes.c_void_p),
        ("stat_scale", ctypes.c_void_p), ("stat_shift", ctypes.c_void_p), ("stat_relu", ctypes.c_int32),
    ]


# Argument signatures of include/mdhs_b200.h: p = pointer, i = int32, l = int64, f = float, u = uint64.
# Every function returns int status and takes the stream as its last pointer.
SIGNATURES = {
    "mdhs_gemm_bf16": "pp",
    "mdhs_layernorm_fwd": "pilpppl" "ppp" "iiffup",
    "mdhs_layernorm_bwd": "pilpilppp" "plppppp" "iifufup",
    "mdhs_bn_finalize": "pplppppffppppiip",
    "mdhs_bn_apply": "pppppliip",
    "mdhs_bn_fwd": "ppppppp" "ff" "ppppppp" "liii" "p",
    "mdhs_bn_bwd": "pppppppppppppppp" "liiii" "p",
    "mdhs_col_stats": "plppplip",
    "mdhs_im2col_nchw_f32": "ppiiiiiiiiip",
    "mdhs_im2col_nchw_f32_tta": "ppiiiiiiiiiiip",
    "mdhs_im2col_nhwc": "ppiiiiiiiip",
    "mdhs_col2im_nhwc": "pppiiiiiiiip",
    "mdhs_maxpool3x3s2_fwd": "pppiiiip",
    "mdhs_maxpool3x3s2_bwd": "pppiiiip",
    "mdhs_tta_expand": "ppiiiiiip",
    "mdhs_mean_tokens_fwd": "pppiiifip",
    "mdhs_mean_tokens_bwd": "pppiiifp",
    "mdhs_conv_weight_pack": "ppiiiiip",
    "mdhs_conv_weight_pack_dgrad": "ppiiiip",
    "mdhs_conv_wgrad_unpack": "ppiiiiip",
    "mdhs_cast_f32_bf16": "pplp",
    "mdhs_cast_bf16_f32": "pplp",
    "mdhs_nhwc_bf16_to_nchw_f32": "ppiiiip",
    "mdhs_nchw_f32_to_nhwc_bf16": "ppiiiip",
    "mdhs_attention_fwd": "plplplplpp" "iiiii" "ffup",
    "mdhs_attention_bwd": "plplplpplpp" "ppppp" "iiiii" "ffup",
    "mdhs_attention_bwd_workspace": "iii",
    "mdhs_embed_gather": "ppppppiiiip",
    "mdhs_embed_scatter": "ppppppiiiip",
    "mdhs_linear_f32_fwd": "plpppliiiip",
    "mdhs_linear_f32_bwd": "plplppli" "pp" "iiip",
    "mdhs_ce_loss": "plppppiififp",
    "mdhs_axpby_f32": "pplpffp",
    "mdhs_supcon_loss": "plppplppp" "iifp",
    "mdhs_act_dropout_bwd": "ppplifup",
    "mdhs_relu_bwd_f32": "ppplp",
    "mdhs_mul_f32": "ppplp",
    "mdhs_sum64_to_grad": "pppip",
    "mdhs_dropout_f32": "pplfup",
    "mdhs_level_mix_fwd": "ppplip",
    "mdhs_level_mix_bwd": "pppppplip",
    "mdhs_axpby_bf16": "ppplffp",
    "mdhs_global_local": "ppiiiifp",
    "mdhs_preprocess_u8": "ppppiiiiippp",
    "mdhs_lstm_cell_fwd": "pppppiip",
    "mdhs_lstm_cell_bwd": "pppppppiip",
    "mdhs_gru_cell_fwd": "pppppiip",
    "mdhs_gru_cell_bwd": "pppppppiip",
    "mdhs_ibfa_fwd": "plplppiiip",
    "mdhs_ibfa_bwd": "plplppppiiip",
    "mdhs_mp_loss": "ppppppppiip",
    "mdhs_kan_basis_fwd": "plpplii" "p",
    "mdhs_kan_basis_bwd": "plppplii" "i" "p",
    "mdhs_kan_weight_pack": "ppppiiii" "p",
    "mdhs_kan_wgrad_unpack": "ppppppiii" "p",
    "mdhs_moe_gate_fwd": "ppppppppppp" "pp" "iiiii" "p",
    "mdhs_moe_loss": "pppppif" "p",
    "mdhs_moe_gate_bwd": "ppppppp" "f" "pppppppp" "pp" "iiiii" "p",
    "mdhs_randn_f32": "plup",
    "mdhs_moe_combine_fwd": "pppiiii" "p",
    "mdhs_moe_combine_bwd": "pppppiiii" "p",
    "mdhs_dwconv7_fwd": "ppppiiiiip",
    "mdhs_dwconv7_wgrad": "ppppiiiip",
    "mdhs_layer_scale_fwd": "pppplii" "fup",
    "mdhs_layer_scale_bwd": "pppppplii" "fup",
    "mdhs_sq_attn_fwd": "plplplpp" "iiifp",
    "mdhs_sq_attn_bwd": "plplplpp" "plplpl" "iiifp",
    "mdhs_adam_flat": "pppppplfffffifiippip",
    "mdhs_sgd_flat": "ppppplffffiippip",
    "mdhs_step_begin": "pp",
    "mdhs_set_sm_reserve": "i",
    "mdhs_set_gemm_dynamic": "i",
}
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int32, "l": ctypes.c_int64, "f": ctypes.c_float, "u": ctypes.c_uint64}


def lib():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MdhsError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or eager fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mdhs_abi_version.restype = ctypes.c_int
        _lib.mdhs_launch_count.restype = ctypes.c_int64
        for name, sig in SIGNATURES.items():
            fn = getattr(_lib, name)  # AttributeError here = header / library mismatch
            fn.argtypes = [_CT[c] for c in sig]
            fn.restype = ctypes.c_int
    return _lib


def call(name, *args):
    """Invoke one C-ABI entry point; raises MdhsError on a non-zero status."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise MdhsError(f"{name} failed with status {rc}")


def check(rc, what):
    if rc != 0:
        raise MdhsError(f"{what} failed with status {rc}")


def launch_count():
    return int(lib().mdhs_launch_count())
