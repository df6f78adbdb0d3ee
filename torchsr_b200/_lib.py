"""ctypes binding of libtorchsr_b200.so (include/torchsr_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100, every entry
point raises. The library is built in-tree by ``torchsr_b200/build.py`` (``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSR_LIB_PATH") or os.path.join(_HERE, "lib", "libtorchsr_b200.so")   # override: A/B of two builds

MAX_TAPS = 81
OUT_LINEAR, OUT_SHUFFLE, OUT_UNSHUFFLE, OUT_GEMM_T_ATOMIC, OUT_GATHER_W = 0, 1, 2, 3, 4
ACT_NONE, ACT_PRELU, ACT_LEAKY, ACT_RELU = 0, 1, 2, 3

(E_IM2ROW, E_GATHER_OUT, E_NCHW2NHWC, E_NHWC2NCHW, E_BN_FINALIZE, E_BN_EVAL_COEF, E_BN_ACT, E_BN_BWD_REDUCE,
 E_BN_BWD_FINALIZE, E_BN_BWD_APPLY, E_ACT_BWD, E_COLSUM_FINALIZE, E_SUM_FINALIZE, E_PACK_W, E_UNPACK_G,
 E_LINEAR_WGRAD, E_LOSS, E_ZERO, E_UPSAMPLE2X, E_UPSAMPLE2X_BWD, E_HEAD, E_HEAD_BWD, E_AXPBY, E_MAXPOOL2,
 E_MAXPOOL2_BWD, E_CAST, E_ADAM, E_CHANSUM_NCHW, E_GAN_LOSS, E_AXPBY_F32, E_FEAT_T, E_CROP_LR, E_PACK_GATHER) = range(1, 34)

PK_FWD, PK_T, PK_ROWK, PK_ROWN, PK_ROWN_T, PK_FULLK, PK_LINEAR = range(7)


class ConvDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p),
        ("N", C.c_int64), ("H", C.c_int64), ("W", C.c_int64), ("C", C.c_int64),
        ("x_ld", C.c_int64), ("Ho", C.c_int64), ("Wo", C.c_int64),
        ("gemm_M", C.c_int64), ("gemm_K", C.c_int64), ("w_rows", C.c_int64), ("w_ld", C.c_int64),
        ("a_mode", C.c_int32), ("stride", C.c_int32),
        ("lower_h", C.c_int32), ("lower_w", C.c_int32), ("upper_h", C.c_int32), ("upper_w", C.c_int32),
        ("num_taps", C.c_int32), ("block_k", C.c_int32), ("block_n", C.c_int32), ("cout_pad", C.c_int32),
        ("a_c0", C.c_int32), ("splits", C.c_int32),
        ("tap_off", C.c_uint16 * MAX_TAPS), ("tap_wrow", C.c_uint16 * MAX_TAPS),
        ("out", C.c_void_p), ("out_preact", C.c_void_p), ("bias", C.c_void_p), ("prelu", C.c_void_p),
        ("res", C.c_void_p), ("bwd_z", C.c_void_p), ("dalpha_partial", C.c_void_p), ("stats_partial", C.c_void_p),
        ("os_n", C.c_int64), ("os_h", C.c_int64), ("os_w", C.c_int64),
        ("aux_n", C.c_int64), ("aux_h", C.c_int64), ("aux_w", C.c_int64),
        ("out_mode", C.c_int32), ("out_f32", C.c_int32), ("out_ch_off", C.c_int32), ("aux_ch_off", C.c_int32),
        ("n_valid", C.c_int32), ("act", C.c_int32), ("bwd_act", C.c_int32), ("stats_ld", C.c_int32),
        ("shuf_c", C.c_int32),
        ("acc_scale", C.c_float), ("leaky_slope", C.c_float),
        ("res2", C.c_void_p), ("res_scale", C.c_float), ("res2_scale", C.c_float), ("res_cols", C.c_int32),
        ("w_static", C.c_int32),
        ("bnr_x", C.c_void_p), ("bnr_coef", C.c_void_p), ("bnr_prelu", C.c_void_p), ("bnr_act", C.c_int32),
        ("bnr_c", C.c_int32),
        ("ws", C.c_void_p), ("tile_counters", C.c_void_p), ("ws_ld", C.c_int32), ("side", C.c_int32),
        ("trace", C.c_void_p),
        ("group_rows", C.c_int32), ("bnf_mode", C.c_int32),
        ("bnf_counter", C.c_void_p), ("bnf_gamma", C.c_void_p), ("bnf_beta", C.c_void_p), ("bnf_rm", C.c_void_p),
        ("bnf_rv", C.c_void_p), ("bnf_nbt", C.c_void_p), ("bnf_coef", C.c_void_p),
        ("bnf_count", C.c_int64), ("bnf_c", C.c_int32), ("w_chunk_rows", C.c_int32),
        ("bnf_eps", C.c_float), ("bnf_momentum", C.c_float),
        ("bnr_apply", C.c_int32), ("_pad3", C.c_int32), ("bnr_dx", C.c_void_p),
        ("bnr_gamma", C.c_void_p), ("bnr_dgamma", C.c_void_p), ("bnr_dbeta", C.c_void_p), ("bnr_dalpha", C.c_void_p),
        ("bnr_count", C.c_int64),
        ("gather_bias", C.c_void_p), ("gather_k", C.c_int32), ("gather_pad", C.c_int32), ("gather_c", C.c_int32),
        ("gather_rows", C.c_int32),
        ("out_rep2x", C.c_void_p), ("rep_n", C.c_int64), ("rep_h", C.c_int64), ("rep_w", C.c_int64),
        ("rep_ch_off", C.c_int32), ("stride_w", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("out", C.c_void_p),
        ("N", C.c_int64), ("H", C.c_int64), ("W", C.c_int64), ("C", C.c_int64), ("x_ld", C.c_int64),
        ("Ho", C.c_int64), ("Wo", C.c_int64), ("dy_ld", C.c_int64), ("dy_c", C.c_int64),
        ("stride", C.c_int32), ("lower_h", C.c_int32), ("lower_w", C.c_int32), ("upper_h", C.c_int32),
        ("upper_w", C.c_int32), ("num_taps", C.c_int32), ("chan_block", C.c_int32), ("dy_block", C.c_int32),
        ("block_n", C.c_int32), ("cout_valid", C.c_int32), ("x_c0", C.c_int32), ("dy_c0", C.c_int32),
        ("splits", C.c_int32),
        ("tap_off", C.c_uint16 * MAX_TAPS),
    ]


class EltDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("side", C.c_int32), ("p", C.c_void_p * 12), ("i", C.c_int64 * 16),
                ("f", C.c_float * 8)]


class PackEntry(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("mode", C.c_int32),
        ("cout", C.c_int32), ("cin", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("rows_pad", C.c_int32), ("cols_pad", C.c_int32), ("shuffle", C.c_int32),
        ("block_start", C.c_int64), ("count", C.c_int64),
    ]


AD_PLAIN, AD_CONV, AD_LINEAR, AD_CONV_TILE = 0, 1, 2, 3


class AdamEntry(C.Structure):
    _fields_ = [
        ("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
        ("dst_fwd", C.c_void_p), ("dst_t", C.c_void_p), ("numel", C.c_int64), ("block_start", C.c_int64),
        ("mode", C.c_int32), ("cout", C.c_int32), ("cin", C.c_int32), ("kk", C.c_int32),
        ("rows_fwd", C.c_int32), ("cols_fwd", C.c_int32), ("rows_t", C.c_int32), ("cols_t", C.c_int32),
        ("shuffle", C.c_int32), ("_pad", C.c_int32),
    ]


EXPORTS = [
    "tsr_init", "tsr_last_error", "tsr_version", "tsr_conv", "tsr_wgrad", "tsr_elt", "tsr_prog_create",
    "tsr_prog_destroy", "tsr_prog_add_conv", "tsr_prog_add_conv_group", "tsr_prog_add_wgrad", "tsr_prog_add_elt", "tsr_prog_size",
    "tsr_prog_run", "tsr_launch_count", "tsr_check_watchdog", "tsr_conv_bnf_capacity",
]

_lib = None


class TorchSRB200Error(RuntimeError):
    pass


def load():
    """Loads the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TorchSRB200Error(
            f"{LIB_PATH} is missing: the sm_100a kernels are not built. Run `python -m torchsr_b200.build` "
            "(there is no CPU or PyTorch fallback for this path).")
    lib = C.CDLL(LIB_PATH)
    lib.tsr_last_error.restype = C.c_char_p
    lib.tsr_launch_count.restype = C.c_int64
    lib.tsr_prog_create.restype = C.c_void_p
    lib.tsr_prog_destroy.argtypes = [C.c_void_p]
    lib.tsr_prog_destroy.restype = None
    lib.tsr_conv.argtypes = [C.POINTER(ConvDesc), C.c_void_p]
    lib.tsr_wgrad.argtypes = [C.POINTER(WgradDesc), C.c_void_p]
    lib.tsr_elt.argtypes = [C.POINTER(EltDesc), C.c_void_p]
    lib.tsr_prog_add_conv.argtypes = [C.c_void_p, C.POINTER(ConvDesc)]
    lib.tsr_prog_add_wgrad.argtypes = [C.c_void_p, C.POINTER(WgradDesc)]
    lib.tsr_prog_add_conv_group.argtypes = [C.c_void_p, C.POINTER(ConvDesc), C.c_int]
    lib.tsr_prog_add_elt.argtypes = [C.c_void_p, C.POINTER(EltDesc)]
    lib.tsr_prog_size.argtypes = [C.c_void_p]
    lib.tsr_prog_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.tsr_check_watchdog.argtypes = [C.c_void_p]
    lib.tsr_conv_bnf_capacity.argtypes = [C.POINTER(ConvDesc), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _lib = lib
    return lib


def check(code: int):
    if code < 0 or code > 0:
        raise TorchSRB200Error(f"torchsr_b200 error {code}: {load().tsr_last_error().decode()}")


def launch_count() -> int:
    return int(load().tsr_launch_count())
