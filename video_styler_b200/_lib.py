"""ctypes binding of libwvd.so (include/wvd.h).  Fails loudly: there is no CPU or PyTorch fallback.

The library is built in-tree by ``python -m video_styler_b200.build`` (nvcc, sm_100a).  Loading works on a
CPU-only box (symbol checks); any compute call needs a B200.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwvd.so")

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> argtypes ; every function returns int except wvd_last_error
SIGNATURES = {
    "wvd_version": [],
    "wvd_sm_arch": [],
    "wvd_debug_flags": [ctypes.POINTER(ctypes.c_ulonglong)],
    "wvd_debug_attention_resident_ctas": [ctypes.c_int],
    "wvd_ln_modulate": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                        c_float, c_int, c_void_p],
    "wvd_qk_rmsnorm_rope": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                            c_int64, c_int64, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int64,
                            c_int, c_void_p],
    "wvd_scale_add": [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_int, c_void_p],
    "wvd_cfg_euler_step": [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_int64, c_int, c_void_p],
    "wvd_tile_blend": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "wvd_tile_finalize": [c_void_p, c_void_p, c_int, c_int64, c_int, c_float, c_float, c_int, c_void_p],
    "wvd_editor_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int,
                        c_int, c_int, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "wvd_attention_bias_fwd": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                               c_int64, c_int, c_int64, c_int64, c_int, c_float, c_int, c_void_p],
    "wvd_gate_residual": [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p],
    "wvd_gemm_bf16": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                      c_int, c_void_p, c_void_p, c_int64, c_void_p],
    "wvd_gemm_bf16_select": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                             c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "wvd_gemm_bf16_grouped": [c_void_p, c_int64, ctypes.POINTER(c_void_p), c_int64, ctypes.POINTER(c_void_p), c_void_p,
                              c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p],
    "wvd_gemm_f32": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                     c_int, c_void_p, c_void_p, c_int64, c_void_p],
    "wvd_attention_fwd": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int64,
                          c_int64, c_int, c_float, c_void_p],
    "wvd_attention_fwd_select": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int64,
                                 c_int64, c_int, c_float, c_int, c_void_p],
    "wvd_attention_fwd_f32": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int,
                              c_int64, c_int64, c_int, c_float, c_void_p],
    "wvd_ulysses_pack_qkv": [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p],
    "wvd_ulysses_unpack_out": [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p],
    "wvd_ulysses_scatter_qkv": [c_void_p, c_int64, ctypes.POINTER(c_void_p), c_int64, c_int, c_int, c_int, c_int, c_void_p],
    "wvd_ulysses_scatter_v": [c_void_p, c_int64, ctypes.POINTER(c_void_p), c_int64, c_int, c_int, c_int, c_int, c_void_p],
    "wvd_qk_rmsnorm_rope_scatter": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_int64,
                                    c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int, c_void_p],
    "wvd_attention_fwd_scatter": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, ctypes.POINTER(c_void_p), c_int64,
                                  c_int64, c_int64, c_int, c_int, c_int64, c_int64, c_int, c_float, c_int, c_void_p],
}
MAX_PEERS = 8

WVD_BF16, WVD_F32 = 0, 1
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_BIAS_GATE_RES, EPI_BIAS_MUL, EPI_BIAS_GELU_T5 = 0, 1, 2, 3, 4, 5
ATTN_AUTO, ATTN_TWO_TILE, ATTN_PAIR, ATTN_CG2, ATTN_ONE_TILE, ATTN_CG2_PERSISTENT = 0, 1, 2, 3, 4, 5
GEMM_AUTO, GEMM_1CTA, GEMM_2CTA, GEMM_2CTA_M512 = 0, 1, 2, 3

_lib = None


class WvdError(RuntimeError):
    pass


def load():
    """Load libwvd.so once.  Raises WvdError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WvdError(f"{LIB_PATH} not found: build it with `python -m video_styler_b200.build` "
                       "(there is no CPU / PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.wvd_last_error.restype = ctypes.c_char_p
    lib.wvd_last_error.argtypes = []
    lib.wvd_build_info.restype = ctypes.c_char_p
    lib.wvd_build_info.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = c_int
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().wvd_last_error().decode("utf-8", "replace")
        raise WvdError(f"{what} failed ({rc}): {msg}")


def debug_flags():
    """Synchronise and read (then clear) the in-kernel watchdog record: dict(timeouts, tag, block, thread)."""
    arr = (ctypes.c_ulonglong * 8)()
    check(load().wvd_debug_flags(arr), "wvd_debug_flags")
    return dict(timeouts=int(arr[0]), tag=int(arr[1]), block=int(arr[2]), thread=int(arr[3]),
                gemm_timeouts=int(arr[4]), attn_timeouts=int(arr[5]))
