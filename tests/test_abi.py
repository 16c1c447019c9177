"""The C-ABI library loads on a CPU-only box and exports every symbol include/wvd.h declares; argument validation
returns error codes + messages without touching the GPU."""
import ctypes
import os
import re

import pytest

from video_styler_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "wvd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wvd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from video_styler_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wvd.h but not exported"
    assert set(names) == set(_lib.SIGNATURES) | {"wvd_last_error", "wvd_build_info"}


def test_library_was_built_from_the_sources_in_the_tree():
    """The GPU box runs the libwvd.so that travelled with the snapshot: its baked-in source hash must be the hash of
    the csrc/ + include/ files next to it (a stale prebuilt library fails here, on the CPU suite and on the box)."""
    from video_styler_b200 import build
    info = _lib.load().wvd_build_info().decode()
    assert info.endswith("src=" + build.source_hash()), (info, build.source_hash())


def test_version_and_arch():
    lib = _lib.load()
    assert lib.wvd_version() >= 200
    assert lib.wvd_sm_arch() == 100


def test_argument_validation_without_gpu():
    lib = _lib.load()
    rc = lib.wvd_gemm_bf16(None, 8, None, 8, None, None, 8, 1, 8, 8, 0, None, None, 0, None)
    assert rc == -1 and b"null" in lib.wvd_last_error()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.addressof(buf)
    p = (p + 15) & ~15
    rc = lib.wvd_gemm_bf16(p, 8, p, 8, None, p, 8, 4, 7, 8, 0, None, None, 0, None)      # N not a multiple of 8
    assert rc == -1 and b"multiples of 8" in lib.wvd_last_error()
    rc = lib.wvd_gemm_bf16(p, 8, p, 8, None, p, 8, 4, 8, 8, 3, None, None, 0, None)      # gate epilogue without gate
    assert rc == -1
    rc = lib.wvd_attention_fwd(p, 128, p, 128, p, 128, p, 128, 1, 8, 8, 64, 0.1, None)   # head_dim 64
    assert rc == -1 and b"head_dim" in lib.wvd_last_error()
    rc = lib.wvd_ln_modulate(p, 256, p, None, None, None, p, 256, 4, 256, 1e-6, 0, None)  # shift without scale
    assert rc == -1
    rc = lib.wvd_ulysses_pack_qkv(p, 3 * 5 * 128, p, 4, 5, 128, 2, None)                 # heads not divisible
    assert rc == -1 and b"divide" in lib.wvd_last_error()
    ptrs = (ctypes.c_void_p * 8)(*([p] * 8))
    rc = lib.wvd_ulysses_scatter_qkv(p, 3 * 4 * 128, ptrs, 4, 4, 128, 9, 0, None)             # more than WVD_MAX_PEERS ranks
    assert rc == -1 and b"world" in lib.wvd_last_error()
    rc = lib.wvd_attention_fwd_scatter(p, 128, p, 128, p, 128, ptrs, 256, 4, 200, 2, 1, 8, 8, 128, 0.1, 0, None)   # columns past ldo
    assert rc == -1 and b"col_offset" in lib.wvd_last_error()
    rc = lib.wvd_attention_fwd_select(p, 128, p, 128, p, 128, p, 128, 1, 8, 8, 128, 0.1, 7, None)               # unknown kernel
    assert rc == -1 and b"selector" in lib.wvd_last_error()
    # empty inputs are accepted as no-ops
    assert lib.wvd_ln_modulate(p, 256, None, None, None, None, p, 256, 0, 256, 1e-6, 0, None) == 0
    assert lib.wvd_scale_add(p, p, 1.0, p, 0, 0, None) == 0
    with pytest.raises(_lib.WvdError):
        _lib.check(-1, "x")
