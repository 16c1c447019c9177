"""Developer tool: A/B timing of several builds of libwvd (tools/build_attn_variants.sh) on the c3 self-attention shape,
interleaved in one process so that all variants see the same thermal / power state.  Also times torch SDPA (cuDNN).

    python tools/attn_ab.py [variant ...]        # names under video_styler_b200/variants/, default: all + the main lib
"""
import ctypes, glob, os, statistics, sys
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_styler_b200 import _lib

n, h = int(os.environ.get("N", 29640)), int(os.environ.get("H", 40))
reps, inner = int(os.environ.get("REPS", 5)), int(os.environ.get("INNER", 20))
names = sys.argv[1:] or [os.path.basename(p)[len("libwvd_"):-3] for p in sorted(glob.glob(os.path.join(ROOT, "video_styler_b200/variants/libwvd_*.so")))]
libs = {"main": ctypes.CDLL(_lib.LIB_PATH)}
for nm in names:
    libs[nm] = ctypes.CDLL(os.path.join(ROOT, "video_styler_b200/variants", f"libwvd_{nm}.so"))
for lib in libs.values():
    lib.wvd_attention_fwd.restype = ctypes.c_int
    lib.wvd_attention_fwd.argtypes = _lib.SIGNATURES["wvd_attention_fwd"]

d = h * 128
qkv = torch.randn(n, 3 * d, device="cuda").bfloat16()
out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
q4 = qkv[:, :d].reshape(1, n, h, 128).transpose(1, 2)
k4 = qkv[:, d:2 * d].reshape(1, n, h, 128).transpose(1, 2)
v4 = qkv[:, 2 * d:].reshape(1, n, h, 128).transpose(1, 2)
stream = torch.cuda.current_stream().cuda_stream


def run(lib):
    rc = lib.wvd_attention_fwd(qkv.data_ptr(), 3 * d, qkv.data_ptr() + 2 * d, 3 * d, qkv.data_ptr() + 4 * d, 3 * d,
                               out.data_ptr(), d, h, n, n, 128, 128 ** -0.5, stream)
    assert rc == 0


fns = {nm: (lambda lib=lib: run(lib)) for nm, lib in libs.items()}
fns["sdpa"] = lambda: F.scaled_dot_product_attention(q4, k4, v4)
ref = None
times = {nm: [] for nm in fns}
for r in range(reps + 1):
    for nm, fn in fns.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            times[nm].append(e0.elapsed_time(e1) / inner)
flops = 4.0 * n * n * 128 * h
for nm, t in times.items():
    med = statistics.median(t)
    print(f"{nm:12s} median {med:7.3f} ms  min {min(t):7.3f}  max {max(t):7.3f}   {flops / med / 1e9:7.0f} TFLOP/s")
