"""Developer tool: CTA timelines of the short-key attention kernels at the text cross-attention shape.
Needs WVD_NVCC_FLAGS=-DWVD_ATTN_PROF python -m video_styler_b200.build --force."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops

n, h, sk = 29640, 40, 512
d = h * 128
q = torch.randn(n, d, device="cuda").bfloat16()
k = torch.randn(sk, d, device="cuda").bfloat16()
v = torch.randn(sk, d, device="cuda").bfloat16()
out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()
import ctypes
lib.wvd_debug_attention_profile.argtypes = [ctypes.c_void_p]
for which, name in ((_lib.ATTN_ONE_TILE, "one_tile"), (_lib.ATTN_TWO_TILE, "two_tile")):
    print(name, "resident CTAs/SM:", lib.wvd_debug_attention_resident_ctas(which))
    for _ in range(3):
        ops.attention(q, k, v, h, out=out, kernel=which)
    buf = torch.zeros(256, dtype=torch.int64, device="cuda")
    lib.wvd_debug_attention_profile(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention(q, k, v, h, out=out, kernel=which)
    e1.record()
    torch.cuda.synchronize()
    lib.wvd_debug_attention_profile(None)
    b = buf.cpu().tolist()
    print(f"  kernel {e0.elapsed_time(e1):.3f} ms; cycles since CTA entry: setup done | first S seen | main loop done | O complete | stores done | after syncthreads ; end globaltimer ns (rel) ; smid")
    t0 = min(b[128 + 8 * c + 6] for c in range(16) if b[128 + 8 * c + 6])
    for c in range(16):
        o = b[128 + 8 * c:128 + 8 * c + 8]
        print("  CTA", c, " ".join(f"{x:7d}" for x in o[:6]), f"{o[6] - t0:8d}", o[7])
