"""Developer tool: per-phase cycle counters of the softmax warps of one CTA of the experimental CTA-pair attention kernel.
Needs WVD_NVCC_FLAGS=-DWVD_ATTN_PROF python -m video_styler_b200.build --force and WVD_ATTN_KERNEL=2."""
import os, sys
import ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops
_lib.load().wvd_debug_attention_profile.argtypes = [ctypes.c_void_p]      # a bare int would be truncated to 32 bits
n, h = int(os.environ.get("N", 29640)), int(os.environ.get("H", 40))
d = h * 128
q = torch.randn(n, 3 * d, device="cuda").bfloat16()
out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(q[:, :d], q[:, d:2 * d], q[:, 2 * d:], h, out=out)
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
_lib.load().wvd_debug_attention_profile(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(q[:, :d], q[:, d:2 * d], q[:, 2 * d:], h, out=out)
e1.record()
torch.cuda.synchronize()
_lib.load().wvd_debug_attention_profile(None)
b = buf.cpu().tolist()
print(f"kernel {e0.elapsed_time(e1):.3f} ms")
for w in range(8):
    o = b[w * 8:(w + 1) * 8]
    it = max(o[6], 1)
    print(f"softmax warp {w} (warpgroup {w // 4}): per own tile: wait_S {o[0]/it:7.1f} ld {o[1]/it:6.1f} rowmax {o[2]/it:6.1f} m-handoff {o[3]/it:6.1f} decide {o[4]/it:6.1f} exp+handover {o[5]/it:7.1f}  total {sum(o[:6])/it:7.1f} over {o[6]} tiles")
o = b[72:80]
it = max(o[5], 1)
print(f"MMA issuer: per tile: wait K {o[0]/it:6.1f}  wait V {o[1]/it:6.1f}  wait P hand-over0 {o[2]/it:7.1f}  hand-over1 {o[3]/it:7.1f}  issue+commit {o[4]/it:6.1f}  total {sum(o[:5])/it:7.1f}")
