"""Developer tool: per-phase cycle counters of one attention CTA (see wvd_debug_attention_profile).
Needs a library built with WVD_NVCC_FLAGS=-DWVD_ATTN_PROF python -m video_styler_b200.build --force."""
import os, sys
import ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops
_lib.load().wvd_debug_attention_profile.argtypes = [ctypes.c_void_p]      # a bare int would be truncated to 32 bits

n, h = int(os.environ.get("N", 29640)), int(os.environ.get("H", 40))
q = torch.randn(n, 3 * h * 128, device="cuda").bfloat16()
d = h * 128
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
print(f"kernel {e0.elapsed_time(e1):.3f} ms  emu={os.environ.get('WVD_ATTN_EMU')}")
for w in range(16):
    o = b[w * 8:(w + 1) * 8]
    it = max(o[5], 1)
    print(f"softmax warp {w:2d} (tile {w // 8} half {(w // 4) % 2}): per-iter cycles wait_S {o[0]/it:7.1f} ld {o[1]/it:7.1f} rowmax {(o[7]>>32)/it:6.1f} xchg {(o[7]&0xffffffff)/it:6.1f} vote {o[2]/it:6.1f} turn {o[6]/it:7.1f} exp+handover {o[3]/it:7.1f}  total {(sum(o[:4])+o[6]+(o[7]>>32)+(o[7]&0xffffffff))/it:7.1f}  reference moves {o[4]} in {o[5]} steps")
