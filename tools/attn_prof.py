"""Developer tool: per-phase cycle counters of one attention CTA (see wvd_debug_attention_profile)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops

n, h = int(os.environ.get("N", 29640)), int(os.environ.get("H", 40))
q = torch.randn(n, 3 * h * 128, device="cuda").bfloat16()
d = h * 128
out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(q[:, :d], q[:, d:2 * d], q[:, 2 * d:], h, out=out)
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
_lib.load().wvd_debug_attention_profile(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(q[:, :d], q[:, d:2 * d], q[:, 2 * d:], h, out=out)
e1.record()
torch.cuda.synchronize()
_lib.load().wvd_debug_attention_profile(None)
b = buf.cpu().tolist()
print(f"kernel {e0.elapsed_time(e1):.3f} ms  emu={os.environ.get('WVD_ATTN_EMU')}")
for w in range(8):
    o = b[w * 8:(w + 1) * 8]
    it = max(o[5], 1)
    print(f"softmax warp {w} (tile {w // 4}): per-iter cycles wait_S {o[0]/it:7.1f} ld {o[1]/it:7.1f} max {o[2]/it:7.1f} exp {o[3]/it:7.1f} st+arrive {o[4]/it:7.1f}  total {sum(o[:5])/it:7.1f}")
o = b[64:72]
it = max(o[4], 1)
print(f"mma issuer: per-iter cycles total {o[0]/it:7.1f} sleeping {o[1]/it:7.1f}  poll loops/iter {o[2]/it:6.1f} sleeps/iter {o[3]/it:6.1f}  cycles per sleep {o[1]/max(o[3],1):7.1f}")
