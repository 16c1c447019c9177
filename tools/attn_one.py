"""Developer tool: a few launches of the c3-shaped self-attention (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import ops
n, h = int(os.environ.get("N", 29640)), int(os.environ.get("H", 40))
q = torch.randn(n, 3 * h * 128, device="cuda").bfloat16()
d = h * 128
out = torch.empty(n, d, device="cuda", dtype=torch.bfloat16)
for _ in range(int(os.environ.get("REPS", 3))):
    ops.attention(q[:, :d], q[:, d:2 * d], q[:, 2 * d:], h, out=out)
torch.cuda.synchronize()
print("ok")
