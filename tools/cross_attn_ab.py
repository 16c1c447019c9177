"""Text cross-attention (Sq = 29,640 video tokens, Sk = 512 text tokens, 40 heads, 0.311 TFLOP) A/B on one B200:
the one-tile kernel (two CTAs per SM), the two-tile kernel (one CTA per SM) and torch SDPA (cuDNN), each timed
back to back over many launches (CUDA events), rotating order.  `python tools/cross_attn_ab.py [--sq 29640] [--heads 40]`"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sq", type=int, default=29640)
ap.add_argument("--sk", type=int, default=512)
ap.add_argument("--heads", type=int, default=40)
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
dev = "cuda"
h, d = a.heads, a.heads * 128
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(a.sq, d, device=dev, generator=g).bfloat16()
k = torch.randn(a.sk, d, device=dev, generator=g).bfloat16()
v = torch.randn(a.sk, d, device=dev, generator=g).bfloat16()
out = torch.empty_like(q)
qh, kh, vh = (t.view(-1, h, 128).transpose(0, 1)[None] for t in (q, k, v))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def sdpa():
    return F.scaled_dot_product_attention(qh, kh, vh)


contenders = [("one_tile(2 CTA/SM)", lambda: ops.attention(q, k, v, h, out=out, kernel=_lib.ATTN_ONE_TILE)),
              ("two_tile", lambda: ops.attention(q, k, v, h, out=out, kernel=_lib.ATTN_TWO_TILE)),
              ("cg2_persistent", lambda: ops.attention(q, k, v, h, out=out, kernel=_lib.ATTN_CG2_PERSISTENT)),
              ("torch sdpa", sdpa)]
ref = sdpa()[0].transpose(0, 1).reshape(a.sq, d).float()
for name, fn in contenders[:3]:
    fn()
    torch.cuda.synchronize()
    err = float((out.float() - ref).norm() / ref.norm())
    print(f"{name}: relL2 vs sdpa {err:.2e}  timeouts {_lib.debug_flags()['timeouts']}")
fl = 4.0 * a.sq * a.sk * h * 128
for rnd in range(4):
    line = []
    for name, fn in contenders[rnd % 4:] + contenders[:rnd % 4]:
        for _ in range(5):
            fn()
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        line.append(f"{name} {ms:.3f} ms ({fl / ms / 1e9:.0f} TFLOP/s)")
    print(" | ".join(line), flush=True)
