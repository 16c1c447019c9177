import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from video_styler_b200 import _lib, ops
torch.manual_seed(0)
for (sq, sk, h) in [(512, 512, 1), (384, 512, 1), (1280, 1280, 12), (1000, 777, 3), (300, 512, 2), (72, 72, 2), (4096, 4096, 8)]:
    q = torch.randn(sq, h * 128, device="cuda").bfloat16(); k = torch.randn(sk, h * 128, device="cuda").bfloat16(); v = torch.randn(sk, h * 128, device="cuda").bfloat16()
    t0 = time.time()
    o = ops.attention(q, k, v, h)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ref = F.scaled_dot_product_attention(q.view(1, sq, h, 128).transpose(1, 2), k.view(1, sk, h, 128).transpose(1, 2), v.view(1, sk, h, 128).transpose(1, 2)).transpose(1, 2).reshape(sq, h * 128)
    rel = float((o.float() - ref.float()).norm() / ref.float().norm())
    # per 128-row tile error
    errs = [float((o[i:i+128].float() - ref[i:i+128].float()).norm() / ref[i:i+128].float().norm()) for i in range(0, sq, 128)]
    print(f"sq {sq} sk {sk} h {h}: rel {rel:.3e} per-tile {['%.2e' % e for e in errs]} time {dt:.2f}s flags {_lib.debug_flags()}", flush=True)
