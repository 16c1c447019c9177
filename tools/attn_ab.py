"""Sustained A/B of the self-attention kernels at the c3 shape (29,640 tokens, 40 heads) under the power cap: each
contender runs back to back for --secs seconds in rotating order (tools/gemm_sustained.py explains why).  cuDNN / flash
SDPA (torch F.scaled_dot_product_attention) runs beside them."""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_styler_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--secs", type=float, default=0.7)
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--tokens", type=int, default=29640)
ap.add_argument("--heads", type=int, default=40)
ap.add_argument("--out", default="gpurun_out/attn_ab.txt")
ap.add_argument("--lib", default=None, help="developer: another build of libwvd.so (e.g. a -D variant)")
a = ap.parse_args()
if a.lib:
    _lib.LIB_PATH = os.path.abspath(a.lib)
DEV = "cuda"
n, h = a.tokens, a.heads
d = h * 128
g = torch.Generator(device=DEV).manual_seed(0)
qkv = torch.randn(n, 3 * d, device=DEV, generator=g).bfloat16()
out = torch.empty(n, d, device=DEV, dtype=torch.bfloat16)
q4 = qkv[:, :d].reshape(1, n, h, 128).transpose(1, 2)
k4 = qkv[:, d:2 * d].reshape(1, n, h, 128).transpose(1, 2)
v4 = qkv[:, 2 * d:].reshape(1, n, h, 128).transpose(1, 2)
cont = [("sdpa", lambda: F.scaled_dot_product_attention(q4, k4, v4)),
        ("wvd-pair(cg1)", lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], h, out=out, kernel=_lib.ATTN_PAIR)),
        ("wvd-cg2", lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], h, out=out, kernel=_lib.ATTN_CG2)),
        ("wvd-cg2-persistent", lambda: ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], h, out=out, kernel=_lib.ATTN_CG2_PERSISTENT))]
fl = 4.0 * n * n * d
ref = F.scaled_dot_product_attention(q4, k4, v4).transpose(1, 2).reshape(n, d)
lines = []
for nm, fn in cont[1:]:
    fn(); torch.cuda.synchronize()
    rel = float((out.float() - ref.float()).norm() / ref.float().norm())
    lines.append(f"# {nm}: rel_l2 vs SDPA {rel:.2e}")
res = {nm: [] for nm, _ in cont}
for r in range(a.rounds):
    order = cont[r % len(cont):] + cont[:r % len(cont)]
    for nm, fn in order:
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        cnt = 0
        while time.perf_counter() - t0 < a.secs:
            for _ in range(5):
                fn()
            cnt += 5
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        res[nm].append(e0.elapsed_time(e1) / cnt)
lines.append(f"# sustained ms per launch ({n} tokens, {h} heads; {a.secs} s per contender, rotating order) and TFLOP/s")
for nm, vs in res.items():
    lines.append("  %-16s %s ms   %s TFLOP/s" % (nm, " / ".join("%.3f" % v for v in vs), " / ".join("%.0f" % (fl / (v * 1e-3) / 1e12) for v in vs)))
print("\n".join(lines))
print("flags", _lib.debug_flags())
os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
open(a.out, "w").write("\n".join(lines) + "\n")
